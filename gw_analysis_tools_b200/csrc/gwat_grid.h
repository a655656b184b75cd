// Host-side preparation of the per-grid tables the kernels stream (done once per gwat_b200_set_network).
//
// The reference evaluates pow(M*f, 1./6.) and log(pi*M*f) for every (walker, bin) (src/IMRPhenomD.cpp:882, 1054).  Both
// separate into a per-walker and a per-bin factor, so the per-bin factors are tabulated here once per grid:
//   sf = f^(fl(1/6)) as a double-double (hi + lo, ~64 significant bits from x87 powl) -- multiplied on the device with
//        the walker's M^(fl(1/6)) double-double and rounded once, which reproduces the correctly-rounded glibc pow of
//        the reference to within one unit in the last place (identical in ~9 cases out of 10);
//   logf = ln f.
// The quadrature rule is folded into one weight per (detector, bin): Simpson's 1,4,2,...,4,1 pattern exactly as
// simpsons_sum applies it (include/gwat/util.h:843-856: by index parity, whatever the parity of L) or the
// Gauss-Legendre weights (src/mcmc_gw.cpp:821-833), divided by the PSD.
#ifndef GWAT_GRID_H
#define GWAT_GRID_H

#include <cmath>
#include <vector>

#include "gwat_hd.h"

namespace gwat {

inline void build_frequency_tables(const double *f, int L, std::vector<double> &sf_hi, std::vector<double> &sf_lo,
                                   std::vector<double> &logf)
{
	sf_hi.resize(L);
	sf_lo.resize(L);
	logf.resize(L);
	const long double e6 = (long double)GWAT_SIXTH;
	for (int i = 0; i < L; i++) {
		const long double s = powl((long double)f[i], e6);
		const double hi = (double)s;
		sf_hi[i] = hi;
		sf_lo[i] = (double)(s - (long double)hi);
		logf[i] = (double)logl((long double)f[i]);
	}
}

// Quadrature coefficient of bin i, without the PSD and without the global prefactor (see quadrature_prefactor).
inline double quadrature_coefficient(int i, int L, bool gaussleg, bool log10F, const double *weights, const double *f)
{
	if (gaussleg) {
		double w = weights[i];
		if (log10F) w = w * f[i] * std::log(10.);
		return w;
	}
	if (i == 0 || i == L - 1) return 1.0;
	return (i % 2 == 0) ? 2.0 : 4.0;
}
// HH = prefactor * sum_i coef_i |r_i|^2 / S_i  (4 * delta_f / 3 for Simpson with the reference's delta_f choice, 4 for GLQ)
inline double quadrature_prefactor(int L, bool gaussleg, const double *f, bool fisher_convention)
{
	if (gaussleg) return 4.0;
	// Log_Likelihood_internal takes delta_f from the middle of the array (src/mcmc_gw.cpp:811), the Fisher and SNR
	// routines from its start (src/fisher.cpp:2747)
	const double delta_f = fisher_convention ? (f[1] - f[0]) : (f[L / 2] - f[L / 2 - 1]);
	return 4. * (delta_f / 3.);
}

}  // namespace gwat
#endif
