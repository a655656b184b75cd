// Host-side input preparation next to the path: Gauss-Legendre frequency grids.
//
// The reference builds its GAUSSLEG grids with gauleg(log10(f_lower), log10(f_upper), freqs, weights, n) and then
// freqs[i] = pow(10, freqs[i]) (src/ortho_basis.cpp:14-48, src/waveform_util.cpp:3113-3116); the likelihood is then called
// with integration_method "GAUSSLEG" and log10F = true, which multiplies each weight by f ln 10 (src/mcmc_gw.cpp:821-833).
// gwat_b200_set_network takes exactly those arrays; this file provides the grid itself so that a caller does not need GWAT
// for it.  No GPU work here.
#include <cmath>

#include "../../include/gwat_b200.h"

namespace {

// Abscissas and weights of the n-point Gauss-Legendre rule on [x1, x2]: Newton iteration on P_n(z) from the Chebyshev-like
// guess z_i = cos(pi (i - 1/4) / (n + 1/2)), P_n and P_{n-1} by the three-term recurrence, P_n' = n (z P_n - P_{n-1}) / (z^2 - 1),
// w_i = 2 / ((1 - z_i^2) P_n'(z_i)^2).  Roots come in +- pairs, so half of them are computed.
void gauss_legendre(double x1, double x2, int n, double *x, double *w)
{
	const double mid = 0.5 * (x2 + x1), half = 0.5 * (x2 - x1);
	const int m = (n + 1) / 2;
	for (int i = 1; i <= m; i++) {
		// the reference seeds with pi truncated to 3.141592654 and stops at |dz| <= 1e-10; Newton converges quadratically, so the
		// roots agree to rounding whatever the seed -- the same constants are used anyway
		double z = std::cos(3.141592654 * (i - 0.25) / (n + 0.5));
		double dp = 1, z_prev;
		do {
			double p_n = 1.0, p_nm1 = 0.0;
			for (int j = 1; j <= n; j++) {
				const double p_nm2 = p_nm1;
				p_nm1 = p_n;
				p_n = ((2.0 * j - 1.0) * z * p_nm1 - (j - 1.0) * p_nm2) / j;
			}
			dp = n * (z * p_n - p_nm1) / (z * z - 1.0);
			z_prev = z;
			z = z_prev - p_n / dp;
		} while (std::fabs(z - z_prev) > 1e-10);
		x[i - 1] = mid - half * z;
		x[n - i] = mid + half * z;
		w[i - 1] = 2.0 * half / ((1.0 - z * z) * dp * dp);
		w[n - i] = w[i - 1];
	}
}

}  // namespace

extern "C" int gwat_b200_gauss_legendre_grid(double f_lower, double f_upper, int n, int log10F, double *frequencies, double *weights)
{
	if (n < 1 || !frequencies || !weights || !(f_upper > f_lower) || (log10F && !(f_lower > 0))) return GWAT_B200_ERR_ARG;
	if (log10F) {
		gauss_legendre(std::log10(f_lower), std::log10(f_upper), n, frequencies, weights);
		for (int i = 0; i < n; i++) frequencies[i] = std::pow(10., frequencies[i]);
	} else {
		gauss_legendre(f_lower, f_upper, n, frequencies, weights);
	}
	return GWAT_B200_OK;
}
