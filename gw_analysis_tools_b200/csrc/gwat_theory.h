// Theory -> ppE mapping evaluated per walker on the device (GWAT_HD code).
//
// Reference being replaced (host code with std::function lambdas and new[] per likelihood call):
//   assign_mapping                       src/ppE_utilities.cpp:158-359   theory name -> (b_i, beta_i(source)), inspiral/IMR
//   prep_source_parameters, theory block src/waveform_generator.cpp:1383-1410
//   dCS_beta / dCS_phase_factor          src/ppE_utilities.cpp:473-524
//   EdGB_beta / EdGB_phase_factor        src/ppE_utilities.cpp:573-617
//   Z_from_DL, cosmology_interpolation_function   src/util.cpp:356-382, 502-512 (PLANCK15, the gen_params default)
#ifndef GWAT_THEORY_H
#define GWAT_THEORY_H

#include "gwat_model.h"

namespace gwat {

struct DzTable {
	const double *boundaries;   // [segments + 1]
	const double (*coeffs)[12];  // [segments][12]
	int segments;
};

GWAT_HD double pow_int_seq(double base, int power)
{
	// pow_int of the reference: a sequential product starting from 1 (src/util.cpp:1585-1597)
	if (power == 0) return 1.;
	double prod = 1;
	const int n = power < 0 ? -power : power;
	for (int i = 0; i < n; i++) prod = prod * base;
	return power > 0 ? prod : 1. / prod;
}

// redshift from luminosity distance in Mpc: piecewise series in half-integer powers of D_L
GWAT_HD double z_from_dl(double DL_mpc, const DzTable &t)
{
	for (int i = 0; i < t.segments; i++) {
		if (DL_mpc < t.boundaries[i + 1]) {
			const double *c = t.coeffs[i];
			double sum = c[0];
			const double rootx = sqrt(DL_mpc);
			for (int k = 1; k < 12; k++) sum += c[k] * pow_int_seq(rootx, k);
			return sum;
		}
	}
	return -1;
}

GWAT_HD double dcs_phase_factor(const SrcQ &s, double m1, double m2)
{
	const double eta = s.eta;
	const double m = m1 + m2;
	const double chi1 = s.chi_s + s.chi_a, chi2 = s.chi_s - s.chi_a;
	const double s1temp = 2. + 2. * pow_int_seq(chi1, 4) - 2. * sqrt((1. - chi1 * chi1)) - chi1 * chi1 * ((3. - 2. * sqrt(1. - chi1 * chi1)));
	const double s2temp = 2. + 2. * pow_int_seq(chi2, 4) - 2. * sqrt((1. - chi2 * chi2)) - chi2 * chi2 * ((3. - 2. * sqrt(1. - chi2 * chi2)));
	double s1 = fabs(chi1) < 1e-10 ? 0 : s1temp / (2. * chi1 * chi1 * chi1);
	double s2 = fabs(chi2) < 1e-10 ? 0 : s2temp / (2. * chi2 * chi2 * chi2);
	if (s.NSflag1) s1 = 0;  // neutron stars carry no scalar charge
	if (s.NSflag2) s2 = 0;
	double g = 0;
	g += (-5. / 8192.) / (pow(eta, 14. / 5.)) * pow((m1 * s2 - m2 * s1), 2.) / (m * m);
	g += (15075. / 114688.) / (pow(eta, 14. / 5.)) * (m2 * m2 * chi1 * chi1 - 350. / 201. * m1 * m2 * chi1 * chi2 + m1 * m1 * chi2 * chi2) / (m * m);
	return g;
}

GWAT_HD double edgb_phase_factor(const SrcQ &s, double m1, double m2)
{
	const double chi1 = s.chi_s + s.chi_a, chi2 = s.chi_s - s.chi_a;
	const double temp1 = 2. * (sqrt(1. - chi1 * chi1) - 1. + chi1 * chi1);
	const double temp2 = 2. * (sqrt(1. - chi2 * chi2) - 1. + chi2 * chi2);
	double s1 = fabs(chi1) < 1e-10 ? 0 : temp1 / (chi1 * chi1);
	double s2 = fabs(chi2) < 1e-10 ? 0 : temp2 / (chi2 * chi2);
	if (s.NSflag1) s1 = 0;
	if (s.NSflag2) s2 = 0;
	return (-5. / 7168.) * pow_int_seq((m1 * m1 * s2 - m2 * m2 * s1), 2) / (pow_int_seq(s.M, 4) * pow(s.eta, (18. / 5)));
}

// Replace the walker's (beta, b) by what the theory dictates; betappe[0] on entry is the coupling alpha^2 [s^4].
// `theory` uses the TheoryId values of gwat_method.h (1 = dCS, 2 = EdGB).
GWAT_HD void apply_theory(int theory, const DzTable &dz, SrcQ &s)
{
	if (theory == 0) return;
	const double etapow = pow(s.eta, 3. / 5);
	const double root = sqrt(1. - 4 * s.eta);
	const double m1 = 1. / 2 * (s.chirpmass / etapow + root * s.chirpmass / etapow);
	const double m2 = 1. / 2 * (s.chirpmass / etapow - root * s.chirpmass / etapow);
	const double Z = z_from_dl(s.DL / GWAT_MPC_SEC, dz);
	const double unredshiftedM = s.M / (1. + Z);
	const double coupling = s.betappe[0];
	const double factor = theory == 1 ? dcs_phase_factor(s, m1, m2) : edgb_phase_factor(s, m1, m2);
	s.betappe[0] = 16. * GWAT_PI * coupling / (pow_int_seq(unredshiftedM, 4)) * factor;
	s.bppe[0] = theory == 1 ? -1 : -7;
	s.Nmod = 1;
}

}  // namespace gwat
#endif
