// Theory -> ppE mapping evaluated per walker on the device (GWAT_HD code).
//
// Reference being replaced (host code with std::function lambdas and new[] per likelihood call):
//   assign_mapping                       src/ppE_utilities.cpp:158-359   theory name -> (b_i, beta_i(source)), inspiral/IMR
//   prep_source_parameters, theory block src/waveform_generator.cpp:1383-1410
//   dCS_beta / dCS_phase_factor          src/ppE_utilities.cpp:473-524
//   EdGB_beta / EdGB_phase_factor        src/ppE_utilities.cpp:573-617
//   Z_from_DL, cosmology_interpolation_function   src/util.cpp:356-382, 502-512 (all six cosmologies of include/gwat/D_Z_Config.h)
#ifndef GWAT_THEORY_H
#define GWAT_THEORY_H

#include "gwat_model.h"

namespace gwat {

struct DzTable {
	const double (*boundaries)[4];   // [cosmology][segments + 1]
	const double (*coeffs)[3][12];   // [cosmology][segments][12]
	int segments;
	int n_cosmologies;
	// modified-dispersion distance D_alpha(z) (ModDispersion theory): alphas[n], z boundaries [n][4], coefficients [n][3][17]
	const double *md_alphas;
	const double (*md_boundaries_z)[4];
	const double (*md_coeffs)[3][17];
	int md_n;
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
// (cosmology = index into the reference's cosmos[]: PLANCK15, PLANCK13, WMAP9, WMAP7, WMAP5, TESTING_COSMOLOGY; -1 as the
// reference returns it for a name it does not know)
GWAT_HD double z_from_dl(double DL_mpc, int cosmology, const DzTable &t)
{
	if (cosmology < 0 || cosmology >= t.n_cosmologies) return -1;
	for (int i = 0; i < t.segments; i++) {
		if (DL_mpc < t.boundaries[cosmology][i + 1]) {
			const double *c = t.coeffs[cosmology][i];
			double sum = c[0];
			const double rootx = sqrt(DL_mpc);
			for (int k = 1; k < 12; k++) sum += c[k] * pow_int_seq(rootx, k);
			return sum;
		}
	}
	return -1;
}

// DL_from_Z_MD + cosmology_interpolation_function_MD + dispersion_lookup (src/ppE_utilities.cpp:785-835): Mpc, -1 if alpha or z
// is outside the tables
GWAT_HD double dl_from_z_md(double Z, double alpha, const DzTable &t)
{
	int idx = -1;
	for (int i = 0; i < t.md_n; i++)
		if (fabs(alpha - t.md_alphas[i]) < 1e-10) {
			idx = i;
			break;
		}
	if (idx < 0) return -1;
	for (int i = 0; i < 3; i++) {
		if (Z < t.md_boundaries_z[idx][i + 1]) {
			double result = 0;
			for (int j = 0; j < 17; j++) result += t.md_coeffs[idx][i][j] * sm::pow(Z, -3.5 + j * 0.5);
			return result;
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
	g += (-5. / 8192.) / (sm::pow(eta, 14. / 5.)) * pow((m1 * s2 - m2 * s1), 2.) / (m * m);
	g += (15075. / 114688.) / (sm::pow(eta, 14. / 5.)) * (m2 * m2 * chi1 * chi1 - 350. / 201. * m1 * m2 * chi1 * chi2 + m1 * m1 * chi2 * chi2) / (m * m);
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
	return (-5. / 7168.) * pow_int_seq((m1 * m1 * s2 - m2 * m2 * s1), 2) / (pow_int_seq(s.M, 4) * sm::pow(s.eta, (18. / 5)));
}

// Theory ids (gwat_method.h parses the method string into one of these).  "EdGB_HO_<model>" without "_LO" is THEORY_EDGB: in
// assign_mapping the `if (EdGB_HO)` block is followed by an independent `if (EdGB_HO_LO) ... else if (GHOv1..3) ... else`
// chain whose final else overwrites the three-term mapping with the plain EdGB one (src/ppE_utilities.cpp:171-221).
enum TheoryId {
	THEORY_NONE = 0, THEORY_DCS = 1, THEORY_EDGB = 2, THEORY_EDGB_HO_LO = 3, THEORY_EDGB_GHOV1 = 4, THEORY_EDGB_GHOV2 = 5,
	THEORY_EDGB_GHOV3 = 6, THEORY_EXTRADIM = 7, THEORY_BHEVAP = 8, THEORY_TVG = 9, THEORY_DIPRAD = 10, THEORY_NONCOMM = 11,
	THEORY_PNSERIES = 12, THEORY_PPEALT = 13, THEORY_MODDISP = 14
};
// MCMC_prep_params converts sqrt(alpha)[km] -> alpha^2 [s^4] for every method whose name contains dCS or EdGB (src/mcmc_gw.cpp:2560)
GWAT_HD bool theory_alpha_units(int theory) { return theory >= THEORY_DCS && theory <= THEORY_EDGB_GHOV3; }

// Replace the walker's (beta_i, b_i) by what the theory dictates (prep_source_parameters' theory block,
// src/waveform_generator.cpp:1383-1410: every beta function sees the INPUT betappe, results are installed afterwards).
GWAT_HD void apply_theory(int theory, const DzTable &dz, SrcQ &s)
{
	if (theory == THEORY_NONE) return;
	const double etapow = sm::pow(s.eta, 3. / 5);
	const double root = sqrt(1. - 4 * s.eta);
	const double m1 = 1. / 2 * (s.chirpmass / etapow + root * s.chirpmass / etapow);
	const double m2 = 1. / 2 * (s.chirpmass / etapow - root * s.chirpmass / etapow);
	const double Z = z_from_dl(s.DL / GWAT_MPC_SEC, s.cosmology, dz);
	const double unredshiftedM = s.M / (1. + Z);
	const double in0 = s.betappe[0], in1 = s.betappe[1];
	const double eta = s.eta;
	double beta[GWAT_B200_MAX_MOD], b[GWAT_B200_MAX_MOD];
	int n = 1;
	switch (theory) {
	case THEORY_DCS:  // dCS_beta (:473-484)
		beta[0] = 16. * GWAT_PI * in0 / (pow_int_seq(unredshiftedM, 4)) * dcs_phase_factor(s, m1, m2);
		b[0] = -1;
		break;
	case THEORY_EDGB:  // EdGB_beta (:573-583)
		beta[0] = 16. * GWAT_PI * in0 / (pow_int_seq(unredshiftedM, 4)) * edgb_phase_factor(s, m1, m2);
		b[0] = -7;
		break;
	case THEORY_EDGB_HO_LO: {  // EdGB_HO_0PN_beta (:529-539)
		const double alphaSq = 16. * GWAT_PI * in0;
		beta[0] = -5. * alphaSq / pow_int_seq(unredshiftedM, 4) / 7168. / sm::pow(eta, 18. / 5.) * (4 * eta - 1);
		b[0] = -7;
		break;
	}
	case THEORY_EDGB_GHOV1:  // EdGB_beta + EdGB_GHO_betav1 (:619-630)
	case THEORY_EDGB_GHOV2:  // (:633-645)
	case THEORY_EDGB_GHOV3: {  // (:648-661)
		const double pf = edgb_phase_factor(s, m1, m2);
		const double lead = 16. * GWAT_PI * in0 / (pow_int_seq(unredshiftedM, 4));
		beta[0] = lead * pf;
		if (theory == THEORY_EDGB_GHOV1) beta[1] = lead * in1;
		else if (theory == THEORY_EDGB_GHOV2) beta[1] = lead * pf * in1;
		else beta[1] = lead * in1 * pf * (3715. / 756. + 55. * eta / 9.);
		b[0] = -7;
		b[1] = -5;
		n = 2;
		break;
	}
	case THEORY_EXTRADIM: {  // ExtraDimension_beta (:664-678); s.mass1/2 are in seconds
		const double ten_micrometer = 10.e-6 / GWAT_C_SI;
		const double T_year = 31557600.;
		const double m1dot = -2.8e-7 * pow_int_seq(GWAT_MSOL_SEC * (1 + Z) / s.mass1, 2) * in0 / pow_int_seq(ten_micrometer, 2) * GWAT_MSOL_SEC / T_year;
		const double m2dot = -2.8e-7 * pow_int_seq(GWAT_MSOL_SEC * (1 + Z) / s.mass2, 2) * in0 / pow_int_seq(ten_micrometer, 2) * GWAT_MSOL_SEC / T_year;
		beta[0] = (m1dot + m2dot) * (25. / 851968.) * ((3. - 26. * eta + 34. * eta * eta) / (sm::pow(eta, 2. / 5.) * (1 - 2 * eta)));
		b[0] = -13;
		break;
	}
	case THEORY_BHEVAP:  // BHEvaporation_beta (:682-691)
		beta[0] = (in0) * (25. / 851968.) * ((3. - 26. * eta + 34. * eta * eta) / (sm::pow(eta, 2. / 5.) * (1 - 2 * eta)));
		b[0] = -13;
		break;
	case THEORY_TVG:  // TVG_beta (:695-703)
		beta[0] = (-25. / 65526.) * (in0 * s.chirpmass / (1 + Z));
		b[0] = -13;
		break;
	case THEORY_DIPRAD:  // DipRad_beta (:707-713)
		beta[0] = (-3. / 224.) * sm::pow(eta, 2. / 5.) * in0;
		b[0] = -7;
		break;
	case THEORY_NONCOMM:  // NonComm_beta (:717-723)
		beta[0] = (-75. / 256.) * sm::pow(eta, -4. / 5.) * (2. * eta - 1.) * in0;
		b[0] = -1;
		break;
	case THEORY_PNSERIES:  // PNSeries_beta (:420-435): basis M f instead of Mc f; terms > 0 are relative to the first
	case THEORY_PPEALT: {  // ppEAlt_beta (:399-414)
		// calculate_chirpmass(mass1, mass2) (src/util.cpp:1492): the same value as s.chirpmass up to the rounding of its pow calls
		const double chirp = sm::pow(s.mass1 * s.mass2, 3. / 5) / sm::pow(s.mass1 + s.mass2, 1. / 5);
		const double total_m = s.mass1 + s.mass2;
		n = s.Nmod;
		for (int i = 0; i < n; i++) {
			b[i] = s.bppe[i];
			const double conv = sm::pow(total_m / chirp, s.bppe[i] / 3.);
			if (i == 0 || theory == THEORY_PPEALT) beta[i] = s.betappe[i] * conv;
			else beta[i] = s.betappe[0] * s.betappe[i] * conv;
		}
		break;
	}
	case THEORY_MODDISP: {  // ModDispersion_beta (:750-769), arXiv:1110.2720; b = 3 alpha - 3 comes from the caller
		const double alpha = (s.bppe[0] + 3.) / 3.;
		const double Dalpha = dl_from_z_md(Z, alpha, dz) * GWAT_MPC_SEC;
		const double h_planck = 4.135667696e-15;
		beta[0] = (sm::pow(GWAT_PI, 2. - alpha) / (1. - alpha)) * (Dalpha * in0 / sm::pow(h_planck, 2. - alpha)) *
		          (sm::pow(s.chirpmass, 1. - alpha) / sm::pow(1 + Z, 1. - alpha));
		b[0] = s.bppe[0];
		break;
	}
	default:
		beta[0] = NAN;
		b[0] = 0;
	}
	s.Nmod = n;
	for (int i = 0; i < n; i++) {
		s.betappe[i] = beta[i];
		s.bppe[i] = b[i];
	}
}

}  // namespace gwat
#endif
