// Sampling vector -> physical parameter record, per walker, on the device (GWAT_HD code).
//
// Reference being replaced (run on the host, once per likelihood call, with heap allocation of the beta arrays):
//   MCMC_prep_params                         src/mcmc_gw.cpp:2492-2568   flags, f_ref = 20, dCS/EdGB unit change
//   repack_parameters<double>("MCMC_"+m)     src/fisher.cpp:2167-2507    vector -> gen_params
//   tc_ref = T - tc                          src/mcmc_gw.cpp:2466-2473   (T explicit here, see gwat_b200.h)
// and, for the Fisher stencil, the inverse map and the non-MCMC parameterisation:
//   unpack_parameters                        src/fisher.cpp:1841-2162
//   repack_non_parameter_options             src/fisher.cpp:2513-2572
#ifndef GWAT_REPACK_H
#define GWAT_REPACK_H

#include "gwat_model.h"

namespace gwat {

// What the host resolved from the method string + MCMC_modification_struct, passed to the kernels by value.
struct RepackPlan {
	int dimension;
	int pv2;            // 15-dimensional precessing layout instead of the 11-dimensional aligned one
	int nrt;            // tidal parameter(s) at index 11 (12)
	int ppe;            // last ppE_Nmod entries are betas (ppE_* and theory-mapped methods)
	int gimr;           // last (phi+sigma+beta+alpha) entries are fractional deviations
	int alpha_unit_fix; // dCS/EdGB: entry [base] is sqrt(alpha) in km -> alpha^2 in s^4 (src/mcmc_gw.cpp:2560-2565)
	int mcmc;           // "MCMC_" parameterisation (sin DEC, cos iota, ln DL, ln Mc); else the physical one
	int sky;            // sky-averaged IMRPhenomD set: ln A0, phic, tc, ln Mc, ln eta, chi_s, chi_a (src/fisher.cpp:40, 2015-2032), then the modifications
	                    // sky && mcmc: the INTRINSIC sampling sets of the reference's tc/phic-maximised runs ("MCMC_" + method with
	                    // sky_average): ln Mc, eta, chi1, chi2 [, ln tidal_s | ln tidal1, ln tidal2] (src/fisher.cpp:2000-2013, 2061-2078)
	                    // or, IMRPhenomPv2, ln Mc, eta, a1, a2, cos tilt1, cos tilt2, phi1, phi2 (:1968-1990), then the modifications
	gwat_b200_mod mod;
};

GWAT_HD void source_defaults(gwat_b200_source &s)
{
	// member defaults of gen_params_base (include/gwat/util.h:125-285)
	s.mass1 = s.mass2 = s.Luminosity_Distance = 0;
	for (int i = 0; i < 3; i++) s.spin1[i] = s.spin2[i] = 0;
	s.tc = 0;
	s.phiRef = 0;
	s.f_ref = 0;
	s.psi = 0;
	s.incl_angle = 0;
	s.RA = s.DEC = s.gmst = 0;
	s.theta = s.phi = s.theta_l = s.phi_l = 0;
	s.tidal1 = s.tidal2 = s.tidal_s = s.tidal_a = s.tidal_weighted = s.delta_tidal_weighted = -1;
	s.diss_tidal1 = s.diss_tidal2 = s.diss_tidal_s = s.diss_tidal_a = s.diss_tidal_weighted = -1;
	s.chip = -1;
	s.phip = -1;
	for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
		s.betappe[i] = s.bppe[i] = 0;
		s.delta_phi[i] = s.delta_sigma[i] = s.delta_beta[i] = s.delta_alpha[i] = 0;
		s.phii[i] = s.sigmai[i] = s.betai[i] = s.alphai[i] = 0;
	}
	s.Nmod = s.Nmod_phi = s.Nmod_sigma = s.Nmod_beta = s.Nmod_alpha = 0;
	s.PNorder = 35;
	s.shift_time = 1;
	s.shift_phase = 1;
	s.sky_average = 0;
	s.tidal_love = 1;
	s.tidal_love_error = 0;
	s.NSflag1 = s.NSflag2 = 0;
	s.dep_postmerger = 0;
	s.equatorial_orientation = 0;
	s.horizon_coord = 0;
	s.cosmology = 0;  // "PLANCK15"
	s.reserved_ = 0;
}

// calculate_mass1 / calculate_mass2 (src/util.cpp:1516-1540)
GWAT_HD double mass1_of(double chirpmass, double eta)
{
	const double etapow = sm::pow(eta, 3. / 5);
	return 1. / 2 * (chirpmass / etapow + sqrt(1. - 4 * eta) * chirpmass / etapow);
}
GWAT_HD double mass2_of(double chirpmass, double eta)
{
	const double etapow = sm::pow(eta, 3. / 5);
	return 1. / 2 * (chirpmass / etapow - sqrt(1. - 4 * eta) * chirpmass / etapow);
}

// both at once: one exponent-3/5 power and one root instead of two of each (the expressions -- and so the bits -- are those of
// mass1_of / mass2_of; in a latency-bound setup thread every libm call is ~1.3 k cycles)
GWAT_HD void masses_of(double chirpmass, double eta, double &m1, double &m2)
{
	const double etapow = sm::pow(eta, 3. / 5);
	const double a = chirpmass / etapow, b = sqrt(1. - 4 * eta) * chirpmass / etapow;
	m1 = 1. / 2 * (a + b);
	m2 = 1. / 2 * (a - b);
}

GWAT_HD double clamped_acos(double x)
{
	// "Fishers don't necessarily respect the bounds of acos" (src/fisher.cpp:2190-2213)
	if (x > 1) return 0;
	if (x < -1) return GWAT_PI;
	return sm::acos(x);
}

// The modification tails shared by both parameterisations (src/fisher.cpp:2399-2505)
GWAT_HD void repack_tails(const double *v, const RepackPlan &plan, gwat_b200_source &s)
{
	const int dim = plan.dimension;
	if (plan.nrt && !plan.pv2) {
		const int at = (plan.sky && plan.mcmc) ? 4 : 11;  // the intrinsic set keeps them behind chi2 (src/fisher.cpp:2420-2431)
		if (s.tidal_love) {
			s.tidal_s = sm::exp(v[at]);
		} else {
			s.tidal1 = sm::exp(v[at]);
			s.tidal2 = sm::exp(v[at + 1]);
		}
	}
	if (plan.ppe) {
		const int base = dim - s.Nmod;
		for (int i = 0; i < s.Nmod; i++) s.betappe[i] = v[base + i];
	} else if (plan.gimr) {
		const int mods = s.Nmod_phi + s.Nmod_sigma + s.Nmod_beta + s.Nmod_alpha;
		int at = dim - mods;
		for (int i = 0; i < s.Nmod_phi; i++) s.delta_phi[i] = v[at++];
		for (int i = 0; i < s.Nmod_sigma; i++) s.delta_sigma[i] = v[at++];
		for (int i = 0; i < s.Nmod_beta; i++) s.delta_beta[i] = v[at++];
		for (int i = 0; i < s.Nmod_alpha; i++) s.delta_alpha[i] = v[at++];
	}
}

// The non-parameter options MCMC_prep_params sets (src/mcmc_gw.cpp:2494-2559).
GWAT_HD void apply_mod_options(const RepackPlan &plan, gwat_b200_source &s)
{
	const gwat_b200_mod &m = plan.mod;
	s.tidal_love = m.tidal_love;
	s.tidal_love_error = m.tidal_love_error;
	s.NSflag1 = m.NSflag1;
	s.NSflag2 = m.NSflag2;
	if (plan.ppe) {
		s.Nmod = m.ppE_Nmod;
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) s.bppe[i] = m.bppe[i];
	} else if (plan.gimr) {
		s.Nmod_phi = m.gIMR_Nmod_phi;
		s.Nmod_sigma = m.gIMR_Nmod_sigma;
		s.Nmod_beta = m.gIMR_Nmod_beta;
		s.Nmod_alpha = m.gIMR_Nmod_alpha;
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
			s.phii[i] = m.gIMR_phii[i];
			s.sigmai[i] = m.gIMR_sigmai[i];
			s.betai[i] = m.gIMR_betai[i];
			s.alphai[i] = m.gIMR_alphai[i];
		}
	}
}

// repack_parameters, "MCMC_" branches with sky_average = true (src/fisher.cpp:2308-2376): the physical part of the intrinsic sets.
// Everything the set does not hold is a constant of the reference's choosing (the maximised likelihoods override most of them).
GWAT_HD void repack_intrinsic_physical(const double *v, bool pv2, gwat_b200_source &s)
{
	masses_of(sm::exp(v[0]), v[1], s.mass1, s.mass2);
	if (pv2) {
		const double th1 = clamped_acos(v[4]), th2 = clamped_acos(v[5]);
		// transform_sph_cart (src/util.cpp:1909-1914)
		s.spin1[0] = v[2] * sm::sin(th1) * sm::cos(v[6]);
		s.spin1[1] = v[2] * sm::sin(th1) * sm::sin(v[6]);
		s.spin1[2] = v[2] * sm::cos(th1);
		s.spin2[0] = v[3] * sm::sin(th2) * sm::cos(v[7]);
		s.spin2[1] = v[3] * sm::sin(th2) * sm::sin(v[7]);
		s.spin2[2] = v[3] * sm::cos(th2);
		s.tc = 0;
		s.phiRef = 0;
		s.RA = 0;
		s.DEC = 0;
		s.psi = 0;
		s.incl_angle = GWAT_PI / 4.;
		s.Luminosity_Distance = 100;
	} else {
		s.Luminosity_Distance = 1000;
		s.spin1[2] = v[2];
		s.spin2[2] = v[3];
		s.phiRef = 0;
		s.tc = 0;
		s.incl_angle = 0;
	}
}

// The transcendental part of the extrinsic repack below, cut into four independent PARTS so that the cooperative setup kernel can give
// each to another warp (k_setup: the repack was 20-25 k of the kernel's ~83 k cycles, the same ~20 libm calls made by every role):
//   part 0: the masses           part 1: distance, declination, inclination, NRT tidal parameters
//   part 2: spin 1 (Pv2)         part 3: spin 2 (Pv2)
// The expressions are the ones repack_mcmc_walker has always used (it now calls all four parts itself): same bits.
struct RepackHeavy {
	double mass1, mass2, DL, DEC, incl, spin1[3], spin2[3], tidal[2];
};
GWAT_HD void repack_heavy_part(int part, const double *v, const RepackPlan &plan, RepackHeavy &h)
{
	if (part == 0) {
		masses_of(sm::exp(v[7]), v[8], h.mass1, h.mass2);
	} else if (part == 1) {
		h.DL = sm::exp(v[6]);
		h.DEC = sm::asin(v[1]);
		h.incl = sm::acos(v[3]);
		if (plan.nrt && !plan.pv2) {  // repack_tails: tidal_s, or tidal1 and tidal2
			h.tidal[0] = sm::exp(v[11]);
			if (!plan.mod.tidal_love) h.tidal[1] = sm::exp(v[12]);
		}
	} else if (plan.pv2) {
		// transform_sph_cart (src/util.cpp:1909-1914)
		const int a = part == 2 ? 9 : 10, ct = part == 2 ? 11 : 12, ph = part == 2 ? 13 : 14;
		double *sp = part == 2 ? h.spin1 : h.spin2;
		const double th = clamped_acos(v[ct]);
		sp[0] = v[a] * sm::sin(th) * sm::cos(v[ph]);
		sp[1] = v[a] * sm::sin(th) * sm::sin(v[ph]);
		sp[2] = v[a] * sm::cos(th);
	}
}

// The rest of the extrinsic repack: flags, copies and the values of `h`.
GWAT_HD void repack_mcmc_assemble(const double *param, const RepackPlan &plan, double gmst, double T_segment, const RepackHeavy &h,
                                  gwat_b200_source &s)
{
	source_defaults(s);
	// MCMC_prep_params
	s.sky_average = 0;
	s.f_ref = 20;
	s.shift_time = 1;
	s.shift_phase = 1;
	s.gmst = gmst;
	s.equatorial_orientation = 0;
	s.horizon_coord = 0;
	apply_mod_options(plan, s);
	// repack_parameters, "MCMC_" branch, sky_average = false
	s.mass1 = h.mass1;
	s.mass2 = h.mass2;
	s.Luminosity_Distance = h.DL;
	s.RA = param[0];
	s.DEC = h.DEC;
	s.psi = param[2];
	s.incl_angle = h.incl;
	s.phiRef = param[4];
	s.tc = param[5];
	if (plan.pv2) {
		for (int i = 0; i < 3; i++) {
			s.spin1[i] = h.spin1[i];
			s.spin2[i] = h.spin2[i];
		}
	} else {
		s.spin1[2] = param[9];
		s.spin2[2] = param[10];
	}
	if (plan.nrt && !plan.pv2) {
		if (s.tidal_love) s.tidal_s = h.tidal[0];
		else {
			s.tidal1 = h.tidal[0];
			s.tidal2 = h.tidal[1];
		}
	}
	// the modification tails (repack_tails) with the dCS / EdGB unit change of MCMC_prep_params
	const int dim = plan.dimension;
	if (plan.ppe) {
		const int base = dim - s.Nmod;
		for (int i = 0; i < s.Nmod; i++) s.betappe[i] = param[base + i];
		if (plan.alpha_unit_fix) {
			const int ab = dim - plan.mod.ppE_Nmod;
			const double x = param[ab] / (GWAT_C_SI / 1000.);
			s.betappe[ab - base] = ((x * x) * x) * x;  // pow_int(x, 4): sequential product (src/util.cpp:1585-1597)
		}
	} else if (plan.gimr) {
		const int mods = s.Nmod_phi + s.Nmod_sigma + s.Nmod_beta + s.Nmod_alpha;
		int at = dim - mods;
		for (int i = 0; i < s.Nmod_phi; i++) s.delta_phi[i] = param[at++];
		for (int i = 0; i < s.Nmod_sigma; i++) s.delta_sigma[i] = param[at++];
		for (int i = 0; i < s.Nmod_beta; i++) s.delta_beta[i] = param[at++];
		for (int i = 0; i < s.Nmod_alpha; i++) s.delta_alpha[i] = param[at++];
	}
	// MCMC_likelihood_extrinsic: tc is measured back from the end of the segment (src/mcmc_gw.cpp:2467,2473)
	s.tc = T_segment - s.tc;
}

// One walker of MCMC_likelihood_wrapper's parameter handling: param[dimension] -> record with tc = T_segment - tc.
GWAT_HD void repack_mcmc_walker(const double *param, const RepackPlan &plan, double gmst, double T_segment,
                                gwat_b200_source &s)
{
	if (plan.sky) {  // the intrinsic sets: no T_segment (the maximised likelihoods set tc themselves, src/mcmc_gw.cpp:2613-2619)
		source_defaults(s);
		// MCMC_prep_params
		s.sky_average = 1;  // mcmc_intrinsic (src/mcmc_gw.cpp:2494)
		s.f_ref = 20;
		s.shift_time = 1;
		s.shift_phase = 1;
		s.gmst = gmst;
		s.equatorial_orientation = 0;
		s.horizon_coord = 0;
		apply_mod_options(plan, s);
		double v[GWAT_B200_MAX_DIM];
		for (int i = 0; i < plan.dimension; i++) v[i] = param[i];
		if (plan.alpha_unit_fix) {
			const int base = plan.dimension - plan.mod.ppE_Nmod;
			const double x = v[base] / (GWAT_C_SI / 1000.);
			v[base] = ((x * x) * x) * x;  // pow_int(x, 4): sequential product (src/util.cpp:1585-1597)
		}
		repack_intrinsic_physical(v, plan.pv2 != 0, s);
		repack_tails(v, plan, s);
		return;
	}
	RepackHeavy h;
	for (int part = 0; part < 4; part++) repack_heavy_part(part, param, plan, h);
	repack_mcmc_assemble(param, plan, gmst, T_segment, h, s);
}

// ---- Fisher stencil: physical record <-> parameter vector ------------------------------------------------------------
// transform_cart_sph (src/util.cpp:1882-1891)
GWAT_HD void cart_to_sph(const double *c, double *sph)
{
	sph[0] = sqrt(c[0] * c[0] + c[2] * c[2] + c[1] * c[1]);
	sph[1] = sm::acos(c[2] / sph[0]);
	sph[2] = sm::atan2(c[1], c[0]);
	if (sph[2] < 0) sph[2] += 2 * GWAT_PI;
}

GWAT_HD double chirpmass_from(double m1, double m2) { return sm::pow(m1 * m2, 3. / 5) / sm::pow(m1 + m2, 1. / 5); }
GWAT_HD double eta_from(double m1, double m2) { return (m1 * m2) / ((m1 + m2) * (m1 + m2)); }

// unpack_parameters, non-sky-averaged branches (src/fisher.cpp:1843-1966, 2038-2160).  `logf[i]` != 0 marks the
// parameters whose derivative is multiplied by the parameter itself afterwards (d/d ln x).
// A0_from_DL / DL_from_A0 (src/util.cpp:1269-1295): the same expression both ways round, in seconds
GWAT_HD double a0_dl_conversion(double chirpmass_sec, double other, bool sky_average)
{
	const double pref = sky_average ? sqrt(GWAT_PI / 30) : sqrt(GWAT_PI * 40. / 192.);
	return pref * chirpmass_sec * chirpmass_sec / other * sm::pow(GWAT_PI * chirpmass_sec, -7. / 6);
}

// the modification parameters at the end of the vector (src/fisher.cpp:2082-2160)
GWAT_HD void unpack_fisher_mods(const gwat_b200_source &in, const RepackPlan &plan, double *v)
{
	const int dim = plan.dimension;
	if (plan.ppe) {
		const int base = dim - in.Nmod;
		for (int i = 0; i < in.Nmod; i++) v[base + i] = in.betappe[i];
	} else if (plan.gimr) {
		int at = dim - (in.Nmod_phi + in.Nmod_sigma + in.Nmod_beta + in.Nmod_alpha);
		for (int i = 0; i < in.Nmod_phi; i++) v[at++] = in.delta_phi[i];
		for (int i = 0; i < in.Nmod_sigma; i++) v[at++] = in.delta_sigma[i];
		for (int i = 0; i < in.Nmod_beta; i++) v[at++] = in.delta_beta[i];
		for (int i = 0; i < in.Nmod_alpha; i++) v[at++] = in.delta_alpha[i];
	}
}

GWAT_HD void unpack_fisher(const gwat_b200_source &in, const RepackPlan &plan, double *v, int *logfac)
{
	const int dim = plan.dimension;
	for (int i = 0; i < dim; i++) logfac[i] = 0;
	if (plan.sky && plan.mcmc) {  // the intrinsic sets (src/fisher.cpp:1968-1990, 2000-2013, 2061-2078): no logarithmic factors
		v[0] = sm::log(chirpmass_from(in.mass1, in.mass2));
		v[1] = eta_from(in.mass1, in.mass2);
		if (plan.pv2) {
			double s1[3], s2[3];
			cart_to_sph(in.spin1, s1);
			cart_to_sph(in.spin2, s2);
			v[2] = s1[0];
			v[3] = s2[0];
			v[4] = sm::cos(s1[1]);
			v[5] = sm::cos(s2[1]);
			v[6] = s1[2];
			v[7] = s2[2];
		} else {
			v[2] = in.spin1[2];
			v[3] = in.spin2[2];
			if (plan.nrt) {
				if (in.tidal_love) v[4] = sm::log(in.tidal_s);
				else {
					v[4] = sm::log(in.tidal1);
					v[5] = sm::log(in.tidal2);
				}
			}
		}
		unpack_fisher_mods(in, plan, v);
		return;
	}
	if (plan.sky) {  // src/fisher.cpp:2015-2032
		logfac[0] = logfac[3] = logfac[4] = 1;
		v[3] = chirpmass_from(in.mass1, in.mass2);
		v[0] = a0_dl_conversion(v[3] * GWAT_MSOL_SEC, in.Luminosity_Distance * GWAT_MPC_SEC, in.sky_average != 0);
		v[1] = in.phiRef;
		v[2] = in.tc;
		v[4] = eta_from(in.mass1, in.mass2);
		v[5] = (in.spin1[2] + in.spin2[2]) / 2.;
		v[6] = (in.spin1[2] - in.spin2[2]) / 2.;
		unpack_fisher_mods(in, plan, v);  // ppE betas / gIMR deviations behind the seven (:2082-2160)
		return;
	}
	v[0] = in.RA;
	// equatorial_orientation: the direction of L (theta_l, phi_l) takes the place of (psi, iota) (src/fisher.cpp:1851-1858,1887-1894,
	// 1916-1923,1944-1951); incl_angle and psi of a stencil point are then derived by transform_orientation_coords (gwat_orient.h)
	const bool eq = in.equatorial_orientation != 0;
	v[2] = eq ? in.theta_l : in.psi;
	v[4] = in.phiRef;
	v[5] = in.tc;
	v[8] = eta_from(in.mass1, in.mass2);
	if (plan.mcmc) {
		v[1] = sm::sin(in.DEC);
		v[3] = eq ? in.phi_l : sm::cos(in.incl_angle);
		v[6] = sm::log(in.Luminosity_Distance);
		v[7] = sm::log(chirpmass_from(in.mass1, in.mass2));
	} else {
		logfac[6] = 1;
		logfac[7] = 1;
		v[1] = in.DEC;
		v[3] = eq ? in.phi_l : in.incl_angle;
		v[6] = in.Luminosity_Distance;
		v[7] = chirpmass_from(in.mass1, in.mass2);
	}
	if (plan.pv2) {
		if (plan.mcmc) {
			double s1[3], s2[3];
			cart_to_sph(in.spin1, s1);
			cart_to_sph(in.spin2, s2);
			v[9] = s1[0];
			v[10] = s2[0];
			v[11] = sm::cos(s1[1]);
			v[12] = sm::cos(s2[1]);
			v[13] = s1[2];
			v[14] = s2[2];
		} else {
			v[9] = in.spin1[2];
			v[10] = in.spin2[2];
			v[11] = in.chip;
			v[12] = in.phip;
		}
	} else {
		v[9] = in.spin1[2];
		v[10] = in.spin2[2];
	}
	if (plan.nrt && !plan.pv2) {
		if (in.tidal_love) {
			logfac[11] = plan.mcmc ? 0 : 1;
			v[11] = sm::log(in.tidal_s);
		} else {
			logfac[11] = logfac[12] = plan.mcmc ? 0 : 1;
			v[11] = sm::log(in.tidal1);
			v[12] = sm::log(in.tidal2);
		}
	}
	unpack_fisher_mods(in, plan, v);
}

// One stencil point: repack_non_parameter_options + repack_parameters (src/fisher.cpp:2513-2572, 2167-2507) starting
// from a default-constructed record, as calculate_derivatives does (src/fisher.cpp:176-177, 385).
GWAT_HD void repack_fisher_point(const double *v, const gwat_b200_source &orig, const RepackPlan &plan,
                                 gwat_b200_source &s)
{
	source_defaults(s);
	// repack_non_parameter_options
	s.sky_average = orig.sky_average;
	s.tidal_love = orig.tidal_love;
	s.tidal_love_error = orig.tidal_love_error;
	s.f_ref = orig.f_ref;
	s.gmst = orig.gmst;
	s.horizon_coord = orig.horizon_coord;
	s.equatorial_orientation = orig.equatorial_orientation;
	s.NSflag1 = orig.NSflag1;
	s.NSflag2 = orig.NSflag2;
	s.shift_time = 0;
	s.shift_phase = orig.shift_phase;
	s.dep_postmerger = orig.dep_postmerger;
	if (plan.ppe) {
		s.Nmod = orig.Nmod;
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) s.bppe[i] = orig.bppe[i];
	} else if (plan.gimr) {
		s.Nmod_phi = orig.Nmod_phi;
		s.Nmod_sigma = orig.Nmod_sigma;
		s.Nmod_beta = orig.Nmod_beta;
		s.Nmod_alpha = orig.Nmod_alpha;
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
			s.phii[i] = orig.phii[i];
			s.sigmai[i] = orig.sigmai[i];
			s.betai[i] = orig.betai[i];
			s.alphai[i] = orig.alphai[i];
		}
	}
	// repack_parameters
	if (plan.sky && plan.mcmc) {  // src/fisher.cpp:2308-2376
		repack_intrinsic_physical(v, plan.pv2 != 0, s);
		repack_tails(v, plan, s);
		return;
	}
	if (plan.sky) {  // src/fisher.cpp:2379-2393
		masses_of(v[3], v[4], s.mass1, s.mass2);
		s.Luminosity_Distance = a0_dl_conversion(v[3] * GWAT_MSOL_SEC, v[0], s.sky_average != 0) / GWAT_MPC_SEC;
		s.tc = v[2];
		s.phiRef = v[1];
		s.spin1[2] = v[5] + v[6];
		s.spin2[2] = v[5] - v[6];
		repack_tails(v, plan, s);  // (never with plan.nrt: the reference's sky-averaged NRT layout collides with ln eta, refused upstream)
		return;
	}
	s.RA = v[0];
	const bool eq = s.equatorial_orientation != 0;  // (src/fisher.cpp:2180-2187, 2242-2249, 2268-2275, 2293-2300)
	if (eq) {
		s.theta_l = v[2];
		s.phi_l = v[3];
	} else
		s.psi = v[2];
	s.phiRef = v[4];
	s.tc = v[5];
	if (plan.mcmc) {
		masses_of(sm::exp(v[7]), v[8], s.mass1, s.mass2);
		s.Luminosity_Distance = sm::exp(v[6]);
		s.DEC = sm::asin(v[1]);
		if (!eq) s.incl_angle = sm::acos(v[3]);
	} else {
		masses_of(v[7], v[8], s.mass1, s.mass2);
		s.Luminosity_Distance = v[6];
		s.DEC = v[1];
		if (!eq) s.incl_angle = v[3];
	}
	if (plan.pv2) {
		if (plan.mcmc) {
			const double th1 = clamped_acos(v[11]), th2 = clamped_acos(v[12]);
			s.spin1[0] = v[9] * sm::sin(th1) * sm::cos(v[13]);
			s.spin1[1] = v[9] * sm::sin(th1) * sm::sin(v[13]);
			s.spin1[2] = v[9] * sm::cos(th1);
			s.spin2[0] = v[10] * sm::sin(th2) * sm::cos(v[14]);
			s.spin2[1] = v[10] * sm::sin(th2) * sm::sin(v[14]);
			s.spin2[2] = v[10] * sm::cos(th2);
		} else {
			s.spin1[2] = v[9];
			s.spin2[2] = v[10];
			s.chip = v[11];
			s.phip = v[12];
		}
	} else {
		s.spin1[2] = v[9];
		s.spin2[2] = v[10];
	}
	repack_tails(v, plan, s);
}

}  // namespace gwat
#endif
