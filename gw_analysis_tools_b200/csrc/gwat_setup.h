// Per-walker setup: physical parameters -> everything the per-bin kernels need (GWAT_HD code).
//
// Reference path being replaced, per walker and per likelihood call:
//   source_parameters::populate_source_parameters      src/util.cpp:997-1028
//   prep_source_parameters                              src/waveform_generator.cpp:1255-1415
//   IMRPhenomD<double>::construct_waveform (setup)      src/IMRPhenomD.cpp:404-469
//   inclination factors                                 src/waveform_generator.cpp:183-199
//   DTOA_DETECTOR, detector_response_functions_equatorial   src/detector_util.cpp:676-788, 900-1179
#ifndef GWAT_SETUP_H
#define GWAT_SETUP_H

#include "gwat_nrt.h"
#include "gwat_theory.h"

namespace gwat {

// ---- populate_source_parameters --------------------------------------------------------------------------------------
GWAT_HD double chirpmass_of(double m1, double m2) { return sm::pow(m1 * m2, 3. / 5) / sm::pow(m1 + m2, 1. / 5); }
GWAT_HD double eta_of(double m1, double m2) { return (m1 * m2) / ((m1 + m2) * (m1 + m2)); }
// A0_from_DL (src/util.cpp:1271-1279)
GWAT_HD double a0_from_dl(double chirpmass, double DL, bool sky_average)
{
	const double pref = sky_average ? sqrt(GWAT_PI / 30) : sqrt(GWAT_PI * 40. / 192.);
	return pref * chirpmass * chirpmass / DL * sm::pow(GWAT_PI * chirpmass, -7. / 6);
}

GWAT_HD void populate_source(const gwat_b200_source &in, SrcQ &s)
{
	s.mass1 = in.mass1 * GWAT_MSOL_SEC;
	s.mass2 = in.mass2 * GWAT_MSOL_SEC;
	s.spin1x = in.spin1[0];
	s.spin2x = in.spin2[0];
	s.spin1y = in.spin1[1];
	s.spin2y = in.spin2[1];
	s.spin1z = in.spin1[2];
	s.spin2z = in.spin2[2];
	s.chi_s = (1. / 2) * (s.spin1z + s.spin2z);
	s.chi_a = (1. / 2) * (s.spin1z - s.spin2z);
	s.chirpmass = chirpmass_of(s.mass1, s.mass2);
	s.eta = eta_of(s.mass1, s.mass2);
	s.M = s.mass1 + s.mass2;
	s.chi_eff = (s.mass1 * s.spin1z + s.mass2 * s.spin2z) / s.M;
	s.chi_pn = s.chi_eff - (38 * s.eta / 113) * (2 * s.chi_s);
	s.DL = in.Luminosity_Distance * GWAT_MPC_SEC;
	s.delta_mass = sqrt(1. - 4 * s.eta);
	s.phiRef = in.phiRef;
	s.tc = in.tc;
	s.sky_average = in.sky_average != 0;
	s.A0 = a0_from_dl(s.chirpmass, s.DL, s.sky_average);
	// prep_source_parameters: plain copies
	s.incl_angle = in.incl_angle;
	s.f_ref = in.f_ref;
	s.shift_time = in.shift_time != 0;
	s.shift_phase = in.shift_phase != 0;
	s.NSflag1 = in.NSflag1 != 0;
	s.NSflag2 = in.NSflag2 != 0;
	s.dep_postmerger = in.dep_postmerger != 0;
	s.chip = in.chip;
	s.phip = in.phip;
	s.q = 0;
	s.cosmology = in.cosmology;
	s.Nmod = 0;
	s.Nmod_phi = s.Nmod_sigma = s.Nmod_beta = s.Nmod_alpha = 0;
	s.tidal1 = s.tidal2 = s.tidal_weighted = s.delta_tidal_weighted = s.diss_tidal_weighted = -1;
}

// ---- detectors -------------------------------------------------------------------------------------------------------
// Sky-position trigonometry shared by the antenna patterns and the arrival-time differences of all detectors.  The
// reference recomputes cos/sin of the same (gmst - ra), dec for every detector and again in DTOA_earth_centered_coord
// (src/detector_util.cpp:780-784, 915-922); identical inputs give identical values, so they are evaluated once.
struct SkyTrig {
	double cosgha, singha, cosdec, sindec, cospsi, sinpsi;
};
GWAT_HD SkyTrig sky_trig(double ra, double dec, double psi, double gmst)
{
	SkyTrig t;
	const double gha = gmst - ra;
	sm::sincos(gha, &t.singha, &t.cosgha);
	sm::sincos(dec, &t.sindec, &t.cosdec);
	sm::sincos(psi, &t.sinpsi, &t.cospsi);
	return t;
}

// Antenna patterns of an interferometer with response tensor D (row-major 3x3).  Same construction as LAL's
// XLALComputeDetAMResponse, which the reference transcribes (src/detector_util.cpp:900-1013).
GWAT_HD void antenna_pattern(const double *D, double geometric_factor, const SkyTrig &t, double &Fplus, double &Fcross)
{
	double X[3], Y[3];
	X[0] = -t.cospsi * t.singha - t.sinpsi * t.cosgha * t.sindec;
	X[1] = -t.cospsi * t.cosgha + t.sinpsi * t.singha * t.sindec;
	X[2] = t.sinpsi * t.cosdec;
	Y[0] = t.sinpsi * t.singha - t.cospsi * t.cosgha * t.sindec;
	Y[1] = t.sinpsi * t.cosgha + t.cospsi * t.singha * t.sindec;
	Y[2] = t.cospsi * t.cosdec;
	double fp = 0, fc = 0;
	for (int i = 0; i < 3; i++) {
		const double DX = D[3 * i + 0] * X[0] + D[3 * i + 1] * X[1] + D[3 * i + 2] * X[2];
		const double DY = D[3 * i + 0] * Y[0] + D[3 * i + 1] * Y[1] + D[3 * i + 2] * Y[2];
		fp += X[i] * DX - Y[i] * DY;
		fc += X[i] * DY + Y[i] * DX;
	}
	Fplus = fp * geometric_factor;
	Fcross = fc * geometric_factor;
}

// Arrival-time difference t(loc1) - t(loc2) of a plane wave from (ra, dec)  (DTOA_earth_centered_coord,
// src/detector_util.cpp:766-788).
GWAT_HD double dtoa_between(const double *loc1, const double *loc2, const SkyTrig &t)
{
	const double dx0 = loc1[0] - loc2[0], dx1 = loc1[1] - loc2[1], dx2 = loc1[2] - loc2[2];
	const double e0 = t.cosdec * t.cosgha;
	const double e1 = t.cosdec * -t.singha;
	const double e2 = t.sindec;
	return (dx0 * e0 + dx1 * e1 + dx2 * e2) / GWAT_C_SI;
}
GWAT_HD double dtoa_between(const double *loc1, const double *loc2, double ra, double dec, double gmst)
{
	return dtoa_between(loc1, loc2, sky_trig(ra, dec, 0.0, gmst));
}

// The detector network as the kernels see it: rows of the generated detector table (tensor, location, factor).
struct Network {
	int D;
	int horizon_mode;  // 1: a source with horizon_coord set gets the horizon-frame patterns (the single-detector response path only)
	double row[GWAT_B200_MAX_DETECTORS][13];
};

GWAT_HD void detector_setup(const Network &net, double ra, double dec, double psi, double gmst, DetCoef *out)
{
	const SkyTrig t = sky_trig(ra, dec, psi, gmst);
	for (int d = 0; d < net.D; d++) {
		antenna_pattern(net.row[d], net.row[d][12], t, out[d].Fplus, out[d].Fcross);
		const double dtoa = dtoa_between(net.row[0] + 9, net.row[d] + 9, t);
		// tc = -DTOA; tc *= 2*M_PI;   (src/waveform_util.cpp:173-174)
		out[d].tshift = (-dtoa) * (2 * GWAT_PI);
	}
}

// Antenna patterns in the detector's horizon frame   (right_interferometer; the geometric factor as fourier_detector_response_horizon applies it)
GWAT_HD void horizon_patterns(double theta, double phi, double psi, double geometric_factor, double &Fplus, double &Fcross)
{
	const double ct = cos(theta);
	const double fp = (1. / 2) * (1 + ct * ct) * cos(2. * phi);
	const double fc = ct * sin(2. * phi);
	const double c2psi = cos(2. * psi), s2psi = sin(2. * psi);
	Fplus = (fp * c2psi - fc * s2psi) * geometric_factor;
	Fcross = (fp * s2psi + fc * c2psi) * geometric_factor;
}

// The detector constants of one source: equatorial sky position, or -- on the single-detector response path, for a source with
// horizon_coord set -- the detector's own frame (fourier_detector_response_horizon, src/waveform_util.cpp:684-720: no time shift).
GWAT_HD void detector_setup_source(const Network &net, const gwat_b200_source &src, DetCoef *out)
{
	if (net.horizon_mode && src.horizon_coord) {
		for (int d = 0; d < net.D; d++) {
			horizon_patterns(src.theta, src.phi, src.psi, net.row[d][12], out[d].Fplus, out[d].Fcross);
			out[d].tshift = 0.0;
		}
		return;
	}
	detector_setup(net, src.RA, src.DEC, src.psi, src.gmst, out);
}

// ---- IMRPhenomD carrier setup ----------------------------------------------------------------------------------------
// The setup is written as four stages so that the cooperative setup kernel (k_setup, gwat_engine.cu) can give them to different
// warps of a CTA; phenomd_setup below runs them one after the other on one thread (Fisher stencil points, host harness) with the
// very same arithmetic, so both give the same bits.
//   remnant : ringdown / damping frequency (QNM spline)                         <- s
//   common  : region boundaries, mass scalings                                   <- s, fRD, fdamp, f3
//   amp     : rows 0-6 of the fit, PN amplitude coefficients, f3, the collocation of the intermediate amplitude
//   phase   : rows 7-18 of the fit, PN phase coefficients, modifications, C1 matching, reference phase and time
// Ownership of DCoef: `amp` fills the contiguous block A0 .. mr_w2, `common` + `phase` everything else.

// ringdown and damping frequency: spline(QNM table, a_final) / (1 - E_rad) / M      (calc_fring / calc_fdamp)
template <class Fam>
GWAT_HD void phenomd_setup_remnant(SrcQ &s, const double (*qnm)[5], int qnm_n)
{
	const double a_final = remnant_spin<Fam>(s);
	const double erad = erad_rational_0815(s.eta, s.spin1z, s.spin2z);
	s.fRD = (qnm_eval(qnm, qnm_n, a_final, 1) / (1.0 - erad)) / s.M;
	s.fdamp = (qnm_eval(qnm, qnm_n, a_final, 3) / (1.0 - erad)) / s.M;
}

// fpeak (:1312-1326) and the fixed region boundaries; needs gamma[1], gamma[2], fRD, fdamp
GWAT_HD void phenomd_setup_boundaries(SrcQ &s, double g2, double g3)
{
	s.f1_phase = 0.018 / s.M;
	s.f2_phase = s.fRD / 2.;
	s.f1 = 0.014 / s.M;
	double pk;
	if (g2 > 1) pk = s.fRD + (s.fdamp * (-1.) * g3) / g2;
	else pk = s.fRD + s.fdamp * g3 * (sqrt(1 - g2 * g2) - 1) / g2;
	s.f3 = sqrt(pk * pk);
}

GWAT_HD void phenomd_setup_common(const SrcQ &s, DCoef &c)
{
	const double M = s.M;
	c.fcut = .2 / M;
	c.f1a = s.f1;
	c.f3a = s.f3;
	c.f1p = s.f1_phase;
	c.f2p = s.f2_phase;
	c.M = M;
	{
		// M^(fl(1/6)): the per-bin sixth root is this times the grid's f^(fl(1/6)) table, both in double-double
		const dd r = pow_sixth_dd(M);
		c.sM_hi = r.hi;
		c.sM_lo = r.lo;
	}
	c.logM = sm::log(M);
	c.logpiM = sm::log(GWAT_PI * M);
	c.pichirp = GWAT_PI * s.chirpmass;
	c.fRD = s.fRD;
	c.fdamp = s.fdamp;
	c.inv_fdamp = 1. / s.fdamp;
	c.inv_eta = 1. / s.eta;
}

// Rows [first, first + count) of the fit table into v (one rolled loop: setup code is fetch-bound, gwat_hd.h)
GWAT_HD void phenomd_fit_rows(const double (*fit)[11], double eta, double chi_pn, int first, int count, double *v)
{
	GWAT_SETUP_LOOP
	for (int i = 0; i < count; i++) v[i] = phenomd_fit_element(fit, first + i, eta, chi_pn);
}

// The amplitude block of DCoef (A0 .. mr_w2).  Needs s.fRD, s.fdamp, s.f1, s.f3 and c.M (phenomd_setup_common); `lam` brings rho,
// v2, gamma.
GWAT_HD void phenomd_setup_amp(const SrcQ &s, const Lambda &lam, DCoef &c)
{
	const double M = s.M;
	const PiPowers pi = pi_powers();
	double camp[7];
	pn_amplitude_coeffs(s, camp);
	c.A0 = s.A0 * sm::pow(M, 7. / 6.);
	c.ains[0] = camp[0];
	c.ains[1] = camp[1] * pi.third;
	c.ains[2] = camp[2] * pi.two3;
	c.ains[3] = camp[3] * GWAT_PI;
	c.ains[4] = camp[4] * pi.four3;
	c.ains[5] = camp[5] * pi.five3;
	c.ains[6] = camp[6] * pi.sq;
	for (int i = 0; i < 3; i++) c.rho[i] = lam.rho[i];
	c.mr_num = lam.gamma[0] * lam.gamma[2] * s.fdamp / M;
	c.mr_rate = lam.gamma[1] / (lam.gamma[2] * s.fdamp);
	c.mr_w2 = (lam.gamma[2] * s.fdamp) * (lam.gamma[2] * s.fdamp);
	// Collocation of the intermediate amplitude: value+slope at f1, value at the midpoint, value+slope at f3
	// (amp_connection_coeffs, :1679-1702; the reference expands the solution into closed-form monomial
	// coefficients, here it stays in Newton divided-difference form in x = M f).
	MfPowers p1;
	mf_powers(M, s.f1, sixth_root_approx(M, s.f1), p1);  // amplitude only
	const double v1 = phenomd_amp_ins(c, p1);
	const double v2 = lam.v2;
	const double v3 = phenomd_amp_mr(c, s.f3);
	// d/df of the inspiral amplitude at f1
	double dA1 = 0;
	{
		const double u = sm::cbrt(GWAT_PI * M * s.f1);
		double uk = 1;
		for (int k = 0; k < 7; k++) {
			dA1 += camp[k] * uk * (k / 3.);
			uk *= u;
		}
		const double m13 = sm::cbrt(M * s.f1);
		const double m73 = m13 * m13 * m13 * m13 * m13 * m13 * m13;
		dA1 += lam.rho[0] * m73 * (7. / 3.) + lam.rho[1] * m73 * m13 * (8. / 3.) + lam.rho[2] * m73 * m13 * m13 * 3.;
		dA1 /= s.f1;
	}
	// d/df of the merger-ringdown amplitude at f3
	double dA3;
	{
		const double df = s.f3 - s.fRD;
		const double den = df * df + c.mr_w2;
		dA3 = -c.mr_num * sm::exp(-c.mr_rate * df) * (c.mr_rate * den + 2 * df) / (den * den);
	}
	const double x1 = M * s.f1, x3 = M * s.f3, x2 = M * ((s.f1 + s.f3) / 2.);
	const double s1 = dA1 / M, s3 = dA3 / M;  // slopes with respect to x
	// divided differences on [x1,x1,x2,x3,x3]
	const double f01 = s1;
	const double f12 = (v2 - v1) / (x2 - x1);
	const double f23 = (v3 - v2) / (x3 - x2);
	const double f34 = s3;
	const double f012 = (f12 - f01) / (x2 - x1);
	const double f123 = (f23 - f12) / (x3 - x1);
	const double f234 = (f34 - f23) / (x3 - x2);
	const double f0123 = (f123 - f012) / (x3 - x1);
	const double f1234 = (f234 - f123) / (x3 - x1);
	const double f01234 = (f1234 - f0123) / (x3 - x1);
	c.ix1 = x1;
	c.ix2 = x2;
	c.ix3 = x3;
	c.ic[0] = v1;
	c.ic[1] = f01;
	c.ic[2] = f012;
	c.ic[3] = f0123;
	c.ic[4] = f01234;
}

// What the phase stage computes before it needs the remnant: fit rows 7-18, PN phase coefficients, NRT moments, gIMR rescalings.
struct PhasePrep {
	double cph[12];
	double c8_gr, alpha1_fit;
};
template <class Fam>
GWAT_HD void phenomd_setup_phase_prep(SrcQ &s, const double (*fit)[11], Lambda &lam, PhasePrep &pp)
{
	double v[12];
	phenomd_fit_rows(fit, s.eta, s.chi_pn, 7, 12, v);
	lam.sigma[0] = 0;
	for (int i = 0; i < 4; i++) lam.sigma[i + 1] = v[i];
	lam.beta[0] = 0;
	for (int i = 0; i < 3; i++) lam.beta[i + 1] = v[i + 4];
	lam.alpha[0] = 0;
	for (int i = 0; i < 5; i++) lam.alpha[i + 1] = v[i + 7];
	pn_phase_coeffs(s, pp.cph);
	if (Fam::nrt) {
		nrt_moments(s);
		nrt_adjust_pn_phase(s, pp.cph);
	}
	pp.c8_gr = pp.cph[8];
	pp.alpha1_fit = lam.alpha[1];  // the reference re-evaluates fit element 14 for the time shift (:456)
	apply_gimr<Fam>(s, lam, pp.cph);
}

// The phase side of DCoef.  Needs the common block of c, s.fRD, s.fdamp, s.f3 and lam.sigma/beta/alpha.
template <class Fam>
GWAT_HD void phenomd_setup_phase(SrcQ &s, Lambda &lam, const PhasePrep &pp, DCoef &c)
{
	const double M = s.M, eta = s.eta;
	const PiPowers pi = pi_powers();
	const double *cph = pp.cph;
	c.k1 = cph[1] * pi.third;
	c.k2 = cph[2] * pi.two3;
	c.k3 = (cph[3] * GWAT_PI) * M;
	c.k4 = cph[4] * pi.four3;
	c.k7 = cph[7] * pi.seven3;
	c.c8 = cph[8];
	c.c9 = cph[9];
	c.c10 = cph[10];
	c.c11 = cph[11];
	c.pi53 = pi.five3;
	c.pi2 = pi.sq;
	c.tf2 = 3. / (128. * eta) * pi.m53;
	c.k128 = 3. / (128. * eta);
	c.sig1M = lam.sigma[1] * M;
	c.sig2q = (3. / 4.) * lam.sigma[2];
	c.sig3q = (3. / 5) * lam.sigma[3];
	c.sig4q = (1. / 2.) * lam.sigma[4];
	c.Nmod = 0;
	c.n_gimr_neg = 0;
	setup_family_extras<Fam>(s, c);
	if (Fam::nrt) nrt_setup(s, c);

	// C1 matching of the three phase regions (phase_connection_coefficients, :1614-1674): the connection coefficients are
	// found in sequence, each with the not-yet-known ones at zero.
	InsDerivIn din;
	for (int k = 0; k < 12; k++) din.c[k] = cph[k];
	din.c8_gr = pp.c8_gr;
	for (int k = 0; k < 5; k++) din.sigma[k] = lam.sigma[k];
	din.M = M;
	din.eta = eta;
	lam.beta[0] = lam.beta[1] = lam.alpha[0] = lam.alpha[1] = 0;
	const double f1p = s.f1_phase, f2p = s.f2_phase;
	const double log_f1p = sm::log(f1p), log_f2p = sm::log(f2p);
	auto sync_int_mr = [&]() {
		c.beta0 = lam.beta[0];
		c.beta1 = lam.beta[1];
		c.beta2 = lam.beta[2];
		c.beta3_3 = lam.beta[3] / 3.;
		c.alpha0 = lam.alpha[0];
		c.alpha1 = lam.alpha[1];
		c.alpha2 = lam.alpha[2];
		c.alpha3_43 = (4. / 3) * lam.alpha[3];
		c.alpha4 = lam.alpha[4];
		c.alpha5fRD = lam.alpha[5] * s.fRD;
	};
	sync_int_mr();
	{
		const double Dins = dphase_ins_df(din, f1p) + dphase_ins_extra<Fam>(s, f1p);
		const double Dint = dphase_int_df(lam, M, eta, f1p) + dphase_imr_extra<Fam>(s, f1p);
		lam.beta[1] = (eta / M) * Dins - (eta / M) * Dint;
	}
	sync_int_mr();
	{
		MfPowers p;
		mf_powers(M, f1p, sixth_root_direct(M, f1p), p);
		// at exactly f1p the per-bin code would take the intermediate branch; the matching needs the inspiral expression
		const double ins = phenomd_phase_ins<Fam>(c, f1p, p, log_f1p);
		const double intv = phenomd_phase_int<Fam>(c, f1p, log_f1p, p.sixth);
		lam.beta[0] = eta * ins - eta * intv;
	}
	sync_int_mr();
	{
		const double Dint = dphase_int_df(lam, M, eta, f2p) + dphase_imr_extra<Fam>(s, f2p);
		const double Dmr = dphase_mr_df(lam, M, eta, s.fRD, s.fdamp, f2p) + dphase_imr_extra<Fam>(s, f2p);
		lam.alpha[1] = (eta / M) * Dint - eta / M * Dmr;
	}
	sync_int_mr();
	{
		const double root2 = sixth_root_approx(M, f2p);
		const double intv = phenomd_phase_int<Fam>(c, f2p, log_f2p, root2);
		const double mr = phenomd_phase_mr<Fam>(c, f2p, root2);
		lam.alpha[0] = eta * intv - eta * mr;
	}
	sync_int_mr();

	// ---- reference phase and coalescence time (:439-466) ---------------------------------------------------------------
	if (Fam::base == BASE_D) {
		double f_ref, phic;
		if (s.shift_phase) {
			f_ref = s.f_ref;
			// (the phase alone: the amplitude block may belong to another warp, see the note at the top)
			const double root = sixth_root_direct(M, f_ref), lg = sm::log(f_ref);
			typedef Family<BASE_D, Fam::ppe, Fam::gimr, false> Plain;
			double phi_shift;
			if (f_ref < c.f1p) {
				MfPowers p;
				mf_powers(M, f_ref, root, p);
				phi_shift = phenomd_phase_ins<Plain>(c, f_ref, p, lg);
			} else if (f_ref > c.f2p) phi_shift = phenomd_phase_mr<Plain>(c, f_ref, root);
			else phi_shift = phenomd_phase_int<Plain>(c, f_ref, lg, root);
			phic = 2 * s.phiRef + phi_shift;
		} else {
			f_ref = 0;
			phic = s.phiRef;
		}
		double tc_shift = 0;
		if (s.shift_time) {
			tc_shift = dphase_mr_df(lam, M, eta, s.fRD, s.fdamp, s.f3) + dphase_imr_extra<Fam>(s, s.f3) +
			           (-lam.alpha[1] + pp.alpha1_fit) * M / eta;
		}
		c.tc_shift = tc_shift;
		c.tc = phenomd_time_coefficient(s.tc, tc_shift);
		c.f_ref = f_ref;
		c.phic = phic;
	}
}

// The four stages on one thread.
template <class Fam>
GWAT_HD void phenomd_setup(SrcQ &s, const double (*fit)[11], const double (*qnm)[5], int qnm_n, DCoef &c)
{
	Lambda lam;
	{
		double v[7];
		phenomd_fit_rows(fit, s.eta, s.chi_pn, 0, 7, v);
		for (int i = 0; i < 3; i++) lam.rho[i] = v[i];
		lam.v2 = v[3];
		for (int i = 0; i < 3; i++) lam.gamma[i] = v[i + 4];
	}
	PhasePrep pp;
	phenomd_setup_phase_prep<Fam>(s, fit, lam, pp);
	phenomd_setup_remnant<Fam>(s, qnm, qnm_n);
	phenomd_setup_boundaries(s, lam.gamma[1], lam.gamma[2]);
	phenomd_setup_common(s, c);
	phenomd_setup_amp(s, lam, c);
	phenomd_setup_phase<Fam>(s, lam, pp, c);
}

// ---- one walker, start to finish --------------------------------------------------------------------------------------
struct Tables {
	const double (*fit)[11];
	const double (*qnm)[5];
	int qnm_n;
	DzTable dz;
};

// prep_source_parameters: hand the modification arrays over (src/waveform_generator.cpp:1296-1314)
template <class Fam>
GWAT_HD void copy_modifications(const gwat_b200_source &in, SrcQ &s)
{
	if (Fam::ppe != PPE_NONE) {
		s.Nmod = in.Nmod < GWAT_B200_MAX_MOD ? in.Nmod : GWAT_B200_MAX_MOD;
		for (int i = 0; i < s.Nmod; i++) {
			s.betappe[i] = in.betappe[i];
			s.bppe[i] = in.bppe[i];
		}
	}
	if (Fam::gimr) {
		s.Nmod_phi = in.Nmod_phi;
		s.Nmod_sigma = in.Nmod_sigma;
		s.Nmod_beta = in.Nmod_beta;
		s.Nmod_alpha = in.Nmod_alpha;
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
			s.phii[i] = in.phii[i];
			s.sigmai[i] = in.sigmai[i];
			s.betai[i] = in.betai[i];
			s.alphai[i] = in.alphai[i];
			s.delta_phi[i] = in.delta_phi[i];
			s.delta_sigma[i] = in.delta_sigma[i];
			s.delta_beta[i] = in.delta_beta[i];
			s.delta_alpha[i] = in.delta_alpha[i];
		}
	}
}

}  // namespace gwat
#endif
