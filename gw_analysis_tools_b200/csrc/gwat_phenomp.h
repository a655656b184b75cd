// IMRPhenomPv2: precessing twist-up of the IMRPhenomD carrier (GWAT_HD code).
//
// Reference being replaced:
//   per walker  PhenomPv2_Param_Transform[_reduced]      src/IMRPhenomP.cpp:822-938, 1082-1211   (chi_p, S_L, S_P, theta_JN, alpha0, zeta)
//               XLALSpinWeightedSphericalHarmonic(-2,2,m) src/util.cpp:1953-2000
//               calculate_euler_coeffs                    src/IMRPhenomP.cpp:65-144
//               setup part of construct_waveform          :195-289
//               calculate_time_shift                      :550-606  (10-point natural spline of the phase around fRD)
//   per bin     WignerD, L2PN                             :665-693, 1214-1220
//               calculate_euler_angles                    :717-735
//               calculate_twistup                         :696-715
//               the two loops of construct_waveform       :290-362
//               rotation by 2 zeta                        src/waveform_generator.cpp:257-266
#ifndef GWAT_PHENOMP_H
#define GWAT_PHENOMP_H

#include "gwat_setup.h"

namespace gwat {

// Orbital angular momentum to 2PN, non-spinning, per unit M^2, from x = (pi M f)^(2/3).
GWAT_HD double l2pn(double eta, double x, double sqrt_x)
{
	const double x2 = x * x, eta2 = eta * eta;
	return (eta * (1.0 + (1.5 + eta / 6.0) * x + (3.375 - (19.0 * eta) / 8. - eta2 / 24.0) * x2)) / sqrt_x;
}

struct Vec3 {
	double x, y, z;
};
// Rotations about z and y by an angle given through its cosine and sine.  The reference's ROTATEZ/ROTATEY macros
// (include/gwat/IMRPhenomP.h:17-27) call cos/sin of the same three angles up to four times each; here each angle's
// cosine and sine are evaluated once and reused (identical inputs, identical values).
struct CosSin {
	double c, s;
};
GWAT_HD CosSin cos_sin(double angle)
{
	CosSin r;
	sm::sincos(angle, &r.s, &r.c);
	return r;
}
GWAT_HD void rot_z(const CosSin &a, Vec3 &v)
{
	const double t1 = v.x * a.c - v.y * a.s, t2 = v.x * a.s + v.y * a.c;
	v.x = t1;
	v.y = t2;
}
GWAT_HD void rot_y(const CosSin &a, Vec3 &v)
{
	const double t1 = v.x * a.c + v.z * a.s, t2 = -v.x * a.s + v.z * a.c;
	v.x = t1;
	v.z = t2;
}

// Source-frame spins -> the PhenomP parameters.  `reduced` selects the (chi_p, phi_p) input convention.
// First the spin projections (chi_l, chi_p, S_perp, S_L: all the carrier's remnant needs), then the frame angles.
GWAT_HD void phenompv2_spin_projection(SrcQ &s, bool reduced)
{
	const double chi1_l = s.spin1z, chi2_l = s.spin2z;
	const double q = s.mass1 / s.mass2;
	const double chi_eff = (s.mass1 * chi1_l + s.mass2 * chi2_l) / s.M;
	s.chil = (1.0 + q) / q * chi_eff;
	const double m1_2 = s.mass1 * s.mass1, m2_2 = s.mass2 * s.mass2;
	if (!reduced) {
		const double S1_perp = m1_2 * sqrt(s.spin1y * s.spin1y + s.spin1x * s.spin1x);
		const double S2_perp = m2_2 * sqrt(s.spin2y * s.spin2y + s.spin2x * s.spin2x);
		const double A1 = 2 + (3 * s.mass2) / (2 * s.mass1);
		const double A2 = 2 + (3 * s.mass1) / (2 * s.mass2);
		const double ASp1 = A1 * S1_perp, ASp2 = A2 * S2_perp;
		const double num = (ASp2 > ASp1) ? ASp2 : ASp1;
		const double denom = (s.mass2 > s.mass1) ? A2 * m2_2 : A1 * m1_2;
		s.chip = num / denom;
	}
	const double m1 = q / (1 + q), m2 = 1. / (1 + q);
	s.SP = s.chip * m1 * m1;
	s.SL = chi1_l * m1 * m1 + chi2_l * m2 * m2;
}
GWAT_HD void phenompv2_frame_angles(SrcQ &s, bool reduced)
{
	const double m1_2 = s.mass1 * s.mass1, m2_2 = s.mass2 * s.mass2;

	// L at f_ref (the reference fills the inspiral power table at f_ref: src/IMRPhenomP.cpp:1117-1123)
	const PiPowers pi = pi_powers();
	const double sixth = sixth_root_approx(s.M, s.f_ref);  // feeds L(f_ref), an angle-level quantity
	const double mf_third = mul_rn(sixth, sixth);
	const double mf_two3 = mul_rn(mf_third, mf_third);
	const double x = mf_two3 * pi.two3;
	const double L0 = s.M * s.M * l2pn(s.eta, x, sqrt(x));

	double J0x, J0y;
	if (reduced) {
		J0x = m1_2 * s.chip * sm::cos(s.phip);
		J0y = m1_2 * s.chip * sm::sin(s.phip);
	} else {
		J0x = m1_2 * s.spin1x + m2_2 * s.spin2x;
		J0y = m1_2 * s.spin1y + m2_2 * s.spin2y;
	}
	const double J0z = L0 + m1_2 * s.spin1z + m2_2 * s.spin2z;
	const double J0 = sqrt(J0x * J0x + J0y * J0y + J0z * J0z);
	const double thetaJ = sm::acos(J0z / J0);
	const double phiJ = sm::atan2(J0y, J0x);
	s.phi_aligned = -phiJ;

	const double incl = s.incl_angle, phiRef = s.phiRef;
	const CosSin ci = cos_sin(incl), cp = cos_sin(phiRef), cq = cos_sin(GWAT_PI / 2. - phiRef);
	const CosSin rJz = cos_sin(-phiJ), rJy = cos_sin(-thetaJ);
	const Vec3 N{ci.s * cq.c, ci.s * cq.s, ci.c};
	Vec3 t = N;
	rot_z(rJz, t);
	rot_y(rJy, t);
	const double kappa = -sm::atan2(t.y, t.x);
	const CosSin rK = cos_sin(kappa);

	t = Vec3{0., 0., 1.};
	rot_z(rJz, t);
	rot_y(rJy, t);
	rot_z(rK, t);
	s.alpha0 = sm::atan2(t.y, t.x);

	t = N;
	rot_z(rJz, t);
	rot_y(rJy, t);
	rot_z(rK, t);
	const double Nx_Jf = t.x, Nz_Jf = t.z;
	s.thetaJN = sm::acos(Nz_Jf);

	// polarisation-frame mismatch angle zeta between the (P,Q,N) triad of PhenomP and the LAL wave frame
	t = Vec3{-ci.c * cp.s, -ci.c * cp.c, ci.s};
	rot_z(rJz, t);
	rot_y(rJy, t);
	rot_z(rK, t);
	const double XdotP = t.x * 0. + t.y * -1. + t.z * 0.;
	const double XdotQ = t.x * Nz_Jf + t.y * 0. + t.z * -Nx_Jf;
	s.zeta_polariz = sm::atan2(XdotQ, XdotP);
}

GWAT_HD void phenompv2_param_transform(SrcQ &s, bool reduced)
{
	phenompv2_spin_projection(s, reduced);
	phenompv2_frame_angles(s, reduced);
}

// Coefficients of the PN expansions of the precession angles alpha(omega), epsilon(omega) (LAL's
// ComputeNNLOanglecoeffs, transcribed by the reference at src/IMRPhenomP.cpp:65-144), in powers of omega^(1/3):
//   angle = c1/omega + c2/omega^(2/3) + c3/omega^(1/3) + c4 log(omega) + c5 omega^(1/3)
GWAT_HD void euler_angle_coeffs(double q, double chil, double chip, double *a, double *e)
{
	const double m2 = q / (1. + q), m1 = 1. / (1. + q);
	const double dm = m1 - m2;
	const double eta = m1 * m2, eta2 = eta * eta, eta3 = eta2 * eta, eta4 = eta3 * eta;
	const double chil2 = chil * chil, chip2 = chip * chip, chip4 = chip2 * chip2;
	const double dm2 = dm * dm, dm3 = dm2 * dm;
	const double m2_2 = m2 * m2, m2_3 = m2_2 * m2, m2_4 = m2_3 * m2, m2_5 = m2_4 * m2, m2_6 = m2_5 * m2, m2_7 = m2_6 * m2,
	             m2_8 = m2_7 * m2;
	const double pi = GWAT_PI;
	// total mass is 1 in these units, so its powers are dropped
	a[0] = (-0.18229166666666666 - (5 * dm) / (64. * m2));
	a[1] = ((-15 * dm * m2 * chil) / (128. * eta) - (35 * m2_2 * chil) / (128. * eta));
	a[2] = (-1.7952473958333333 - (4555 * dm) / (7168. * m2) - (15 * chip2 * dm * m2_3) / (128. * eta2) -
	        (35 * chip2 * m2_4) / (128. * eta2) - (515 * eta) / 384. - (15 * dm2 * eta) / (256. * m2_2) -
	        (175 * dm * eta) / (256. * m2));
	a[3] = -(35 * pi) / 48. - (5 * dm * pi) / (16. * m2) + (5 * dm2 * chil) / (16.) + (5 * dm * m2 * chil) / (3.) +
	       (2545 * m2_2 * chil) / (1152.) - (5 * chip2 * dm * m2_5 * chil) / (128. * eta3) -
	       (35 * chip2 * m2_6 * chil) / (384. * eta3) + (2035 * dm * m2 * chil) / (21504. * eta) +
	       (2995 * m2_2 * chil) / (9216. * eta);
	a[4] = (4.318908476114694 + (27895885 * dm) / (2.1676032e7 * m2) - (15 * chip4 * dm * m2_7) / (512. * eta4) -
	        (35 * chip4 * m2_8) / (512. * eta4) - (485 * chip2 * dm * m2_3) / (14336. * eta2) +
	        (475 * chip2 * m2_4) / (6144. * eta2) + (15 * chip2 * dm2 * m2_2) / (256. * eta) +
	        (145 * chip2 * dm * m2_3) / (512. * eta) + (575 * chip2 * m2_4) / (1536. * eta) + (39695 * eta) / 86016. +
	        (1615 * dm2 * eta) / (28672. * m2_2) - (265 * dm * eta) / (14336. * m2) + (955 * eta2) / 576. +
	        (15 * dm3 * eta2) / (1024. * m2_3) + (35 * dm2 * eta2) / (256. * m2_2) + (2725 * dm * eta2) / (3072. * m2) -
	        (15 * dm * m2 * pi * chil) / (16. * eta) - (35 * m2_2 * pi * chil) / (16. * eta) +
	        (15 * chip2 * dm * m2_7 * chil2) / (128. * eta4) + (35 * chip2 * m2_8 * chil2) / (128. * eta4) +
	        (375 * dm2 * m2_2 * chil2) / (256. * eta) + (1815 * dm * m2_3 * chil2) / (256. * eta) +
	        (1645 * m2_4 * chil2) / (192. * eta));
	e[0] = (-0.18229166666666666 - (5 * dm) / (64. * m2));
	e[1] = ((-15 * dm * m2 * chil) / (128. * eta) - (35 * m2_2 * chil) / (128. * eta));
	e[2] = (-1.7952473958333333 - (4555 * dm) / (7168. * m2) - (515 * eta) / 384. - (15 * dm2 * eta) / (256. * m2_2) -
	        (175 * dm * eta) / (256. * m2));
	e[3] = -(35 * pi) / 48. - (5 * dm * pi) / (16. * m2) + (5 * dm2 * chil) / (16.) + (5 * dm * m2 * chil) / (3.) +
	       (2545 * m2_2 * chil) / (1152.) + (2035 * dm * m2 * chil) / (21504. * eta) + (2995 * m2_2 * chil) / (9216. * eta);
	e[4] = (4.318908476114694 + (27895885 * dm) / (2.1676032e7 * m2) + (39695 * eta) / 86016. +
	        (1615 * dm2 * eta) / (28672. * m2_2) - (265 * dm * eta) / (14336. * m2) + (955 * eta2) / 576. +
	        (15 * dm3 * eta2) / (1024. * m2_3) + (35 * dm2 * eta2) / (256. * m2_2) + (2725 * dm * eta2) / (3072. * m2) -
	        (15 * dm * m2 * pi * chil) / (16. * eta) - (35 * m2_2 * pi * chil) / (16. * eta) +
	        (375 * dm2 * m2_2 * chil2) / (256. * eta) + (1815 * dm * m2_3 * chil2) / (256. * eta) +
	        (1645 * m2_4 * chil2) / (192. * eta));
}

// alpha and epsilon at one frequency; omega_cbrt = (pi M f)^(1/3), log_omega = ln(pi M f)
GWAT_HD void euler_angles(const double *a, const double *e, double omega_cbrt, double log_omega, double &alpha,
                          double &epsilon)
{
	const double oc2 = omega_cbrt * omega_cbrt;
	const double omega = oc2 * omega_cbrt;
	alpha = (a[0] / omega + a[1] / oc2 + a[2] / omega_cbrt + a[3] * log_omega + a[4] * omega_cbrt);
	epsilon = (e[0] / omega + e[1] / oc2 + e[2] / omega_cbrt + e[3] * log_omega + e[4] * omega_cbrt);
}

// -2Y_{2m}(theta, 0), m = -2..2: real for zero azimuth
GWAT_HD void spin_weighted_y2(double theta, double *Y)
{
	const double ct = sm::cos(theta), st = sm::sin(theta);
	Y[0] = sqrt(5.0 / (64.0 * GWAT_PI)) * (1.0 - ct) * (1.0 - ct);
	Y[1] = sqrt(5.0 / (16.0 * GWAT_PI)) * st * (1.0 - ct);
	Y[2] = sqrt(15.0 / (32.0 * GWAT_PI)) * st * st;
	Y[3] = sqrt(5.0 / (16.0 * GWAT_PI)) * st * (1.0 + ct);
	Y[4] = sqrt(5.0 / (64.0 * GWAT_PI)) * (1.0 + ct) * (1.0 + ct);
}

// Derivative at x0 of the natural cubic spline through n <= 16 points: the reference calls gsl_spline_eval_deriv on a
// freshly built 10-point cspline (src/IMRPhenomP.cpp:595-599); same tridiagonal recurrence as GSL's, see tools/gen_tables.py.
GWAT_HD double natural_spline_deriv(const double *xa, const double *ya, int n, double x0)
{
	double c[16], g[16], diag[16], off[16], gam[16], alp[16], z[16];
	const int N = n - 2;
	for (int i = 0; i < n; i++) c[i] = 0;
	GWAT_SETUP_LOOP
	for (int i = 0; i < N; i++) {
		const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
		const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
		const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
		off[i] = h_ip1;
		diag[i] = mul_rn(2.0, add_rn(h_ip1, h_i));
		g[i] = mul_rn(3.0, sub_rn(mul_rn(yd_ip1, g_ip1), mul_rn(yd_i, g_i)));
	}
	alp[0] = diag[0];
	gam[0] = off[0] / alp[0];
	GWAT_SETUP_LOOP
	for (int i = 1; i < N - 1; i++) {
		alp[i] = sub_rn(diag[i], mul_rn(off[i - 1], gam[i - 1]));
		gam[i] = off[i] / alp[i];
	}
	alp[N - 1] = sub_rn(diag[N - 1], mul_rn(off[N - 2], gam[N - 2]));
	z[0] = g[0];
	GWAT_SETUP_LOOP
	for (int i = 1; i < N; i++) z[i] = sub_rn(g[i], mul_rn(gam[i - 1], z[i - 1]));
	GWAT_SETUP_LOOP
	for (int i = 0; i < N; i++) z[i] = z[i] / alp[i];
	c[N] = z[N - 1];
	GWAT_SETUP_LOOP
	for (int i = N - 2; i >= 0; i--) c[i + 1] = sub_rn(z[i], mul_rn(gam[i], c[i + 2]));
	// interval by bisection, like gsl_interp_bsearch
	int lo = 0, hi = n - 1;
	while (hi > lo + 1) {
		const int mid = (hi + lo) / 2;
		if (xa[mid] > x0) hi = mid; else lo = mid;
	}
	const double dx = xa[lo + 1] - xa[lo], dy = ya[lo + 1] - ya[lo];
	const double delx = x0 - xa[lo];
	const double b_i = sub_rn(dy / dx, mul_rn(dx, add_rn(c[lo + 1], mul_rn(2.0, c[lo]))) / 3.0);
	const double d_i = sub_rn(c[lo + 1], c[lo]) / mul_rn(3.0, dx);
	return add_rn(b_i, mul_rn(delx, add_rn(mul_rn(2.0, c[lo]), mul_rn(mul_rn(3.0, d_i), delx))));
}

// Per-walker setup of the twist-up.  Everything but the time shift is independent of the carrier block (phenomp_setup_angles);
// the time shift (calculate_time_shift) is the slope at fRD of the natural spline through kTimeShiftSamples samples of -phase on
// [0.8, 1.2] fRD of the FINISHED carrier block: phenomp_time_shift_sample gives sample j, phenomp_time_shift_finish the slope.
constexpr int kTimeShiftSamples = 10;
GWAT_HD void phenomp_setup_angles(const SrcQ &s, PCoef &p)
{
	double Y[5];
	spin_weighted_y2(s.thetaJN, Y);
	for (int i = 0; i < 5; i++) p.Y[i] = Y[i];
	p.tw[0] = Y[3] - Y[1];
	p.tw[1] = Y[3] + Y[1];
	p.tw[2] = (0.5 * 2.44948974278317788) * Y[2];
	p.tw[3] = Y[4] + Y[0];
	p.tw[4] = Y[4] - Y[0];
	p.tw[5] = 0.5 * p.tw[3];
	p.tw[6] = 0.5 * p.tw[4];
	p.SP2 = s.SP * s.SP;
	p.A0 = s.A0 * sm::pow(s.M, 7. / 6.) / (2. * sqrt(5. / (64. * GWAT_PI)));
	p.SP = s.SP;
	p.SL = s.SL;
	p.eta = s.eta;
	p.lc1 = 1.5 + s.eta / 6.0;
	p.lc2 = 3.375 - (19.0 * s.eta) / 8. - (s.eta * s.eta) / 24.0;
	const double q = s.mass1 / s.mass2;
	euler_angle_coeffs(q, s.chil, s.chip, p.acoef, p.ecoef);
	// offsets of the angles at f_ref; the reference forms (M f_ref)^(1/3) with pow(x, 1./3.)
	const PiPowers pi = pi_powers();
	const double mf_third_ref = sm::pow(s.M * s.f_ref, 1. / 3.);
	const double oc_ref = mf_third_ref * pi.third;
	double alpha_off, eps_off;
	euler_angles(p.acoef, p.ecoef, oc_ref, sm::log((oc_ref * oc_ref) * oc_ref), alpha_off, eps_off);
	p.alpha_const = s.alpha0 - alpha_off;
	p.epsilon_offset = eps_off;
	p.c2z = sm::cos(2. * s.zeta_polariz);
	p.s2z = sm::sin(2. * s.zeta_polariz);
	p.phic = 2 * s.phi_aligned;
	p.tc = phenomp_time_coefficient(s.tc);
	p.f_ref = s.f_ref;
}
// x_j and y_j = -phase(x_j); false when the window is degenerate (the reference bails out with 0 as well, :567-572)
template <class Fam>
GWAT_HD bool phenomp_time_shift_sample(const DCoef &c, int j, double &x, double &y)
{
	const double f_final = c.fRD;
	const double start = .8 * f_final, stop = 1.2 * f_final;
	const double step = (stop - start) / (kTimeShiftSamples - 1);
	if (!(step > 0)) return false;
	const double f = start + j * step;
	double a_unused, ph;
	// the samples straddle fRD > f2p: merger-ringdown phase, where the sixth root only enters through (Mf)^(3/4);
	// below f1p (never for physical parameters) the exact root is used
	const double root = f < c.f1p ? sixth_root_direct(c.M, f) : sixth_root_approx(c.M, f);
	// ln f feeds the inspiral and intermediate phases only: not evaluated for a merger-ringdown sample
	const double lg = f > c.f2p ? 0.0 : sm::log(f);
	phenomd_bin<Family<BASE_D, Fam::ppe, Fam::gimr, false>>(c, f, root, lg, a_unused, ph);
	x = f;
	y = -ph;
	return true;
}
GWAT_HD double phenomp_time_shift_finish(const DCoef &c, const double *xs, const double *ys)
{
	return 2 * GWAT_PI * (natural_spline_deriv(xs, ys, kTimeShiftSamples, c.fRD) / (2. * GWAT_PI));
}
template <class Fam>
GWAT_HD void phenomp_setup(const SrcQ &s, WalkerCoef &w)
{
	PCoef &p = w.p;
	const DCoef &c = w.d;
	phenomp_setup_angles(s, p);
	p.tcorr_2pi = 2 * GWAT_PI * 0.0;
	if (s.shift_time) {
		double xs[kTimeShiftSamples], ys[kTimeShiftSamples];
		bool ok = true;
		// (a rolled loop: the setup kernels are bound by instruction fetch, see gwat_hd.h; ten inlined copies of the carrier
		// evaluation were a quarter of the kernel's code)
		GWAT_SETUP_LOOP
		for (int j = 0; j < kTimeShiftSamples; j++) ok = phenomp_time_shift_sample<Fam>(c, j, xs[j], ys[j]) && ok;
		if (ok) p.tcorr_2pi = phenomp_time_shift_finish(c, xs, ys);
	}
}

// One bin of IMRPhenomPv2: both polarisations, rotated by 2 zeta (fourier_waveform semantics).
// What a bin's polarisations are made of before the coalescence time enters: carrier amplitude and phase, and (PhenomPv2)
// the twist factors.  The Fisher stencil evaluates this once per stencil point and finishes it per detector, because the
// reference re-times some stencil points per detector (src/fisher.cpp:436-453) and nothing else depends on the detector.
struct PolParts {
	double amp, phase;
	cplx hpf, hcf;
	bool zero;  // above the model's cutoff: the polarisations are exactly zero
};

template <class Fam>
GWAT_HD void phenomp_polarization_parts(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, PolParts &pp)
{
	const DCoef &c = w.d;
	const PCoef &p = w.p;
	pp.zero = f > c.fcut;
	if (pp.zero) return;
	const double sixth = bin_sixth_root(c, sf_hi, sf_lo);
	MfPowers mp;
	mf_powers(c.M, f, sixth, mp);
	double shape;
	if (f < c.f1a) shape = phenomd_amp_ins(c, mp);
	else if (f > c.f3a) shape = phenomd_amp_mr(c, f);
	else shape = phenomd_amp_int(c, mp.Mf);
	const double amp = (p.A0 * (shape / mp.seven6)) / 2.;
	double phase;
	if (f < c.f1p) phase = phenomd_phase_ins<Fam>(c, f, mp, logf);
	else if (f > c.f2p) phase = phenomd_phase_mr<Fam>(c, f, sixth);
	else phase = phenomd_phase_int<Fam>(c, f, logf, sixth);

	// Wigner d^2_{m,+-2}(beta): tan(beta) = S_perp / (L + S_parallel)
	const PiPowers pi = pi_powers();
	const double oc = mp.third * pi.third;   // omega^(1/3)
	const double x = mp.two3 * pi.two3;      // omega^(2/3)
	const double L = l2pn(p.eta, x, sqrt(x));
	const double sb = p.SP / (L + p.SL);
	const double cos_beta = 1. / sqrt(1.0 + sb * sb);
	const double ch = sqrt((1.0 + cos_beta) / 2.0), sh = sqrt((1.0 - cos_beta) / 2.0);
	const double c2 = ch * ch, s2 = sh * sh, c3 = c2 * ch, s3 = s2 * sh, c4 = c3 * ch, s4 = s3 * sh;
	double d2[5];
	d2[0] = s4;
	d2[1] = 2 * ch * s3;
	d2[2] = 2.44948974278317788 * s2 * c2;  // sqrt(6) as the reference spells it (include/gwat/IMRPhenomP.h:37)
	d2[3] = 2 * c3 * sh;
	d2[4] = c4;
	const double dm2[5] = {d2[4], -d2[3], d2[2], -d2[1], d2[0]};

	double alpha, epsilon;
	euler_angles(p.acoef, p.ecoef, oc, log((oc * oc) * oc), alpha, epsilon);
	alpha = alpha + p.alpha_const;
	epsilon = epsilon - p.epsilon_offset;

	// twist-up: sum over m of exp(-+ i m alpha) d^2 Y
	double sa, ca;
	fast_sincos(alpha, &sa, &ca);
	// e^{i k alpha}, k = -2..2.  exp(-i alpha) is formed as 1/exp(i alpha) in the reference; |e^{i alpha}| = 1 to rounding.
	const double inv = 1. / (ca * ca + sa * sa);
	const cplx e1{ca, sa}, em1{ca * inv, -sa * inv};
	const cplx e2{e1.re * e1.re - e1.im * e1.im, 2 * e1.re * e1.im};
	const cplx em2{em1.re * em1.re - em1.im * em1.im, 2 * em1.re * em1.im};
	const cplx ek[5] = {em2, em1, cplx{1., 0.}, e1, e2};
	cplx hpf{0., 0.}, hcf{0., 0.};
#pragma unroll
	for (int m = -2; m <= 2; m++) {
		const cplx ea = ek[-m + 2], eb = ek[m + 2];
		const double wa = dm2[m + 2] * p.Y[m + 2], wb = d2[m + 2] * p.Y[m + 2];
		const cplx T2m{ea.re * wa, ea.im * wa}, Tm2m{eb.re * wb, eb.im * wb};
		hpf.re += T2m.re + Tm2m.re;
		hpf.im += T2m.im + Tm2m.im;
		// i (T2m - Tm2m)
		hcf.re += -(T2m.im - Tm2m.im);
		hcf.im += (T2m.re - Tm2m.re);
	}
	pp.amp = amp;
	pp.phase = add_rn(phase, mul_rn(2., epsilon));
	pp.hpf = hpf;
	pp.hcf = hcf;
}

// `tc`: the coefficient of (f - f_ref), p.tc unless a caller re-times the point.
template <class Fam>
GWAT_HD void phenomp_polarizations_finish(const WalkerCoef &w, const PolParts &pp, double tc, double f, cplx &hp, cplx &hc)
{
	const PCoef &p = w.p;
	if (pp.zero) {
		hp = cplx{0.0, 0.0};
		hc = cplx{0.0, 0.0};
		return;
	}
	const double amp = pp.amp;
	const cplx hpf = pp.hpf, hcf = pp.hcf;
	// exp(-i (phase - tc (f - f_ref) - phic + 2 pi t_corr f))       (src/IMRPhenomP.cpp:354-362)
	double arg = sub_rn(pp.phase, mul_rn(tc, sub_rn(f, p.f_ref)));
	arg = sub_rn(arg, p.phic);
	arg = add_rn(arg, mul_rn(p.tcorr_2pi, f));
	double sn, cs;
	fast_sincos(arg, &sn, &cs);
	const cplx carrier{amp * cs, -(amp * sn)};
	const cplx hplus{carrier.re * hpf.re - carrier.im * hpf.im, carrier.re * hpf.im + carrier.im * hpf.re};
	const cplx hcross{carrier.re * hcf.re - carrier.im * hcf.im, carrier.re * hcf.im + carrier.im * hcf.re};
	hp = cplx{p.c2z * hplus.re + p.s2z * hcross.re, p.c2z * hplus.im + p.s2z * hcross.im};
	hc = cplx{p.c2z * hcross.re - p.s2z * hplus.re, p.c2z * hcross.im - p.s2z * hplus.im};
}

template <class Fam>
GWAT_HD void phenomp_polarizations_bin(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, cplx &hp,
                                       cplx &hc)
{
	PolParts pp;
	phenomp_polarization_parts<Fam>(w, f, sf_hi, sf_lo, logf, pp);
	phenomp_polarizations_finish<Fam>(w, pp, w.p.tc, f, hp, hc);
}

// ---- one walker, start to finish (all families) ---------------------------------------------------------------------
template <class Fam>
GWAT_HD void walker_setup(const gwat_b200_source &src, const Network &net, const Tables &t, int theory, WalkerCoef &w,
                          bool allow_sky_average = false)
{
	SrcQ s;
	populate_source(src, s);
	copy_modifications<Fam>(src, s);
	if (Fam::ppe != PPE_NONE) apply_theory(theory, t.dz, s);
	if (Fam::nrt) nrt_prepare_source(src, s);
	if (Fam::base == BASE_P) {
		// prep_source_parameters, src/waveform_generator.cpp:1271-1283: chip given -> reduced transform
		const bool reduced = (src.chip + 1) > 1e-10;
		phenompv2_param_transform(s, reduced);
	}
	phenomd_setup<Fam>(s, t.fit, t.qnm, t.qnm_n, w.d);
	if (Fam::base == BASE_P) {
		phenomp_setup<Fam>(s, w);
		w.cfac = 0;
		w.pfac = 0;
	} else {
		const double ci = sm::cos(s.incl_angle);
		w.cfac = ci;
		w.pfac = .5 * (1. + ci * ci);
	}
	detector_setup_source(net, src, w.det);
	for (int d = 0; d < net.D; d++) {
		DetCoef &dc = w.det[d];
		if (Fam::base == BASE_P) {
			dc.ga = dc.Fplus * w.p.c2z - dc.Fcross * w.p.s2z;
			dc.gb = dc.Fplus * w.p.s2z + dc.Fcross * w.p.c2z;
		} else {
			dc.ga = dc.Fplus * w.pfac;
			dc.gb = dc.Fcross * w.cfac;
		}
	}
	// an option of the reference that is outside this path is refused loudly (NaN), never silently approximated: the wall-clock-seeded
	// tidal_love_error draw.  sky_average changes the amplitude prefactor A0 and nothing else in the reference's waveform path
	// (populate_source_parameters, src/util.cpp:1024-1025) -- the records of the intrinsic samplers carry it (src/mcmc_gw.cpp:2494).
	// equatorial_orientation / horizon_coord are handled where the reference handles them (gwat_orient.h) and ignored elsewhere.
	(void)allow_sky_average;
	if (Fam::nrt && src.tidal_love_error) w.d.A0 = NAN;
	w.valid = 1;
}

}  // namespace gwat
#endif
