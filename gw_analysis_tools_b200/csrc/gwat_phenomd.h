// IMRPhenomD carrier: per-walker setup and per-bin amplitude/phase, as GWAT_HD code.
//
// What it has to reproduce (reference, relative to the GWAT root):
//   per walker  IMRPhenomD<double>::construct_waveform, setup part           src/IMRPhenomD.cpp:404-469
//               assign_lambda_param                                           :833-870 (+ include/gwat/IMRPhenomD.h:190-270)
//               post_merger_variables / calc_fring / calc_fdamp              :1091-1219
//               FinalSpin0815, EradRational0815, fpeak                       :1234-1326
//               assign_pn_amplitude_coeff, assign_static_pn_phase_coeff      :955-1032
//               amp_connection_coeffs, phase_connection_coefficients         :1614-1785
//   per bin     precalc_powers_ins, build_amp, build_phase                   :877-892, 747-821, 1336-1399, 1468-1503, 1560-1589
// The structure here is different from the reference's (no virtual dispatch, no per-call heap, the 1003-knot QNM spline is
// solved once at build time, the collocation polynomial is kept in Newton form, derivatives are written analytically
// with shared powers instead of ~150 libm pow() calls), the *values* are the same to rounding.
#ifndef GWAT_PHENOMD_H
#define GWAT_PHENOMD_H

#include "gwat_model.h"

namespace gwat {

// pi^(1/3) exactly as glibc evaluates pow(M_PI, 1./3.) in precalc_powers_PI (src/IMRPhenomD.cpp:943); the other powers
// of pi are formed from it by the same products the reference uses (:941-948).
#define GWAT_PI_THIRD 1.4645918875615231
struct PiPowers {
	double third, two3, four3, five3, seven3, sq, cubep, m53;
};
GWAT_HD PiPowers pi_powers()
{
	PiPowers p;
	p.sq = GWAT_PI * GWAT_PI;
	p.cubep = GWAT_PI * GWAT_PI * GWAT_PI;
	p.third = GWAT_PI_THIRD;
	p.two3 = p.third * p.third;
	p.four3 = p.third * GWAT_PI;
	p.five3 = p.two3 * GWAT_PI;
	p.seven3 = p.four3 * GWAT_PI;
	p.m53 = 1. / p.five3;
	return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// sampling-independent physics helpers
// ---------------------------------------------------------------------------------------------------------------------

// Final spin and radiated energy fits of arXiv:1508.07250 eqs. (3.6)-(3.8) (reference: src/IMRPhenomD.cpp:1244-1299).
GWAT_HD double final_spin_0815_s(double eta, double s)
{
	const double eta2 = eta * eta, eta3 = eta2 * eta;
	const double s2 = s * s, s3 = s2 * s;
	return eta * (3.4641016151377544 - 4.399247300629289 * eta + 9.397292189321194 * eta2 - 13.180949901606242 * eta3 +
	              s * ((1.0 / eta - 0.0850917821418767 - 5.837029316602263 * eta) +
	                   (0.1014665242971878 - 2.0967746996832157 * eta) * s +
	                   (-1.3546806617824356 + 4.108962025369336 * eta) * s2 +
	                   (-0.8676969352555539 + 2.064046835273906 * eta) * s3));
}
GWAT_HD double final_spin_0815(double eta, double chi1, double chi2)
{
	const double Seta = sqrt(1.0 - 4.0 * eta);
	const double m1 = 0.5 * (1.0 + Seta), m2 = 0.5 * (1.0 - Seta);
	const double s = (m1 * m1 * chi1 + m2 * m2 * chi2);
	return final_spin_0815_s(eta, s);
}
GWAT_HD double erad_rational_0815(double eta, double chi1, double chi2)
{
	const double Seta = sqrt(1.0 - 4.0 * eta);
	const double m1 = 0.5 * (1.0 + Seta), m2 = 0.5 * (1.0 - Seta);
	const double m1s = m1 * m1, m2s = m2 * m2;
	const double s = (m1s * chi1 + m2s * chi2) / (m1s + m2s);
	const double eta2 = eta * eta, eta3 = eta2 * eta;
	return (eta * (0.055974469826360077 + 0.5809510763115132 * eta - 0.9606726679372312 * eta2 + 3.352411249771192 * eta3) *
	        (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * eta2) * s)) /
	       (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * eta2) * s);
}

// Natural-cubic-spline lookup into the precomputed QNM table (tools/gen_tables.py): gsl_spline_eval semantics
// (interval by bisection, Horner form of cspline_eval).  `col` 1 = M*f_ring, 3 = M*f_damp.
GWAT_HD double qnm_eval(const double (*knots)[5], int n, double a, int col)
{
	if (!(a >= knots[0][0] && a <= knots[n - 1][0])) return NAN;
	// gsl_interp_bsearch semantics (knots[lo] <= a < knots[lo+1], last interval closed), reached from a guess instead of a
	// bisection: the table is uniform (0.002) except for a few 0.001 steps at either end, so the walk is 0-3 steps long
	int lo = (int)((a - knots[0][0]) * ((n - 1) / (knots[n - 1][0] - knots[0][0])));
	lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
	while (lo > 0 && knots[lo][0] > a) lo--;
	while (lo < n - 2 && !(knots[lo + 1][0] > a)) lo++;
	const double x_lo = knots[lo][0], x_hi = knots[lo + 1][0];
	const double dx = x_hi - x_lo;
	const double y_lo = knots[lo][col], y_hi = knots[lo + 1][col];
	const double dy = y_hi - y_lo;
	const double delx = a - x_lo;
	const double c_i = knots[lo][col + 1], c_ip1 = knots[lo + 1][col + 1];
	const double b_i = sub_rn(dy / dx, mul_rn(dx, add_rn(c_ip1, mul_rn(2.0, c_i))) / 3.0);
	const double d_i = sub_rn(c_ip1, c_i) / mul_rn(3.0, dx);
	return add_rn(y_lo, mul_rn(delx, add_rn(b_i, mul_rn(delx, add_rn(c_i, mul_rn(delx, d_i))))));
}

// The 19 phenomenological coefficients: bi-polynomial fits in eta and (chi_PN - 1).
GWAT_HD double phenomd_fit_element(const double (*fit)[11], int i, double eta, double chi_pn)
{
	const double *p = fit[i];
	const double xi = chi_pn - 1;
	const double xi2 = xi * xi, xi3 = xi2 * xi;  // the reference uses pow(xi,2), pow(xi,3): correctly rounded -> same to 1 ulp
	return p[0] + p[1] * eta + xi * (p[2] + p[3] * eta + p[4] * eta * eta) + xi2 * (p[5] + p[6] * eta + p[7] * eta * eta) +
	       xi3 * (p[8] + p[9] * eta + p[10] * eta * eta);
}
// TaylorF2 3PN amplitude coefficients (reference: assign_pn_amplitude_coeff, src/IMRPhenomD.cpp:955-986).
GWAT_HD void pn_amplitude_coeffs(const SrcQ &s, double *a)
{
	const double dm = s.delta_mass, eta = s.eta, xa = s.chi_a, xs = s.chi_s;
	const double eta2 = eta * eta, eta3 = eta2 * eta, xs2 = xs * xs, xa2 = xa * xa;
	const double pi = GWAT_PI, pi2 = GWAT_PI * GWAT_PI;
	a[0] = 1.;
	a[1] = 0.;
	a[2] = (-323. / 224 + 451 * eta / 168);
	a[3] = (27. * dm * xa / 8 + (27. / 8 - 11. * eta / 6) * xs);
	a[4] = (-27312085. / 8128512 - 1975055 * eta / 338688 + 105271 * eta2 / 24192 + (-81. / 32 + 8 * eta) * xa2 -
	        81 * dm * xa * xs / 16 + (-81. / 32 + 17 * eta / 8) * xs2);
	a[5] = 1. * (-85 * pi / 64 + 85 * pi * eta / 16 + dm * (285197. / 16128 - 1579 * eta / 4032) * xa +
	             (285197. / 16128 - 15317 * eta / 672 - 2227 * eta2 / 1008) * xs);
	a[6] = 1. * (-177520268561. / 8583708672 + (545384828789. / 5007163392 - 205 * pi2 / 48) * eta -
	             3248849057 * eta2 / 178827264 + 34473079 * eta3 / 6386688 +
	             (1614569. / 64512 - 1873643. * eta / 16128 + 2167 * eta2 / 42) * xa2 +
	             (31 * pi / 12 - 7 * pi * eta / 3) * xs + (1614569. / 64512 - 61391 * eta / 1344 + 57451 * eta2 / 4032) * xs2 +
	             dm * xa * (31 * pi / 12 + (1614569. / 32256 - 165961 * eta / 2688) * xs));
}

// TaylorF2 3.5PN phase coefficients.  c[5], c[6] are frequency dependent: c5 = c[8] + c[9] ln(pi M f),
// c6 = c[10] - c[11] ln(pi M f)   (reference: assign_static_pn_phase_coeff / assign_nonstatic_pn_phase_coeff,
// src/IMRPhenomD.cpp:992-1060).  The reference hard-codes ln 64 as 4.15888308336 (:25); so do we.
GWAT_HD void pn_phase_coeffs(const SrcQ &s, double *c)
{
	const double dm = s.delta_mass, eta = s.eta, xa = s.chi_a, xs = s.chi_s;
	const double eta2 = eta * eta, eta3 = eta2 * eta, xs2 = xs * xs, xa2 = xa * xa;
	const double pi = GWAT_PI, pi2 = GWAT_PI * GWAT_PI;
	const double ln64_as_in_reference = 4.15888308336;
	c[0] = 1.;
	c[1] = 0.;
	c[2] = 3715. / 756 + 55. * eta / 9;
	c[3] = -16. * pi + 113. * dm * xa / 3 + (113. / 3 - 76. * eta / 3) * xs;
	c[4] = 15293365. / 508032 + 27145. * eta / 504 + 3085. * eta2 / 72 + (-405. / 8 + 200. * eta) * xa2 -
	       (405. / 4) * dm * xa * xs + (-405. / 8 + 5. * eta / 2) * xs2;
	c[5] = 0;
	c[6] = 0;
	c[7] = 77096675. * pi / 254016 + 378515. * pi * eta / 1512 - 74045. * pi * eta2 / 756 +
	       dm * (-25150083775. / 3048192 + 26804935. * eta / 6048 - 1985. * eta2 / 48) * xa +
	       (-25150083775. / 3048192 + 10566655595. * eta / 762048 - 1042165. * eta2 / 3024 + 5345. * eta3 / 36) * xs;
	c[8] = (38645. * pi / 756 - 65. * pi * eta / 9 + dm * (-732985. / 2268 - 140. * eta / 9) * xa +
	        (-732985. / 2268 + 24260. * eta / 81 + 340. * eta2 / 9) * xs);
	c[9] = c[8];
	c[10] = 11583231236531. / 4694215680 - 6848. * GWAT_GAMMA_E / 21 - 640. * pi2 / 3 +
	        (-15737765635. / 3048192 + 2255. * pi2 / 12) * eta + 76055. * eta2 / 1728 - 127825. * eta3 / 1296 +
	        2270. * dm * xa * pi / 3 + (2270. * pi / 3 - 520. * pi * eta) * xs - 6848. * (ln64_as_in_reference) / 63;
	c[11] = 6848. / 63.;
}

// ---------------------------------------------------------------------------------------------------------------------
// per-bin evaluation
// ---------------------------------------------------------------------------------------------------------------------

// Powers of M f the reference precomputes per bin (precalc_powers_ins, src/IMRPhenomD.cpp:877-892), built from the sixth
// root by the same products in the same order.
struct MfPowers {
	double Mf, sixth, seven6, third, two3, four3, five3, seven3, eight3, sq, cubep, m53;
};
GWAT_HD void mf_powers(double M, double f, double sixth, MfPowers &p)
{
	p.Mf = mul_rn(M, f);
	p.sq = mul_rn(p.Mf, p.Mf);
	p.cubep = mul_rn(p.sq, p.Mf);
	p.sixth = sixth;
	p.seven6 = mul_rn(mul_rn(sixth, M), f);
	p.third = mul_rn(sixth, sixth);
	p.two3 = mul_rn(p.third, p.third);
	p.four3 = mul_rn(p.third, p.Mf);
	p.five3 = mul_rn(p.two3, p.Mf);
	p.seven3 = mul_rn(p.four3, p.Mf);
	p.eight3 = mul_rn(p.five3, p.Mf);
	p.m53 = 1. / p.five3;
}

// Per-bin sixth root: walker factor times grid factor, both double-double, rounded once (see gwat_grid.h).
GWAT_HD double bin_sixth_root(const DCoef &c, double sf_hi, double sf_lo)
{
	return dd_mul_to_double(c.sM_hi, c.sM_lo, sf_hi, sf_lo);
}

// (M f)^(fl(1/6)) the way the reference gets it, pow(M*f, 1./6.), for setup-time evaluations at single frequencies.
GWAT_HD double sixth_root_direct(double M, double f)
{
	const dd r = pow_sixth_dd(mul_rn(M, f));
	return add_rn(r.hi, r.lo);
}

// ppE phase terms sum_i beta_i (pi Mc f)^(b_i/3) (reference: src/ppE_IMRPhenomD.cpp:25-38,54-99).  The reference forms
// u = pow(pi*Mc*f, 1./3.) and then pow(u, b_i) per term; so do we (the phases involved are small, libm-vs-CUDA pow
// differences are far below the tolerance).
GWAT_HD double ppe_phase_terms(const DCoef &c, double mf_third, double acc)
{
	// u = (pi Mc f)^(1/3) from the already available (M f)^(1/3); integer exponents (every theory mapping and almost
	// every ppE use has b in -13..6) by repeated multiplication, anything else through pow().  The ppE phases are O(1..100)
	// rad, so the few-ulp difference to the reference's pow(pow(pi*Mc*f, 1./3.), b) is < 1e-13 rad.
	const double u = c.ppe_scale * mf_third;
	double ru = 0.0;
	for (int i = 0; i < c.Nmod; i++) {
		double t;
		const int b = c.bint[i];
		if (b == 9999) {
			t = pow(u, c.bppe[i]);
		} else {
			const int n = b < 0 ? -b : b;
			double base = u;
			if (b < 0) {
				if (ru == 0.0) ru = fast_rcp(u);
				base = ru;
			}
			t = 1.0;
			for (int k = n; k > 0; k >>= 1) {
				if (k & 1) t *= base;
				base *= base;
			}
		}
		acc = add_rn(acc, mul_rn(t, c.betappe[i]));
	}
	return acc;
}
// gIMR negative-PN-order inspiral terms 3/(128 eta) dphi_i (pi M f)^((i-5)/3), i = -4..-1 (src/gIMRPhenomD.cpp:149-166)
GWAT_HD double gimr_negative_pn_terms(const DCoef &c, double f, double acc)
{
	if (c.n_gimr_neg == 0) return acc;
	const double u = pow(GWAT_PI * c.M * f, 1. / 3.);
	for (int j = 0; j < c.n_gimr_neg; j++) {
		double prod = 1;
		for (int k = 0; k < -c.gimr_neg_pow[j]; k++) prod = mul_rn(prod, u);
		acc = add_rn(acc, mul_rn(c.gimr_neg_coef[j], 1. / prod));
	}
	return acc;
}
// Extra inspiral-phase terms of the modified families, accumulated onto `ph` in the reference's order.
// (M f)^(1/6) to a few ulp, for setup-time evaluations where the sixth root only feeds amplitudes or the (M f)^(3/4) term of
// the merger-ringdown phase (never the leading TaylorF2 term).
GWAT_HD double sixth_root_approx(double M, double f) { return sqrt(sm::cbrt(M * f)); }

template <class Fam>
GWAT_HD double phase_ins_extra(const DCoef &c, double f, double mf_third, double ph)
{
	if (Fam::ppe != PPE_NONE) ph = ppe_phase_terms(c, mf_third, ph);
	if (Fam::gimr) ph = gimr_negative_pn_terms(c, f, ph);
	return ph;
}

// Inspiral phase, TaylorF2 + sigma terms (reference: phase_ins, src/IMRPhenomD.cpp:1365-1399).  `logf` = ln f.
template <class Fam>
GWAT_HD double phenomd_phase_ins(const DCoef &c, double f, const MfPowers &p, double logf)
{
	const double logF = add_rn(c.logpiM, logf);
	const double c5 = add_rn(c.c8, mul_rn(logF, c.c9));
	const double c6 = sub_rn(c.c10, mul_rn(c.c11, logF));
	double pn = 1.0;
	if (Fam::gimr) pn = add_rn(pn, mul_rn(c.k1, p.third));
	pn = add_rn(pn, mul_rn(c.k2, p.two3));
	pn = add_rn(pn, mul_rn(c.k3, f));
	pn = add_rn(pn, mul_rn(c.k4, p.four3));
	pn = add_rn(pn, mul_rn(mul_rn(c5, c.pi53), p.five3));
	pn = add_rn(pn, mul_rn(mul_rn(c6, c.pi2), p.sq));
	pn = add_rn(pn, mul_rn(c.k7, p.seven3));
	const double tf2 = add_rn(-GWAT_PI / 4., mul_rn(mul_rn(c.tf2, p.m53), pn));
	double sg = mul_rn(c.sig1M, f);
	sg = add_rn(sg, mul_rn(c.sig2q, p.four3));
	sg = add_rn(sg, mul_rn(c.sig3q, p.five3));
	sg = add_rn(sg, mul_rn(c.sig4q, p.sq));
	double ph = add_rn(tf2, mul_rn(c.inv_eta, sg));
	if (Fam::ppe != PPE_NONE || Fam::gimr) ph = phase_ins_extra<Fam>(c, f, p.third, ph);
	return ph;
}

// Intermediate phase (reference: phase_int, src/IMRPhenomD.cpp:1577-1589).  log(Mf) is formed as ln M + ln f.
template <class Fam>
GWAT_HD double phenomd_phase_int(const DCoef &c, double f, double logf, double sixth)
{
	const double Mf = mul_rn(c.M, f);
	const double Mf3 = mul_rn(mul_rn(Mf, Mf), Mf);
	const double lg = add_rn(c.logM, logf);
	double t = add_rn(c.beta0, mul_rn(c.beta1, Mf));
	t = add_rn(t, mul_rn(c.beta2, lg));
	t = sub_rn(t, mul_rn(c.beta3_3, fast_rcp(Mf3)));
	double ph = mul_rn(c.inv_eta, t);
	if (Fam::ppe == PPE_IMR) ph = ppe_phase_terms(c, sixth * sixth, ph);
	return ph;
}

// Merger-ringdown phase (reference: phase_mr, src/IMRPhenomD.cpp:1485-1503).
template <class Fam>
GWAT_HD double phenomd_phase_mr(const DCoef &c, double f, double sixth)
{
	const double Mf = mul_rn(c.M, f);
	// (M f)^(3/4) = s^4 sqrt(s) with s = (M f)^(1/6): one square root instead of the reference's sqrt(sqrt((Mf)^3))
	const double s2 = sixth * sixth;
	const double Mf34 = (s2 * s2) * fast_sqrt(sixth);
	double t = add_rn(c.alpha0, mul_rn(c.alpha1, Mf));
	t = sub_rn(t, mul_rn(c.alpha2, fast_rcp(Mf)));
	t = add_rn(t, mul_rn(c.alpha3_43, Mf34));
	t = add_rn(t, mul_rn(c.alpha4, fast_atan(mul_rn(sub_rn(f, c.alpha5fRD), c.inv_fdamp))));
	double ph = mul_rn(c.inv_eta, t);
	if (Fam::ppe == PPE_IMR) ph = ppe_phase_terms(c, s2, ph);
	return ph;
}

GWAT_HD double phenomd_amp_ins(const DCoef &c, const MfPowers &p)
{
	return c.ains[0] + c.ains[1] * p.third + c.ains[2] * p.two3 + c.ains[3] * p.Mf + c.ains[4] * p.four3 +
	       c.ains[5] * p.five3 + c.ains[6] * p.sq + (c.rho[0] * p.seven3 + c.rho[1] * p.eight3 + c.rho[2] * p.cubep);
}
GWAT_HD double phenomd_amp_int(const DCoef &c, double Mf)
{
	// Newton form on the nodes [x1, x1, x2, x3, x3]
	const double d1 = Mf - c.ix1, d2 = Mf - c.ix2, d3 = Mf - c.ix3;
	return c.ic[0] + d1 * (c.ic[1] + d1 * (c.ic[2] + d2 * (c.ic[3] + d3 * c.ic[4])));
}
GWAT_HD double phenomd_amp_mr(const DCoef &c, double f)
{
	const double df = f - c.fRD;
	return c.mr_num * exp(-c.mr_rate * df) * fast_rcp(df * df + c.mr_w2);
}

// Amplitude (scaled by A0 M^{7/6}) and phase of the carrier at one bin with f <= fcut.
//   sixth = (M f)^(fl(1/6)),  logf = ln f
// Reference: the loop body of construct_waveform, src/IMRPhenomD.cpp:484-494.
template <class Fam>
GWAT_HD void phenomd_bin(const DCoef &c, double f, double sixth, double logf, double &amp, double &phase, MfPowers &p)
{
	if (f < c.f1p || f < c.f1a || Fam::base == BASE_P || Fam::nrt) {
		mf_powers(c.M, f, sixth, p);
	} else {
		p.Mf = mul_rn(c.M, f);
		p.sixth = sixth;
		p.seven6 = mul_rn(mul_rn(sixth, c.M), f);
	}
	double shape;
	if (f < c.f1a) shape = phenomd_amp_ins(c, p);
	else if (f > c.f3a) shape = phenomd_amp_mr(c, f);
	else shape = phenomd_amp_int(c, p.Mf);
	amp = c.A0 * (shape / p.seven6);

	if (f < c.f1p) phase = phenomd_phase_ins<Fam>(c, f, p, logf);
	else if (f > c.f2p) phase = phenomd_phase_mr<Fam>(c, f, sixth);
	else phase = phenomd_phase_int<Fam>(c, f, logf, sixth);
}

template <class Fam>
GWAT_HD void phenomd_bin(const DCoef &c, double f, double sixth, double logf, double &amp, double &phase)
{
	MfPowers p;
	phenomd_bin<Fam>(c, f, sixth, logf, amp, phase, p);
}

// The coefficient of (f - f_ref) in the carrier phase: 2 pi t_c plus the shift that puts the peak at t_c
// (src/IMRPhenomD.cpp:452-466); for IMRPhenomPv2 the shift is applied separately (src/IMRPhenomP.cpp:354-362).
// One definition so that every caller rounds it the same way.
GWAT_HD double phenomd_time_coefficient(double tc_seconds, double tc_shift) { return 2 * GWAT_PI * tc_seconds + tc_shift; }
GWAT_HD double phenomp_time_coefficient(double tc_seconds) { return 2 * GWAT_PI * tc_seconds; }

// phase -= tc (f - f_ref) + phic     (src/IMRPhenomD.cpp:497), unfused like the reference
GWAT_HD double phenomd_apply_time_phase(const DCoef &c, double tc, double f, double phase)
{
	return sub_rn(phase, add_rn(mul_rn(tc, sub_rn(f, c.f_ref)), c.phic));
}
GWAT_HD double phenomd_apply_time_phase(const DCoef &c, double f, double phase) { return phenomd_apply_time_phase(c, c.tc, f, phase); }

// ---------------------------------------------------------------------------------------------------------------------
// per-walker setup
// ---------------------------------------------------------------------------------------------------------------------

// Frequency derivatives used only for the C1 matching of the regions.
struct InsDerivIn {
	double c[12];     // phase coefficients (static ones; 8..11 the log pieces)
	double c8_gr;     // the GR value of c[8]: the reference's derivative of the log terms ignores gIMR rescalings
	double sigma[5];
	double M, eta;
};
GWAT_HD double dphase_ins_df(const InsDerivIn &in, double f)
{
	const double x = GWAT_PI * in.M * f;
	const double u = sm::cbrt(x);
	const double logx = sm::log(x);
	double c[8];
	for (int k = 0; k < 8; k++) c[k] = in.c[k];
	c[5] = in.c[8] + logx * in.c[9];
	c[6] = in.c[10] - in.c[11] * logx;
	double uk = 1, P = 0, dP = 0;
	for (int k = 0; k < 8; k++) {
		P += c[k] * uk;
		dP += c[k] * uk * (k / 3.);
		uk *= u;
	}
	const double u5 = u * u * u * u * u;
	// explicit f-dependence of the two log coefficients (GR values, as the reference's
	// assign_nonstatic_pn_phase_coeff_deriv recomputes them: src/IMRPhenomD.cpp:1068-1084)
	dP += in.c8_gr * u5 + (-6848. / 63.) * u5 * u;
	const double K = 3. / (128. * in.eta);
	const double xm53 = 1. / u5;
	const double tf2 = K * xm53 * (dP - (5. / 3.) * P) / f;
	const double Mf13 = sm::cbrt(in.M * f);
	const double sig = in.sigma[1] * in.M + in.M * Mf13 * (in.sigma[2] + Mf13 * (in.sigma[3] + Mf13 * in.sigma[4]));
	return tf2 + sig / in.eta;
}
GWAT_HD double dphase_int_df(const Lambda &l, double M, double eta, double f)
{
	const double Mf = M * f;
	return (l.beta[1] * M + l.beta[2] / f + l.beta[3] / (Mf * Mf * Mf * f)) / eta;
}
GWAT_HD double dphase_mr_df(const Lambda &l, double M, double eta, double fRD, double fdamp, double f)
{
	const double y = (f - l.alpha[5] * fRD) / fdamp;
	return (l.alpha[4] / (fdamp * (1. + y * y)) + l.alpha[2] / (f * f * M) + M * (l.alpha[1] + l.alpha[3] / sqrt(sqrt(f * M)))) /
	       eta;
}

GWAT_HD int find_id(int key, const int *list, int n)
{
	for (int i = 0; i < n; i++)
		if (list[i] == key) return i;
	return -1;
}
// Extra terms in d(phase)/df of the modified families, used only by the connection coefficients and the time shift.
// ppE: sum_i (b_i/3) f^(b_i/3-1) (pi Mc)^(b_i/3) beta_i    (src/ppE_IMRPhenomD.cpp:40-50,66-87)
GWAT_HD double ppe_dphase_terms(const SrcQ &s, double f)
{
	double acc = 0;
	for (int i = 0; i < s.Nmod; i++) {
		const double b3 = s.bppe[i] / 3.;
		acc += b3 * sm::pow(f, b3 - 1.) * sm::pow(s.chirpmass * GWAT_PI, b3) * s.betappe[i];
	}
	return acc;
}
template <class Fam>
GWAT_HD double dphase_ins_extra(const SrcQ &s, double f)
{
	double acc = 0;
	if (Fam::ppe != PPE_NONE) acc += ppe_dphase_terms(s, f);
	if (Fam::gimr && s.Nmod_phi != 0) {
		// src/gIMRPhenomD.cpp:168-185
		const double pimcube = sm::pow(GWAT_PI * s.M, 1. / 3.);
		for (int i = -4; i < 0; i++) {
			const int id = find_id(i, s.phii, s.Nmod_phi);
			if (id == -1) continue;
			double prod = 1;
			for (int k = 0; k < 5 - i; k++) prod *= pimcube;
			acc += 3. / (128. * s.eta) * s.delta_phi[id] * (1. / prod) * (i - 5.) / 3. * sm::pow(f, ((i - 5.) / 3. - 1));
		}
	}
	return acc;
}
template <class Fam>
GWAT_HD double dphase_imr_extra(const SrcQ &s, double f)
{
	return Fam::ppe == PPE_IMR ? ppe_dphase_terms(s, f) : 0.0;
}
// Copy the per-bin pieces of the modifications into the coefficient block.
template <class Fam>
GWAT_HD void setup_family_extras(const SrcQ &s, DCoef &c)
{
	c.Nmod = 0;
	c.n_gimr_neg = 0;
	if (Fam::ppe != PPE_NONE) {
		c.Nmod = s.Nmod;
		c.ppe_scale = sm::cbrt(GWAT_PI * s.chirpmass / s.M);
		for (int i = 0; i < s.Nmod && i < GWAT_B200_MAX_MOD; i++) {
			c.betappe[i] = s.betappe[i];
			c.bppe[i] = s.bppe[i];
			const double r = rint(s.bppe[i]);
			c.bint[i] = (r == s.bppe[i] && fabs(r) <= 32) ? (int)r : 9999;
		}
	}
	if (Fam::gimr && s.Nmod_phi != 0) {
		for (int i = -4; i < 0; i++) {
			const int id = find_id(i, s.phii, s.Nmod_phi);
			if (id == -1) continue;
			c.gimr_neg_coef[c.n_gimr_neg] = 3. / (128. * s.eta) * s.delta_phi[id];
			c.gimr_neg_pow[c.n_gimr_neg] = i - 5;
			c.n_gimr_neg++;
		}
	}
}

// Spin of the remnant: plain IMRPhenomD uses the aligned-spin fit; IMRPhenomPv2 overrides it to put the in-plane spin
// on the larger body (src/IMRPhenomP.cpp:1232-1260).
template <class Fam>
GWAT_HD double remnant_spin(const SrcQ &s)
{
	if (Fam::base == BASE_P) {
		const double m1 = s.mass1, m2 = s.mass2;
		const double M = m1 + m2;
		const double eta = m1 * m2 / (M * M);
		double af_parallel, q_factor;
		if (m1 >= m2) {
			q_factor = m1 / M;
			af_parallel = final_spin_0815(eta, s.spin1z, s.spin2z);
		} else {
			q_factor = m2 / M;
			af_parallel = final_spin_0815(eta, s.spin2z, s.spin1z);
		}
		const double Sperp = s.chip * q_factor * q_factor;
		// sign(af_parallel) the way the reference's copysign_internal forms it (src/util.cpp:566-569): not exactly +-1
		const double sgn = sqrt(1.0 * 1.0 / (af_parallel * af_parallel)) * af_parallel;
		return sgn * sqrt(Sperp * Sperp + af_parallel * af_parallel);
	}
	return final_spin_0815(s.eta, s.spin1z, s.spin2z);
}

// gIMR rescalings of the fit coefficients and of the static PN phase coefficients (src/gIMRPhenomD.cpp:60-147).
template <class Fam>
GWAT_HD void apply_gimr(const SrcQ &s, Lambda &l, double *c)
{
	if (!Fam::gimr) return;
	for (int i = 2; i < 4; i++) {
		const int id = find_id(i, s.betai, s.Nmod_beta);
		if (id != -1) l.beta[i] *= (1. + s.delta_beta[id]);
	}
	for (int i = 2; i < 6; i++) {
		const int id = find_id(i, s.alphai, s.Nmod_alpha);
		if (id != -1) l.alpha[i] *= (1. + s.delta_alpha[id]);
	}
	for (int i = 2; i < 5; i++) {
		const int id = find_id(i, s.sigmai, s.Nmod_sigma);
		if (id != -1) l.sigma[i] *= (1. + s.delta_sigma[id]);
	}
	if (s.Nmod_phi != 0) {
		for (int i = 0; i < 8; i++) {
			const int id = find_id(i, s.phii, s.Nmod_phi);
			if (id == -1) continue;
			if (i == 6) c[10] *= (1. + s.delta_phi[id]);
			else if (i == 5) c[8] *= (1. + s.delta_phi[id]);
			else if (i == 1) c[1] = s.delta_phi[id];
			else c[i] *= (1. + s.delta_phi[id]);
		}
		int id = find_id(8, s.phii, s.Nmod_phi);
		if (id != -1) c[9] *= (1. + s.delta_phi[id]);
		id = find_id(9, s.phii, s.Nmod_phi);
		if (id != -1) c[11] *= (1. + s.delta_phi[id]);
	}
}

// Fill the per-bin coefficient block from the physical quantities.  Tables come in as pointers so the same code runs
// on the device (constant/global memory) and in the host test harness.
template <class Fam>
GWAT_HD void phenomd_setup(SrcQ &s, const double (*fit)[11], const double (*qnm)[5], int qnm_n, DCoef &c);

}  // namespace gwat
#endif
