// NRTidal-v2 additions to the IMRPhenomD carrier (GWAT_HD code).
//
// Reference being replaced:
//   per walker  prep_source_parameters, NRT block          src/waveform_generator.cpp:1316-1358 (binary Love, weighted tides)
//               binary_love_relation                        src/IMRPhenomD_NRT.cpp:83-128
//               calculate_quad/oct_moment                   :360-384
//               calculate_spin_coefficients_3p5             :387-414
//               assign_static_pn_phase_coeff (override)     :131-172
//               calculate_NRT_amp_coefficient               :512-516
//               setup part of construct_waveform            :594-611
//   per bin     Pade / phase_ins_NRT                        :177-246
//               phase_spin_NRT                              :417-509
//               phase_ins_NRT_D (dissipative tide)          :252-264
//               amp_ins_NRT                                 :519-541
//               taper                                       :546-588
//               loop body of construct_waveform             :697-746
// Quirk kept on purpose: with default gen_params the dissipative-tide phase is evaluated
// with diss_tidal_weighted = -1 (the member default is never overwritten).
#ifndef GWAT_NRT_H
#define GWAT_NRT_H

#include "gwat_phenomd.h"

namespace gwat {

// Pade coefficients of the NRTidal-v2 phase as LAL spells them (the reference copies them: include/gwat/IMRPhenomD_NRT.h:73-74)
#define GWAT_NRT_N0 -12.615214237993088
#define GWAT_NRT_N1 19.0537346970349
#define GWAT_NRT_N2 -21.166863146081035
#define GWAT_NRT_N3 90.55082156324926
#define GWAT_NRT_N4 -60.25357801943598
#define GWAT_NRT_D0 -15.111207827736678
#define GWAT_NRT_D1 22.195327350624694
#define GWAT_NRT_D2 8.064109635305156

// lambda_a from lambda_s and the mass ratio (binary Love relation, arXiv:1903.03909 eqs. 11-13)
GWAT_HD void binary_love(double tidal_s, double mass1, double mass2, double &tidal1, double &tidal2)
{
	const double n_fit = 0.743;
	const double b[3][2] = {{-14.40, 14.45}, {31.36, -32.25}, {-22.44, 20.35}};
	const double cc[3][2] = {{-15.25, 15.37}, {37.33, -43.20}, {-29.93, 35.18}};
	const double q = mass2 / mass1;
	const double Q = sm::pow(q, 10. / (3. - n_fit));
	const double F = (1. - Q) / (1. + Q);
	double num = 1, den = 1;
	const double qp[2] = {q, q * q};
	double lp[3];
	lp[0] = sm::pow(tidal_s, -1. / 5.);
	lp[1] = lp[0] * lp[0];
	lp[2] = lp[0] * lp[1];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 2; j++) {
			num += b[i][j] * qp[j] * lp[i];
			den += cc[i][j] * qp[j] * lp[i];
		}
	const double tidal_a = F * (num / den) * tidal_s;
	tidal1 = tidal_s - tidal_a;
	tidal2 = tidal_s + tidal_a;
}

// spin-induced quadrupole and octupole moments of a neutron star from its tidal deformability (arXiv:1608.02582 eq. 15)
GWAT_HD double ns_quad_moment(double lambda)
{
	const double l = sm::log(lambda);
	return sm::exp(0.1940 + 0.09163 * l + 0.04812 * pow(l, 2.) + -0.004283 * pow(l, 3.) + 0.00012450 * pow(l, 4.));
}
GWAT_HD double ns_oct_moment(double quad)
{
	const double l = sm::log(quad);
	return sm::exp(0.003131 + 2.071 * l + -0.7152 * pow(l, 2.) + 0.2458 * pow(l, 3.) + -0.03309 * pow(l, 4.));
}

// prep_source_parameters' NRT block: tidal deformabilities -> mass-weighted combinations.
GWAT_HD void nrt_prepare_source(const gwat_b200_source &in, SrcQ &s)
{
	double t1 = in.tidal1, t2 = in.tidal2;
	s.tidal1 = s.tidal2 = -1;
	s.tidal_weighted = -1;
	s.delta_tidal_weighted = -1;
	s.diss_tidal_weighted = -1;
	if (in.tidal_love) {
		binary_love(in.tidal_s, s.mass1, s.mass2, t1, t2);
		s.tidal1 = t1;
		s.tidal2 = t2;
	}
	if ((t1 < 0 || t2 < 0) && in.tidal_weighted >= 0) {
		s.tidal_weighted = in.tidal_weighted;
	} else if (t1 >= 0 && t2 >= 0) {
		s.tidal1 = t1;
		s.tidal2 = t2;
		const double eta = s.eta;
		s.tidal_weighted = 8. / 13. * ((1. + 7. * eta - 31. * eta * eta) * (t1 + t2) +
		                               sqrt(1. - 4. * eta) * (1. + 9. * eta - 11. * eta * eta) * (t1 - t2));
		s.delta_tidal_weighted =
		    1. / 2. * (sqrt(1. - 4. * eta) * (1. - 13272. / 1319. * eta + 8944. / 1319. * eta * eta) * (t1 + t2) +
		               (1. - 15910. / 1319. * eta + 32850. / 1319. * eta * eta + 3380. / 1319. * eta * eta * eta) * (t1 - t2));
	}
	if ((in.diss_tidal1 < 0 || in.diss_tidal2 < 0) && in.diss_tidal_weighted >= 0) {
		s.diss_tidal_weighted = in.diss_tidal_weighted;
	} else if (in.diss_tidal1 >= 0 && in.diss_tidal2 >= 0) {
		const double eta = s.eta;
		s.diss_tidal_weighted = (2.0 * pow(eta, 2) - 4.0 * eta + 1.0) * 0.5 * (in.diss_tidal1 + in.diss_tidal2) -
		                        sqrt(1.0 - 4.0 * eta) * (1.0 - 2.0 * eta) * 0.5 * (in.diss_tidal1 - in.diss_tidal2);
	}
}

// Moments and spin-spin coefficients; called before the PN phase coefficients are assembled.
GWAT_HD void nrt_moments(SrcQ &s)
{
	if (s.tidal1 <= 0) { s.oct1 = 1; s.quad1 = 1; }
	else { s.quad1 = ns_quad_moment(s.tidal1); s.oct1 = ns_oct_moment(s.quad1); }
	if (s.tidal2 <= 0) { s.oct2 = 1; s.quad2 = 1; }
	else { s.quad2 = ns_quad_moment(s.tidal2); s.oct2 = ns_oct_moment(s.quad2); }
}
// quadrupole-monopole terms added to the 2PN and 3PN phase coefficients (arXiv:1905.06011 eq. 27)
GWAT_HD void nrt_adjust_pn_phase(const SrcQ &s, double *c)
{
	const double XA = s.mass1 / s.M, XB = s.mass2 / s.M;
	const double XA2 = XA * XA, XB2 = XB * XB;
	const double c1 = s.spin1z, c2 = s.spin2z;
	const double c1s = c1 * c1, c2s = c2 * c2;
	const double ssA2 = -50 * (s.quad1 - 1.) * XA2 * c1s;
	const double ssB2 = -50 * (s.quad2 - 1.) * XB2 * c2s;
	const double ssA3 = (5 / 84.) * (9407 + 8218 * XA - 2016 * XA2) * (s.quad1 - 1.) * XA2 * c1s;
	const double ssB3 = (5 / 84.) * (9407 + 8218 * XB - 2016 * XB2) * (s.quad2 - 1.) * XB2 * c2s;
	c[4] += ssA2 + ssB2;
	c[10] += ssA3 + ssB3;
}

GWAT_HD void nrt_setup(const SrcQ &s, DCoef &c)
{
	c.nrt_phase_coeff = -(3. / 16.) * s.tidal_weighted * (39. / (16. * s.eta));
	{
		const double XA = s.mass1 / s.M, XB = s.mass2 / s.M;
		const double XA2 = XA * XA, XB2 = XB * XB;
		const double c1 = s.spin1z, c2 = s.spin2z;
		const double c1s = c1 * c1, c2s = c2 * c2;
		const double A = 10 * ((XA2 + (308. / 3.) * XA) * c1 + (XB2 - (89 / 3.) * XB) * c2 - 40 * GWAT_PI) * (s.quad1 - 1.) * XA2 * c1s -
		                 440 * (s.oct1 - 1.) * XA2 * XA * c1s * c1;
		const double B = 10 * ((XB2 + (308. / 3.) * XB) * c2 + (XA2 - (89 / 3.) * XA) * c1 - 40 * GWAT_PI) * (s.quad2 - 1.) * XB2 * c2s -
		                 440 * (s.oct2 - 1.) * XB2 * XB * c2s * c2;
		c.nrt_ss_coeff = A + B;
	}
	c.nrt_amp_coeff = -sqrt(5 * GWAT_PI * s.eta / 24.) * (9 * s.M * s.M / s.DL) * (3. / 16.) * s.tidal_weighted;
	c.nrt_diss_coeff = -(75. / 512.) * (1.0 / s.eta) * s.diss_tidal_weighted;
	{
		// merger frequency of the tidal fit (arXiv:1804.02235 eq. 11) with LAL's digits, as in the reference's taper()
		const double kappa = (3. / 16.) * s.tidal_weighted;
		const double a0 = 0.3586, n1 = 3.35411203e-2, n2 = 4.31460284e-5, d1 = 7.54224145e-2, d2 = 2.23626859e-4;
		c.nrt_fmerger = (1. / (2. * s.M * GWAT_PI)) * a0 * sqrt(s.mass2 / s.mass1) * (1.0 + n1 * kappa + n2 * kappa * kappa) /
		                (1.0 + d1 * kappa + d2 * kappa * kappa);
		c.nrt_fmerger12 = c.nrt_fmerger * 1.2;
	}
}

// Tidal additions to amplitude and phase at one bin (before the time/phase shift), and the Planck-taper factor.
GWAT_HD void nrt_bin(const DCoef &c, double f, const MfPowers &p, double logf, double &amp, double &phase)
{
	const PiPowers pi = pi_powers();
	const double x = mul_rn(p.two3, pi.two3);
	const double x32 = mul_rn(mul_rn(c.M, f), GWAT_PI);
	const double x2 = mul_rn(x, x), x52 = mul_rn(x, x32), x3 = mul_rn(mul_rn(x, x), x);
	double P = 1;
	P = add_rn(P, mul_rn(GWAT_NRT_N0, x));
	P = add_rn(P, mul_rn(GWAT_NRT_N1, x32));
	P = add_rn(P, mul_rn(GWAT_NRT_N2, x2));
	P = add_rn(P, mul_rn(GWAT_NRT_N3, x52));
	P = add_rn(P, mul_rn(GWAT_NRT_N4, x3));
	double Pd = 1;
	Pd = add_rn(Pd, mul_rn(GWAT_NRT_D0, x));
	Pd = add_rn(Pd, mul_rn(GWAT_NRT_D1, x32));
	Pd = add_rn(Pd, mul_rn(GWAT_NRT_D2, x2));
	const double pade = mul_rn(x52, P) / Pd;
	phase = add_rn(phase, mul_rn(c.nrt_phase_coeff, pade));
	// 3.5PN quadrupole-monopole / octupole spin term
	const double xm52 = mul_rn(p.m53, pi.m53);
	const double x72 = mul_rn(p.seven3, pi.seven3);
	const double coeff = mul_rn(c.k128, xm52);  // 3/(128 eta) * x^(-5/2)
	phase = add_rn(phase, mul_rn(coeff, mul_rn(c.nrt_ss_coeff, x72)));
	// tidal amplitude (already in strain units: it is multiplied by the same A0 M^{7/6} prefactor in the reference)
	// x = (pi M f)^(2/3):  x^(13/4) = x^3 (pi M f)^(1/6),  x^2.89 = exp(2.89 * 2/3 * ln(pi M f)); ln(pi M f) = ln(pi M) + ln f
	// comes from the walker constant and the grid table.  (The reference calls pow() twice and log() once per bin here.)
	const double x4 = (x2 * x2);
	const double log_piMf = c.logpiM + logf;
	const double x134 = x3 * (1.2102032422537643 * p.sixth);  // pi^(1/6) (M f)^(1/6)
	const double x289 = exp((2.89 * (2. / 3.)) * log_piMf);
	const double ampNRT = c.nrt_amp_coeff * x134 * (1 + (449. / 108) * x + (22672. / 9.) * x289) * fast_rcp(1 + 13477.8 * x4);
	amp += c.A0 * ampNRT;
	// dissipative tide
	const double piMf = mul_rn(GWAT_PI * c.M, f);
	phase = add_rn(phase, mul_rn(mul_rn(c.nrt_diss_coeff, piMf), log_piMf));
}

GWAT_HD double nrt_taper_factor(const DCoef &c, double f)
{
	const double fm = c.nrt_fmerger, fm12 = c.nrt_fmerger12;
	double taper;
	if (f < fm) taper = 0.0;
	else if (fm < f && f < fm12) {
		const double z = (fm - fm12) / (f - fm) + (fm - fm12) / (f - fm12);
		taper = 1.0 / (exp(-z) + 1.0);
	} else if (f > fm12) taper = 1.0;
	else taper = -1.0;  // f exactly on a boundary: the reference returns -1 there (src/IMRPhenomD_NRT.cpp:587)
	return 1.0 - taper;
}

}  // namespace gwat
#endif
