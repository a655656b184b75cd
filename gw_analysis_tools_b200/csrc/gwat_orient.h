// Orientation conventions of the single-detector entry points (GWAT_HD code, run on the host before a batch is uploaded).
//
//   equatorial_orientation: the source is described by the direction of L (theta_l, phi_l) in equatorial coordinates instead of
//       (incl_angle, psi); transform_orientation_coords (src/waveform_util.cpp:1535-1595) derives incl_angle and psi from it, for
//       IMRPhenomPv2 through the direction of J (PhenomPv2_JSF_from_params src/IMRPhenomP.cpp:776-816, equatorial_from_SF
//       src/util.cpp:1440-1485, terr_pol_iota_from_equat_sph src/util.cpp:1362-1378).  The reference applies it in
//       fourier_detector_response (one detector) and calculate_snr; the coherent network response and therefore the likelihood
//       read incl_angle / psi as given (create_coherent_GW_detection_reuse_WF, src/waveform_util.cpp:153-184), and so does this
//       library.
//   horizon_coord: the sky position is (theta, phi) in the detector's own frame; fourier_detector_response then uses the
//       right-angle interferometer patterns (right_interferometer, src/detector_util.cpp:497-531) and no arrival-time shift.
// Terrestrial detectors only (the reference's LISA / ecliptic branch is outside this path).
#ifndef GWAT_ORIENT_H
#define GWAT_ORIENT_H

#include "gwat_phenomp.h"

namespace gwat {

// psi and the inclination of the direction (thetaj, phij) for a source at (RA, DEC)   (terr_pol_iota_from_equat_sph)
GWAT_HD void terr_pol_iota_from_equat_sph(double RA, double DEC, double thetaj, double phij, double &pol, double &iota)
{
	const double temp_pol = atan(cos(DEC) * 1. / tan(thetaj) * 1. / sin(phij - RA) - 1. / tan(phij - RA) * sin(DEC));
	pol = GWAT_PI / 2. - temp_pol;
	iota = acos(-(cos(thetaj) * sin(DEC)) - cos(DEC) * cos(phij - RA) * sin(thetaj));
}

// Unit vector of J in the source frame (L along z) at f_ref, from (chip, phip)   (PhenomPv2_JSF_from_params)
GWAT_HD void phenompv2_jsf(const gwat_b200_source &p, double *JSF)
{
	const double chi1_l = p.spin1[2], chi2_l = p.spin2[2];
	const double eta = (p.mass1 * p.mass2) / pow(p.mass1 + p.mass2, 2);
	const double m1_2 = p.mass1 * p.mass1, m2_2 = p.mass2 * p.mass2;
	// L2PN from the powers of pi M f_ref the reference tabulates (precalc_powers_PI / precalc_powers_ins, src/IMRPhenomD.cpp:846-892)
	const double Mf = (p.mass1 + p.mass2) * GWAT_MSOL_SEC * p.f_ref;
	const double sixth = pow(Mf, 1. / 6.);
	const double mf_third = sixth * sixth, mf_two3 = mf_third * mf_third;
	const PiPowers pi = pi_powers();
	const double x = mf_two3 * pi.two3;
	const double L0 = ((p.mass1 + p.mass2) * (p.mass1 + p.mass2)) * l2pn(eta, x, sqrt(x));
	const double J0x = m1_2 * p.chip * cos(p.phip), J0y = m1_2 * p.chip * sin(p.phip);
	const double J0z = L0 + m1_2 * chi1_l + m2_2 * chi2_l;
	const double J0 = sqrt(J0x * J0x + J0y * J0y + J0z * J0z);
	JSF[0] = J0x / J0;
	JSF[1] = J0y / J0;
	JSF[2] = J0z / J0;
}

// Source-frame vector -> equatorial frame   (equatorial_from_SF: the rotation fixed by L, N and L x N in both frames)
GWAT_HD void equatorial_from_sf(const double *SF, double thetal, double phil, double thetas, double phis, double iota, double phi_ref, double *EQ)
{
	const double cp = cos(GWAT_PI / 2. - phi_ref), sp = sin(GWAT_PI / 2. - phi_ref);
	const double ci = cos(iota), si = sin(iota);
	const double ctl = cos(thetal), stl = sin(thetal), cts = cos(thetas), sts = sin(thetas);
	const double cps = cos(phis), sps = sin(phis), cpl = cos(phil), spl = sin(phil);
	const double Jxn = SF[0], Jyn = SF[1], Jzn = SF[2];
	const double cp2 = cp * cp;  // pow_int(cp, 2)
	EQ[0] = (cp2 * cpl * Jzn * si * stl - ci * cpl * (cp * Jxn + Jyn * sp) * stl - cp * (cts * Jyn * spl * stl + cps * Jxn * sts - ctl * Jyn * sps * sts) +
	         sp * (cpl * Jzn * si * sp * stl + cts * Jxn * spl * stl - (cps * Jyn + ctl * Jxn * sps) * sts)) /
	        si;
	EQ[1] = (cp2 * Jzn * si * spl * stl +
	         sp * (-(cpl * cts * Jxn * stl) - ci * Jyn * spl * stl + Jzn * si * sp * spl * stl + cps * ctl * Jxn * sts - Jyn * sps * sts) +
	         cp * (cpl * cts * Jyn * stl - ci * Jxn * spl * stl - (cps * ctl * Jyn + Jxn * sps) * sts)) /
	        si;
	EQ[2] = (ctl * Jzn * si - Jxn * (ci * cp * ctl + cp * cts + sp * (cps * spl - cpl * sps) * stl * sts) -
	         Jyn * (ci * ctl * sp + cts * sp + cp * (-(cps * spl) + cpl * sps) * stl * sts)) /
	        si;
}

// transform_orientation_coords for a terrestrial detector: incl_angle and psi of `p` from (theta_l, phi_l).
GWAT_HD void transform_orientation_coords(gwat_b200_source &p, bool pv2)
{
	const double theta_s = GWAT_PI / 2. - p.DEC, phi_s = p.RA;
	// N points from the source to the detector
	const double Neq[3] = {-sin(theta_s) * cos(phi_s), -sin(theta_s) * sin(phi_s), -cos(theta_s)};
	const double Leq[3] = {sin(p.theta_l) * cos(p.phi_l), sin(p.theta_l) * sin(p.phi_l), cos(p.theta_l)};
	p.incl_angle = acos(Neq[0] * Leq[0] + Neq[1] * Leq[1] + Neq[2] * Leq[2]);
	if (pv2) {
		double JSF[3], Jeq[3];
		phenompv2_jsf(p, JSF);
		equatorial_from_sf(JSF, p.theta_l, p.phi_l, theta_s, phi_s, p.incl_angle, p.phiRef, Jeq);
		const double r = sqrt(Jeq[0] * Jeq[0] + Jeq[2] * Jeq[2] + Jeq[1] * Jeq[1]);  // transform_cart_sph (src/util.cpp:1882-1890)
		const double theta_j = acos(Jeq[2] / r);
		double phi_j = atan2(Jeq[1], Jeq[0]);
		if (phi_j < 0) phi_j += 2 * GWAT_PI;
		double iota_j;  // (the inclination of J is not used: incl_angle stays the one of L)
		terr_pol_iota_from_equat_sph(p.RA, p.DEC, theta_j, phi_j, p.psi, iota_j);
	} else {
		terr_pol_iota_from_equat_sph(p.RA, p.DEC, p.theta_l, p.phi_l, p.psi, p.incl_angle);
	}
}

}  // namespace gwat
#endif
