// Per-(walker, bin) evaluation on top of the carrier: polarisations, detector responses, likelihood integrands.
// GWAT_HD code shared by the CUDA kernels (gwat_engine.cu) and the CPU-side test harness.
//
// Reference being replaced, per bin:
//   waveform[j] = amp * exp(-i phase)                                   src/IMRPhenomD.cpp:497-498
//   h+ = (1+cos^2 i)/2 h,  hx = -i cos(i) h                             src/waveform_generator.cpp:188-199
//   r_d = F+ h+ + Fx hx                                                 src/waveform_util.cpp:871-872
//   r_d *= exp(i (-2 pi DTOA_d) f)                                      src/waveform_util.cpp:177-179
//   |r|^2/S and Re(d conj(r))/S                                         src/mcmc_gw.cpp:815,838
#ifndef GWAT_BINS_H
#define GWAT_BINS_H

#include "gwat_phenomp.h"

namespace gwat {

// What the kernels read per frequency bin (built once per grid by the host, see gwat_grid.h).
struct BinGrid {
	const double *f;      // [L] Hz
	const double *sf_hi;  // [L] f^(fl(1/6)) double-double
	const double *sf_lo;
	const double *logf;   // [L] ln f
};


// The (2,2) carrier h = A exp(-i phi) of the IMRPhenomD families before the time and phase shifts: amplitude (NRT: already
// tapered) and phase.
template <class Fam>
GWAT_HD void carrier_parts(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, PolParts &pp)
{
	const DCoef &c = w.d;
	// NRT: the Planck taper multiplies everything above 1.2 f_merger by exactly zero (src/IMRPhenomD_NRT.cpp:582-585, 745)
	pp.zero = f > c.fcut || (Fam::nrt && f > c.nrt_fmerger12);
	if (pp.zero) return;
	double amp, phase;
	MfPowers p;
	phenomd_bin<Fam>(c, f, bin_sixth_root(c, sf_hi, sf_lo), logf, amp, phase, p);
	if (Fam::nrt) {
		nrt_bin(c, f, p, logf, amp, phase);
		amp *= nrt_taper_factor(c, f);
	}
	pp.amp = amp;
	pp.phase = phase;
}

// Polarisations of one bin in two steps (all families): the parts that do not involve the coalescence time, and the finish
// with the coefficient `tc` of (f - f_ref) -- w.d.tc / w.p.tc unless the caller re-times the point (Fisher stencil).
template <class Fam>
GWAT_HD void polarization_parts(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, PolParts &pp)
{
	if (Fam::base == BASE_P) phenomp_polarization_parts<Fam>(w, f, sf_hi, sf_lo, logf, pp);
	else carrier_parts<Fam>(w, f, sf_hi, sf_lo, logf, pp);
}
template <class Fam>
GWAT_HD double walker_time_coefficient(const WalkerCoef &w)
{
	return Fam::base == BASE_P ? w.p.tc : w.d.tc;
}
template <class Fam>
GWAT_HD void polarizations_finish(const WalkerCoef &w, const PolParts &pp, double tc, double f, cplx &hp, cplx &hc)
{
	if (Fam::base == BASE_P) {
		phenomp_polarizations_finish<Fam>(w, pp, tc, f, hp, hc);
		return;
	}
	if (pp.zero) {
		hp = cplx{0.0, 0.0};
		hc = cplx{0.0, 0.0};
		return;
	}
	// waveform[j] = amp exp(-i phase); h+ = (1+cos^2 i)/2 h, hx = -i cos(i) h
	double sn, cs;
	fast_sincos(phenomd_apply_time_phase(w.d, tc, f, pp.phase), &sn, &cs);
	const cplx h{pp.amp * cs, -(pp.amp * sn)};
	hp = cplx{h.re * w.pfac, h.im * w.pfac};
	hc = cplx{h.im * w.cfac, -h.re * w.cfac};
}

// fourier_waveform semantics.
template <class Fam>
GWAT_HD void polarizations_bin(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, cplx &hp, cplx &hc)
{
	PolParts pp;
	polarization_parts<Fam>(w, f, sf_hi, sf_lo, logf, pp);
	polarizations_finish<Fam>(w, pp, walker_time_coefficient<Fam>(w), f, hp, hc);
}

// fourier_amplitude / fourier_phase values of one bin for the IMRPhenomD families (construct_amplitude / construct_phase,
// src/IMRPhenomD.cpp:604-740): amplitude zero above 0.2/M, phase -(phi - tc (f - f_ref) - phic) with no cutoff.
template <class Fam>
GWAT_HD void amplitude_phase_bin(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, double &amp, double &phase)
{
	MfPowers p;
	phenomd_bin<Fam>(w.d, f, bin_sixth_root(w.d, sf_hi, sf_lo), logf, amp, phase, p);
	if (f > w.d.fcut) amp = 0.0;
	phase = -phenomd_apply_time_phase(w.d, f, phase);
}

// One bin of the sky-averaged derivative (calculate_derivatives, src/fisher.cpp:183-338): central differences of amplitude
// and phase; order 2 forms dA + i A dphi, order 4 dA - i A dphi (as the reference has it; the Fisher matrix does not
// depend on that sign).  a[k], ph[k]: the stencil points (+eps, -eps, +2eps, -2eps); amp0: the unperturbed amplitude.
GWAT_HD cplx sky_derivative_bin(int npts, const double *a, const double *ph, double amp0)
{
	const double epsilon = 1e-8;
	if (npts == 2) {
		const double da = (a[0] - a[1]) / (2 * epsilon), dp = (ph[0] - ph[1]) / (2 * epsilon);
		return cplx{da, dp * amp0};
	}
	const double da = (((-a[2] + 8. * a[0]) - 8. * a[1]) + a[3]) / (12. * epsilon);
	const double dp = (((-ph[2] + 8. * ph[0]) - 8. * ph[1]) + ph[3]) / (12. * epsilon);
	return cplx{da, -(dp * amp0)};
}

// Response of detector d including the arrival-time phase (create_coherent_GW_detection_reuse_WF semantics);
// with_shift = false gives fourier_detector_response semantics.
GWAT_HD cplx project_bin(const DetCoef &dc, cplx hp, cplx hc, double f, bool with_shift)
{
	cplx r{dc.Fplus * hp.re + dc.Fcross * hc.re, dc.Fplus * hp.im + dc.Fcross * hc.im};
	if (with_shift && dc.tshift != 0.0) {  // (a zero shift multiplies by exactly 1: the reference detector, and most stencil points)
		double sn, cs;
		fast_sincos(mul_rn(dc.tshift, f), &sn, &cs);
		r = cplx{r.re * cs - r.im * sn, r.re * sn + r.im * cs};
	}
	return r;
}

}  // namespace gwat
#endif
