// Cooperative per-walker setup: the stages of walker_setup (gwat_phenomp.h) dealt out over ROLES that only meet at a few values.
//
// One thread per walker is a dependent FP64 chain of ~12 k instructions: 69 us (IMRPhenomPv2) / 41 us (IMRPhenomD) for a
// single warp per SM whatever the batch size -- 13 % / 25 % of the cfg2 / cfg1 step.  The chain is not one chain, though
// (profiles/r02_c_setup_stages_before.json): ringdown frequencies + amplitude collocation, phase coefficients + C1 matching,
// twist-up angles and detector geometry are independent up to a handful of numbers.  In k_setup (gwat_engine.cu) a CTA sets up 32
// walkers, lane = walker, WARP = ROLE:
//   role 0  phase   : fit rows 7-18, PN phase coefficients, modifications | fRD, fdamp <- role 1 | common block, matching, t/phi reference
//   role 1  amp     : spin projections, QNM spline -> fRD, fdamp | fit rows 0-6, PN amplitude coefficients, collocation
//   role 2  detector: antenna patterns, arrival times, refusal flags; ga/gb once the twist block is there
//   role 3  twist   : (IMRPhenomPv2) frame angles, Euler-angle coefficients, offsets; the spline of the time shift
// joined through one SetupRec per walker in shared memory, with a barrier between the steps below.  The 10 phase samples of
// the IMRPhenomPv2 time shift are dealt out over all roles once the carrier block is complete.
// Every value is computed by the same expressions as in walker_setup, which the Fisher stencil still uses: same bits
// (tests/test_host_math.py runs the steps role by role on the host, with everything a role does not own poisoned).
#ifndef GWAT_SETUP_COOP_H
#define GWAT_SETUP_COOP_H

#include <stddef.h>

#include "gwat_phenomp.h"

namespace gwat {

constexpr int kCoefWords = (int)(sizeof(WalkerCoef) / sizeof(double));
template <class Fam>
GWAT_HD constexpr int setup_roles()
{
	return Fam::base == BASE_P ? 4 : 3;
}
enum SetupRole { ROLE_PHASE = 0, ROLE_AMP = 1, ROLE_DETECTOR = 2, ROLE_TWIST = 3 };

struct SetupRec {
	WalkerCoef w;
	double fRD, fdamp;
	double ys[kTimeShiftSamples];
	int ts_ok, refused;
	double pad_;
};
static_assert(sizeof(SetupRec) % 16 == 8, "an odd number of doubles per record keeps lane-strided 8-byte accesses free of bank conflicts");
// words [kAmpBegin, kAmpEnd) of DCoef belong to the amplitude role
constexpr int kAmpBegin = (int)(offsetof(DCoef, A0) / sizeof(double)), kAmpEnd = (int)(offsetof(DCoef, fRD) / sizeof(double));

// What a role carries from step 1 to step 2 (registers / local memory of its thread).
struct SetupCarry {
	SrcQ q;
	Lambda lam;
	PhasePrep pp;
};

GWAT_HD void copy_words(const void *src, void *dst, int begin, int end)
{
	const double *s = static_cast<const double *>(src);
	double *d = static_cast<double *>(dst);
	for (int i = begin; i < end; i++) d[i] = s[i];
}

// ---- step 1: everything that needs nothing from another role -------------------------------------------------------------
template <class Fam>
GWAT_HD void setup_step1(int role, const gwat_b200_source &s, const Network &net, const Tables &t, int theory, SetupCarry &k, SetupRec &r)
{
	constexpr bool kP = Fam::base == BASE_P;
	SrcQ &q = k.q;
	if (role == ROLE_PHASE) {
		populate_source(s, q);
		copy_modifications<Fam>(s, q);
		if (Fam::ppe != PPE_NONE) apply_theory(theory, t.dz, q);
		if (Fam::nrt) nrt_prepare_source(s, q);
		phenomd_setup_phase_prep<Fam>(q, t.fit, k.lam, k.pp);
		double g[2];
		phenomd_fit_rows(t.fit, q.eta, q.chi_pn, 5, 2, g);  // gamma[1], gamma[2]: fpeak
		k.lam.gamma[1] = g[0];
		k.lam.gamma[2] = g[1];
	} else if (role == ROLE_AMP) {
		populate_source(s, q);
		if (kP) phenompv2_spin_projection(q, (s.chip + 1) > 1e-10);  // chi_p enters the remnant spin
		phenomd_setup_remnant<Fam>(q, t.qnm, t.qnm_n);
		r.fRD = q.fRD;
		r.fdamp = q.fdamp;
	} else if (role == ROLE_DETECTOR) {
		detector_setup_source(net, s, r.w.det);
		if (!kP) {
			const double ci = sm::cos(s.incl_angle);
			r.w.cfac = ci;
			r.w.pfac = .5 * (1. + ci * ci);
		} else {
			r.w.cfac = 0;
			r.w.pfac = 0;
		}
		r.w.pad_ = 0;
		// options of the reference that are outside this path are refused loudly (NaN), never silently approximated (walker_setup)
		r.refused = (Fam::nrt && s.tidal_love_error) ? 1 : 0;
	} else {
		populate_source(s, q);
		phenompv2_param_transform(q, (s.chip + 1) > 1e-10);
		phenomp_setup_angles(q, r.w.p);
	}
}

// ---- step 2 (fRD, fdamp published): the two halves of the carrier block ----------------------------------------------------
template <class Fam>
GWAT_HD void setup_step2(int role, const Tables &t, SetupCarry &k, SetupRec &r)
{
	SrcQ &q = k.q;
	DCoef c;
	if (role == ROLE_PHASE) {
		q.fRD = r.fRD;
		q.fdamp = r.fdamp;
		phenomd_setup_boundaries(q, k.lam.gamma[1], k.lam.gamma[2]);
		phenomd_setup_common(q, c);
		phenomd_setup_phase<Fam>(q, k.lam, k.pp, c);
		copy_words(&c, &r.w.d, 0, kAmpBegin);
		copy_words(&c, &r.w.d, kAmpEnd, (int)(sizeof(DCoef) / sizeof(double)));
	} else if (role == ROLE_AMP) {
		double v[7];
		phenomd_fit_rows(t.fit, q.eta, q.chi_pn, 0, 7, v);
		for (int i = 0; i < 3; i++) k.lam.rho[i] = v[i];
		k.lam.v2 = v[3];
		for (int i = 0; i < 3; i++) k.lam.gamma[i] = v[i + 4];
		phenomd_setup_boundaries(q, k.lam.gamma[1], k.lam.gamma[2]);
		c.M = q.M;  // the two fields of the common block the amplitude expressions read
		c.fRD = q.fRD;
		phenomd_setup_amp(q, k.lam, c);
		copy_words(&c, &r.w.d, kAmpBegin, kAmpEnd);
	}
}

// ---- step 3 (carrier block, twist angles, detector geometry in the record) ---------------------------------------------------
template <class Fam>
GWAT_HD void setup_step3(int role, const Network &net, SetupRec &r)
{
	constexpr bool kP = Fam::base == BASE_P;
	if (kP) {
		// time shift: samples role, role + 4, ... from the finished carrier block
		bool ok = true;
		for (int j = role; j < kTimeShiftSamples; j += setup_roles<Fam>()) {
			double x, y;
			ok = phenomp_time_shift_sample<Fam>(r.w.d, j, x, y) && ok;
			r.ys[j] = y;
		}
		if (role == ROLE_PHASE) r.ts_ok = ok ? 1 : 0;  // (the window is degenerate for every sample or for none)
	}
	if (role == ROLE_DETECTOR) {
		for (int d = 0; d < net.D; d++) {
			DetCoef &dc = r.w.det[d];
			if (kP) {
				dc.ga = dc.Fplus * r.w.p.c2z - dc.Fcross * r.w.p.s2z;
				dc.gb = dc.Fplus * r.w.p.s2z + dc.Fcross * r.w.p.c2z;
			} else {
				dc.ga = dc.Fplus * r.w.pfac;
				dc.gb = dc.Fcross * r.w.cfac;
			}
		}
	}
}

// ---- step 4 (IMRPhenomPv2; samples in the record): the spline of the time shift ------------------------------------------------
template <class Fam>
GWAT_HD void setup_step4(int role, bool shift_time, SetupRec &r)
{
	if (Fam::base != BASE_P || role != ROLE_TWIST) return;
	double tc2pi = 2 * GWAT_PI * 0.0;
	if (shift_time && r.ts_ok) {
		const double f_final = r.w.d.fRD;
		const double start = .8 * f_final, stop = 1.2 * f_final;
		const double step = (stop - start) / (kTimeShiftSamples - 1);
		double xs[kTimeShiftSamples], ys[kTimeShiftSamples];
		for (int j = 0; j < kTimeShiftSamples; j++) {
			xs[j] = start + j * step;
			ys[j] = r.ys[j];
		}
		tc2pi = phenomp_time_shift_finish(r.w.d, xs, ys);
	}
	r.w.p.tcorr_2pi = tc2pi;
}

}  // namespace gwat
#endif
