// gwat_b200 engine: CUDA kernels for sm_100a and the C ABI of include/gwat_b200.h.
//
// Data layout in HBM (per context = per GPU):
//   grid tables      f[L], sf_hi[L], sf_lo[L], logf[L]                      shared by every walker, L2-resident
//   network tables   wq[D][L] = quadrature coefficient / PSD,  data_re[D][L], data_im[D][L]
//   per call         params[W][P] -> WalkerCoef[W] (setup kernel) -> partial[W][units][8 warps][2] -> logL[W]
// Kernels:
//   k_setup_mcmc / k_setup_src   one thread per walker: sampling vector or physical record -> WalkerCoef
//   k_loglike                    grid (walkers, bin chunks); threads stride over consecutive bins (coalesced table reads,
//                                region branches diverge only at the per-walker boundaries); amplitude/phase in FP64,
//                                detector projection and PSD weighting in registers, per-CTA seed tables for the
//                                recurrences, warp-shuffle reduction to one partial per (walker, unit, warp)
//   k_finish                     fixed-order sum of a walker's partials (one warp per walker) -> logL
//   k_waveform / k_response      the same per-bin code writing polarisations / responses (API parity entry points)
// There is no CPU fallback anywhere in this file.
#include <atomic>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <cctype>
#include <string>
#include <vector>

#define GWAT_TABLE_QUALIFIER static __device__ const
#include "gwat_tables.inc"
#undef GWAT_TABLE_QUALIFIER
namespace hosttab {  // the same generated tables for host-side use (detector rows)
#define GWAT_TABLE_QUALIFIER static const
#include "gwat_tables.inc"
#undef GWAT_TABLE_QUALIFIER
}  // namespace hosttab

#include "gwat_like.h"
#include "gwat_grid.h"
#include "gwat_method.h"
#include "gwat_repack.h"
#include "gwat_setup_coop.h"
#include "gwat_orient.h"

using namespace gwat;

namespace {

constexpr int kThreads = 256;
// Kernel experiments only (never the shipped library): -DGWAT_EXPERIMENT_SLIM compiles 4 of the 11 families and 2 of the 5
// detector counts so that a variant builds in a minute.
#ifdef GWAT_EXPERIMENT_SLIM
#define GWAT_FULL_ONLY(...)
#else
#define GWAT_FULL_ONLY(...) __VA_ARGS__
#endif
#ifndef GWAT_LOGLIKE_THREADS
#define GWAT_LOGLIKE_THREADS 256
#endif
constexpr int kLikeThreads = GWAT_LOGLIKE_THREADS;
// bin tiles (of kThreads bins) one CTA of the Fisher derivative kernel evaluates with one staging of its coefficient blocks
constexpr int kFisherTilesPerCta = 4;
// detectors one Fisher pass handles (sizes FisherPlan and the kernels' staging arrays)
constexpr int kFisherMaxDetectors = 5;

// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ Tables device_tables()
{
	return Tables{gwat_phenomd_fit, gwat_qnm_knots, GWAT_QNM_N, DzTable{gwat_dz_boundaries, gwat_dz_coeffs, GWAT_DZ_SEGMENTS, GWAT_NUM_COSMOLOGIES, gwat_md_alphas, gwat_md_boundaries_z, gwat_md_coeffs, GWAT_MD_ALPHAS}};
}

typedef LikeGrid GridPtrs;


__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// Block-wide sum of two values; result valid in thread 0.
template <int NT = kThreads>
__device__ __forceinline__ void block_sum2(double &a, double &b)
{
	constexpr int kThreads = NT;
	__shared__ double sa[kThreads / 32], sb[kThreads / 32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	a = warp_sum(a);
	b = warp_sum(b);
	if (lane == 0) {
		sa[wid] = a;
		sb[wid] = b;
	}
	__syncthreads();
	if (wid == 0) {
		a = lane < kThreads / 32 ? sa[lane] : 0.0;
		b = lane < kThreads / 32 ? sb[lane] : 0.0;
		a = warp_sum(a);
		b = warp_sum(b);
	}
}

__device__ __forceinline__ bool coef_is_finite(const WalkerCoef &w, int D, bool pv2)
{
	const DCoef &c = w.d;
	bool ok = isfinite(c.fcut) && isfinite(c.A0) && isfinite(c.fRD) && isfinite(c.fdamp) && isfinite(c.beta0) &&
	          isfinite(c.alpha0) && isfinite(c.ic[4]) && isfinite(w.pfac);
	if (pv2) {
		const PCoef &p = w.p;
		ok = ok && isfinite(p.A0) && isfinite(p.tc) && isfinite(p.phic) && isfinite(p.tcorr_2pi) && isfinite(p.alpha_const) &&
		     isfinite(p.epsilon_offset) && isfinite(p.c2z) && isfinite(p.acoef[4]) && isfinite(p.Y[0]) && isfinite(p.SP);
	} else {
		ok = ok && isfinite(c.tc) && isfinite(c.phic);
	}
	for (int d = 0; d < D; d++) ok = ok && isfinite(w.det[d].Fplus) && isfinite(w.det[d].Fcross) && isfinite(w.det[d].tshift);
	return ok;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------

// Per-walker setup, cooperative (gwat_setup_coop.h): one CTA sets up kSetupWalkers walkers, lane = walker, warp = role; the
// finished records leave with coalesced stores.  params != NULL: sampling vectors (repack_mcmc_walker), else physical records.
constexpr int kSetupWalkers = 32;
constexpr size_t kSetupSmemBytes = (sizeof(SetupRec) + sizeof(RepackHeavy)) * kSetupWalkers;  // 53 KB: dynamic shared memory, opted in once per instantiation (setup_smem_opt_in)
static_assert(sizeof(RepackHeavy) % 16 == 8, "an odd number of doubles per record keeps lane-strided accesses free of bank conflicts");
static_assert(kSetupSmemBytes <= 96 * 1024, "shared memory of k_setup");

// kernel experiments only (-DGWAT_SETUP_PROFILE): clock64 stamps of lane 0 of every role of block 0
#ifdef GWAT_SETUP_PROFILE
__device__ long long g_setup_stamp[4][16];
#define GWAT_KSTAMP(i)                                                                    \
	do {                                                                                    \
		if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_setup_stamp[threadIdx.x >> 5][i] = clock64(); \
	} while (0)
#else
#define GWAT_KSTAMP(i)
#endif
template <class Fam>
__global__ void __launch_bounds__(kSetupWalkers * setup_roles<Fam>()) k_setup(const double *__restrict__ params, const gwat_b200_source *__restrict__ src_in, int W,
                                                                              RepackPlan plan, Network net, int theory, double gmst, double T_segment,
                                                                              WalkerCoef *__restrict__ out, unsigned long long *__restrict__ active_zero)
{
	constexpr bool kP = Fam::base == BASE_P;
	extern __shared__ __align__(16) unsigned char setup_smem[];
	SetupRec *recs = reinterpret_cast<SetupRec *>(setup_smem);
	if (active_zero && blockIdx.x == 0 && threadIdx.x == 0) *active_zero = 0;  // the pass's active-bin counter (k_finish / k_loglike add to it)
	const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
	const int w = blockIdx.x * kSetupWalkers + lane;
	const bool active = w < W;
	SetupRec &r = recs[lane];
	RepackHeavy *heavy = reinterpret_cast<RepackHeavy *>(setup_smem + sizeof(SetupRec) * kSetupWalkers);
	const Tables t = device_tables();
	gwat_b200_source s;
	SetupCarry k;
	GWAT_KSTAMP(0);
	// step 0 (sampling vectors): the ~20 libm calls of the repack, a quarter per role (masses / distance and angles / spin 1 / spin 2);
	// every role used to make all of them -- 20-25 k of the kernel's ~83 k cycles (profiles/r02_c_setup_roles_coop.json).  The
	// assembly of the record from these values is cheap and stays with every role.
	// IMRPhenomPv2 only (43.5 -> 39.5 us): the aligned-spin repack has 7 such calls, and splitting them costs more in the extra barrier
	// and the second code stream than it saves (measured: 31.4 -> 33.0 us), so those families keep one copy of the repack for all roles
	const bool split = kP && params && !plan.sky;
	if (split) {
		if (active) repack_heavy_part(role, params + (size_t)w * plan.dimension, plan, heavy[lane]);
		__syncthreads();
	}
	if (active) {
		if (params) {
			if (split) repack_mcmc_assemble(params + (size_t)w * plan.dimension, plan, gmst, T_segment, heavy[lane], s);
			else repack_mcmc_walker(params + (size_t)w * plan.dimension, plan, gmst, T_segment, s);
		} else s = src_in[w];
		setup_step1<Fam>(role, s, net, t, theory, k, r);
		GWAT_KSTAMP(1);
	}
	GWAT_KSTAMP(2);
	if (role <= ROLE_AMP) {
		// fRD, fdamp go from the amplitude role to the phase role: a barrier of these two warps alone (the twist role is the longest
		// of step 1 and nobody needs it before step 3)
		asm volatile("bar.sync 1, 64;" ::: "memory");
		GWAT_KSTAMP(3);
		if (active) setup_step2<Fam>(role, t, k, r);
	}
	GWAT_KSTAMP(4);
	__syncthreads();
	GWAT_KSTAMP(5);
	if (active) setup_step3<Fam>(role, net, r);
	GWAT_KSTAMP(6);
	if (kP) {
		__syncthreads();
		GWAT_KSTAMP(7);
		if (active) setup_step4<Fam>(role, s.shift_time != 0, r);
		__syncthreads();
	}
	GWAT_KSTAMP(8);
	if (active && role == ROLE_PHASE) {
		if (r.refused) r.w.d.A0 = NAN;
		r.w.valid = coef_is_finite(r.w, net.D, kP) ? 1 : 0;
	}
	__syncthreads();
	GWAT_KSTAMP(9);
	if (out) {
		const int n = min(kSetupWalkers, W - blockIdx.x * kSetupWalkers);
		double *o = reinterpret_cast<double *>(out + (size_t)blockIdx.x * kSetupWalkers);
		for (int j = threadIdx.x; j < n * kCoefWords; j += blockDim.x) {
			const int l = j / kCoefWords;
			o[j] = reinterpret_cast<const double *>(&recs[l].w)[j - l * kCoefWords];
		}
	}
	GWAT_KSTAMP(10);
}

// Stage one walker's coefficient block in shared memory.
__device__ __forceinline__ void load_walker(const WalkerCoef *__restrict__ src, WalkerCoef &dst)
{
	static_assert(sizeof(WalkerCoef) % sizeof(double) == 0, "WalkerCoef must be a whole number of doubles");
	const double *s = reinterpret_cast<const double *>(src);
	double *d = reinterpret_cast<double *>(&dst);
	for (int i = threadIdx.x; i < (int)(sizeof(WalkerCoef) / sizeof(double)); i += blockDim.x) d[i] = s[i];
	__syncthreads();
}

// Resident CTAs per SM the bin kernel is compiled for.  64 registers (4 CTAs, 32 warps/SM) spill a little but hide the
// FP64 latency better for the PhenomD and PhenomPv2 families (measured: cfg1 -5 %, cfg2 -2 %, cfg4 -6 %); the NRTidal bin is
// the heaviest and runs best with 80 registers (3 CTAs).
template <class Fam>
constexpr int like_min_ctas()
{
#ifdef GWAT_LOGLIKE_MIN_CTAS
	return GWAT_LOGLIKE_MIN_CTAS;
#else
	return Fam::nrt ? 3 : 4;
#endif
}

// grid (walkers, chunks): a CTA evaluates `units_per_cta` consecutive units of one walker (see gwat_like.h for the cut) and
// writes one partial sum per (unit, warp):  partial[walker][unit][warp][2] = {sum, active bins}.
// When one CTA holds all units of its walker (fin.logL != NULL: every BASELINE grid up to 16384 bins) it also finishes the walker:
// the same fixed-order sum as k_finish -- lane l adds partials l, l + 32, ... in order, then the shuffle tree -- over the partials
// in shared memory, so logL has the same bits whichever kernel forms it, and the third launch of the pass disappears.
struct LikeFinish {
	double *logL;  // NULL: partials go to `partial` and k_finish adds them
	unsigned long long *active_total;
	double prefactor;
	int snr_mode;
};
template <class Fam, int D>
__global__ void __launch_bounds__(kLikeThreads, like_min_ctas<Fam>()) k_loglike(const WalkerCoef *__restrict__ coefs, GridPtrs g, int unit_bins,
                                                                              int units_per_cta, int units_total, double *__restrict__ partial,
                                                                              LikeFinish fin)
{
	static_assert(kLikeThreads == kUnitThreads && kSeedSlots * (D + 1) <= kLikeThreads, "seed table layout");
	__shared__ WalkerCoef w;
	__shared__ CtaSeeds<D> seeds;
	__shared__ double fin_part[2 * kUnitWarps * kMaxUnitsPerCta];
	const int unit0 = blockIdx.y * units_per_cta;
	const int n_units = min(units_per_cta, units_total - unit0);
	const int begin = unit0 * unit_bins;
	double *pout = partial + 2 * kUnitWarps * ((size_t)blockIdx.x * units_total + unit0);
	if (g.uniform) {
		// ascending grid: a chunk that starts above the walker's cutoff contributes exactly zero.  Decided from two words of
		// the coefficient record before anything is staged, so such CTAs cost one L2 round trip.
		const WalkerCoef &wg = coefs[blockIdx.x];
		if (wg.valid && g.f[begin] > walker_fmax<Fam>(wg)) {
			if (fin.logL) {
				if (threadIdx.x == 0) fin.logL[blockIdx.x] = fin.snr_mode ? sqrt(fin.prefactor * 0.0) : -0.5 * (fin.prefactor * 0.0);
			} else if (threadIdx.x < 2 * kUnitWarps * n_units) pout[threadIdx.x] = 0.0;
			return;
		}
	}
	load_walker(coefs + blockIdx.x, w);
	const bool seeded = g.uniform != 0 && w.valid;  // the same for the whole CTA
	if (seeded) {
		if (threadIdx.x < kSeedSlots * (D + 1)) cta_seed_slot<D>(w, g, begin, unit_bins, n_units, kLikeThreads, threadIdx.x, seeds);
		__syncthreads();
	}
	const double fmax = walker_fmax<Fam>(w);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	bool live = w.valid != 0;
	for (int u = 0; u < n_units; u++) {
		double acc = 0.0;
		int nact = 0;
		const int ub = begin + u * unit_bins;
		if (live) live = loglike_unit<Fam, D>(w, g, ub + threadIdx.x, min(g.L, ub + unit_bins), kLikeThreads, fmax, acc, nact,
		                                      &seeds, u, threadIdx.x);
		acc = warp_sum(acc);
		nact = __reduce_add_sync(0xffffffffu, nact);
		if (lane == 0) {
			double *p = (fin.logL ? fin_part : pout) + 2 * (u * kUnitWarps + wid);
			p[0] = w.valid ? acc : NAN;
			p[1] = (double)nact;
		}
	}
	if (fin.logL) {
		__syncthreads();
		if (wid == 0) {
			const int entries = n_units * kUnitWarps;
			double s = 0, n = 0;
			for (int e = lane; e < entries; e += 32) {
				s += fin_part[2 * e];
				n += fin_part[2 * e + 1];
			}
			s = warp_sum(s);
			n = warp_sum(n);
			if (lane == 0) {
				fin.logL[blockIdx.x] = fin.snr_mode ? sqrt(fin.prefactor * s) : -0.5 * (fin.prefactor * s);
				if (fin.active_total) atomicAdd(fin.active_total, (unsigned long long)n);
			}
		}
	}
}

// logL[w] = -1/2 * prefactor * sum of the walker's partials   (Log_Likelihood_internal: -0.5*(HH - 2*DH), src/mcmc_gw.cpp:866)
// One warp per walker: lane l adds entries l, l + 32, ... in order, then the shuffle tree -- a fixed order for a given grid.
// snr_mode: the partials are sums of |r|^2 / S alone (data replaced by zeros) and the output is sqrt(prefactor * sum).
__global__ void k_finish(const double *__restrict__ partial, int W, int entries, double prefactor, int snr_mode, double *__restrict__ logL,
                         unsigned long long *__restrict__ active_total)
{
	const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= W) return;
	const double *p = partial + 2 * (size_t)w * entries;
	double s = 0, n = 0;
	for (int e = lane; e < entries; e += 32) {
		s += p[2 * e];
		n += p[2 * e + 1];
	}
	s = warp_sum(s);
	n = warp_sum(n);
	if (lane == 0) {
		logL[w] = snr_mode ? sqrt(prefactor * s) : -0.5 * (prefactor * s);
		if (active_total) atomicAdd(active_total, (unsigned long long)n);
	}
}

template <class Fam>
__global__ void __launch_bounds__(kThreads) k_waveform(const WalkerCoef *__restrict__ coefs, GridPtrs g, double *hp_re,
                                                      double *hp_im, double *hc_re, double *hc_im)
{
	__shared__ WalkerCoef w;
	load_walker(coefs + blockIdx.x, w);
	const int i = blockIdx.y * kThreads + threadIdx.x;
	if (i >= g.L) return;
	cplx hp{NAN, NAN}, hc{NAN, NAN};
	if (w.valid) polarizations_bin<Fam>(w, g.f[i], g.sf_hi[i], g.sf_lo[i], g.logf[i], hp, hc);
	const size_t k = (size_t)blockIdx.x * g.L + i;
	if (hp_re) hp_re[k] = hp.re;
	if (hp_im) hp_im[k] = hp.im;
	if (hc_re) hc_re[k] = hc.re;
	if (hc_im) hc_im[k] = hc.im;
}

// Amplitude and phase of the (2,2) carrier of the IMRPhenomD families, as the reference's fourier_amplitude / fourier_phase
// return them (IMRPhenomD::construct_amplitude / construct_phase, src/IMRPhenomD.cpp:604-740): amplitude A0 M^(7/6) x shape,
// zero above 0.2/M; phase phi(f) - tc (f - f_ref) - phic at EVERY frequency (construct_phase applies no cutoff; it returns
// the negative, and fourier_phase negates that again: src/waveform_generator.cpp:704-707).
template <class Fam>
__global__ void __launch_bounds__(kThreads) k_amp_phase(const WalkerCoef *__restrict__ coefs, GridPtrs g, double *amp_out, double *phase_out)
{
	__shared__ WalkerCoef w;
	load_walker(coefs + blockIdx.x, w);
	const int i = blockIdx.y * kThreads + threadIdx.x;
	if (i >= g.L) return;
	double amp = NAN, phase = NAN;
	if (w.valid) {
		const double f = g.f[i];
		MfPowers p;
		phenomd_bin<Fam>(w.d, f, bin_sixth_root(w.d, g.sf_hi[i], g.sf_lo[i]), g.logf[i], amp, phase, p);
		if (f > w.d.fcut) amp = 0.0;
		phase = phenomd_apply_time_phase(w.d, f, phase);
	}
	const size_t k = (size_t)blockIdx.x * g.L + i;
	if (amp_out) amp_out[k] = amp;
	if (phase_out) phase_out[k] = phase;
}

// responses of detectors [d0, d0+nd) of the network; out shape [W][nd][L]
template <class Fam>
__global__ void __launch_bounds__(kThreads) k_response(const WalkerCoef *__restrict__ coefs, GridPtrs g, int d0, int nd,
                                                      int with_shift, double *re, double *im)
{
	__shared__ WalkerCoef w;
	load_walker(coefs + blockIdx.x, w);
	const int i = blockIdx.y * kThreads + threadIdx.x;
	if (i >= g.L) return;
	cplx hp{NAN, NAN}, hc{NAN, NAN};
	const double f = g.f[i];
	if (w.valid) polarizations_bin<Fam>(w, f, g.sf_hi[i], g.sf_lo[i], g.logf[i], hp, hc);
	for (int d = 0; d < nd; d++) {
		const cplx r = project_bin(w.det[d0 + d], hp, hc, f, with_shift != 0);
		const size_t k = ((size_t)blockIdx.x * nd + d) * g.L + i;
		re[k] = r.re;
		im[k] = r.im;
	}
}

// ---- Fisher stencil ------------------------------------------------------------------------------------------------------
// calculate_derivatives, non-sky-averaged branch (src/fisher.cpp:340-557): for every source and parameter i, the detector
// response at theta_i +- eps (and +- 2 eps for order 4), eps = 1e-8 absolute.
struct FisherPlan {
	RepackPlan rp;
	int npts;       // 2 or 4 stencil points per parameter
	int theory;
	int nd;               // detectors handled by one pass (1..5)
	int det_is_ref[kFisherMaxDetectors];    // detector == reference detector: no arrival-time handling at all
	double det_row[kFisherMaxDetectors][13], ref_row[13];
};

// One thread per (source, parameter, stencil point): the perturbed source's coefficient block with the antenna patterns of
// all detectors of the pass, and per detector the two quantities through which the detector enters the response besides
// F+/Fx -- the arrival-time phase (+-eps points) or the re-timed carrier (+-2 eps points):
//   tshift[d]   -2 pi DTOA(reference, d) for k < 2, else 0                    (:403-408, 425-430)
//   tcoef[d]    coefficient of (f - f_ref) with t_c - DTOA for k >= 2         (:436-438, 451-453)  [reference quirk, kept]
// Everything else of a response is the same for all detectors, so the carrier is evaluated once per stencil point
// (k_fisher_deriv) instead of once per detector as the reference does.
template <class Fam>
__global__ void __launch_bounds__(128) k_fisher_setup(const gwat_b200_source *__restrict__ src, int S, FisherPlan fp,
                                                     WalkerCoef *__restrict__ coefs, double *__restrict__ tcoef,
                                                     double *__restrict__ scale, int *__restrict__ eta_bc)
{
	const int dim = fp.rp.dimension;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= S * dim * fp.npts) return;
	const int k = t % fp.npts, i = (t / fp.npts) % dim, sidx = t / (fp.npts * dim);
	const double epsilon = 1e-8;
	const gwat_b200_source orig = src[sidx];
	double v[GWAT_B200_MAX_DIM];
	int logfac[GWAT_B200_MAX_DIM];
	unpack_fisher(orig, fp.rp, v, logfac);
	const bool bc = (i == 8 && v[8] > .25 - epsilon);  // eta at its upper boundary: one-sided difference (:369-383)
	if (k == 0) {
		scale[(size_t)sidx * dim + i] = logfac[i] ? v[i] : 1.0;
		// bit 0: one-sided difference; bit 1: the stencil points of RA, DEC, psi do NOT share their intrinsic part -- IMRPhenomPv2 with
		// equatorial_orientation, where the sky position moves theta_JN through the derived inclination
		// ... and the intrinsic set (fp.rp.sky), whose first parameters are ln Mc, eta, a1
		eta_bc[(size_t)sidx * dim + i] = (bc ? 1 : 0) | (((Fam::base == BASE_P && orig.equatorial_orientation) || fp.rp.sky) ? 2 : 0);
	}
	const double step = (k == 0) ? epsilon : (k == 1) ? -epsilon : (k == 2) ? 2 * epsilon : -2 * epsilon;
	const double base = v[i];
	if (step > 0 && bc) v[i] = base;
	else v[i] = base + step;
	gwat_b200_source sp;
	repack_fisher_point(v, orig, fp.rp, sp);
	// fourier_detector_response_equatorial derives incl_angle and psi from (theta_l, phi_l) at every stencil point (src/waveform_util.cpp:947-949)
	if (sp.equatorial_orientation) transform_orientation_coords(sp, Fam::base == BASE_P);
	Network net;
	net.D = fp.nd;
	net.horizon_mode = 0;
	for (int d = 0; d < fp.nd; d++)
		for (int j = 0; j < 13; j++) net.row[d][j] = fp.det_row[d][j];
	WalkerCoef wc;
	walker_setup<Fam>(sp, net, device_tables(), fp.theory, wc);
	for (int d = 0; d < fp.nd; d++) {
		double tshift = 0, tc_seconds = sp.tc;
		if (!fp.det_is_ref[d]) {
			const double dtoa = dtoa_between(fp.ref_row + 9, fp.det_row[d] + 9, sp.RA, sp.DEC, sp.gmst);
			if (k < 2) tshift = (-2 * GWAT_PI) * dtoa;
			else tc_seconds = sp.tc - dtoa;
		}
		wc.det[d].tshift = tshift;
		tcoef[(size_t)t * fp.nd + d] =
		    Fam::base == BASE_P ? phenomp_time_coefficient(tc_seconds) : phenomd_time_coefficient(tc_seconds, wc.d.tc_shift);
	}
	wc.valid = coef_is_finite(wc, fp.nd, Fam::base == BASE_P) ? 1 : 0;
	coefs[t] = wc;
}

// One bin of the stencil of one (source, parameter): the derivative of every detector's response with respect to the parameter,
// times the log-parameter factor (:462-484, 548-554), handed to store(d, value) detector by detector.
//   w[k], tc_s[k * tcs_stride + d]   the coefficient blocks and per-detector time coefficients of the npts stencil points
//   shared_parts                     the intrinsic part of the source is bit-identical at all points (RA, DEC, psi)
// Returns false when every point is exactly zero at this bin (above all cutoffs); zeros are stored then.
template <class Fam, class Store>
__device__ __forceinline__ bool fisher_deriv_bin(const WalkerCoef *w, const double *tc_s, int tcs_stride, int npts, int nd, bool shared_parts,
                                                 bool bc, double sc, double f, double hi, double lo, double lg, const Store &store)
{
	const double epsilon = 1e-8;
	PolParts pp[4];
	// The "above the cutoff" and "invalid record" flags of the points live in two bit masks: as members of pp[] the compiler kept them
	// in LOCAL memory (byte loads and stores, one 32-byte sector each, 44 % of them missing L1: 5.5 GB of L2 traffic per 5000 sources
	// and 5.5 % of the kernel's stall samples on the tests below; profiles/r02_j_*)
	unsigned zero_mask = 0, invalid_mask = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) {  // (unrolled with a guard so that pp[] and r[] live in registers)
		if (k >= npts) continue;
		if (!w[k].valid) invalid_mask |= 1u << k;
		if (k > 0 && shared_parts) {
			pp[k] = pp[0];
			zero_mask |= (zero_mask & 1u) << k;
		} else {
			polarization_parts<Fam>(w[k], f, hi, lo, lg, pp[k]);
			if (pp[k].zero) zero_mask |= 1u << k;
		}
		pp[k].zero = false;  // (read through the mask from here on)
	}
	const unsigned all_points = (1u << npts) - 1u;
	const bool live = ((invalid_mask | ~zero_mask) & all_points) != 0;
	if (!live) {
		for (int d = 0; d < nd; d++) store(d, cplx{0.0, 0.0});
		return false;
	}
	for (int d = 0; d < nd; d++) {
		cplx r[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (k >= npts) continue;
			if ((invalid_mask >> k) & 1u) {
				r[k] = cplx{NAN, NAN};
				continue;
			}
			if ((zero_mask >> k) & 1u) {  // zero polarisations project to zero
				r[k] = cplx{0.0, 0.0};
				continue;
			}
			// (keeping each point's finished polarisations and re-finishing only when the time coefficient changes -- 8
			// instead of 12 finishes for three detectors -- measured 11 % slower: 40 more live registers)
			cplx hp, hc;
			polarizations_finish<Fam>(w[k], pp[k], tc_s[k * tcs_stride + d], f, hp, hc);
			r[k] = project_bin(w[k].det[d], hp, hc, f, true);
		}
		cplx dv;
		if (npts == 2) {
			const double den = bc ? epsilon : 2. * epsilon;
			dv = cplx{(r[0].re - r[1].re) / den, (r[0].im - r[1].im) / den};
		} else {
			const double den = bc ? 6. * epsilon : 12. * epsilon;
			dv = cplx{(((-r[2].re + 8. * r[0].re) - 8. * r[1].re) + r[3].re) / den,
			          (((-r[2].im + 8. * r[0].im) - 8. * r[1].im) + r[3].im) / den};
		}
		store(d, cplx{dv.re * sc, dv.im * sc});
	}
	return true;
}

// deriv[d][s][i][bin] = stencil combination of the responses, times the log-parameter factor (:462-484, 548-554).
// grid (bin tiles, sources x parameters): the time-independent parts of the 2 or 4 stencil points are evaluated once per bin
// -- once for ALL points when the parameter is RA, DEC or psi, which only move the antenna patterns and arrival times --
// and finished per detector.
template <class Fam>
__global__ void __launch_bounds__(kThreads, 3) k_fisher_deriv(const WalkerCoef *__restrict__ coefs, const double *__restrict__ tcoef,
                                                          GridPtrs g, int npts, int nd, int dim, const double *__restrict__ scale,
                                                          const int *__restrict__ eta_bc, double *__restrict__ dre, double *__restrict__ dim_,
                                                          int *__restrict__ bin_limit)
{
	__shared__ WalkerCoef w[4];
	__shared__ double tc_s[4][kFisherMaxDetectors];
	const int ntiles = (g.L + kThreads - 1) / kThreads;
	const int tile0 = blockIdx.x * kFisherTilesPerCta, tile1 = min(ntiles, tile0 + kFisherTilesPerCta);
	auto zero_tile = [&](int t) {
		const int bin = t * kThreads + threadIdx.x;
		if (bin < g.L)
			for (int d = 0; d < nd; d++) {
				const size_t o = ((size_t)d * gridDim.y + blockIdx.y) * g.L + bin;
				dre[o] = 0.0;
				dim_[o] = 0.0;
			}
	};
	// ascending grid: a tile that starts above the cutoff of every stencil point is exactly zero -- decided from a few words
	// of the coefficient records before anything is staged (about half of the tiles of a BBH population end here)
	double fmax_all = INFINITY;
	if (g.uniform) {
		const WalkerCoef *wg = coefs + (size_t)blockIdx.y * npts;
		bool all_valid = true;
		double fm = 0.0;
		for (int k = 0; k < npts; k++) {
			all_valid = all_valid && wg[k].valid;
			fm = fmax(fm, walker_fmax<Fam>(wg[k]));
		}
		if (all_valid) fmax_all = fm;
		if (g.f[tile0 * kThreads] > fmax_all) {
			for (int t = tile0; t < tile1; t++) zero_tile(t);
			return;
		}
	}
	{
		const double *sp = reinterpret_cast<const double *>(coefs + (size_t)blockIdx.y * npts);
		double *dp = reinterpret_cast<double *>(&w[0]);
		for (int i = threadIdx.x; i < (int)(npts * sizeof(WalkerCoef) / sizeof(double)); i += blockDim.x) dp[i] = sp[i];
		if (threadIdx.x < npts * nd) tc_s[threadIdx.x / nd][threadIdx.x % nd] = tcoef[(size_t)blockIdx.y * npts * nd + threadIdx.x];
		__syncthreads();
	}
	// parameters 0..2 are RA, DEC (or sin DEC) and psi in every parameterisation (src/fisher.cpp:46-66): the intrinsic part of
	// the source, and with it everything in PolParts, is bit-identical at all stencil points
	const bool shared_parts = (int)(blockIdx.y % dim) < 3 && (eta_bc[blockIdx.y] & 2) == 0;
	const bool bc = (eta_bc[blockIdx.y] & 1) != 0;
	const double sc = scale[blockIdx.y];
	// several tiles per CTA: the 4 coefficient blocks (5.7 KB) are staged once for 1024 bins instead of once per 256
	for (int t = tile0; t < tile1; t++) {
		if (g.f[t * kThreads] > fmax_all) {  // (the same for the whole CTA)
			zero_tile(t);
			continue;
		}
		const int bin_raw = t * kThreads + threadIdx.x;
		const bool in_grid = bin_raw < g.L;
		const int bin = in_grid ? bin_raw : g.L - 1;  // (tail threads shadow the last bin and store nothing: everyone reaches the barrier below)
		const double f = g.f[bin], hi = g.sf_hi[bin], lo = g.sf_lo[bin], lg = g.logf[bin];
		auto store = [&](int d, const cplx &dv) {
			if (!in_grid) return;
			const size_t o = ((size_t)d * gridDim.y + blockIdx.y) * g.L + bin;
			dre[o] = dv.re;
			dim_[o] = dv.im;
		};
		// Bins above every stencil point's cutoff have exactly zero responses: nothing to evaluate, and k_fisher_assemble need
		// not read them -- the highest live bin of each source is recorded (one atomic per tile).  Invalid points (NaN) keep every bin.
		const bool live = fisher_deriv_bin<Fam>(w, &tc_s[0][0], kFisherMaxDetectors, npts, nd, shared_parts, bc, sc, f, hi, lo, lg, store);
		const int any_live = __syncthreads_or(live && in_grid);
		if (threadIdx.x == 0 && any_live) atomicMax(&bin_limit[blockIdx.y / dim], min(g.L, (t + 1) * kThreads));
	}
}

// ---- fused stencil + assembly -------------------------------------------------------------------------------------------------
// One CTA per source, one WARP per parameter: the dim x npts coefficient blocks of the source are staged in shared memory
// once (dim 11, order 4: 62 KB), then the CTA walks the live part of the grid in tiles of 32 bins.  Per tile every warp
// evaluates its parameter's derivative of all detector responses (lane = bin; exactly the arithmetic of k_fisher_deriv) and
// parks sqrt(w_d) x (Re, Im) in a [detector, Re/Im, bin][parameter] tile; after a barrier the whole CTA adds the tile's
// contribution to the dim (dim + 1) / 2 inner products F_jk = sum w_d Re(d_j conj d_k) (calculate_fisher_elements,
// src/fisher.cpp:2704-2781), each thread keeping ONE running sum in a register over all tiles.  The derivative rows never go
// to HBM (the unfused path writes and re-reads 2 D dim L 8 B = 2.2 MB per source), nothing is atomic, and a source's matrix
// does not depend on the batch it is evaluated in: the order of every sum is a function of (dim, D, L) alone.
constexpr int kFusedTileBins = 32;

struct FusedLayout {
	int K;           // nd * 2 * kFusedTileBins values per parameter and tile
	int Kp;          // row stride of the derivative tile Z[parameter][k]: K + 1, odd -- lanes store consecutive k (no bank conflict), and the
	                 // threads of a warp that read the same k of different parameters hit different banks (the first layout, Z[k][parameter]
	                 // with 12 doubles per k, stored with an 8-way conflict: 113 M conflicts per 5000 sources, short-scoreboard 2.1 per issue)
	int npairs, slices;
	size_t off_tc, off_z, off_red, bytes;
};
__host__ __device__ inline FusedLayout fused_layout(int dim, int npts, int nd)
{
	FusedLayout l;
	l.K = nd * 2 * kFusedTileBins;
	l.Kp = l.K + 1;
	l.npairs = dim * (dim + 1) / 2;
	l.slices = (dim * 32) / l.npairs;  // >= 2 for every dim <= 32
	size_t o = (size_t)dim * npts * sizeof(WalkerCoef);
	l.off_tc = o;
	o += sizeof(double) * (size_t)dim * npts * nd;
	l.off_z = o;
	o += sizeof(double) * 2 * (size_t)l.Kp * dim;  // two tiles: the products of tile t overlap the derivatives of tile t + 1
	l.off_red = o;
	o += sizeof(double) * (size_t)l.slices * l.npairs;
	l.bytes = o;
	return l;
}

template <class Fam, int MAXW, int MINB>
__global__ void __launch_bounds__(MAXW * 32, MINB) k_fisher_fused(const WalkerCoef *__restrict__ coefs, const double *__restrict__ tcoef, GridPtrs g,
                                                                  const double *__restrict__ wq, int npts, int nd, int dim,
                                                                  const double *__restrict__ scale, const int *__restrict__ eta_bc,
                                                                  double prefactor, double *__restrict__ out)
{
	extern __shared__ __align__(16) unsigned char fused_smem[];
	const FusedLayout lay = fused_layout(dim, npts, nd);
	WalkerCoef *w_all = reinterpret_cast<WalkerCoef *>(fused_smem);
	double *tc_all = reinterpret_cast<double *>(fused_smem + lay.off_tc);
	double *Z = reinterpret_cast<double *>(fused_smem + lay.off_z);
	double *red = reinterpret_cast<double *>(fused_smem + lay.off_red);
	const int s = blockIdx.x, lane = threadIdx.x & 31, param = threadIdx.x >> 5;
	{
		const double *sp = reinterpret_cast<const double *>(coefs + (size_t)s * dim * npts);
		double *dp = reinterpret_cast<double *>(w_all);
		for (int i = threadIdx.x; i < (int)(dim * npts * sizeof(WalkerCoef) / sizeof(double)); i += blockDim.x) dp[i] = sp[i];
		for (int i = threadIdx.x; i < dim * npts * nd; i += blockDim.x) tc_all[i] = tcoef[(size_t)s * dim * npts * nd + i];
		__syncthreads();
	}
	// the live part of an ascending grid ends at the highest cutoff of any stencil point (all points valid), else it is all of it
	double fmax_all = INFINITY;
	if (g.uniform) {
		bool all_valid = true;
		double fm = 0.0;
		for (int k = 0; k < dim * npts; k++) {
			all_valid = all_valid && w_all[k].valid;
			fm = fmax(fm, walker_fmax<Fam>(w_all[k]));
		}
		if (all_valid) fmax_all = fm;
	}
	const WalkerCoef *w = w_all + param * npts;
	const double *tc_s = tc_all + param * npts * nd;
	const bool shared_parts = param < 3 && (eta_bc[(size_t)s * dim + param] & 2) == 0;  // RA, DEC (or sin DEC), psi: see k_fisher_deriv
	const bool bc = (eta_bc[(size_t)s * dim + param] & 1) != 0;
	const double sc = scale[(size_t)s * dim + param];
	// inner-product ownership: thread -> (pair, slice); the threads of a warp hold consecutive pairs of one slice, so the
	// reads of a k-row are broadcasts out of one or two 128-byte lines.  (Measured and rejected in round 2, tools/gpu_runs/gpurun_r2_29/30/33:
	// four accumulators per thread +1 %; 2 x 2 register blocks of pairs -- half the shared loads -- -1 %; the products of tile t - 1 formed
	// by the warps of RA, DEC, psi while the others differentiate tile t: +-0.  The loop shows up with 17 % of the stall samples, but the
	// other CTA of the SM fills those slots.)
	const int gp = threadIdx.x % lay.npairs, gs = threadIdx.x / lay.npairs;
	int pj = 0, pk = gp;
	while (pk > pj) {
		pk -= pj + 1;
		pj++;
	}
	double acc = 0.0;
	const int ntiles = (g.L + kFusedTileBins - 1) / kFusedTileBins;
	const size_t tile_words = (size_t)lay.Kp * dim;
	for (int t = 0; t < ntiles; t++) {
		const int bin0 = t * kFusedTileBins;
		if (g.f[bin0] > fmax_all) break;  // (the same for the whole CTA)
		const int bin_raw = bin0 + lane;
		const bool in_grid = bin_raw < g.L;
		const int bin = in_grid ? bin_raw : g.L - 1;
		const double f = g.f[bin], hi = g.sf_hi[bin], lo = g.sf_lo[bin], lg = g.logf[bin];
		double *Zt = Z + (t & 1) * tile_words;
		double *zrow = Zt + (size_t)param * lay.Kp;
		auto store = [&](int d, const cplx &dv) {
			// sqrt(w) on both factors of the product: w >= 0 (quadrature coefficient / PSD); bins of the last tile beyond L weigh 0
			const double rw = in_grid ? wq[(size_t)d * g.ld + bin] : 0.0;  // (wq: the table of square roots)
			zrow[(d * 2 + 0) * kFusedTileBins + lane] = in_grid ? rw * dv.re : 0.0;
			zrow[(d * 2 + 1) * kFusedTileBins + lane] = in_grid ? rw * dv.im : 0.0;
		};
		fisher_deriv_bin<Fam>(w, tc_s, nd, npts, nd, shared_parts, bc, sc, f, hi, lo, lg, store);
		// one barrier per tile: the next tile's derivatives go to the other buffer, and nobody writes THIS buffer again before every
		// thread has passed the next barrier, i.e. has finished the products below
		__syncthreads();
		if (gs < lay.slices) {
			const double *zj = Zt + (size_t)pj * lay.Kp, *zk = Zt + (size_t)pk * lay.Kp;
			for (int k = gs; k < lay.K; k += lay.slices) acc = fma(zj[k], zk[k], acc);
		}
	}
	if (gs < lay.slices) red[gs * lay.npairs + gp] = acc;
	__syncthreads();
	if (threadIdx.x < lay.npairs) {
		double total = 0.0;
		for (int q = 0; q < lay.slices; q++) total += red[q * lay.npairs + threadIdx.x];
		total *= prefactor;
		double *o = out + (size_t)s * dim * dim;
		o[pj * dim + pk] = total;
		o[pk * dim + pj] = total;
	}
}

// ---- sky-averaged Fisher (calculate_derivatives, src/fisher.cpp:183-338) ---------------------------------------------------
// IMRPhenomD in the 7-parameter set ln A0, phic, tc, ln Mc, ln eta, chi_s, chi_a: derivatives of amplitude and phase instead
// of the detector response.  One coefficient block per (source, parameter, stencil point) plus one for the unperturbed
// source (the reference multiplies the phase derivative by the UNPERTURBED amplitude, :187-190, 303-320).
template <class Fam>
__global__ void __launch_bounds__(128) k_fisher_setup_sky(const gwat_b200_source *__restrict__ src, int S, FisherPlan fp,
                                                         WalkerCoef *__restrict__ coefs, double *__restrict__ scale)
{
	const int dim = fp.rp.dimension, nblk = fp.npts + 1;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= S * dim * nblk) return;
	const int k = t % nblk, i = (t / nblk) % dim, sidx = t / (nblk * dim);
	const double epsilon = 1e-8;
	const gwat_b200_source orig = src[sidx];
	gwat_b200_source sp = orig;
	if (k < fp.npts) {
		double v[GWAT_B200_MAX_DIM];
		int logfac[GWAT_B200_MAX_DIM];
		unpack_fisher(orig, fp.rp, v, logfac);
		// The reference keeps the "eta at its boundary" rule of the 11-parameter set -- parameter 8 above 1/4 - eps gets a one-sided
		// difference -- in this branch too (src/fisher.cpp:218-232, 303-316), where parameter 8 is the second modification: the
		// +eps / +2 eps points stay at the value and the quotient loses its factor 2 (folded into `scale`, exact).
		const bool one_sided = i == 8 && v[i] > .25 - epsilon;
		if (k == 0) scale[(size_t)sidx * dim + i] = (logfac[i] ? v[i] : 1.0) * (one_sided ? 2.0 : 1.0);
		const double step = (k == 0) ? epsilon : (k == 1) ? -epsilon : (k == 2) ? 2 * epsilon : -2 * epsilon;
		if (!(one_sided && step > 0)) v[i] += step;
		repack_fisher_point(v, orig, fp.rp, sp);
	}
	Network net;
	net.D = 1;
	net.horizon_mode = 0;
	for (int j = 0; j < 13; j++) net.row[0][j] = fp.det_row[0][j];
	WalkerCoef wc;
	walker_setup<Fam>(sp, net, device_tables(), fp.theory, wc, true);
	wc.valid = coef_is_finite(wc, 1, false) ? 1 : 0;
	coefs[t] = wc;
}

template <class Fam>
__global__ void __launch_bounds__(kThreads) k_fisher_deriv_sky(const WalkerCoef *__restrict__ coefs, GridPtrs g, int npts, int dim,
                                                              const double *__restrict__ scale, double *__restrict__ dre,
                                                              double *__restrict__ dim_, int *__restrict__ bin_limit)
{
	__shared__ WalkerCoef w[5];
	{
		const double *sp = reinterpret_cast<const double *>(coefs + (size_t)blockIdx.y * (npts + 1));
		double *dp = reinterpret_cast<double *>(&w[0]);
		for (int i = threadIdx.x; i < (int)((npts + 1) * sizeof(WalkerCoef) / sizeof(double)); i += blockDim.x) dp[i] = sp[i];
		__syncthreads();
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&bin_limit[blockIdx.y / dim], g.L);  // the phase has no cutoff: every bin counts
	const int bin = blockIdx.x * kThreads + threadIdx.x;
	if (bin >= g.L) return;
	const double f = g.f[bin], hi = g.sf_hi[bin], lo = g.sf_lo[bin], lg = g.logf[bin];
	cplx dv{NAN, NAN};
	bool ok = w[npts].valid != 0;
	for (int k = 0; k < npts; k++) ok = ok && w[k].valid;
	if (ok) {
		double a[4], ph[4], a0, p0;
		amplitude_phase_bin<Fam>(w[npts], f, hi, lo, lg, a0, p0);
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (k < npts) amplitude_phase_bin<Fam>(w[k], f, hi, lo, lg, a[k], ph[k]);
		dv = sky_derivative_bin(npts, a, ph, a0);
	}
	const double sc = scale[blockIdx.y];
	const size_t o = (size_t)blockIdx.y * g.L + bin;
	dre[o] = dv.re * sc;
	dim_[o] = dv.im * sc;
}

// F_jk = sum_d prefactor * sum_bins coef_d * Re(d_j conj(d_k)) / S_d   (calculate_fisher_elements, src/fisher.cpp:2704-2781, and the
// detector loop of the callers: one partial Fisher per detector, added in detector order)
__global__ void __launch_bounds__(kThreads) k_fisher_assemble(const double *__restrict__ dre, const double *__restrict__ dim_,
                                                             const double *__restrict__ wq, int ld, int L, int dim, int ns, int nd,
                                                             double prefactor, const int *__restrict__ bin_limit, double *__restrict__ out)
{
	const int Llive = bin_limit[blockIdx.y];  // derivatives are exactly zero from here on (k_fisher_deriv)
	// blockIdx.x enumerates the pairs j >= k, blockIdx.y the source
	int j = 0, k = blockIdx.x;
	while (k > j) {
		k -= j + 1;
		j++;
	}
	double total = 0;
	for (int d = 0; d < nd; d++) {
		const size_t sbase = ((size_t)d * ns + blockIdx.y) * dim;
		const double *ajr = dre + (sbase + j) * L, *aji = dim_ + (sbase + j) * L;
		const double *akr = dre + (sbase + k) * L, *aki = dim_ + (sbase + k) * L;
		const double *w = wq + (size_t)d * ld;
		double acc = 0, unused = 0;
		for (int i = threadIdx.x; i < Llive; i += kThreads) acc += w[i] * (ajr[i] * akr[i] + aji[i] * aki[i]);
		block_sum2(acc, unused);
		if (threadIdx.x == 0) {
			const double val = prefactor * acc;
			total = (d == 0) ? val : total + val;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		double *o = out + (size_t)blockIdx.y * dim * dim;
		o[j * dim + k] = total;
		o[k * dim + j] = total;
	}
}

// Log_Likelihood_internal for a response the caller supplies (src/mcmc_gw.cpp:801-868): sum_i w_i (|r_i|^2 - 2 Re(d_i conj r_i)) with
// w_i = quadrature coefficient / S_i.  One CTA, fixed order: thread t adds bins t, t + 1024, ...; then the block tree.
__global__ void __launch_bounds__(1024) k_inner_product(int L, const double *__restrict__ w, const double *__restrict__ dre,
                                                        const double *__restrict__ dim, const double *__restrict__ rre,
                                                        const double *__restrict__ rim, double *__restrict__ out)
{
	double hh = 0, dh = 0;
	for (int i = threadIdx.x; i < L; i += 1024) {
		hh += w[i] * (rre[i] * rre[i] + rim[i] * rim[i]);
		dh += w[i] * (dre[i] * rre[i] + dim[i] * rim[i]);
	}
	block_sum2<1024>(hh, dh);
	if (threadIdx.x == 0) {
		out[0] = hh;
		out[1] = dh;
	}
}

__global__ void k_antenna(int W, const double *RA, const double *DEC, const double *psi, double gmst, Network net,
                          double *Fp, double *Fc, double *dt)
{
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	DetCoef dc[GWAT_B200_MAX_DETECTORS];
	detector_setup(net, RA[w], DEC[w], psi[w], gmst, dc);
	for (int d = 0; d < net.D; d++) {
		Fp[(size_t)w * net.D + d] = dc[d].Fplus;
		Fc[(size_t)w * net.D + d] = dc[d].Fcross;
		dt[(size_t)w * net.D + d] = dtoa_between(net.row[0] + 9, net.row[d] + 9, RA[w], DEC[w], gmst);
	}
}

// FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so the DFMA issue peak is measured here with 8
// independent dependent-FMA chains per thread (enough ILP to cover the 4-cycle-class latency at 8 warps per scheduler).
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double seed)
{
	double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; i++) {
		a0 = fma(a0, m, c);
		a1 = fma(a1, m, c);
		a2 = fma(a2, m, c);
		a3 = fma(a3, m, c);
		a4 = fma(a4, m, c);
		a5 = fma(a5, m, c);
		a6 = fma(a6, m, c);
		a7 = fma(a7, m, c);
	}
	const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
	if (r == 12345.678) out[0] = r;  // never true; keeps the chains alive
}

}  // namespace


// ---------------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------------

#include "gwat_engine_internal.h"

namespace {

std::string g_create_error;

int fail(gwat_b200_ctx *c, int code, const std::string &msg)
{
	if (c) c->err = msg;
	else g_create_error = msg;
	return code;
}
#define CUDA_TRY(ctx, expr)                                                                              \
	do {                                                                                                   \
		cudaError_t e_ = (expr);                                                                             \
		if (e_ != cudaSuccess)                                                                               \
			return fail(ctx, GWAT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
	} while (0)

template <class T>
int grow(gwat_b200_ctx *c, T *&ptr, size_t &cap, size_t need)
{
	if (need <= cap) return 0;
	if (ptr) CUDA_TRY(c, cudaFree(ptr));
	ptr = nullptr;
	cap = 0;
	const size_t n = need + need / 4;
	CUDA_TRY(c, cudaMalloc((void **)&ptr, n * sizeof(T)));
	cap = n;
	return 0;
}

GridPtrs grid_ptrs(const gwat_b200_ctx *c, const double *zero_data = nullptr)
{
	GridPtrs g;
	const size_t L = c->ld, DL = (size_t)c->D * c->ld;
	g.f = c->d_grid;
	g.sf_hi = c->d_grid + L;
	g.sf_lo = c->d_grid + 2 * L;
	g.logf = c->d_grid + 3 * L;
	g.ld = c->ld;
	g.wq = c->d_net;
	g.dre = zero_data ? zero_data : c->d_net + DL;  // zero_data: D*ld zeros standing in for the strain (SNR: <h|h> alone)
	g.dim = zero_data ? zero_data : c->d_net + 2 * DL;
	g.L = c->L;
	g.uniform = c->uniform ? 1 : 0;
	g.df = c->df;
	return g;
}

// intrinsic: the sets of the reference's tc/phic-maximised runs (PTMCMC_method_specific_prep, src/mcmc_gw.cpp:1880-1985): ln Mc, eta and the
// spins (4 for the IMRPhenomD family, + 1 or 2 tidal parameters for NRT; 8 for IMRPhenomPv2), then the modifications
int make_plan(const MethodDesc &desc, const gwat_b200_mod *mod, int dimension, RepackPlan &plan, bool intrinsic = false)
{
	std::memset(&plan, 0, sizeof(plan));
	plan.dimension = dimension;
	plan.pv2 = desc.pv2;
	plan.nrt = desc.nrt;
	plan.ppe = desc.ppe || desc.theory != THEORY_NONE;
	plan.gimr = desc.gimr && !plan.ppe;
	plan.alpha_unit_fix = theory_alpha_units(desc.theory);
	plan.mcmc = 1;
	if (mod) plan.mod = *mod;
	else {
		std::memset(&plan.mod, 0, sizeof(plan.mod));
		plan.mod.tidal_love = 1;
	}
	if (dimension < 1 || dimension > GWAT_B200_MAX_DIM) return -1;
	if (plan.mod.ppE_Nmod < 0 || plan.mod.ppE_Nmod > GWAT_B200_MAX_MOD) return -1;
	int base = intrinsic ? (desc.pv2 ? 8 : 4) : (desc.pv2 ? 15 : 11);
	plan.sky = intrinsic ? 1 : 0;
	if (desc.nrt && !desc.pv2) base += plan.mod.tidal_love ? 1 : 2;
	int mods = 0;
	if (plan.ppe) mods = plan.mod.ppE_Nmod;
	else if (plan.gimr)
		mods = plan.mod.gIMR_Nmod_phi + plan.mod.gIMR_Nmod_sigma + plan.mod.gIMR_Nmod_beta + plan.mod.gIMR_Nmod_alpha;
	if (dimension != base + mods) return -1;
	return 0;
}

// Run the statement given as trailing arguments with `Fam` bound to the kernel family of `desc`.
#define GWAT_DISPATCH_FAMILY(desc, ...)                                                                                  \
	do {                                                                                                                   \
		switch ((desc).family_id) {                                                                                          \
		case FAM_D: { typedef Family<BASE_D, PPE_NONE, false, false> Fam; __VA_ARGS__; break; }                              \
		case FAM_D_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, false> Fam; __VA_ARGS__; break; }                  \
		GWAT_FULL_ONLY(case FAM_D_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, false> Fam; __VA_ARGS__; break; }) \
		GWAT_FULL_ONLY(case FAM_D_GIMR: { typedef Family<BASE_D, PPE_NONE, true, false> Fam; __VA_ARGS__; break; }) \
		case FAM_D_NRT: { typedef Family<BASE_D, PPE_NONE, false, true> Fam; __VA_ARGS__; break; }                           \
		GWAT_FULL_ONLY(case FAM_D_NRT_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, true> Fam; __VA_ARGS__; break; }) \
		GWAT_FULL_ONLY(case FAM_D_NRT_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, true> Fam; __VA_ARGS__; break; }) \
		case FAM_P: { typedef Family<BASE_P, PPE_NONE, false, false> Fam; __VA_ARGS__; break; }                              \
		GWAT_FULL_ONLY(case FAM_P_PPE_INS: { typedef Family<BASE_P, PPE_INSPIRAL, false, false> Fam; __VA_ARGS__; break; }) \
		GWAT_FULL_ONLY(case FAM_P_PPE_IMR: { typedef Family<BASE_P, PPE_IMR, false, false> Fam; __VA_ARGS__; break; }) \
		GWAT_FULL_ONLY(case FAM_P_GIMR: { typedef Family<BASE_P, PPE_NONE, true, false> Fam; __VA_ARGS__; break; }) \
		default: return fail(ctx, GWAT_B200_ERR_METHOD, std::string("generation_method not implemented: ") + (desc).base);   \
		}                                                                                                                    \
	} while (0)

struct LikeCut {
	int unit_bins, units_total, units_per_cta, chunks;
};

template <class Fam>
int launch_loglike(gwat_b200_ctx *ctx, int W, const LikeCut &cut, cudaStream_t st, const double *zero_data, const LikeFinish &fin)
{
	const GridPtrs g = grid_ptrs(ctx, zero_data);
	// walkers vary fastest: CTAs in flight together work on the same stretch of the grid tables, so a tile is fetched from
	// HBM once per pass even when the tables (cfg5: 109 MB) do not fit in L2
	const dim3 grid(W, cut.chunks);
#define GWAT_LAUNCH_LIKE(DD)                                                                                                      \
	case DD:                                                                                                                      \
		k_loglike<Fam, DD><<<grid, kLikeThreads, 0, st>>>(ctx->d_coef, g, cut.unit_bins, cut.units_per_cta, cut.units_total, ctx->d_partial, fin); \
		return 0;
	switch (ctx->D) {
		GWAT_FULL_ONLY(GWAT_LAUNCH_LIKE(1))
		GWAT_LAUNCH_LIKE(2)
		GWAT_LAUNCH_LIKE(3)
		GWAT_FULL_ONLY(GWAT_LAUNCH_LIKE(4))
		GWAT_FULL_ONLY(GWAT_LAUNCH_LIKE(5))
	default: return -1;
	}
#undef GWAT_LAUNCH_LIKE
}

// The cut of the bin axis into units is a property of the grid (gwat_like.h: unit_bins_for); only the number of consecutive
// units a CTA evaluates depends on W: runs of up to 64 bins per thread for big ensembles (the per-CTA start-up -- coefficient
// load, seed table, first table reads -- is worth about one bin per thread), shorter runs when few walkers have to fill
// 148 SMs.  Results do not depend on that choice.
LikeCut choose_cut(int W, int L)
{
	static const int upc_env = getenv("GWAT_B200_UNITS_PER_CTA") ? atoi(getenv("GWAT_B200_UNITS_PER_CTA")) : 0;  // experiments only
	LikeCut c;
	c.unit_bins = unit_bins_for(L);
	c.units_total = (L + c.unit_bins - 1) / c.unit_bins;
	int upc = (64 * kUnitThreads) / c.unit_bins;  // 2 for grids up to 2^20 bins, 1 for longer ones
	auto chunks_for = [&](int n) { return (c.units_total + n - 1) / n; };
	while (upc > 1 && (long long)W * chunks_for(upc) < 148LL * 12) upc /= 2;
	if (upc_env > 0 && upc_env <= kMaxUnitsPerCta) upc = upc_env;
	c.units_per_cta = upc;
	c.chunks = chunks_for(upc);
	return c;
}

// The shared tail of every likelihood entry point: coefficients are in ctx->d_coef.
// snr_mode: the strain is replaced by zeros and the output is sqrt(<h|h>) instead of -1/2 (<h|h> - 2 <d|h>).
int run_loglike(gwat_b200_ctx *ctx, const MethodDesc &desc, int W, double *d_logL, cudaStream_t st, cudaStream_t st_heavy = nullptr,
                cudaEvent_t ev_a = nullptr, cudaEvent_t ev_b = nullptr, bool snr_mode = false)
{
	const double *zero_data = nullptr;
	if (snr_mode) {
		const size_t n = (size_t)ctx->D * ctx->ld;
		if (ctx->cap_zero < n) {
			if (ctx->d_zero) cudaFree(ctx->d_zero);
			ctx->d_zero = nullptr;
			ctx->cap_zero = 0;
			CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_zero, sizeof(double) * n));
			CUDA_TRY(ctx, cudaMemset(ctx->d_zero, 0, sizeof(double) * n));
			ctx->cap_zero = n;
		}
		zero_data = ctx->d_zero;
	}
	const LikeCut cut = choose_cut(W, ctx->L);
	const int entries = cut.units_total * kUnitWarps;  // partial sums per walker
	if (grow(ctx, ctx->d_partial, ctx->cap_partial, (size_t)2 * W * entries)) return GWAT_B200_ERR_CUDA;
	cudaStream_t sl = st;
	if (st_heavy && st_heavy != st && ev_a && ev_b) {
		CUDA_TRY(ctx, cudaEventRecord(ev_a, st));
		CUDA_TRY(ctx, cudaStreamWaitEvent(st_heavy, ev_a, 0));
		sl = st_heavy;
	}
	// (the pass's active-bin counter was zeroed by k_setup, which every path runs first on `st`)
	const bool fused = cut.chunks == 1;  // one CTA per walker: it finishes the walker itself
	const LikeFinish fin{fused ? d_logL : nullptr, ctx->d_active, ctx->pref_like, snr_mode ? 1 : 0};
	// CUDA events around the bin kernel cost ~3 us each on the stream (measured: cfg1 0.158 -> 0.148 ms per call without them): only on
	// request (gwat_b200_set_kernel_timing; bench.py's resident loop asks for them, the roofline needs the kernel's own duration)
	if (ctx->kernel_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, sl));
	GWAT_DISPATCH_FAMILY(desc, if (launch_loglike<Fam>(ctx, W, cut, sl, zero_data, fin)) return fail(
	                               ctx, GWAT_B200_ERR_STATE, "unsupported detector count"));
	if (ctx->kernel_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, sl));
	if (sl != st) {
		CUDA_TRY(ctx, cudaEventRecord(ev_b, sl));
		CUDA_TRY(ctx, cudaStreamWaitEvent(st, ev_b, 0));
	}
	if (!fused) k_finish<<<(W + 3) / 4, 128, 0, st>>>(ctx->d_partial, W, entries, ctx->pref_like, snr_mode ? 1 : 0, d_logL, ctx->d_active);
	ctx->launches += fused ? 1 : 2;
	CUDA_TRY(ctx, cudaGetLastError());
	return 0;
}

int collect_stats(gwat_b200_ctx *ctx, cudaStream_t st)
{
	// the active-bin count lands in a pinned word of the context: an asynchronous copy behind the caller's own result copy, one
	// synchronisation for both (a pageable destination would make this copy a second, blocking round trip)
	if (!ctx->h_active) CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->h_active, sizeof(unsigned long long), cudaHostAllocDefault));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_active, ctx->d_active, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	const unsigned long long act = *ctx->h_active;
	float ms = 0;
	if (ctx->kernel_timing) CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->last_ms = ms;
	ctx->last_active = (long long)act;
	return 0;
}

int check_ready(gwat_b200_ctx *ctx, bool need_data)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (ctx->L <= 0) return fail(ctx, GWAT_B200_ERR_STATE, "gwat_b200_set_network has not been called");
	if (need_data && !ctx->have_data) return fail(ctx, GWAT_B200_ERR_STATE, "the network was set without data");
	return 0;
}

// k_setup's records take more than the 48 KB a kernel gets without asking: opt in once per instantiation and device.
template <class Fam>
void setup_smem_opt_in()
{
	static std::atomic<bool> done[64];
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (done[dev].load(std::memory_order_acquire)) return;
	cudaFuncSetAttribute(k_setup<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSetupSmemBytes);
	done[dev].store(true, std::memory_order_release);
}
template <class Fam>
void launch_setup(int W, cudaStream_t st, const double *d_params, const gwat_b200_source *d_src, const RepackPlan &plan, const Network &net, int theory,
                  double gmst, double T_segment, WalkerCoef *d_coef, unsigned long long *d_active)
{
	setup_smem_opt_in<Fam>();
	k_setup<Fam><<<(W + kSetupWalkers - 1) / kSetupWalkers, kSetupWalkers * setup_roles<Fam>(), kSetupSmemBytes, st>>>(d_params, d_src, W, plan, net, theory, gmst,
	                                                                                                            T_segment, d_coef, d_active);
}
template <class Fam>
void launch_setup_src(gwat_b200_ctx *ctx, int W, const gwat_b200_source *d_src, int theory, cudaStream_t st, const Network &net)
{
	launch_setup<Fam>(W, st, nullptr, d_src, RepackPlan{}, net, theory, 0.0, 0.0, ctx->d_coef, ctx->d_active);
}

// single_detector: the semantics of the reference's one-detector entry points (fourier_detector_response, calculate_snr), which
// honour equatorial_orientation (incl_angle and psi derived from theta_l, phi_l: gwat_orient.h) and horizon_coord; the coherent
// network response and the likelihoods read incl_angle / psi / RA / DEC as given, as create_coherent_GW_detection does.
int setup_from_sources(gwat_b200_ctx *ctx, const MethodDesc &desc, int W, const gwat_b200_source *h_src, cudaStream_t st,
                       bool single_detector = false)
{
	if (grow(ctx, ctx->d_src, ctx->cap_src, (size_t)W)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_coef, ctx->cap_walkers, (size_t)W)) return GWAT_B200_ERR_CUDA;
	Network net = ctx->net;
	std::vector<gwat_b200_source> oriented;
	if (single_detector) {
		net.horizon_mode = 1;
		bool any = false;
		for (int w = 0; w < W && !any; w++) any = h_src[w].equatorial_orientation != 0;
		if (any) {
			oriented.assign(h_src, h_src + W);
			for (gwat_b200_source &s : oriented)
				if (s.equatorial_orientation) transform_orientation_coords(s, desc.pv2 != 0);
			h_src = oriented.data();
		}
	}
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_src, h_src, sizeof(gwat_b200_source) * W, cudaMemcpyHostToDevice, st));
	if (!oriented.empty()) CUDA_TRY(ctx, cudaStreamSynchronize(st));  // (the staging vector is local)
	GWAT_DISPATCH_FAMILY(desc, launch_setup_src<Fam>(ctx, W, ctx->d_src, desc.theory, st, net));
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	return 0;
}

// ---- Fisher passes shared by the host-source entry point and the sampler's device-parameter entry point ------------------
// The fused kernel (one CTA per source, one warp per parameter) is used whenever its shared-memory footprint fits; otherwise
// (dimension > 16) the stencil and the assembly run as two kernels with the derivative rows in HBM.
bool fisher_fused_fits(int dim, int npts, int nd)
{
	return dim <= 16 && fused_layout(dim, npts, nd).bytes <= (size_t)227 * 1024;
}

int fisher_chunk_size(const gwat_b200_ctx *ctx, int S, int dim, int npts, int nd, bool fused)
{
	if (fused) {
		// sources per pass: bounded by 1 GiB of coefficient blocks (cfg3: 17 k sources)
		const size_t per_source = (size_t)dim * npts * sizeof(WalkerCoef);
		return (int)std::max<size_t>(1, std::min<size_t>((size_t)S, ((size_t)1024 << 20) / per_source));
	}
	// sources per pass: bounded by a 2 GiB derivative buffer
	const size_t per_source = (size_t)dim * ctx->L * 16 * nd;
	const size_t by_memory = ((size_t)2048 << 20) / per_source, by_grid = 65535 / (size_t)dim;  // gridDim.y = sources * dim
	return (int)std::max<size_t>(1, std::min<size_t>((size_t)S, std::min(by_memory, by_grid)));
}

int fisher_reserve(gwat_b200_ctx *ctx, int chunk, int dim, int npts, int nd, bool fused)
{
	if (grow(ctx, ctx->d_src, ctx->cap_src, (size_t)chunk)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_coef, ctx->cap_walkers, (size_t)chunk * dim * npts)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_tcoef, ctx->cap_tcoef, (size_t)chunk * dim * npts * nd)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_binlim, ctx->cap_binlim, (size_t)chunk)) return GWAT_B200_ERR_CUDA;
	if (!fused && grow(ctx, ctx->d_deriv, ctx->cap_deriv, (size_t)2 * chunk * dim * ctx->L * nd)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_scale, ctx->cap_scale, (size_t)chunk * dim)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_bc, ctx->cap_bc, (size_t)chunk * dim)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_fisher, ctx->cap_fisher, (size_t)chunk * dim * dim)) return GWAT_B200_ERR_CUDA;
	return 0;
}

template <class Fam>
int launch_fisher_fused(gwat_b200_ctx *ctx, const FisherPlan &fp, int ns, const GridPtrs &g, const double *wq, double *d_out, cudaStream_t st)
{
	const int dim = fp.rp.dimension;
	const FusedLayout lay = fused_layout(dim, fp.npts, fp.nd);
#define GWAT_FUSED_LAUNCH(MAXW, MINB)                                                                                                    \
	do {                                                                                                                                   \
		auto kern = k_fisher_fused<Fam, MAXW, MINB>;                                                                                         \
		CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));                              \
		kern<<<ns, dim * 32, lay.bytes, st>>>(ctx->d_coef, ctx->d_tcoef, g, wq, fp.npts, fp.nd, dim, ctx->d_scale, ctx->d_bc, ctx->pref_fisher, d_out); \
	} while (0)
	static const bool wide = getenv("GWAT_B200_FISHER_WIDE") != nullptr;  // experiments only: 128 registers, one CTA per SM
	if (dim <= 12 && !wide) GWAT_FUSED_LAUNCH(12, 2);
	else GWAT_FUSED_LAUNCH(16, 1);
#undef GWAT_FUSED_LAUNCH
	return 0;
}

// d_src[0..ns) -> d_out[ns][dim][dim], detectors d0..d1-1 summed; one launch of each kernel for all detectors
int fisher_chunk(gwat_b200_ctx *ctx, const MethodDesc &desc, FisherPlan &fp, int ns, int chunk, int d0, int d1,
                 int reference_index, cudaStream_t st, const gwat_b200_source *d_src = nullptr, double *d_out = nullptr)
{
	const int L = ctx->L, dim = fp.rp.dimension, nd = d1 - d0;
	if (!d_src) d_src = ctx->d_src;
	if (!d_out) d_out = ctx->d_fisher;
	// FisherPlan::det_row and the kernels' staging arrays hold kFisherMaxDetectors detectors per pass (as the likelihood
	// kernels are instantiated for 1..5): larger networks are refused, never truncated
	if (nd < 1 || nd > kFisherMaxDetectors)
		return fail(ctx, GWAT_B200_ERR_UNSUPPORTED, "fisher: at most 5 detectors per pass (pass detector_index >= 0 for larger networks)");
	const GridPtrs g = grid_ptrs(ctx);
	const double *wq_fisher_all = ctx->d_net + 3 * (size_t)ctx->D * ctx->ld;
	const int npairs = dim * (dim + 1) / 2;
	fp.nd = nd;
	std::memcpy(fp.ref_row, ctx->net.row[reference_index], sizeof(fp.ref_row));
	for (int d = 0; d < nd; d++) {
		std::memcpy(fp.det_row[d], ctx->net.row[d0 + d], sizeof(fp.det_row[d]));
		fp.det_is_ref[d] = (std::memcmp(fp.det_row[d], fp.ref_row, sizeof(fp.ref_row)) == 0) ? 1 : 0;
	}
	const int nthreads = ns * dim * fp.npts;
	GWAT_DISPATCH_FAMILY(desc, k_fisher_setup<Fam><<<(nthreads + 127) / 128, 128, 0, st>>>(d_src, ns, fp, ctx->d_coef, ctx->d_tcoef,
	                                                                                        ctx->d_scale, ctx->d_bc));
	static const bool no_fused = getenv("GWAT_B200_FISHER_UNFUSED") != nullptr;  // experiments / A-B tests only
	if (!no_fused && fisher_fused_fits(dim, fp.npts, nd)) {
		const double *rwq_fisher_all = ctx->d_net + 4 * (size_t)ctx->D * ctx->ld;  // sqrt(wq_fisher), see gwat_b200_set_network
		GWAT_DISPATCH_FAMILY(desc, if (int rc = launch_fisher_fused<Fam>(ctx, fp, ns, g, rwq_fisher_all + (size_t)d0 * ctx->ld, d_out, st)) return rc);
		ctx->launches += 2;
		CUDA_TRY(ctx, cudaGetLastError());
		return 0;
	}
	if (grow(ctx, ctx->d_deriv, ctx->cap_deriv, (size_t)2 * chunk * dim * L * nd)) return GWAT_B200_ERR_CUDA;
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_binlim, 0, sizeof(int) * ns, st));
	const int ntiles = (L + kThreads - 1) / kThreads;
	const dim3 gd((ntiles + kFisherTilesPerCta - 1) / kFisherTilesPerCta, ns * dim);
	double *dre = ctx->d_deriv, *dim_ = ctx->d_deriv + (size_t)chunk * dim * L * nd;
	GWAT_DISPATCH_FAMILY(desc, k_fisher_deriv<Fam><<<gd, kThreads, 0, st>>>(ctx->d_coef, ctx->d_tcoef, g, fp.npts, nd, dim, ctx->d_scale,
	                                                                         ctx->d_bc, dre, dim_, ctx->d_binlim));
	k_fisher_assemble<<<dim3(npairs, ns), kThreads, 0, st>>>(dre, dim_, wq_fisher_all + (size_t)d0 * ctx->ld, ctx->ld, L, dim, ns, nd,
	                                                          ctx->pref_fisher, ctx->d_binlim, d_out);
	ctx->launches += 3;
	CUDA_TRY(ctx, cudaGetLastError());
	return 0;
}

// Sky-averaged pass: ctx->d_src[0..ns) -> ctx->d_fisher[ns][7][7] for the PSD of detector `det`.
template <class Fam>
int fisher_chunk_sky(gwat_b200_ctx *ctx, FisherPlan &fp, int ns, int chunk, int det, cudaStream_t st)
{
	const int L = ctx->L, dim = fp.rp.dimension;
	const GridPtrs g = grid_ptrs(ctx);
	const double *wq_fisher_all = ctx->d_net + 3 * (size_t)ctx->D * ctx->ld;
	fp.nd = 1;
	std::memcpy(fp.det_row[0], ctx->net.row[det], sizeof(fp.det_row[0]));
	const int nthreads = ns * dim * (fp.npts + 1);
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_binlim, 0, sizeof(int) * ns, st));
	k_fisher_setup_sky<Fam><<<(nthreads + 127) / 128, 128, 0, st>>>(ctx->d_src, ns, fp, ctx->d_coef, ctx->d_scale);
	double *dre = ctx->d_deriv, *dim_ = ctx->d_deriv + (size_t)chunk * dim * L;
	k_fisher_deriv_sky<Fam><<<dim3((L + kThreads - 1) / kThreads, ns * dim), kThreads, 0, st>>>(ctx->d_coef, g, fp.npts, dim, ctx->d_scale,
	                                                                                           dre, dim_, ctx->d_binlim);
	k_fisher_assemble<<<dim3(dim * (dim + 1) / 2, ns), kThreads, 0, st>>>(dre, dim_, wq_fisher_all + (size_t)det * ctx->ld, ctx->ld, L, dim, ns, 1,
	                                                                    ctx->pref_fisher, ctx->d_binlim, ctx->d_fisher);
	ctx->launches += 3;
	CUDA_TRY(ctx, cudaGetLastError());
	return 0;
}

// Swap one of the extra scratch sets into the context for the duration of a call (the caller holds ctx->mu).
struct LaneSwap {
	gwat_b200_ctx *c;
	LikeLane *l;
	LaneSwap(gwat_b200_ctx *ctx, int lane) : c(ctx), l(lane > 0 ? &ctx->extra[lane - 1] : nullptr) { swap(); }
	~LaneSwap() { swap(); }
	void swap()
	{
		if (!l) return;
		std::swap(c->d_coef, l->d_coef);
		std::swap(c->cap_walkers, l->cap_walkers);
		std::swap(c->d_partial, l->d_partial);
		std::swap(c->cap_partial, l->cap_partial);
		std::swap(c->d_active, l->d_active);
		std::swap(c->ev0, l->ev0);
		std::swap(c->ev1, l->ev1);
	}
};

// sampling vector -> physical record, tc left as sampled (repack_parameters without the likelihood's tc flip)
__global__ void k_repack_only(const double *__restrict__ params, int W, RepackPlan plan, double gmst, gwat_b200_source *__restrict__ out)
{
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	gwat_b200_source s;
	repack_mcmc_walker(params + (size_t)w * plan.dimension, plan, gmst, 0.0, s);
	if (!plan.sky) s.tc = -s.tc;
	out[w] = s;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------

extern "C" {

int gwat_b200_abi_version(void) { return GWAT_B200_ABI_VERSION; }

void gwat_b200_source_init(gwat_b200_source *src)
{
	if (src) source_defaults(*src);
}

int gwat_b200_transform_orientation_coords(const char *generation_method, int n, gwat_b200_source *sources)
{
	MethodDesc desc;
	if (!sources || n < 0) return GWAT_B200_ERR_ARG;
	if (parse_method(generation_method, desc) != 0) return GWAT_B200_ERR_METHOD;
	for (int i = 0; i < n; i++) transform_orientation_coords(sources[i], desc.pv2 != 0);
	return GWAT_B200_OK;
}

int gwat_b200_cosmology_index(const char *name)
{
	if (!name) return -1;
	static const char *const names[GWAT_NUM_COSMOLOGIES] = GWAT_COSMOLOGY_NAMES;
	std::string up(name);
	for (char &ch : up) ch = (char)std::toupper((unsigned char)ch);
	for (int i = 0; i < GWAT_NUM_COSMOLOGIES; i++)
		if (up == names[i]) return i;
	return -1;
}

#include "gwat_tables_sites.inc"
int gwat_b200_detector_site(const char *detector, double *latitude, double *longitude, double *location, double *response_tensor)
{
	const int id = detector_index(detector);
	if (id < 0 || !latitude || !longitude || !location || !response_tensor) return GWAT_B200_ERR_ARG;
	*latitude = gwat_site_lat_long[id][0];
	*longitude = gwat_site_lat_long[id][1];
	std::memcpy(response_tensor, hosttab::gwat_detector_table[id], sizeof(double) * 9);
	std::memcpy(location, hosttab::gwat_detector_table[id] + 9, sizeof(double) * 3);
	return GWAT_B200_OK;
}

void gwat_b200_mod_init(gwat_b200_mod *mod)
{
	if (!mod) return;
	std::memset(mod, 0, sizeof(*mod));
	mod->tidal_love = 1;
}

int gwat_b200_ctx_create(gwat_b200_ctx **out, int device)
{
	if (!out) return GWAT_B200_ERR_ARG;
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
		return fail(nullptr, GWAT_B200_ERR_CUDA,
		            std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
	if (device < 0 || device >= n) return fail(nullptr, GWAT_B200_ERR_ARG, "device ordinal out of range");
	gwat_b200_ctx *c = new gwat_b200_ctx;
	c->device = device;
	if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
	    (e = cudaMalloc((void **)&c->d_active, sizeof(unsigned long long))) != cudaSuccess) {
		const std::string msg = std::string("context creation: ") + cudaGetErrorString(e);
		delete c;
		return fail(nullptr, GWAT_B200_ERR_CUDA, msg);
	}
	*out = c;
	return GWAT_B200_OK;
}

void gwat_b200_ctx_destroy(gwat_b200_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	cudaFree(c->d_grid);
	cudaFree(c->d_net);
	cudaFree(c->d_coef);
	cudaFree(c->d_partial);
	cudaFree(c->d_params);
	cudaFree(c->d_out);
	cudaFree(c->d_src);
	cudaFree(c->d_active);
	cudaFree(c->d_zero);
	if (c->h_active) cudaFreeHost(c->h_active);
	cudaFree(c->d_tcoef);
	cudaFree(c->d_binlim);
	cudaFree(c->d_deriv);
	cudaFree(c->d_scale);
	cudaFree(c->d_fisher);
	cudaFree(c->d_bc);
	cudaEventDestroy(c->ev0);
	cudaEventDestroy(c->ev1);
	for (LikeLane &l : c->extra) {
		cudaFree(l.d_coef);
		cudaFree(l.d_partial);
		cudaFree(l.d_active);
		if (l.ev0) cudaEventDestroy(l.ev0);
		if (l.ev1) cudaEventDestroy(l.ev1);
		if (l.ev_a) cudaEventDestroy(l.ev_a);
		if (l.ev_b) cudaEventDestroy(l.ev_b);
	}
	for (auto &fs : c->fstage) {
		if (fs.h_src) cudaFreeHost(fs.h_src);
		if (fs.h_out) cudaFreeHost(fs.h_out);
		cudaFree(fs.d_src);
		cudaFree(fs.d_out);
		if (fs.ev_in) cudaEventDestroy(fs.ev_in);
		if (fs.ev_done) cudaEventDestroy(fs.ev_done);
		if (fs.ev_out) cudaEventDestroy(fs.ev_out);
	}
	if (c->copy_in) cudaStreamDestroy(c->copy_in);
	if (c->copy_out) cudaStreamDestroy(c->copy_out);
	cudaStreamDestroy(c->stream);
	delete c;
}

const char *gwat_b200_last_error(const gwat_b200_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int gwat_b200_set_network(gwat_b200_ctx *ctx, int D, const char *const *detectors, int L, const double *f,
                          const double *psd, const double *data_re, const double *data_im, const double *weights,
                          const char *integration_method, int log10F)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (D < 1 || D > GWAT_B200_MAX_DETECTORS || !detectors || L < 4 || !f || !psd)
		return fail(ctx, GWAT_B200_ERR_ARG, "set_network: bad detector count, length or NULL array");
	if ((data_re == nullptr) != (data_im == nullptr)) return fail(ctx, GWAT_B200_ERR_ARG, "set_network: data_re/data_im");
	const std::string integ = integration_method ? integration_method : "SIMPSONS";
	const bool gl = integ == "GAUSSLEG";
	if (!gl && integ != "SIMPSONS") return fail(ctx, GWAT_B200_ERR_ARG, "set_network: integration_method must be SIMPSONS or GAUSSLEG");
	if (gl && !weights) return fail(ctx, GWAT_B200_ERR_ARG, "set_network: GAUSSLEG needs weights");
	Network net{};
	net.D = D;
	for (int d = 0; d < D; d++) {
		const int id = detector_index(detectors[d]);
		if (id < 0) return fail(ctx, GWAT_B200_ERR_ARG, std::string("set_network: unsupported detector ") + (detectors[d] ? detectors[d] : "(null)"));
	}
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	for (int d = 0; d < D; d++) std::memcpy(net.row[d], hosttab::gwat_detector_table[detector_index(detectors[d])], sizeof(double) * 13);

	std::vector<double> hi, lo, lg;
	build_frequency_tables(f, L, hi, lo, lg);
	// device layout: every table padded to `ld`, a whole number of tiles; the padding is inert (f = +inf is above every
	// cutoff, weights and data are zero)
	const int ld = ((L + kLikeThreads - 1) / kLikeThreads) * kLikeThreads;
	std::vector<double> grid((size_t)4 * ld);
	for (int i = 0; i < ld; i++) {
		const bool in = i < L;
		grid[i] = in ? f[i] : INFINITY;
		grid[(size_t)ld + i] = in ? hi[i] : 1.0;
		grid[2 * (size_t)ld + i] = in ? lo[i] : 0.0;
		grid[3 * (size_t)ld + i] = in ? lg[i] : 0.0;
	}
	const size_t DL = (size_t)D * ld;
	std::vector<double> netbuf(5 * DL, 0.0);
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) {
			const size_t k = (size_t)d * ld + i, kin = (size_t)d * L + i;
			netbuf[k] = quadrature_coefficient(i, L, gl, log10F != 0, weights, f) / psd[kin];
			if (data_re) {
				netbuf[DL + k] = data_re[kin];
				netbuf[2 * DL + k] = data_im[kin];
			}
			// the Fisher routines always integrate with Simpson's rule (src/fisher.cpp:128-131)
			netbuf[3 * DL + k] = quadrature_coefficient(i, L, false, false, nullptr, f) / psd[kin];
			// its square root, for the fused Fisher kernel (one factor on each side of the product; IEEE sqrt: the same bits as on the device)
			netbuf[4 * DL + k] = std::sqrt(netbuf[3 * DL + k]);
		}
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	if (ctx->d_grid) cudaFree(ctx->d_grid);
	if (ctx->d_net) cudaFree(ctx->d_net);
	ctx->d_grid = ctx->d_net = nullptr;
	ctx->L = 0;
	CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_grid, sizeof(double) * grid.size()));
	CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_net, sizeof(double) * netbuf.size()));
	CUDA_TRY(ctx, cudaMemcpy(ctx->d_grid, grid.data(), sizeof(double) * grid.size(), cudaMemcpyHostToDevice));
	CUDA_TRY(ctx, cudaMemcpy(ctx->d_net, netbuf.data(), sizeof(double) * netbuf.size(), cudaMemcpyHostToDevice));
	ctx->D = D;
	ctx->L = L;
	ctx->ld = ld;
	ctx->net = net;
	ctx->have_data = data_re != nullptr;
	ctx->gaussleg = gl;
	ctx->log10F = log10F != 0;
	ctx->pref_like = quadrature_prefactor(L, gl, f, false);
	ctx->pref_fisher = quadrature_prefactor(L, false, f, true);
	ctx->h_f.assign(f, f + L);
	// uniform grid?  (then the per-detector arrival-time phase advances by a constant rotation per bin, gwat_like.h)
	{
		const double df = (f[L - 1] - f[0]) / (L - 1);
		bool uni = df > 0;
		for (int i = 0; i < L && uni; i++) uni = std::fabs(f[i] - (f[0] + i * df)) <= 1e-9 * df;
		ctx->uniform = uni;
		ctx->df = df;
	}
	return GWAT_B200_OK;
}

int gwat_b200_loglike_mcmc_batch_dev(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension,
                                     int W, const double *d_params, double gmst, double T_segment, double *d_logL,
                                     void *stream)
{
	if (int rc = check_ready(ctx, true)) return rc;
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return gwat_internal::loglike_mcmc_lane(ctx, 0, method, mod, dimension, W, d_params, gmst, T_segment, d_logL,
	                                        stream ? (cudaStream_t)stream : ctx->stream);
}

int gwat_b200_loglike_mcmc_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                                 const double *params, double gmst, double T_segment, double *logL)
{
	if (int rc = check_ready(ctx, true)) return rc;
	if (W < 0 || (W > 0 && (!params || !logL))) return fail(ctx, GWAT_B200_ERR_ARG, "loglike_mcmc_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	if (dimension < 1 || dimension > GWAT_B200_MAX_DIM) return fail(ctx, GWAT_B200_ERR_ARG, "loglike_mcmc_batch: dimension out of range");
	// one critical section from upload to download: the staging buffers belong to the context, and callers may share it
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if (grow(ctx, ctx->d_params, ctx->cap_params, (size_t)W * dimension)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_out, ctx->cap_out, (size_t)W)) return GWAT_B200_ERR_CUDA;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_params, params, sizeof(double) * W * dimension, cudaMemcpyHostToDevice, ctx->stream));
	if (int rc = gwat_internal::loglike_mcmc_lane(ctx, 0, method, mod, dimension, W, ctx->d_params, gmst, T_segment, ctx->d_out, ctx->stream))
		return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(logL, ctx->d_out, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
	return collect_stats(ctx, ctx->stream);
}

int gwat_b200_loglike_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources, double *logL)
{
	if (int rc = check_ready(ctx, true)) return rc;
	if (W < 0 || (W > 0 && (!sources || !logL))) return fail(ctx, GWAT_B200_ERR_ARG, "loglike_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if (int rc = setup_from_sources(ctx, desc, W, sources, ctx->stream)) return rc;
	if (grow(ctx, ctx->d_out, ctx->cap_out, (size_t)W)) return GWAT_B200_ERR_CUDA;
	if (int rc = run_loglike(ctx, desc, W, ctx->d_out, ctx->stream)) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(logL, ctx->d_out, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
	return collect_stats(ctx, ctx->stream);
}

int gwat_b200_snr_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources, double *snr)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && (!sources || !snr))) return fail(ctx, GWAT_B200_ERR_ARG, "snr_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	// calculate_snr (src/waveform_util.cpp:290-344) is a one-detector routine: orientation / horizon conventions as fourier_detector_response
	if (int rc = setup_from_sources(ctx, desc, W, sources, ctx->stream, ctx->D == 1)) return rc;
	if (grow(ctx, ctx->d_out, ctx->cap_out, (size_t)W)) return GWAT_B200_ERR_CUDA;
	if (int rc = run_loglike(ctx, desc, W, ctx->d_out, ctx->stream, nullptr, nullptr, nullptr, true)) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(snr, ctx->d_out, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
	return collect_stats(ctx, ctx->stream);
}

int gwat_b200_fourier_waveform_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources,
                                     double *hplus_re, double *hplus_im, double *hcross_re, double *hcross_im)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && !sources)) return fail(ctx, GWAT_B200_ERR_ARG, "fourier_waveform_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (int rc = setup_from_sources(ctx, desc, W, sources, st)) return rc;
	const size_t n = (size_t)W * ctx->L;
	if (grow(ctx, ctx->d_out, ctx->cap_out, 4 * n)) return GWAT_B200_ERR_CUDA;
	double *o = ctx->d_out;
	const dim3 grid(W, (ctx->L + kThreads - 1) / kThreads);  // walkers on x: no 65535 limit on the batch
	const GridPtrs g = grid_ptrs(ctx);
	GWAT_DISPATCH_FAMILY(desc, k_waveform<Fam><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n, o + 2 * n, o + 3 * n));
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	double *host[4] = {hplus_re, hplus_im, hcross_re, hcross_im};
	for (int k = 0; k < 4; k++)
		if (host[k]) CUDA_TRY(ctx, cudaMemcpyAsync(host[k], o + k * n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	return GWAT_B200_OK;
}

int gwat_b200_fourier_amplitude_phase_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources,
                                            double *amplitude, double *phase)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && !sources)) return fail(ctx, GWAT_B200_ERR_ARG, "fourier_amplitude_phase_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	if (desc.pv2 || desc.nrt)
		return fail(ctx, GWAT_B200_ERR_UNSUPPORTED, "fourier_amplitude/phase: IMRPhenomD, ppE_IMRPhenomD_* and gIMRPhenomD only");
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (int rc = setup_from_sources(ctx, desc, W, sources, st)) return rc;
	const size_t n = (size_t)W * ctx->L;
	if (grow(ctx, ctx->d_out, ctx->cap_out, 2 * n)) return GWAT_B200_ERR_CUDA;
	double *o = ctx->d_out;
	const dim3 grid(W, (ctx->L + kThreads - 1) / kThreads);  // walkers on x: no 65535 limit on the batch
	const GridPtrs g = grid_ptrs(ctx);
	switch (desc.family_id) {
	case FAM_D: k_amp_phase<Family<BASE_D, PPE_NONE, false, false>><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n); break;
	case FAM_D_PPE_INS: k_amp_phase<Family<BASE_D, PPE_INSPIRAL, false, false>><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n); break;
	GWAT_FULL_ONLY(case FAM_D_PPE_IMR: k_amp_phase<Family<BASE_D, PPE_IMR, false, false>><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n); break;)
	GWAT_FULL_ONLY(case FAM_D_GIMR: k_amp_phase<Family<BASE_D, PPE_NONE, true, false>><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n); break;)
	default: return fail(ctx, GWAT_B200_ERR_UNSUPPORTED, "fourier_amplitude/phase: family not supported");
	}
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	if (amplitude) CUDA_TRY(ctx, cudaMemcpyAsync(amplitude, o, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	if (phase) CUDA_TRY(ctx, cudaMemcpyAsync(phase, o + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	return GWAT_B200_OK;
}

// with_shift == 0 is the one-detector entry point (fourier_detector_response)
static int response_common(gwat_b200_ctx *ctx, const char *method, int d0, int nd, int with_shift, int W,
                           const gwat_b200_source *sources, double *re, double *im)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (int rc = setup_from_sources(ctx, desc, W, sources, st, with_shift == 0)) return rc;
	const size_t n = (size_t)W * nd * ctx->L;
	if (grow(ctx, ctx->d_out, ctx->cap_out, 2 * n)) return GWAT_B200_ERR_CUDA;
	double *o = ctx->d_out;
	const dim3 grid(W, (ctx->L + kThreads - 1) / kThreads);  // walkers on x: no 65535 limit on the batch
	const GridPtrs g = grid_ptrs(ctx);
	GWAT_DISPATCH_FAMILY(desc, k_response<Fam><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, d0, nd, with_shift, o, o + n));
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	CUDA_TRY(ctx, cudaMemcpyAsync(re, o, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(im, o + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	return GWAT_B200_OK;
}

int gwat_b200_coherent_response_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources,
                                      double *resp_re, double *resp_im)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && (!sources || !resp_re || !resp_im))) return fail(ctx, GWAT_B200_ERR_ARG, "coherent_response_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	return response_common(ctx, method, 0, ctx->D, 1, W, sources, resp_re, resp_im);
}

int gwat_b200_fourier_detector_response_batch(gwat_b200_ctx *ctx, const char *method, const char *detector, int W,
                                              const gwat_b200_source *sources, double *resp_re, double *resp_im)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && (!sources || !resp_re || !resp_im))) return fail(ctx, GWAT_B200_ERR_ARG, "fourier_detector_response_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	const int id = detector_index(detector);
	int d0 = -1;
	if (id >= 0) {
		for (int d = 0; d < ctx->D; d++)
			if (std::memcmp(ctx->net.row[d], hosttab::gwat_detector_table[id], sizeof(double) * 13) == 0) d0 = d;
	}
	if (d0 < 0) return fail(ctx, GWAT_B200_ERR_ARG, "fourier_detector_response_batch: detector is not part of the network");
	return response_common(ctx, method, d0, 1, 0, W, sources, resp_re, resp_im);
}

static int repack_mcmc_any(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W, const double *params, double gmst,
                           gwat_b200_source *sources, bool intrinsic);
int gwat_b200_repack_mcmc_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                                const double *params, double gmst, gwat_b200_source *sources)
{
	return repack_mcmc_any(ctx, method, mod, dimension, W, params, gmst, sources, false);
}
int gwat_b200_repack_mcmc_intrinsic_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                                          const double *params, double gmst, gwat_b200_source *sources)
{
	return repack_mcmc_any(ctx, method, mod, dimension, W, params, gmst, sources, true);
}
static int repack_mcmc_any(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W, const double *params, double gmst,
                           gwat_b200_source *sources, bool intrinsic)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (W < 0 || (W > 0 && (!params || !sources))) return fail(ctx, GWAT_B200_ERR_ARG, "repack_mcmc_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	RepackPlan plan;
	if (make_plan(desc, mod, dimension, plan, intrinsic) != 0)
		return fail(ctx, GWAT_B200_ERR_ARG, "repack_mcmc_batch: dimension does not match the method and modification struct");
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (grow(ctx, ctx->d_params, ctx->cap_params, (size_t)W * dimension)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_src, ctx->cap_src, (size_t)W)) return GWAT_B200_ERR_CUDA;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_params, params, sizeof(double) * W * dimension, cudaMemcpyHostToDevice, st));
	// tc is left as sampled: this entry point mirrors repack_parameters alone
	k_repack_only<<<(W + 127) / 128, 128, 0, st>>>(ctx->d_params, W, plan, gmst, ctx->d_src);
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	CUDA_TRY(ctx, cudaMemcpyAsync(sources, ctx->d_src, sizeof(gwat_b200_source) * W, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	return GWAT_B200_OK;
}

// MCMC_fisher_wrapper in an intrinsic run (src/mcmc_gw.cpp:2229-2330): per detector fisher_numerical("MCMC_" + method) of the record
// MCMC_prep_params + repack_parameters build (sky_average: the amplitude / phase branch, one detector's PSD each), summed, then the
// intrinsic branch of MCMC_fisher_transformations (:2163-2179: the prior terms REPLACE the diagonal entries of eta and the spins)
// and the dCS / EdGB unit factor (:2182-2193).
int gwat_b200_mcmc_fisher_intrinsic_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int order, int W,
                                          const double *params, double gmst, double *fisher)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && (!params || !fisher))) return fail(ctx, GWAT_B200_ERR_ARG, "mcmc_fisher_intrinsic_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::vector<gwat_b200_source> src((size_t)W);
	if (int rc = gwat_b200_repack_mcmc_intrinsic_batch(ctx, method, mod, dimension, W, params, gmst, src.data())) return rc;
	const std::string m = std::string("MCMC_") + method;
	const size_t dd = (size_t)dimension * dimension;
	std::vector<double> part((size_t)W * dd);
	std::fill(fisher, fisher + (size_t)W * dd, 0.0);
	const int D = ctx->D;
	for (int d = 0; d < D; d++) {
		if (int rc = gwat_b200_fisher_numerical_batch(ctx, m.c_str(), d, 0, dimension, order, W, src.data(), part.data())) return rc;
		for (size_t i = 0; i < (size_t)W * dd; i++) fisher[i] += part[i];
	}
	const bool alpha_fix = theory_alpha_units(desc.theory);
	for (int w = 0; w < W; w++) {
		double *F = fisher + (size_t)w * dd;
		F[1 * dimension + 1] = 1. / (.25);
		F[2 * dimension + 2] = 1. / (4);
		F[3 * dimension + 3] = 1. / (4);
		if (desc.pv2) {
			F[4 * dimension + 4] = 1. / (4);
			F[5 * dimension + 5] = 1. / (4);
			F[6 * dimension + 6] = 1. / (4 * GWAT_PI * GWAT_PI);
			F[7 * dimension + 7] = 1. / (4 * GWAT_PI * GWAT_PI);
		}
		if (alpha_fix && src[w].Nmod > 0) {
			const int base = dimension - src[w].Nmod;
			double factor = 4 * std::pow(src[w].betappe[0], 3. / 4.);  // temp_params[base]: alpha^2 in s^4 after the unit change
			factor *= 1000 / GWAT_C_SI;
			for (int i = 0; i < dimension; i++) {
				F[base * dimension + i] *= factor;
				F[i * dimension + base] *= factor;
			}
		}
	}
	return GWAT_B200_OK;
}

int gwat_b200_antenna_batch(gwat_b200_ctx *ctx, int W, const double *RA, const double *DEC, const double *psi, double gmst,
                            double *Fplus, double *Fcross, double *dtoa)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (W < 0 || (W > 0 && (!RA || !DEC || !psi || !Fplus || !Fcross || !dtoa))) return fail(ctx, GWAT_B200_ERR_ARG, "antenna_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const size_t nin = 3 * (size_t)W, nout = 3 * (size_t)W * ctx->D;
	if (grow(ctx, ctx->d_params, ctx->cap_params, nin)) return GWAT_B200_ERR_CUDA;
	if (grow(ctx, ctx->d_out, ctx->cap_out, nout)) return GWAT_B200_ERR_CUDA;
	double *in = ctx->d_params, *o = ctx->d_out;
	const size_t WD = (size_t)W * ctx->D;
	CUDA_TRY(ctx, cudaMemcpyAsync(in, RA, sizeof(double) * W, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(in + W, DEC, sizeof(double) * W, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(in + 2 * (size_t)W, psi, sizeof(double) * W, cudaMemcpyHostToDevice, st));
	k_antenna<<<(W + 127) / 128, 128, 0, st>>>(W, in, in + W, in + 2 * (size_t)W, gmst, ctx->net, o, o + WD, o + 2 * WD);
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	CUDA_TRY(ctx, cudaMemcpyAsync(Fplus, o, sizeof(double) * WD, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(Fcross, o + WD, sizeof(double) * WD, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(dtoa, o + 2 * WD, sizeof(double) * WD, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	return GWAT_B200_OK;
}

int gwat_b200_fisher_numerical_batch(gwat_b200_ctx *ctx, const char *method, int detector_index, int reference_index,
                                     int dimension, int order, int S, const gwat_b200_source *sources, double *fisher)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (S < 0 || (S > 0 && (!sources || !fisher))) return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: NULL array");
	if (S == 0) return GWAT_B200_OK;
	if (order != 2 && order != 4) return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: order must be 2 or 4");
	if (detector_index >= ctx->D || reference_index < 0 || reference_index >= ctx->D)
		return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: detector index out of range");
	MethodDesc desc;
	if (parse_method(method, desc) != 0)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	bool any_sky = false, all_sky = true;
	for (int i = 0; i < S; i++) {
		any_sky = any_sky || sources[i].sky_average != 0;
		all_sky = all_sky && sources[i].sky_average != 0;
	}
	// A sky-averaged IMRPhenomPv2 record is not differentiated through amplitude and phase (the reference's test is for "IMRPhenomD",
	// src/fisher.cpp:183): its "MCMC_" set -- the 8 intrinsic parameters, everything else at the constants of repack_parameters
	// (:2308-2352) -- takes the response branch below.
	const bool pv2_intrinsic = any_sky && all_sky && desc.pv2 && desc.mcmc;
	if (any_sky && !pv2_intrinsic) {
		// sky-averaged branch of calculate_derivatives (src/fisher.cpp:183-338): the IMRPhenomD carrier in the 7-parameter set, followed by
		// the ppE betas (ppE_IMRPhenomD_*, and the theories mapped onto them) or the gIMR deviations; one detector's PSD
		if (!all_sky) return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: sky-averaged and pointed sources in one batch");
		const bool ppe = desc.ppe || desc.theory != THEORY_NONE, gimr = desc.gimr && !ppe;
		int mods = 0;
		if (ppe) mods = sources[0].Nmod;
		else if (gimr) mods = sources[0].Nmod_phi + sources[0].Nmod_sigma + sources[0].Nmod_beta + sources[0].Nmod_alpha;
		if (desc.pv2)
			return fail(ctx, GWAT_B200_ERR_UNSUPPORTED,
			            "fisher_numerical_batch: sky-averaged Fishers of amplitude and phase exist for the IMRPhenomD family only (the reference prints "
			            "\"not supported\" for the physical IMRPhenomPv2 set, src/fisher.cpp:1993; pass \"MCMC_IMRPhenomPv2\" and dimension 8 for the intrinsic set)");
		if (desc.nrt)
			return fail(ctx, GWAT_B200_ERR_UNSUPPORTED,
			            desc.mcmc ? "fisher_numerical_batch: sky-averaged NRT Fishers: the reference's construct_phase reads the NRTidal spin/quadrupole "
			                        "coefficients before anything has set them (they are assigned in construct_waveform only, src/IMRPhenomD_NRT.cpp:596-608); not built"
			                      : "fisher_numerical_batch: the reference's sky-averaged NRT layout puts ln(tidal) on top of ln(eta) (src/fisher.cpp:2061-2078); not built");
		// "MCMC_" + method: the intrinsic set ln Mc, eta, chi1, chi2 (src/fisher.cpp:2000-2013) instead of the seven
		const int nbase = desc.mcmc ? 4 : 7;
		if (mods < 0 || mods > GWAT_B200_MAX_MOD || dimension != nbase + mods)
			return fail(ctx, GWAT_B200_ERR_ARG,
			            desc.mcmc ? "fisher_numerical_batch: a sky-averaged MCMC_ Fisher has 4 parameters plus the sources' modifications"
			                      : "fisher_numerical_batch: a sky-averaged Fisher has 7 parameters plus the sources' modifications");
		if (detector_index < 0) return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: a sky-averaged Fisher needs one detector's PSD");
		FisherPlan fp;
		std::memset(&fp, 0, sizeof(fp));
		fp.rp.dimension = dimension;
		fp.rp.sky = 1;
		fp.rp.mcmc = desc.mcmc;
		fp.rp.ppe = ppe;
		fp.rp.gimr = gimr;
		fp.npts = order == 4 ? 4 : 2;
		fp.theory = desc.theory;
		const int dd = dimension * dimension;
		std::lock_guard<std::mutex> lock(ctx->mu);
		CUDA_TRY(ctx, cudaSetDevice(ctx->device));
		cudaStream_t st = ctx->stream;
		const int chunk = fisher_chunk_size(ctx, S, dimension, fp.npts + 1, 1, false);
		if (int rc = fisher_reserve(ctx, chunk, dimension, fp.npts + 1, 1, false)) return rc;
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
		for (int s0 = 0; s0 < S; s0 += chunk) {
			const int ns = std::min(chunk, S - s0);
			CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_src, sources + s0, sizeof(gwat_b200_source) * ns, cudaMemcpyHostToDevice, st));
			switch (desc.family_id) {
			case FAM_D: if (int rc = fisher_chunk_sky<Family<BASE_D, PPE_NONE, false, false>>(ctx, fp, ns, chunk, detector_index, st)) return rc; break;
			case FAM_D_PPE_INS: if (int rc = fisher_chunk_sky<Family<BASE_D, PPE_INSPIRAL, false, false>>(ctx, fp, ns, chunk, detector_index, st)) return rc; break;
			GWAT_FULL_ONLY(case FAM_D_PPE_IMR: if (int rc = fisher_chunk_sky<Family<BASE_D, PPE_IMR, false, false>>(ctx, fp, ns, chunk, detector_index, st)) return rc; break;)
			GWAT_FULL_ONLY(case FAM_D_GIMR: if (int rc = fisher_chunk_sky<Family<BASE_D, PPE_NONE, true, false>>(ctx, fp, ns, chunk, detector_index, st)) return rc; break;)
			default: return fail(ctx, GWAT_B200_ERR_UNSUPPORTED, "fisher_numerical_batch: sky-averaged Fisher: family not built");
			}
			CUDA_TRY(ctx, cudaMemcpyAsync(fisher + (size_t)s0 * dd, ctx->d_fisher, sizeof(double) * ns * dd, cudaMemcpyDeviceToHost, st));
		}
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		float ms = 0;
		CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->last_ms = ms;
		return GWAT_B200_OK;
	}
	for (int i = 0; i < S; i++)
		if (sources[i].horizon_coord)
			return fail(ctx, GWAT_B200_ERR_UNSUPPORTED,
			            "fisher_numerical_batch: horizon_coord sources (the parameter set holds RA and DEC, which a detector-frame response does not read)");
	// the modification layout is taken from the first source (all sources of a batch share it, as they share the method)
	gwat_b200_mod mod;
	gwat_b200_mod_init(&mod);
	mod.ppE_Nmod = sources[0].Nmod;
	mod.gIMR_Nmod_phi = sources[0].Nmod_phi;
	mod.gIMR_Nmod_sigma = sources[0].Nmod_sigma;
	mod.gIMR_Nmod_beta = sources[0].Nmod_beta;
	mod.gIMR_Nmod_alpha = sources[0].Nmod_alpha;
	mod.tidal_love = sources[0].tidal_love;
	FisherPlan fp;
	std::memset(&fp, 0, sizeof(fp));
	RepackPlan &plan = fp.rp;
	plan.dimension = dimension;
	plan.pv2 = desc.pv2;
	plan.nrt = desc.nrt;
	plan.ppe = desc.ppe || desc.theory != THEORY_NONE;
	plan.gimr = desc.gimr && !plan.ppe;
	plan.mcmc = desc.mcmc;
	plan.sky = pv2_intrinsic ? 1 : 0;
	plan.mod = mod;
	{
		int base = pv2_intrinsic ? 8 : desc.pv2 ? (desc.mcmc ? 15 : 13) : 11;
		if (desc.nrt && !desc.pv2) base += mod.tidal_love ? 1 : 2;
		int mods = 0;
		if (plan.ppe) mods = mod.ppE_Nmod;
		else if (plan.gimr) mods = mod.gIMR_Nmod_phi + mod.gIMR_Nmod_sigma + mod.gIMR_Nmod_beta + mod.gIMR_Nmod_alpha;
		if (dimension != base + mods || dimension > GWAT_B200_MAX_DIM)
			return fail(ctx, GWAT_B200_ERR_ARG, "fisher_numerical_batch: dimension does not match the method and the sources' modifications");
	}
	fp.npts = order == 4 ? 4 : 2;
	fp.theory = desc.theory;
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const int L = ctx->L, dim = dimension;
	const int d0 = detector_index < 0 ? 0 : detector_index;
	const int d1 = detector_index < 0 ? ctx->D : detector_index + 1;
	const bool fused = fisher_fused_fits(dim, fp.npts, d1 - d0) && getenv("GWAT_B200_FISHER_UNFUSED") == nullptr;
	// passes of at most `chunk` sources; at least four passes for big batches so that the copies have kernels to hide behind
	int chunk = fisher_chunk_size(ctx, S, dim, fp.npts, d1 - d0, fused);
	if (S >= 4096) chunk = std::min(chunk, (S + 3) / 4);
	if (int rc = fisher_reserve(ctx, chunk, dim, fp.npts, d1 - d0, fused)) return rc;
	// The caller's arrays are pageable: a cudaMemcpyAsync on them is staged by the driver and the D2H one blocks the host
	// until the pass has finished, so the next pass cannot even be enqueued (measured in round 1: 1 ms of idle GPU per 3.4 ms
	// pass).  Two pinned staging sets and two device source/result sets: pass k+1's sources travel and pass k-1's matrices
	// return while pass k computes.
	if (!ctx->copy_in) {
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
		for (auto &fs : ctx->fstage) {
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&fs.ev_in, cudaEventDisableTiming));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&fs.ev_done, cudaEventDisableTiming));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&fs.ev_out, cudaEventDisableTiming));
		}
	}
	const size_t out_per = (size_t)dim * dim;
	for (auto &fs : ctx->fstage) {
		if (fs.cap_src < (size_t)chunk) {
			if (fs.h_src) cudaFreeHost(fs.h_src);
			if (fs.d_src) cudaFree(fs.d_src);
			fs.h_src = nullptr;
			fs.d_src = nullptr;
			fs.cap_src = 0;
			CUDA_TRY(ctx, cudaHostAlloc((void **)&fs.h_src, sizeof(gwat_b200_source) * chunk, cudaHostAllocDefault));
			CUDA_TRY(ctx, cudaMalloc((void **)&fs.d_src, sizeof(gwat_b200_source) * chunk));
			fs.cap_src = chunk;
		}
		if (fs.cap_out < (size_t)chunk * out_per) {
			if (fs.h_out) cudaFreeHost(fs.h_out);
			if (fs.d_out) cudaFree(fs.d_out);
			fs.h_out = nullptr;
			fs.d_out = nullptr;
			fs.cap_out = 0;
			CUDA_TRY(ctx, cudaHostAlloc((void **)&fs.h_out, sizeof(double) * chunk * out_per, cudaHostAllocDefault));
			CUDA_TRY(ctx, cudaMalloc((void **)&fs.d_out, sizeof(double) * chunk * out_per));
			fs.cap_out = (size_t)chunk * out_per;
		}
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
	const int npass = (S + chunk - 1) / chunk;
	auto drain = [&](int pass) -> int {  // pass's matrices: pinned staging -> the caller's array
		auto &fs = ctx->fstage[pass & 1];
		const int s0 = pass * chunk, ns = std::min(chunk, S - s0);
		CUDA_TRY(ctx, cudaEventSynchronize(fs.ev_out));
		std::memcpy(fisher + (size_t)s0 * out_per, fs.h_out, sizeof(double) * ns * out_per);
		return 0;
	};
	for (int pass = 0; pass < npass; pass++) {
		auto &fs = ctx->fstage[pass & 1];
		const int s0 = pass * chunk, ns = std::min(chunk, S - s0);
		if (pass >= 2)
			if (int rc = drain(pass - 2)) return rc;  // (also: the set's previous H2D and kernels are long done)
		std::memcpy(fs.h_src, sources + s0, sizeof(gwat_b200_source) * ns);
		CUDA_TRY(ctx, cudaMemcpyAsync(fs.d_src, fs.h_src, sizeof(gwat_b200_source) * ns, cudaMemcpyHostToDevice, ctx->copy_in));
		CUDA_TRY(ctx, cudaEventRecord(fs.ev_in, ctx->copy_in));
		CUDA_TRY(ctx, cudaStreamWaitEvent(st, fs.ev_in, 0));
		if (int rc = fisher_chunk(ctx, desc, fp, ns, chunk, d0, d1, reference_index, st, fs.d_src, fs.d_out)) return rc;
		CUDA_TRY(ctx, cudaEventRecord(fs.ev_done, st));
		CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_out, fs.ev_done, 0));
		CUDA_TRY(ctx, cudaMemcpyAsync(fs.h_out, fs.d_out, sizeof(double) * ns * out_per, cudaMemcpyDeviceToHost, ctx->copy_out));
		CUDA_TRY(ctx, cudaEventRecord(fs.ev_out, ctx->copy_out));
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
	for (int pass = std::max(0, npass - 2); pass < npass; pass++)
		if (int rc = drain(pass)) return rc;
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	float ms = 0;
	CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->last_ms = ms;
	(void)L;
	return GWAT_B200_OK;
}

int gwat_b200_method_info(const char *generation_method, int *ppe_like, int *gimr, int *alpha_units, int *pv2, int *nrt)
{
	MethodDesc desc;
	if (parse_method(generation_method, desc) != 0) return GWAT_B200_ERR_METHOD;
	if (ppe_like) *ppe_like = (desc.ppe || desc.theory != THEORY_NONE) ? 1 : 0;
	if (gimr) *gimr = (desc.gimr && !(desc.ppe || desc.theory != THEORY_NONE)) ? 1 : 0;
	if (alpha_units) *alpha_units = theory_alpha_units(desc.theory) ? 1 : 0;
	if (pv2) *pv2 = desc.pv2 ? 1 : 0;
	if (nrt) *nrt = desc.nrt ? 1 : 0;
	return GWAT_B200_OK;
}

int gwat_b200_log_likelihood_internal(gwat_b200_ctx *ctx, int L, const double *frequencies, const double *psd, const double *data_re,
                                      const double *data_im, const double *weights, const char *integration_method, int log10F,
                                      const double *response_re, const double *response_im, double *logL)
{
	if (!ctx || L < 4 || !frequencies || !psd || !data_re || !data_im || !response_re || !response_im || !logL) return GWAT_B200_ERR_ARG;
	const std::string integ = integration_method ? integration_method : "SIMPSONS";
	const bool gl = integ == "GAUSSLEG";
	if (!gl && integ != "SIMPSONS") return fail(ctx, GWAT_B200_ERR_ARG, "log_likelihood_internal: integration_method must be SIMPSONS or GAUSSLEG");
	if (gl && !weights) return fail(ctx, GWAT_B200_ERR_ARG, "log_likelihood_internal: GAUSSLEG needs weights");
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	std::vector<double> w(L);
	for (int i = 0; i < L; i++) w[i] = quadrature_coefficient(i, L, gl, log10F != 0, weights, frequencies) / psd[i];
	if (grow(ctx, ctx->d_out, ctx->cap_out, (size_t)5 * L + 2)) return GWAT_B200_ERR_CUDA;
	double *d = ctx->d_out;
	const double *host[5] = {w.data(), data_re, data_im, response_re, response_im};
	for (int k = 0; k < 5; k++) CUDA_TRY(ctx, cudaMemcpyAsync(d + (size_t)k * L, host[k], sizeof(double) * L, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));  // (w is a local buffer)
	k_inner_product<<<1, 1024, 0, st>>>(L, d, d + L, d + 2 * (size_t)L, d + 3 * (size_t)L, d + 4 * (size_t)L, d + 5 * (size_t)L);
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	double sums[2];
	CUDA_TRY(ctx, cudaMemcpyAsync(sums, d + 5 * (size_t)L, sizeof(sums), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	const double pref = quadrature_prefactor(L, gl, frequencies, false);
	*logL = -0.5 * (pref * sums[0] - 2.0 * (pref * sums[1]));
	return GWAT_B200_OK;
}

int gwat_b200_measure_fp64_peak(gwat_b200_ctx *ctx, double *tflops)
{
	if (!ctx || !tflops) return GWAT_B200_ERR_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (grow(ctx, ctx->d_out, ctx->cap_out, 16)) return GWAT_B200_ERR_CUDA;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	const int blocks = sms * 8, iters = 1 << 16;
	double best = 0;
	for (int rep = 0; rep < 5; rep++) {
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
		k_dfma_peak<<<blocks, 256, 0, st>>>(ctx->d_out, iters, 1.0 + rep);
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		float ms = 0;
		CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		const double flops = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
		const double tf = flops / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	ctx->launches += 5;
	*tflops = best;
	return GWAT_B200_OK;
}

#ifdef GWAT_SETUP_PROFILE
int gwat_b200_debug_setup_stamps(long long *out64)
{
	return cudaMemcpyFromSymbol(out64, g_setup_stamp, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif

int gwat_b200_set_kernel_timing(gwat_b200_ctx *c, int on)
{
	if (!c) return GWAT_B200_ERR_ARG;
	std::lock_guard<std::mutex> lock(c->mu);
	c->kernel_timing = on != 0;
	return GWAT_B200_OK;
}

long long gwat_b200_launch_count(const gwat_b200_ctx *c) { return c ? c->launches : 0; }
double gwat_b200_last_kernel_ms(const gwat_b200_ctx *c) { return c ? c->last_ms : 0.0; }
long long gwat_b200_last_active_bins(const gwat_b200_ctx *c) { return c ? c->last_active : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// internal interface (gwat_engine_internal.h)
// ---------------------------------------------------------------------------------------------------------------------
namespace gwat_internal {

int set_error(gwat_b200_ctx *ctx, int code, const std::string &msg) { return fail(ctx, code, msg); }

int loglike_mcmc_lane(gwat_b200_ctx *ctx, int lane, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                      const double *d_params, double gmst, double T_segment, double *d_logL, cudaStream_t st, cudaStream_t st_heavy)
{
	if (int rc = check_ready(ctx, true)) return rc;
	if (lane < 0 || lane > 2) return fail(ctx, GWAT_B200_ERR_ARG, "loglike_mcmc: lane out of range");
	if (W < 0 || (W > 0 && (!d_params || !d_logL))) return fail(ctx, GWAT_B200_ERR_ARG, "loglike_mcmc_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	RepackPlan plan;
	if (make_plan(desc, mod, dimension, plan) != 0)
		return fail(ctx, GWAT_B200_ERR_ARG, "loglike_mcmc_batch: dimension does not match the method and modification struct");
	if (lane > 0) {
		LikeLane &l = ctx->extra[lane - 1];
		if (!l.ev0) {
			CUDA_TRY(ctx, cudaEventCreate(&l.ev0));
			CUDA_TRY(ctx, cudaEventCreate(&l.ev1));
			CUDA_TRY(ctx, cudaMalloc((void **)&l.d_active, sizeof(unsigned long long)));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&l.ev_a, cudaEventDisableTiming));
			CUDA_TRY(ctx, cudaEventCreateWithFlags(&l.ev_b, cudaEventDisableTiming));
		}
	}
	cudaEvent_t ev_a = lane > 0 ? ctx->extra[lane - 1].ev_a : nullptr, ev_b = lane > 0 ? ctx->extra[lane - 1].ev_b : nullptr;
	LaneSwap swap(ctx, lane);
	if (grow(ctx, ctx->d_coef, ctx->cap_walkers, (size_t)W)) return GWAT_B200_ERR_CUDA;
	GWAT_DISPATCH_FAMILY(desc, launch_setup<Fam>(W, st, d_params, nullptr, plan, ctx->net, desc.theory, gmst, T_segment, ctx->d_coef, ctx->d_active));
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	return run_loglike(ctx, desc, W, d_logL, st, st_heavy, ev_a, ev_b);
}

int polarizations_dev(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *h_sources, double **d_out, cudaStream_t st)
{
	if (int rc = check_ready(ctx, false)) return rc;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	if (int rc = setup_from_sources(ctx, desc, W, h_sources, st)) return rc;
	const size_t n = (size_t)W * ctx->L;
	if (grow(ctx, ctx->d_out, ctx->cap_out, 4 * n)) return GWAT_B200_ERR_CUDA;
	double *o = ctx->d_out;
	const dim3 grid(W, (ctx->L + kThreads - 1) / kThreads);  // walkers on x: no 65535 limit on the batch
	const GridPtrs g = grid_ptrs(ctx);
	GWAT_DISPATCH_FAMILY(desc, k_waveform<Fam><<<grid, kThreads, 0, st>>>(ctx->d_coef, g, o, o + n, o + 2 * n, o + 3 * n));
	ctx->launches += 1;
	CUDA_TRY(ctx, cudaGetLastError());
	*d_out = o;
	return GWAT_B200_OK;
}

int fisher_mcmc_dev(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int order, int S,
                    const double *d_params, double gmst, double *d_fisher, cudaStream_t st)
{
	if (int rc = check_ready(ctx, false)) return rc;
	if (S <= 0) return GWAT_B200_OK;
	if (order != 2 && order != 4) return fail(ctx, GWAT_B200_ERR_ARG, "fisher: order must be 2 or 4");
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return fail(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	RepackPlan rp;
	if (make_plan(desc, mod, dimension, rp) != 0)
		return fail(ctx, GWAT_B200_ERR_ARG, "fisher: dimension does not match the method and modification struct");
	FisherPlan fp;
	std::memset(&fp, 0, sizeof(fp));
	fp.rp = rp;
	fp.rp.alpha_unit_fix = 0;  // the stencil works on the physical record; units were fixed by the repack below
	fp.npts = order == 4 ? 4 : 2;
	fp.theory = desc.theory;
	const int dim = dimension;
	const bool fused = fisher_fused_fits(dim, fp.npts, ctx->D) && getenv("GWAT_B200_FISHER_UNFUSED") == nullptr;
	const int chunk = fisher_chunk_size(ctx, S, dim, fp.npts, ctx->D, fused);
	if (int rc = fisher_reserve(ctx, chunk, dim, fp.npts, ctx->D, fused)) return rc;
	for (int s0 = 0; s0 < S; s0 += chunk) {
		const int ns = std::min(chunk, S - s0);
		k_repack_only<<<(ns + 127) / 128, 128, 0, st>>>(d_params + (size_t)s0 * dim, ns, rp, gmst, ctx->d_src);
		ctx->launches += 1;
		if (int rc = fisher_chunk(ctx, desc, fp, ns, chunk, 0, ctx->D, 0, st)) return rc;
		CUDA_TRY(ctx, cudaMemcpyAsync(d_fisher + (size_t)s0 * dim * dim, ctx->d_fisher, sizeof(double) * ns * dim * dim,
		                              cudaMemcpyDeviceToDevice, st));
	}
	return GWAT_B200_OK;
}

}  // namespace gwat_internal
