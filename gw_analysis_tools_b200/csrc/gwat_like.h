// Fused per-bin likelihood terms (GWAT_HD code): what k_loglike runs in its inner loop.
//
// The reference forms, per walker and per detector, three length-L complex arrays and two length-L real integrands
// (src/waveform_generator.cpp:188-199,257-266; src/waveform_util.cpp:865-873,177-179; src/mcmc_gw.cpp:814-857) and sums them.
// Here one bin's contribution to   sum_d  w_d (|r_d|^2 - 2 Re(data_d conj(r_d)))   is formed in registers:
//   r_d      = A e^{-i phi} (ga_d P + gb_d Q) e^{+i t_d f}           t_d = -2 pi DTOA_d
//   |r_d|^2  = A^2 |ga_d P + gb_d Q|^2                                (no trigonometry)
//   Re(data_d conj r_d) = A Re( e^{+i phi} conj(ga_d P + gb_d Q) [data_d e^{-i t_d f}] )
// so the only transcendental per bin is ONE sincos of the carrier phase (plus one of alpha for IMRPhenomPv2), instead of
// 1 + D (2 + D for Pv2) complex exponentials.  On a uniform grid e^{-i t_d f} advances by a constant rotation per step of
// the thread's stride; the state is re-seeded with a full-precision sincos at the start of every thread's run, so the
// accumulated rounding stays below 1e-14 (runs are <= 64 steps).  Non-uniform (Gauss-Legendre) grids evaluate it directly.
// All of this is algebra on the reference's formulas: results agree to rounding (tests: logL <= 1e-9 relative).
#ifndef GWAT_LIKE_H
#define GWAT_LIKE_H

#include "gwat_bins.h"

namespace gwat {

GWAT_HD int min_int(int a, int b) { return a < b ? a : b; }
GWAT_HD cplx cmul(const cplx &a, const cplx &b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

struct LikeGrid {
	const double *f, *sf_hi, *sf_lo, *logf;
	const double *wq, *dre, *dim;  // [D][ld]
	int L;
	int ld;        // leading dimension of the per-detector tables (L padded to a whole number of tiles)
	int uniform;   // f[i] = f[0] + i*df to rounding
	double df;
};

// Carrier amplitude (including the walker's overall scale) and phase argument of e^{-i phi} at one active bin, and for
// IMRPhenomPv2 the twist factors P, Q.  Returns false when the bin contributes nothing.
// `mr_decay`: when > 0 it is exp(-mr_rate (f - fRD)) supplied by the caller (advanced by a constant factor per step on a
// uniform grid) and saves the exponential of the merger-ringdown amplitude.
template <class Fam>
GWAT_HD bool carrier_terms(const WalkerCoef &w, double f, double sf_hi, double sf_lo, double logf, double mr_decay, double &amp,
                           double &arg, cplx &P, cplx &Q)
{
	const DCoef &c = w.d;
	if (f > c.fcut) return false;
	if (Fam::nrt && f > c.nrt_fmerger12) return false;
	const double sixth = bin_sixth_root(c, sf_hi, sf_lo);
	MfPowers p;
	// The full table of powers (with its exact division for the leading TaylorF2 term) only where the inspiral forms are
	// evaluated and for NRTidal; above both inspiral boundaries the carrier needs (M f), its sixth root and -- IMRPhenomPv2 --
	// the two powers of the twist-up.
	const bool full = f < c.f1p || f < c.f1a || Fam::nrt;
	if (full) {
		// mf_powers with the quotient shared: 1/(Mf)^(7/6) = (Mf)^(1/2) / (Mf)^(5/3) -- amplitude side only
		mf_powers(c.M, f, sixth, p);
	} else {
		p.Mf = mul_rn(c.M, f);
		p.sixth = sixth;
		p.seven6 = mul_rn(mul_rn(sixth, c.M), f);
		if (Fam::base == BASE_P) {
			p.third = mul_rn(sixth, sixth);
			p.two3 = mul_rn(p.third, p.third);
		}
	}
	double shape;
	if (f < c.f1a) shape = phenomd_amp_ins(c, p);
	else if (f > c.f3a) {
		const double df = f - c.fRD;
		shape = mr_decay > 0 ? c.mr_num * mr_decay * fast_rcp(df * df + c.mr_w2) : phenomd_amp_mr(c, f);
	} else shape = phenomd_amp_int(c, p.Mf);
	const double inv76 = full ? (p.m53 * (p.third * p.sixth)) : fast_rcp(p.seven6);
	double phase;
	if (f < c.f1p) phase = phenomd_phase_ins<Fam>(c, f, p, logf);
	else if (f > c.f2p) phase = phenomd_phase_mr<Fam>(c, f, sixth);
	else phase = phenomd_phase_int<Fam>(c, f, logf, sixth);

	if (Fam::base == BASE_D) {
		amp = c.A0 * (shape * inv76);
		if (Fam::nrt) {
			nrt_bin(c, f, p, logf, amp, phase);
			amp *= nrt_taper_factor(c, f);
		}
		arg = phenomd_apply_time_phase(c, f, phase);
		P = cplx{1.0, 0.0};
		Q = cplx{0.0, -1.0};
		return true;
	}
	// ---- IMRPhenomPv2 twist-up ----------------------------------------------------------------------------------------
	const PCoef &pc = w.p;
	amp = (pc.A0 * (shape * inv76)) * 0.5;
	const PiPowers pi = pi_powers();
	const double oc = p.third * pi.third;  // omega^(1/3)
	const double x = p.two3 * pi.two3;     // omega^(2/3); sqrt(x) = oc up to rounding
	const double roc = fast_rcp(oc);
	const double x2 = x * x, eta = pc.eta;
	const double Lorb = (eta * (1.0 + pc.lc1 * x + pc.lc2 * x2)) * roc;
	// tan(beta) = SP / (Lorb + SL); the reference takes cos(beta) = 1/sqrt(1 + tan^2) > 0 and positive half-angle roots, i.e.
	// cos(beta) = |u| / sqrt(u^2 + SP^2), sin(beta) = |SP| / sqrt(u^2 + SP^2) with u = Lorb + SL: one reciprocal root, no division
	const double u = Lorb + pc.SL;
	const double rn = fast_rsqrt(u * u + pc.SP2);
	const double cb = fabs(u) * rn, sinb = fabs(pc.SP) * rn;
	// Wigner d^2_{m,2}(beta), m = -2..2, is (s^4, 2 c s^3, sqrt6 s^2 c^2, 2 c^3 s, c^4) in the half angle.  The sums over +-m that
	// the polarisations need are polynomials in cos(beta), sin(beta) themselves:
	//   d_2 - d_-2 = cos b,   d_2 + d_-2 = (1 + cos^2 b)/2,   d_1 + d_-1 = sin b,   d_1 - d_-1 = sin b cos b,   2 d_0 = (sqrt6/2) sin^2 b
	const double sc = sinb * cb, e2 = fma(cb, cb, 1.0), sb2 = sinb * sinb;
	const double roc2 = roc * roc, roc3 = roc2 * roc;
	const double logom = add_rn(c.logpiM, logf);
	const double alpha = ((((pc.acoef[0] * roc3 + pc.acoef[1] * roc2) + pc.acoef[2] * roc) + pc.acoef[3] * logom) + pc.acoef[4] * oc) +
	                     pc.alpha_const;
	const double epsilon = ((((pc.ecoef[0] * roc3 + pc.ecoef[1] * roc2) + pc.ecoef[2] * roc) + pc.ecoef[3] * logom) + pc.ecoef[4] * oc) -
	                       pc.epsilon_offset;
	double s1, c1;
	fast_sincos(alpha, &s1, &c1);
	const double c2 = c1 * c1 - s1 * s1, s2 = 2.0 * s1 * c1;
	// sum_m Y_m (d^2_{-m} e^{-i m alpha} +- d^2_m e^{+i m alpha}) with real Y_m, grouped by |m| (pc.tw: the +- m sums of Y_m)
	const double *tw = pc.tw;
	P = cplx{tw[2] * sb2 + tw[0] * (c1 * sc) + tw[5] * (c2 * e2), tw[0] * (s1 * sinb) + tw[3] * (s2 * cb)};
	Q = cplx{tw[1] * (s1 * sc) + tw[6] * (s2 * e2), -(tw[1] * (c1 * sinb) + tw[4] * (c2 * cb))};
	phase = add_rn(phase, mul_rn(2., epsilon));
	double a = sub_rn(phase, mul_rn(pc.tc, sub_rn(f, pc.f_ref)));
	a = sub_rn(a, pc.phic);
	arg = add_rn(a, mul_rn(pc.tcorr_2pi, f));
	return true;
}

// Per-thread running state of the recurrences (uniform grids): z_d = e^{-i t_d f}, its per-step multiplier, and the
// ringdown-amplitude decay exp(-mr_rate (f - fRD)) with its per-step factor.
// The per-step multipliers are the same for every thread of a CTA.  For the PhenomD / PhenomPv2 families, compiled for 64
// registers, they stay in the CTA's shared seed table and are read where used (E, dstep point there): 14 registers less per
// thread, spills 224 -> 96 bytes, cfg4 -4 %.  The NRTidal kernels (80 registers, no spills) are 3 % faster with private
// copies (E_reg, dstep_reg), which is also what a thread that seeds its own state (host harness) uses.
template <int D>
struct LikeState {
	cplx z[D];
	double decay;
	const cplx *E;
	const double *dstep;
	cplx E_reg[D];
	double dstep_reg;
};
template <class Fam>
GWAT_HD constexpr bool step_in_registers()
{
	return Fam::nrt;
}

// Seed the state at the thread's first frequency f0; `step` is the frequency distance between its consecutive bins and
// f_last the last frequency it may visit.
#if !defined(__CUDA_ARCH__)
template <int D>
GWAT_HD void like_state_init(const WalkerCoef &w, double f0, double step, double f_last, LikeState<D> &st)
{
	st.decay = 0.0;
	st.dstep_reg = 1.0;
	// kept within double range: outside, the exponential is evaluated per bin instead
	const double a0 = -w.d.mr_rate * (f0 - w.d.fRD), a1 = -w.d.mr_rate * (f_last - w.d.fRD);
	if (fabs(a0) < 600.0 && fabs(a1) < 600.0) {
		st.decay = exp(a0);
		st.dstep_reg = exp(-w.d.mr_rate * step);
	}
#pragma unroll
	for (int d = 0; d < D; d++) {
		double sn, cs;
		sincos(mul_rn(w.det[d].tshift, f0), &sn, &cs);
		st.z[d] = cplx{cs, -sn};
		sincos(mul_rn(w.det[d].tshift, step), &sn, &cs);
		st.E_reg[d] = cplx{cs, -sn};
	}
	st.E = st.E_reg;
	st.dstep = &st.dstep_reg;
}
#endif

// One bin: adds  sum_d w_d (|r_d|^2 - 2 Re(d conj r_d))  to acc (and 1 to nact when the bin is below the model's cutoff) and
// advances the recurrences.  wq/dre/dim are the D per-detector values of this bin.
// `tab` supplies the per-detector table values of this bin lazily -- tab.wq(d), tab.dre(d), tab.dim(d) -- so they are
// fetched where they are consumed instead of being held in registers across the carrier evaluation.
template <class Fam, int D, class Tab, class Cnt>
GWAT_HD void like_bin(const WalkerCoef &w, bool uniform, LikeState<D> &st, double f, double sf_hi, double sf_lo, double logf,
                      const Tab &tab, double &acc, Cnt &nact)
{
	double amp, arg;
	cplx P, Q;
	const bool active = carrier_terms<Fam>(w, f, sf_hi, sf_lo, logf, uniform ? st.decay : 0.0, amp, arg, P, Q);
	if (uniform) st.decay *= step_in_registers<Fam>() ? st.dstep_reg : *st.dstep;
	if (!active) {
		if (uniform) {
#pragma unroll
			for (int d = 0; d < D; d++) st.z[d] = cmul(st.z[d], step_in_registers<Fam>() ? st.E_reg[d] : st.E[d]);
		}
		return;
	}
	nact += 1;  // Cnt is an integer type in the kernel: the count stays off the FP64 pipe
	double sn, cs;
	fast_sincos(arg, &sn, &cs);
	double hh = 0.0, Sre = 0.0, Sim = 0.0;
#pragma unroll
	for (int d = 0; d < D; d++) {
		const DetCoef &dc = w.det[d];
		// G = ga P + gb Q   (IMRPhenomD families: P = 1, Q = -i, so G = ga - i gb is a walker constant)
		const double Gre = Fam::base == BASE_D ? dc.ga : dc.ga * P.re + dc.gb * Q.re;
		const double Gim = Fam::base == BASE_D ? -dc.gb : dc.ga * P.im + dc.gb * Q.im;
		const double wq = tab.wq(d);
		hh += wq * (Gre * Gre + Gim * Gim);
		cplx zd;
		if (uniform) {
			zd = st.z[d];
			st.z[d] = cmul(zd, step_in_registers<Fam>() ? st.E_reg[d] : st.E[d]);
		} else {
			double s_, c_;
			fast_sincos(mul_rn(dc.tshift, f), &s_, &c_);
			zd = cplx{c_, -s_};
		}
		// t = data * z ;  S += w * conj(G) * t
		const double dr = tab.dre(d), di = tab.dim(d);
		const double tre = dr * zd.re - di * zd.im, tim = dr * zd.im + di * zd.re;
		Sre += wq * (Gre * tre + Gim * tim);
		Sim += wq * (Gre * tim - Gim * tre);
	}
	// Re(e^{+i arg} S) = cos(arg) Sre - sin(arg) Sim
	const double dh = amp * (cs * Sre - sn * Sim);
	acc += (amp * amp) * hh - 2.0 * dh;
}

// Table access straight from global memory (L2-resident tables).
struct GlobalTab {
	const double *wq_, *dre_, *dim_;
	size_t ld;
	GWAT_HD double wq(int d) const { return wq_[d * ld]; }
	GWAT_HD double dre(int d) const { return dre_[d * ld]; }
	GWAT_HD double dim(int d) const { return dim_[d * ld]; }
};

// Highest frequency at which the walker's model is non-zero.
template <class Fam>
GWAT_HD double walker_fmax(const WalkerCoef &w)
{
	return (Fam::nrt && w.d.nrt_fmerger12 < w.d.fcut) ? w.d.nrt_fmerger12 : w.d.fcut;
}

// ---- the cut of the bin axis --------------------------------------------------------------------------------------------
// A walker's sum over bins is formed in a fixed tree that depends on the grid alone:
//   unit  = kUnitThreads * unit_steps consecutive bins; inside a unit thread t owns bins t, t + 256, ...; the recurrences are
//           re-seeded at every unit, so a unit's value does not depend on what was evaluated before it;
//   per unit one partial sum per warp (shuffle tree), all partials of a walker are added in a fixed order afterwards.
// How many consecutive units one CTA evaluates is a scheduling decision (long runs for big ensembles, short ones when few
// walkers must still fill the GPU) and does NOT change a single bit of the result: a walker's logL is the same in any batch.
constexpr int kUnitThreads = 256;
constexpr int kUnitWarps = kUnitThreads / 32;
#ifndef GWAT_UNIT_STEPS
#define GWAT_UNIT_STEPS 32  // bins per thread and unit on grids of up to 2^20 bins (measured: 16 costs 2 % on cfg2/cfg4, 64 gains 0.5 %)
#endif
constexpr int kMaxUnitsPerCta = 64 / GWAT_UNIT_STEPS;

// Per-CTA seeds of the recurrences (uniform grids).  Seeding every thread's state with its own sincos/exp calls costs
// 2 D + 2 transcendental evaluations per thread and unit -- as much as a whole bin.  Instead the CTA tabulates
// e^{-i t_d df k} for k = 0..15 and k = 16 j, j = 0..15, once (one evaluation per thread, in parallel) and every thread forms
// its seed as a product of three table entries:
//   e^{-i t_d f[unit_begin + t]} = e^{-i t_d f[unit_begin]} * e16[t >> 4] * e1[t & 15]      (3 roundings, < 1e-15)
// and likewise for the ringdown-amplitude decay.  f[i] = f[unit_begin] + (i - unit_begin) df holds to an ulp of f on a
// uniform grid; t_d <= 0.14 s rad/Hz, so the phase differs from the directly evaluated one by < 1e-13 rad.
template <int D>
struct CtaSeeds {
	cplx e1[D][16], e16[D][16], step[D], base[kMaxUnitsPerCta][D];
	double r1[16], r16[16], rstep, rbase[kMaxUnitsPerCta];  // rbase == 0: decay outside double range in that unit -> per-bin exp
	double rstep_unit[kMaxUnitsPerCta];                    // rstep where rbase != 0, else 1 (what the threads multiply by)
};
constexpr int kSeedSlots = 33 + kMaxUnitsPerCta;  // 16 + 16 + step + unit bases per detector, and once more for the decay

// Slot j of the CTA's seed table (0 <= j < kSeedSlots * (D + 1)).  The CTA covers units [0, n_units) of unit_bins bins
// starting at bin `begin`; `stride` is the thread stride of the runs.
template <int D>
GWAT_HD void cta_seed_slot(const WalkerCoef &w, const LikeGrid &g, int begin, int unit_bins, int n_units, int stride, int j,
                           CtaSeeds<D> &sd)
{
	const int d = j / kSeedSlots, k = j % kSeedSlots;
	const int u = k - 33;
	if (u >= n_units) return;
	const int ub = begin + (u > 0 ? u : 0) * unit_bins;                  // first bin of unit u
	const int ue = min_int(g.L, ub + unit_bins) - 1;                     // its last bin
	const double x = k < 16 ? g.df * k : (k < 32 ? g.df * (16 * (k - 16)) : (k == 32 ? g.df * stride : g.f[ub]));
	if (d < D) {
		double sn, cs;
		fast_sincos(mul_rn(w.det[d].tshift, x), &sn, &cs);
		const cplx v{cs, -sn};
		if (k < 16) sd.e1[d][k] = v;
		else if (k < 32) sd.e16[d][k - 16] = v;
		else if (k == 32) sd.step[d] = v;
		else sd.base[u][d] = v;
		return;
	}
	// ringdown decay exp(-mr_rate (f - fRD)); kept within double range, else evaluated per bin
	if (k < 33) {
		const double v = exp(-w.d.mr_rate * x);
		if (k < 16) sd.r1[k] = v;
		else if (k < 32) sd.r16[k - 16] = v;
		else sd.rstep = v;
		return;
	}
	const double a0 = -w.d.mr_rate * (x - w.d.fRD), a1 = -w.d.mr_rate * (g.f[ue] - w.d.fRD);
	const bool in_range = fabs(a0) < 600.0 && fabs(a1) < 600.0;
	sd.rbase[u] = in_range ? exp(a0) : 0.0;
	sd.rstep_unit[u] = in_range ? exp(-w.d.mr_rate * (g.df * stride)) : 1.0;
}

// Seed of the thread whose first bin in unit u is unit_begin + t (t < 256).
template <int D, bool kCopySteps>
GWAT_HD void like_state_from_seeds(const CtaSeeds<D> &sd, int u, int t, LikeState<D> &st)
{
	const int k1 = t & 15, k16 = t >> 4;
#pragma unroll
	for (int d = 0; d < D; d++) {
		st.z[d] = cmul(cmul(sd.base[u][d], sd.e16[d][k16]), sd.e1[d][k1]);
	}
	st.decay = (sd.rbase[u] * sd.r16[k16]) * sd.r1[k1];
	if (kCopySteps) {
#pragma unroll
		for (int d = 0; d < D; d++) st.E_reg[d] = sd.step[d];
		st.dstep_reg = sd.rstep_unit[u];
	}
	st.E = sd.step;  // (never the address of E_reg: that would pin the private copies to local memory)
	st.dstep = &sd.rstep_unit[u];
}

// One thread's share of one unit, read straight from global memory (L2-resident tables): bins first, first + stride, ...
// < end.  `seeds` (uniform grids) is the CTA's seed table, `u` the unit's index in it and `t` the thread's offset; without
// a table the thread seeds its own state.  Returns false once the thread has passed the walker's cutoff on an ascending
// (uniform) grid: everything after that is exactly zero for it.
template <class Fam, int D, class Cnt>
GWAT_HD bool loglike_unit(const WalkerCoef &w, const LikeGrid &g, int first, int end, int stride, double fmax, double &acc,
                          Cnt &nact, const CtaSeeds<D> *seeds = nullptr, int u = 0, int t = 0)
{
	if (first >= end) return true;
	const bool uniform = g.uniform != 0;
	LikeState<D> st;
	if (uniform) {
#if defined(__CUDA_ARCH__)
		like_state_from_seeds<D, step_in_registers<Fam>()>(*seeds, u, t, st);  // device callers always tabulate (no table read here: the cutoff is tested on the first bin below)
#else
		if (seeds) like_state_from_seeds<D, true>(*seeds, u, t, st);
		else {
			const double f0 = g.f[first];
			if (f0 > fmax) return false;
			like_state_init<D>(w, f0, g.df * stride, g.f[min_int(end - 1, g.L - 1)], st);
		}
#endif
	}
	for (int i = first; i < end; i += stride) {
#if defined(__CUDA_ARCH__)
		// The walker's ~150 coefficients live in shared memory.  Without this barrier the compiler hoists all of them into
		// registers as loop invariants (250+ registers, 1 CTA/SM); with it they are re-read (broadcast LDS) where used.
		asm volatile("" ::: "memory");
#endif
		const double f = g.f[i];
		if (uniform && f > fmax) return false;
		const GlobalTab tab{g.wq + i, g.dre + i, g.dim + i, (size_t)g.ld};
		like_bin<Fam, D>(w, uniform, st, f, g.sf_hi[i], g.sf_lo[i], g.logf[i], tab, acc, nact);
	}
	return true;
}

// Bins per unit for a grid of L bins: 16 bins per thread, doubled (up to 64) while that would make more than 128 units.
GWAT_HD int unit_bins_for(int L)
{
	int steps = GWAT_UNIT_STEPS;
	while (steps < 64 && (long long)L > 128LL * steps * kUnitThreads) steps *= 2;
	return steps * kUnitThreads;
}

}  // namespace gwat
#endif
