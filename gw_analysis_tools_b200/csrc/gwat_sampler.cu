// gwat_b200 ensemble sampler: the batched parallel-tempering Metropolis-Hastings step of include/gwat_b200_sampler.h.
//
// Device state (all in HBM, sized for chain_N = C chains of dimension P, history length H):
//   pos[C][P], prop[C][P], ll/lp/llprop/lpprop[C]      current and proposed states
//   hist[C][H][P]                                       per-chain ring buffer for differential evolution
//   fvals[C][P], fvecs[C][P][P]                         Fisher eigen-systems (row i = eigenvector i)
//   widths[C][P+3], counters[C][NCOUNTERS], gauss_ct[C][P][4], type_last[C][4]
// One step of one lane (= contiguous range of chains on its own stream):
//   [Fisher refresh of the chains whose schedule says so]  k_gather -> engine Fisher pass -> k_fisher_eigen
//   k_propose      one thread per chain: draw the step type, build the proposal, evaluate its prior
//   engine         k_setup_mcmc + k_loglike + k_finish on the proposals (lane scratch of the context)
//   k_accept       one thread per chain: Metropolis-Hastings, counters, history, width tuning, cold-chain record
// After every swp_freq steps the lanes join for the swap sweep: k_swap_prepare (thresholds, parallel) -> k_swap_scan (the
// reference's sequential sweep over adjacent chains, one thread, 4 dependent instructions per pair) -> k_swap_apply.
// The host never reads device results inside gwat_b200_sampler_run: the step types are pure functions of (seed, step,
// chain), so the host replays them to know which chains refresh their Fisher matrix at which step.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is opened at run time (see nccl_api)

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "gwat_engine_internal.h"
#include "gwat_method.h"
#include "gwat_sampler_math.h"

using namespace gwat;
using namespace gwat::smp;

namespace {

constexpr int NCT = GWAT_B200_SAMPLER_NCOUNTERS;

struct StepConst {
	uint64_t seed;
	int C, P, H;
	int history_update, check_stepsize_freq;
	int fisher_exist;
	int record_cold;
	int chain_offset;  // global index of chain 0 (several GPUs: every rank draws what one GPU would draw for its chains)
};

struct DevState {
	double *pos, *prop, *ll, *lp, *llprop, *lpprop, *temps;
	double *hist;
	int *hist_pos;
	double *fvals, *fvecs;
	double *widths;
	long long *counters;
	int *gauss_ct;        // [C][P][4]: accept, reject, last accept, last reject
	long long *type_last; // [C][4]: DE last accept/reject, Fisher last accept/reject
	int *info;            // [C]: step type | selected dimension << 8
	int *cold_slot;       // [C]: index among the cold chains, or -1
	double *cold;         // [cap_steps][n_cold][P]
};

__global__ void k_init(DevState d, StepConst k, gwat_b200_prior prior, PriorPlan pp)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= k.C) return;
	const int P = k.P;
	d.lp[c] = standard_log_prior(prior, pp, d.pos + (size_t)c * P);
	for (int i = 0; i < P; i++) {
		d.hist[((size_t)c * k.H) * P + i] = d.pos[(size_t)c * P + i];
		d.widths[(size_t)c * (P + 3) + i] = .05;  // allocate_sampler_mem, src/mcmc_sampler_internals.cpp:1998-2004
		d.fvals[(size_t)c * P + i] = 1;  // (what assign_initial_pos leaves when the first matrix is not finite, :3193-3207)
		for (int j = 0; j < P; j++) d.fvecs[((size_t)c * P + i) * P + j] = (i == j) ? 1.0 : 0.0;
		for (int j = 0; j < 4; j++) d.gauss_ct[((size_t)c * P + i) * 4 + j] = 0;
	}
	d.widths[(size_t)c * (P + 3) + P + 0] = 1;    // DE
	d.widths[(size_t)c * (P + 3) + P + 1] = .05;  // MMALA (unused)
	d.widths[(size_t)c * (P + 3) + P + 2] = .5;   // Fisher
	d.hist_pos[c] = 0;
	for (int j = 0; j < NCT; j++) d.counters[(size_t)c * NCT + j] = 0;
	for (int j = 0; j < 4; j++) d.type_last[(size_t)c * 4 + j] = 0;
}

__global__ void k_propose(DevState d, StepConst k, gwat_b200_prior prior, PriorPlan pp, long long step, int c0, int n)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const int c = c0 + t, P = k.P;
	const double T = d.temps[c];
	double bounds[4];
	step_boundaries(T, k.fisher_exist != 0, step > k.H, bounds);
	double alpha, u_acc, u_pick, u_pick2, n0, n1;
	uniform2(k.seed, (uint64_t)step, (uint32_t)(k.chain_offset + c), DRAW_TYPE_ACCEPT, alpha, u_acc);
	uniform2(k.seed, (uint64_t)step, (uint32_t)(k.chain_offset + c), DRAW_PICK, u_pick, u_pick2);
	uniform2(k.seed, (uint64_t)step, (uint32_t)(k.chain_offset + c), DRAW_NORMAL, n0, n1);
	const double z = normal_from(n0, n1);
	const int type = step_type(alpha, bounds);
	const double *cur = d.pos + (size_t)c * P;
	double *prop = d.prop + (size_t)c * P;
	const double *w = d.widths + (size_t)c * (P + 3);
	int sel = 0;
	if (type == STEP_GAUSS) {
		sel = propose_gaussian(cur, prop, P, w, u_pick, z);
	} else if (type == STEP_DE) {
		int i, j;
		de_pick(k.H, u_pick, u_pick2, i, j);
		double beta, unused;
		uniform2(k.seed, (uint64_t)step, (uint32_t)(k.chain_offset + c), DRAW_DE_SCALE, beta, unused);
		propose_de(cur, prop, P, d.hist + ((size_t)c * k.H + i) * P, d.hist + ((size_t)c * k.H + j) * P, beta, z, w[P + 0]);
	} else {
		propose_fisher(cur, prop, P, d.fvals + (size_t)c * P, d.fvecs + (size_t)c * P * P, T, u_pick, z, w[P + 2]);
	}
	d.lpprop[c] = standard_log_prior(prior, pp, prop);
	d.info[c] = type | (sel << 8);
}

__global__ void k_accept(DevState d, StepConst k, long long step, int c0, int n, long long cold_row)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const int c = c0 + t, P = k.P;
	const double T = d.temps[c];
	double alpha, u_acc;
	uniform2(k.seed, (uint64_t)step, (uint32_t)(k.chain_offset + c), DRAW_TYPE_ACCEPT, alpha, u_acc);
	const int type = d.info[c] & 0xff, sel = d.info[c] >> 8;
	double *cur = d.pos + (size_t)c * P;
	const bool acc = mh_accept(d.ll[c], d.llprop[c], d.lp[c], d.lpprop[c], T, u_acc);
	long long *ct = d.counters + (size_t)c * NCT;
	if (acc) {
		const double *prop = d.prop + (size_t)c * P;
		for (int i = 0; i < P; i++) cur[i] = prop[i];
		d.ll[c] = d.llprop[c];
		d.lp[c] = d.lpprop[c];
	}
	ct[acc ? GWAT_B200_CT_STEP_ACCEPT : GWAT_B200_CT_STEP_REJECT] += 1;
	if (type == STEP_GAUSS) {  // assign_ct_p / assign_ct_m (:2992-3015)
		ct[acc ? GWAT_B200_CT_GAUSS_ACCEPT : GWAT_B200_CT_GAUSS_REJECT] += 1;
		d.gauss_ct[((size_t)c * P + sel) * 4 + (acc ? 0 : 1)] += 1;
	} else if (type == STEP_DE) {
		ct[acc ? GWAT_B200_CT_DE_ACCEPT : GWAT_B200_CT_DE_REJECT] += 1;
	} else {
		ct[acc ? GWAT_B200_CT_FISHER_ACCEPT : GWAT_B200_CT_FISHER_REJECT] += 1;
	}
	// PTMCMC_MH_step_incremental (src/mcmc_sampler.cpp:4603-4642), chain_pos = step + 1 from here on
	const long long chain_pos = step + 1;
	const bool primed = step > k.H;
	if (!primed || chain_pos % k.history_update == 0) {  // update_history (:2198-2219)
		int hp = d.hist_pos[c];
		hp = (hp < k.H - 1) ? hp + 1 : 0;
		d.hist_pos[c] = hp;
		double *h = d.hist + ((size_t)c * k.H + hp) * P;
		for (int i = 0; i < P; i++) h[i] = cur[i];
	}
	if (chain_pos % k.check_stepsize_freq == 0) {  // update_step_widths (:1623-1703)
		double bounds[4];
		step_boundaries(T, k.fisher_exist != 0, primed, bounds);
		const double p_gauss = bounds[0], p_de = bounds[1] - bounds[0], p_fisher = bounds[3] - bounds[2];
		const double lo = .2, hi = .60 - .2 / T;  // :1953-1954
		double *w = d.widths + (size_t)c * (P + 3);
		if (p_gauss != 0) {
			for (int i = 0; i < P; i++) {
				int *g = d.gauss_ct + ((size_t)c * P + i) * 4;
				w[i] = tuned_width(w[i], g[0] - g[2], g[1] - g[3], lo, hi);
				g[2] = g[0];
				g[3] = g[1];
			}
		}
		long long *tl = d.type_last + (size_t)c * 4;
		if (p_de != 0) {
			w[P + 0] = tuned_width(w[P + 0], ct[GWAT_B200_CT_DE_ACCEPT] - tl[0], ct[GWAT_B200_CT_DE_REJECT] - tl[1], lo, hi);
			tl[0] = ct[GWAT_B200_CT_DE_ACCEPT];
			tl[1] = ct[GWAT_B200_CT_DE_REJECT];
		}
		if (p_fisher != 0) {
			w[P + 2] = tuned_width(w[P + 2], ct[GWAT_B200_CT_FISHER_ACCEPT] - tl[2], ct[GWAT_B200_CT_FISHER_REJECT] - tl[3], lo, hi);
			tl[2] = ct[GWAT_B200_CT_FISHER_ACCEPT];
			tl[3] = ct[GWAT_B200_CT_FISHER_REJECT];
		}
	}
	if (k.record_cold && d.cold_slot[c] >= 0 && cold_row >= 0) {
		double *o = d.cold + ((size_t)cold_row + d.cold_slot[c]) * P;
		for (int i = 0; i < P; i++) o[i] = cur[i];
	}
}

// ---- Fisher refresh --------------------------------------------------------------------------------------------------------
__global__ void k_gather(const double *__restrict__ pos, const int *__restrict__ idx, int n, int P, double *__restrict__ out)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n * P) return;
	out[t] = pos[(size_t)idx[t / P] * P + t % P];
}
// transformations + eigen-system of matrix t = blockIdx.x, written to slot t of the outputs; ok[t] = 0 when the result has a NaN.
// One warp per matrix: the cyclic Jacobi sweep of jacobi_eigen (gwat_sampler_math.h) with lane k owning row/column k of each
// rotation -- the same rotations in the same order, 3 warp-synchronous updates per rotation instead of 3 serial loops.
__global__ void __launch_bounds__(32) k_fisher_eigen(double *__restrict__ F, const double *__restrict__ params, int n_mat, int P, int pv2,
                                                    int alpha_unit_fix, int ppE_Nmod, double *__restrict__ out_vals,
                                                    double *__restrict__ out_vecs, int *__restrict__ ok_out)
{
	constexpr int N = GWAT_B200_MAX_DIM;
	__shared__ double A[N * N], V[N * N], vals[N];
	__shared__ int order[N];
	const int t = blockIdx.x, lane = threadIdx.x, n = P;
	if (t >= n_mat) return;
	double *Fg = F + (size_t)t * P * P;
	if (lane == 0) fisher_transformations(Fg, P, pv2 != 0, alpha_unit_fix != 0, ppE_Nmod, params + (size_t)t * P);
	__syncwarp();
	for (int i = lane; i < n * n; i += 32) {
		A[i] = Fg[i];
		V[i] = (i / n == i % n) ? 1.0 : 0.0;
	}
	__syncwarp();
	for (int sweep = 0; sweep < 60; sweep++) {
		double off = 0, diag = 0;
		if (lane < n) {
			diag = A[lane * n + lane] * A[lane * n + lane];
			for (int j = lane + 1; j < n; j++) off += A[lane * n + j] * A[lane * n + j];
		}
		for (int o = 16; o > 0; o >>= 1) {
			off += __shfl_xor_sync(0xffffffffu, off, o);
			diag += __shfl_xor_sync(0xffffffffu, diag, o);
		}
		if (!(off > 1e-32 * diag)) break;
		for (int p = 0; p < n - 1; p++) {
			for (int q = p + 1; q < n; q++) {
				const double apq = A[p * n + q];
				if (apq == 0.0) continue;  // uniform across the warp
				const double app = A[p * n + p], aqq = A[q * n + q];
				const double theta = (aqq - app) / (2.0 * apq);
				const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
				__syncwarp();
				if (lane < n) {
					const double akp = A[lane * n + p], akq = A[lane * n + q];
					A[lane * n + p] = c * akp - sn * akq;
					A[lane * n + q] = sn * akp + c * akq;
				}
				__syncwarp();
				if (lane < n) {
					const double apk = A[p * n + lane], aqk = A[q * n + lane];
					A[p * n + lane] = c * apk - sn * aqk;
					A[q * n + lane] = sn * apk + c * aqk;
					const double vpk = V[p * n + lane], vqk = V[q * n + lane];
					V[p * n + lane] = c * vpk - sn * vqk;
					V[q * n + lane] = sn * vpk + c * vqk;
				}
				__syncwarp();
			}
		}
	}
	if (lane < n) vals[lane] = A[lane * n + lane];
	__syncwarp();
	if (lane == 0) {  // ascending order (Eigen's convention): insertion sort of the indices
		for (int i = 0; i < n; i++) {
			int j = i;
			while (j > 0 && vals[order[j - 1]] > vals[i]) {
				order[j] = order[j - 1];
				j--;
			}
			order[j] = i;
		}
	}
	__syncwarp();
	bool good = true;
	for (int i = lane; i < n * n; i += 32) good = good && (V[i] == V[i]);
	if (lane < n) good = good && (vals[lane] == vals[lane]);
	good = __all_sync(0xffffffffu, good);
	if (good) {
		if (lane < n) out_vals[(size_t)t * P + lane] = vals[order[lane]];
		for (int i = lane; i < n * n; i += 32) out_vecs[(size_t)t * P * P + i] = V[order[i / n] * n + i % n];
	}
	if (ok_out && lane == 0) ok_out[t] = good ? 1 : 0;
}
// staged eigen-systems -> the chains they belong to (update_fisher's tail, :676-711: a NaN result keeps the old system)
__global__ void k_fisher_commit(const int *__restrict__ idx, const int *__restrict__ ok, int n, int P, const double *__restrict__ vals,
                                const double *__restrict__ vecs, double *__restrict__ fvals, double *__restrict__ fvecs,
                                long long *__restrict__ counters)
{
	const int t = blockIdx.x;
	if (t >= n) return;
	const int c = idx[t];
	if (ok[t]) {
		for (int i = threadIdx.x; i < P; i += blockDim.x) fvals[(size_t)c * P + i] = vals[(size_t)t * P + i];
		for (int i = threadIdx.x; i < P * P; i += blockDim.x) fvecs[(size_t)c * P * P + i] = vecs[(size_t)t * P * P + i];
	}
	if (threadIdx.x == 0) counters[(size_t)c * NCT + (ok[t] ? GWAT_B200_CT_FISHER_UPDATES : GWAT_B200_CT_FISHER_NAN)] += 1;
}

// ---- swap sweep --------------------------------------------------------------------------------------------------------------
// Pair i = (slot i, slot i+1).  The reference accepts when !(exp((l1-l2)/T2 - (l1-l2)/T1) < alpha); with l2 = ll[i+1] fixed
// that is a threshold on l1 (the likelihood of whatever state sits in slot i when the sweep arrives), computed here for
// all pairs at once so that the sequential part is one comparison per pair.
//   kind 0: never swap (equal temperatures), 1: swap iff l1 >= thr, 2: swap iff l1 <= thr, 3: always swap
__global__ void k_swap_prepare(const double *__restrict__ ll, const double *__restrict__ temps, uint64_t seed, long long sweep, int C,
                               int chain_offset, double *__restrict__ thr, int *__restrict__ kind)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= C - 1) return;
	double alpha, unused;
	uniform2(seed, (uint64_t)sweep, (uint32_t)(chain_offset + i), DRAW_SWAP, alpha, unused);
	int kd;
	double th;
	swap_threshold(ll[i + 1], temps[i], temps[i + 1], alpha, kd, th);
	kind[i] = kd;
	thr[i] = th;
}
// src[i] = which slot's state ends up in slot i (chain_swap's sweep, src/mcmc_sampler_internals.cpp:1086-1118).
//
// The sweep is sequential as the reference writes it -- whether pair (i, i+1) swaps depends on what pair (i-1, i) did -- but the
// dependence has a simple shape.  A state that the sweep carries upwards from slot s (logL = ll[s]) keeps moving while it passes
// the thresholds of the pairs s, s+1, ...; the first pair e(s) it fails ends its RUN, slot e(s) receives it, and the next run
// starts at slot e(s) + 1 with that slot's own state.  So with nxt[s] = e(s) + 1 (computed for every slot at once, a couple of
// comparisons each: runs are short) the slots where runs start are the orbit of slot 0 under nxt, and
//   accepted[j] = (slot j+1 does not start a run),   src[j] = j + 1 for accepted pairs,   src[nxt[s] - 1] = s for every start s
// (the last run ends at the top slot: nxt = C).  The orbit is marked by pointer doubling: after round k the starts with orbit
// index < 2^k are marked and jump = nxt^(2^k), ceil(log2 C) + 1 rounds in all.  The comparisons are the very ones of swap_scan
// (gwat_sampler_math.h) on the same values, so the decisions are identical; a run longer than kRunCap pairs (practically only
// adversarial inputs) makes the CTA fall back to the staged sequential walk below.
// One CTA; the jump and mark arrays live in shared memory (16-bit / 8-bit) up to kScanSmemSlots slots, else in global scratch (L2).
// Measured before (one thread walking the ladder out of staged shared memory): 0.15 ms per 4096 chains, 1.2 ms per 32768.
constexpr int kScanChunk = 1536, kScanThreads = 1024, kRunCap = 256;
constexpr int kScanSmemSlots = 36000;  // 5 B per slot (two 16-bit jump arrays, byte marks): 176 KB of dynamic shared memory next to the 42 KB of the fallback's staging
__device__ __forceinline__ bool swap_passes(int kd, double th, double carry)
{
	return (kd == 3) || (kd == 1 && carry >= th) || (kd == 2 && carry <= th);
}
// The sequential walk (fallback): one thread walks the ladder; everything that does not depend on the carried state is staged in
// shared memory by the whole CTA in coalesced chunks, and the results leave the same way.
__device__ void swap_scan_staged(const double *__restrict__ ll, const double *__restrict__ thr, const int *__restrict__ kind, int C, int *__restrict__ src,
                                 int *__restrict__ accepted)
{
	__shared__ double s_thr[kScanChunk], s_ll[kScanChunk];
	__shared__ int s_kind[kScanChunk], s_src[kScanChunk], s_acc[kScanChunk];
	__shared__ double s_carry;
	__shared__ int s_carry_src;
	if (threadIdx.x == 0) {
		s_carry = ll[0];
		s_carry_src = 0;
	}
	const int pairs = C - 1;
	for (int base = 0; base < pairs; base += kScanChunk) {
		const int n = min(kScanChunk, pairs - base);
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			s_thr[j] = thr[base + j];
			s_kind[j] = kind[base + j];
			s_ll[j] = ll[base + j + 1];
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			double carry = s_carry;
			int carry_src = s_carry_src;
			for (int j = 0; j < n; j++) {  // the very decisions of swap_scan
				const bool sw = swap_passes(s_kind[j], s_thr[j], carry);
				s_src[j] = sw ? base + j + 1 : carry_src;
				s_acc[j] = sw ? 1 : 0;
				if (!sw) {
					carry = s_ll[j];
					carry_src = base + j + 1;
				}
			}
			s_carry = carry;
			s_carry_src = carry_src;
		}
		__syncthreads();
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			src[base + j] = s_src[j];
			accepted[base + j] = s_acc[j];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) src[C - 1] = s_carry_src;
}
// Rounds of pointer doubling over the slots [0, C): ja = nxt on entry, mk[0] = 1.  Idx/Mark: 16-bit jumps and byte marks in shared
// memory for ladders below 2^16 slots, ints in global scratch otherwise.
template <class Idx, class Mark>
__device__ __forceinline__ void swap_mark_starts(Idx *ja, Idx *jb, Mark *mk, int C)
{
	for (int span = 1; span < 2 * C; span *= 2) {
		// marks first (stores into mk), then the jumps as a pass of loads that nothing in between can alias
		for (int s = threadIdx.x; s < C; s += blockDim.x) {
			if (mk[s]) {
				const int t = ja[s];
				if (t < C) mk[t] = 1;
			}
		}
		{
			const Idx *__restrict__ a = ja;
			Idx *__restrict__ o = jb;
#pragma unroll 8
			for (int s = threadIdx.x; s < C; s += blockDim.x) {
				const int t = a[s];
				o[s] = t < C ? a[t] : (Idx)C;
			}
		}
		__syncthreads();
		Idx *sw = ja;
		ja = jb;
		jb = sw;
	}
}
// work: 4 C ints of global scratch (nxt, and the jump/mark arrays of ladders too long for shared memory); force_sequential: tests only
__global__ void __launch_bounds__(kScanThreads) k_swap_scan(const double *__restrict__ ll, const double *__restrict__ thr, const int *__restrict__ kind, int C,
                                                           int *__restrict__ src, int *__restrict__ accepted, int *__restrict__ work, int force_sequential)
{
	extern __shared__ unsigned char scan_smem[];
	__shared__ int s_slow;
	if (blockIdx.x != 0) return;
	const bool in_smem = C <= kScanSmemSlots;
	int *nxt = work;
	unsigned short *ja16 = reinterpret_cast<unsigned short *>(scan_smem), *jb16 = ja16 + C;
	unsigned char *mk8 = reinterpret_cast<unsigned char *>(jb16 + C);
	int *ja32 = work + C, *jb32 = ja32 + C, *mk32 = jb32 + C;
	const int pairs = C - 1;
	if (threadIdx.x == 0) s_slow = force_sequential;
	__syncthreads();
	// runs: nxt[s] = 1 + the first pair at or after s that the state of slot s does not pass
	for (int s = threadIdx.x; s < C; s += blockDim.x) {
		const double carry = ll[s];
		int j = s;
		while (j < pairs && j - s < kRunCap && swap_passes(kind[j], thr[j], carry)) j++;
		if (j < pairs && j - s >= kRunCap) s_slow = 1;
		nxt[s] = j + 1;
		if (in_smem) {
			ja16[s] = (unsigned short)(j + 1);
			mk8[s] = s == 0 ? 1 : 0;
		} else {
			ja32[s] = j + 1;
			mk32[s] = s == 0 ? 1 : 0;
		}
	}
	__syncthreads();
	if (s_slow) {
		swap_scan_staged(ll, thr, kind, C, src, accepted);
		return;
	}
	// the starts of the runs: the orbit of slot 0 under nxt, by pointer doubling
	if (in_smem) swap_mark_starts(ja16, jb16, mk8, C);
	else swap_mark_starts(ja32, jb32, mk32, C);
	for (int j = threadIdx.x; j < pairs; j += blockDim.x) {
		const int acc = (in_smem ? (int)mk8[j + 1] : mk32[j + 1]) ? 0 : 1;
		accepted[j] = acc;
		if (acc) src[j] = j + 1;
	}
	for (int s = threadIdx.x; s < C; s += blockDim.x)
		if (in_smem ? (int)mk8[s] : mk32[s]) src[nxt[s] - 1] = s;
}
inline size_t scan_smem_bytes(int C) { return C <= kScanSmemSlots ? (size_t)C * 5 : 0; }
inline void scan_smem_opt_in()
{
	static std::atomic<bool> done[64];
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (done[dev].load(std::memory_order_acquire)) return;
	cudaFuncSetAttribute(k_swap_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(kScanSmemSlots));
	done[dev].store(true, std::memory_order_release);
}
// swap counters of the chains [c0, c0 + C) of a ladder of Ct chains: chain g took part in the pairs (g-1, g) and (g, g+1)
__global__ void k_swap_count(const int *__restrict__ accepted, int Ct, int c0, int C, long long *__restrict__ counters)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= C) return;
	const int g = c0 + c;
	long long acc = 0, rej = 0;
	if (g > 0) (accepted[g - 1] ? acc : rej) += 1;
	if (g < Ct - 1) (accepted[g] ? acc : rej) += 1;
	counters[(size_t)c * NCT + GWAT_B200_CT_SWAP_ACCEPT] += acc;
	counters[(size_t)c * NCT + GWAT_B200_CT_SWAP_REJECT] += rej;
}
__global__ void k_swap_apply(const int *__restrict__ src, int C, int P, const double *__restrict__ pos, const double *__restrict__ ll,
                             const double *__restrict__ lp, double *__restrict__ pos2, double *__restrict__ ll2, double *__restrict__ lp2)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= C * P) return;
	const int c = t / P, i = t % P, s = src[c];
	pos2[t] = pos[(size_t)s * P + i];
	if (i == 0) {
		ll2[c] = ll[s];
		lp2[c] = lp[s];
	}
}

// Sharded ladder: one record per chain, [position | logL | logP], what the swap may move
__global__ void k_swap_pack(int C, int P, const double *__restrict__ pos, const double *__restrict__ ll, const double *__restrict__ lp,
                            double *__restrict__ rec)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x, R = P + 2;
	if (t >= C * R) return;
	const int c = t / R, i = t % R;
	rec[t] = i < P ? pos[(size_t)c * P + i] : (i == P ? ll[c] : lp[c]);
}
__global__ void k_swap_global_ll(int Ct, int P, const double *__restrict__ rec, double *__restrict__ g_ll)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < Ct) g_ll[c] = rec[(size_t)c * (P + 2) + P];
}
__global__ void k_swap_take(const int *__restrict__ src, int c0, int C, int P, const double *__restrict__ rec, double *__restrict__ pos,
                            double *__restrict__ ll, double *__restrict__ lp)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x, R = P + 2;
	if (t >= C * R) return;
	const int c = t / R, i = t % R;
	const double v = rec[(size_t)src[c0 + c] * R + i];
	if (i < P) pos[(size_t)c * P + i] = v;
	else if (i == P) ll[c] = v;
	else lp[c] = v;
}

__global__ void k_prior(const double *__restrict__ params, int W, gwat_b200_prior prior, PriorPlan pp, double *__restrict__ out)
{
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	out[w] = standard_log_prior(prior, pp, params + (size_t)w * pp.dimension);
}

#define SCUDA(ctx, expr)                                                                                        \
	do {                                                                                                          \
		cudaError_t e_ = (expr);                                                                                    \
		if (e_ != cudaSuccess)                                                                                      \
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
	} while (0)

template <class T>
cudaError_t dalloc(T *&p, size_t n)
{
	return cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
}

// model layout: which prior, where the modifications start
int make_prior_plan(const char *method, const gwat_b200_mod *mod, int dimension, PriorPlan &pp, MethodDesc &desc, bool &alpha_fix,
                    int &ppE_Nmod)
{
	if (parse_method(method, desc) != 0 || desc.mcmc) return -1;
	const int tidal_love = mod ? mod->tidal_love : 1;
	int base = desc.pv2 ? 15 : 11;
	if (desc.nrt) base += tidal_love ? 1 : 2;
	if (dimension < base || dimension > GWAT_B200_MAX_DIM) return -2;
	pp.pv2 = desc.pv2;
	pp.nrt = desc.nrt;
	pp.tidal_love = tidal_love;
	pp.dimension = dimension;
	pp.first_mod = base;
	alpha_fix = theory_alpha_units(desc.theory);
	ppE_Nmod = mod ? mod->ppE_Nmod : 0;
	return 0;
}

}  // namespace

// ---- NCCL, loaded on demand --------------------------------------------------------------------------------------------------
// Only the sharded sampler needs it, so the library does not link it: libnccl.so.2 is opened the first time a sampler is
// attached to a group of ranks (a process that already loaded NCCL -- e.g. through torch.distributed -- gets that copy).
namespace {
struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string error;
};
NcclApi &nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char *n : names) {
			api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if (api.handle) break;
		}
		if (!api.handle) {
			api.error = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "libnccl.so.2 could not be opened");
			return;
		}
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
		api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
		if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) api.error = "NCCL: missing symbols";
	});
	return api;
}
}  // namespace

struct gwat_b200_sampler {
	gwat_b200_ctx *ctx = nullptr;
	int device = 0;
	std::string method;
	gwat_b200_mod mod;
	gwat_b200_sampler_options opt;
	gwat_b200_prior prior;
	PriorPlan pp;
	MethodDesc desc;
	bool alpha_fix = false;
	int ppE_Nmod = 0;
	double gmst = 0, T_segment = 0;
	StepConst k;
	DevState d{};
	double *pos2 = nullptr, *ll2 = nullptr, *lp2 = nullptr;  // swap double buffers
	double *swap_thr = nullptr;
	int *swap_kind = nullptr, *swap_src = nullptr, *swap_acc = nullptr, *swap_work = nullptr;  // swap_work: 4 C ints, k_swap_scan's arrays when the ladder is too long for shared memory
	// Fisher refresh pipeline, per lane: device staging slots (one per step in flight) and a longer ring of pinned index lists
	struct RefreshLane {
		int ND = 1, NP = 1;
		std::vector<int *> d_idx, d_ok;
		std::vector<double *> d_par, d_mat, d_vals, d_vecs;
		std::vector<cudaEvent_t> ev_gathered, ev_done, ev_h2d;
		std::vector<int> n_in_slot;
		std::vector<char> h_used;
		int *h_idx = nullptr;  // pinned [NP][lane_n]
		long long next_sched = 0, h_counter = 0;
	} rf[2];
	int lookahead = 0;
	// deferred refreshes (fisher_deferred): chains that came due since the last swap-sweep boundary, one staging set, one
	// pass in flight
	bool deferred = false;
	std::vector<int> due;
	std::vector<char> due_mark;
	int *g_idx = nullptr, *g_ok = nullptr;
	double *g_par = nullptr, *g_mat = nullptr, *g_vals = nullptr, *g_vecs = nullptr;
	int g_inflight = 0;
	cudaEvent_t ev_g_gathered = nullptr, ev_g_done = nullptr;
	static constexpr int GNP = 16;
	int *gh_idx = nullptr;  // pinned [GNP][C]
	cudaEvent_t gh_ev[GNP] = {};
	bool gh_used[GNP] = {};
	long long gh_counter = 0;
	int nlanes = 1, lane_c0[2] = {0, 0}, lane_n[2] = {0, 0};
	cudaStream_t st[2] = {nullptr, nullptr}, st_like[2] = {nullptr, nullptr}, st_fisher = nullptr;
	cudaEvent_t ev_lane[2] = {nullptr, nullptr}, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
	// host mirror of the schedule
	std::vector<double> h_temps;
	std::vector<int> h_fisher_ct;
	long long step = 0, sweep = 0;
	int since_swap = 0;
	int n_cold = 0;
	long long cold_cap = 0;
	double last_ms = 0;
	long long last_launches = 0;
	// Sharded ladder (gwat_b200_sampler_attach_ranks): this sampler owns chains [rank * C, (rank + 1) * C) of a ladder of
	// n_ranks * C chains; the swap sweep all-gathers every rank's (position, logL, logP) records over NCCL and every rank runs
	// the same sweep over the whole ladder
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1;
	double *x_send = nullptr, *x_recv = nullptr;  // [C][P + 2] and [n_ranks * C][P + 2]
	double *g_ll = nullptr, *g_temps = nullptr, *g_thr = nullptr;
	int *g_kind = nullptr, *g_src = nullptr, *g_acc = nullptr, *g_work = nullptr;
	static constexpr int NSW = 64;                  // swap sweeps of a run whose exchange is timed
	cudaEvent_t ev_sw0[NSW] = {}, ev_sw1[NSW] = {};
	int n_sw_timed = 0;
	double last_swap_ms = 0;
	long long last_sweeps = 0;
};

namespace {

int ensure_cold(gwat_b200_sampler *s, long long steps_total)
{
	if (!s->opt.record_cold || s->n_cold == 0 || steps_total <= s->cold_cap) return 0;
	long long cap = std::max<long long>(steps_total, s->cold_cap * 2);
	double *p = nullptr;
	const size_t row = (size_t)s->n_cold * s->k.P;
	SCUDA(s->ctx, dalloc(p, (size_t)cap * row));
	if (s->d.cold && s->step > 0)
		SCUDA(s->ctx, cudaMemcpy(p, s->d.cold, (size_t)s->step * row * sizeof(double), cudaMemcpyDeviceToDevice));
	cudaFree(s->d.cold);
	s->d.cold = p;
	s->cold_cap = cap;
	return 0;
}

// chains of lane `ln` whose Fisher eigen-system is refreshed before step `step` (fisher_step, :434-437 and :627)
void fisher_schedule(gwat_b200_sampler *s, int ln, long long step, std::vector<int> &flagged)
{
	flagged.clear();
	if (!s->opt.fisher_exist) return;
	const bool primed = step > s->k.H;
	for (int c = s->lane_c0[ln]; c < s->lane_c0[ln] + s->lane_n[ln]; c++) {
		double bounds[4], alpha, u;
		step_boundaries(s->h_temps[c], true, primed, bounds);
		uniform2(s->k.seed, (uint64_t)step, (uint32_t)(s->k.chain_offset + c), DRAW_TYPE_ACCEPT, alpha, u);
		if (step_type(alpha, bounds) != STEP_FISHER) continue;
		if (s->h_fisher_ct[c] == s->opt.fisher_update_number) {
			flagged.push_back(c);
			s->h_fisher_ct[c] = 0;
		}
		s->h_fisher_ct[c] += 1;
	}
}

// Start the refresh that step `t` of lane `ln` will need: gather the flagged chains' CURRENT positions on the lane's stream,
// then Fisher matrices and eigen-systems on the side stream, into the staging slot of step t.
int refresh_schedule(gwat_b200_sampler *s, int ln, long long t, std::vector<int> &flagged)
{
	gwat_b200_ctx *ctx = s->ctx;
	gwat_b200_sampler::RefreshLane &r = s->rf[ln];
	fisher_schedule(s, ln, t, flagged);
	const int n = (int)flagged.size(), P = s->k.P, ds = (int)(t % r.ND);
	r.n_in_slot[ds] = n;
	if (n == 0) return 0;
	const int ps = (int)(r.h_counter++ % r.NP);
	if (r.h_used[ps]) SCUDA(ctx, cudaEventSynchronize(r.ev_h2d[ps]));  // the copy that last read this pinned slice is done
	int *h = r.h_idx + (size_t)ps * s->lane_n[ln];
	std::memcpy(h, flagged.data(), sizeof(int) * n);
	cudaStream_t st = s->st[ln];
	SCUDA(ctx, cudaMemcpyAsync(r.d_idx[ds], h, sizeof(int) * n, cudaMemcpyHostToDevice, st));
	SCUDA(ctx, cudaEventRecord(r.ev_h2d[ps], st));
	r.h_used[ps] = 1;
	k_gather<<<(n * P + 255) / 256, 256, 0, st>>>(s->d.pos, r.d_idx[ds], n, P, r.d_par[ds]);
	SCUDA(ctx, cudaEventRecord(r.ev_gathered[ds], st));
	cudaStream_t sf = s->st_fisher;
	SCUDA(ctx, cudaStreamWaitEvent(sf, r.ev_gathered[ds], 0));
	if (int rc = gwat_internal::fisher_mcmc_dev(ctx, s->method.c_str(), &s->mod, P, s->opt.fisher_deriv_order, n, r.d_par[ds], s->gmst,
	                                            r.d_mat[ds], sf))
		return rc;
	k_fisher_eigen<<<n, 32, 0, sf>>>(r.d_mat[ds], r.d_par[ds], n, P, s->pp.pv2, s->alpha_fix ? 1 : 0, s->ppE_Nmod, r.d_vals[ds],
	                                             r.d_vecs[ds], r.d_ok[ds]);
	SCUDA(ctx, cudaEventRecord(r.ev_done[ds], sf));
	s->last_launches += 2;
	return 0;
}

// Install the eigen-systems staged for step `t` (no-op when none were flagged).
int refresh_commit(gwat_b200_sampler *s, int ln, long long t)
{
	gwat_b200_ctx *ctx = s->ctx;
	gwat_b200_sampler::RefreshLane &r = s->rf[ln];
	const int ds = (int)(t % r.ND), n = r.n_in_slot[ds], P = s->k.P;
	if (n == 0) return 0;
	cudaStream_t st = s->st[ln];
	SCUDA(ctx, cudaStreamWaitEvent(st, r.ev_done[ds], 0));
	k_fisher_commit<<<n, 64, 0, st>>>(r.d_idx[ds], r.d_ok[ds], n, P, r.d_vals[ds], r.d_vecs[ds], s->d.fvals, s->d.fvecs, s->d.counters);
	r.n_in_slot[ds] = 0;
	s->last_launches += 1;
	return 0;
}

// One Fisher pass for `list` (chain indices), gathered from the current positions on `st`, computed on the side stream into the
// group staging set.
int group_launch(gwat_b200_sampler *s, const std::vector<int> &list, cudaStream_t st)
{
	gwat_b200_ctx *ctx = s->ctx;
	const int n = (int)list.size(), P = s->k.P;
	const int ps = (int)(s->gh_counter++ % gwat_b200_sampler::GNP);
	if (s->gh_used[ps]) SCUDA(ctx, cudaEventSynchronize(s->gh_ev[ps]));
	int *h = s->gh_idx + (size_t)ps * s->k.C;
	std::memcpy(h, list.data(), sizeof(int) * n);
	SCUDA(ctx, cudaMemcpyAsync(s->g_idx, h, sizeof(int) * n, cudaMemcpyHostToDevice, st));
	SCUDA(ctx, cudaEventRecord(s->gh_ev[ps], st));
	s->gh_used[ps] = true;
	k_gather<<<(n * P + 255) / 256, 256, 0, st>>>(s->d.pos, s->g_idx, n, P, s->g_par);
	SCUDA(ctx, cudaEventRecord(s->ev_g_gathered, st));
	cudaStream_t sf = s->st_fisher;
	SCUDA(ctx, cudaStreamWaitEvent(sf, s->ev_g_gathered, 0));
	if (int rc = gwat_internal::fisher_mcmc_dev(ctx, s->method.c_str(), &s->mod, P, s->opt.fisher_deriv_order, n, s->g_par, s->gmst, s->g_mat, sf))
		return rc;
	k_fisher_eigen<<<n, 32, 0, sf>>>(s->g_mat, s->g_par, n, P, s->pp.pv2, s->alpha_fix ? 1 : 0, s->ppE_Nmod, s->g_vals, s->g_vecs, s->g_ok);
	SCUDA(ctx, cudaEventRecord(s->ev_g_done, sf));
	s->g_inflight = n;
	s->last_launches += 2;
	return 0;
}

int group_commit(gwat_b200_sampler *s, cudaStream_t st)
{
	if (s->g_inflight == 0) return 0;
	gwat_b200_ctx *ctx = s->ctx;
	SCUDA(ctx, cudaStreamWaitEvent(st, s->ev_g_done, 0));
	k_fisher_commit<<<s->g_inflight, 64, 0, st>>>(s->g_idx, s->g_ok, s->g_inflight, s->k.P, s->g_vals, s->g_vecs, s->d.fvals, s->d.fvecs,
	                                              s->d.counters);
	s->g_inflight = 0;
	s->last_launches += 1;
	return 0;
}

// Deferred mode, at a swap-sweep boundary (lanes joined on st[0]): install the pass started at the previous boundary, start
// one for the chains that came due since.
int deferred_boundary(gwat_b200_sampler *s)
{
	if (int rc = group_commit(s, s->st[0])) return rc;
	if (s->due.empty()) return 0;
	if (int rc = group_launch(s, s->due, s->st[0])) return rc;
	for (int c : s->due) s->due_mark[c] = 0;
	s->due.clear();
	return 0;
}

int swap_sweep(gwat_b200_sampler *s)
{
	gwat_b200_ctx *ctx = s->ctx;
	const int C = s->k.C, P = s->k.P;
	cudaStream_t st = s->st[0];
	// join: lane 1 -> lane 0
	if (s->nlanes > 1) {
		SCUDA(ctx, cudaEventRecord(s->ev_lane[1], s->st[1]));
		SCUDA(ctx, cudaStreamWaitEvent(st, s->ev_lane[1], 0));
	}
	double gate, unused;
	uniform2(s->k.seed, (uint64_t)s->sweep, 0u, DRAW_SWAP_GATE, gate, unused);
	if (s->comm && gate < s->opt.swap_rate) {
		// The one exchange of the path: all-gather of 8 (P + 2) bytes per chain over NVLink, on the sampler's stream;
		// then the reference's sequential sweep (src/mcmc_sampler_internals.cpp:1086-1184) over the gathered ladder, computed by
		// every rank from the same logL and the same counter-based draws, and each rank takes the records that land in its slots.
		const int Ct = C * s->n_ranks, R = P + 2, c0 = s->rank * C;
		const bool timed = s->n_sw_timed < gwat_b200_sampler::NSW;
		if (timed) SCUDA(ctx, cudaEventRecord(s->ev_sw0[s->n_sw_timed], st));
		k_swap_pack<<<(C * R + 255) / 256, 256, 0, st>>>(C, P, s->d.pos, s->d.ll, s->d.lp, s->x_send);
		const ncclResult_t nr = nccl_api().AllGather(s->x_send, s->x_recv, (size_t)C * R, ncclDouble, s->comm, st);
		if (nr != ncclSuccess)
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string("ncclAllGather: ") + nccl_api().GetErrorString(nr));
		k_swap_global_ll<<<(Ct + 255) / 256, 256, 0, st>>>(Ct, P, s->x_recv, s->g_ll);
		k_swap_prepare<<<(Ct + 255) / 256, 256, 0, st>>>(s->g_ll, s->g_temps, s->k.seed, s->sweep, Ct, 0, s->g_thr, s->g_kind);
		scan_smem_opt_in();
		k_swap_scan<<<1, kScanThreads, scan_smem_bytes(Ct), st>>>(s->g_ll, s->g_thr, s->g_kind, Ct, s->g_src, s->g_acc, s->g_work, 0);  // the WHOLE ladder, the same on every rank
		k_swap_count<<<(C + 255) / 256, 256, 0, st>>>(s->g_acc, Ct, c0, C, s->d.counters);
		k_swap_take<<<(C * R + 255) / 256, 256, 0, st>>>(s->g_src, c0, C, P, s->x_recv, s->d.pos, s->d.ll, s->d.lp);
		if (timed) SCUDA(ctx, cudaEventRecord(s->ev_sw1[s->n_sw_timed++], st));
		s->last_launches += 7;
		s->last_sweeps += 1;
	} else if (!s->comm && gate < s->opt.swap_rate && C > 1) {  // src/mcmc_sampler.cpp:4646-4654
		const bool timed = s->n_sw_timed < gwat_b200_sampler::NSW;
		if (timed) SCUDA(ctx, cudaEventRecord(s->ev_sw0[s->n_sw_timed], st));
		k_swap_prepare<<<(C + 255) / 256, 256, 0, st>>>(s->d.ll, s->d.temps, s->k.seed, s->sweep, C, s->k.chain_offset, s->swap_thr, s->swap_kind);
		scan_smem_opt_in();
		k_swap_scan<<<1, kScanThreads, scan_smem_bytes(C), st>>>(s->d.ll, s->swap_thr, s->swap_kind, C, s->swap_src, s->swap_acc, s->swap_work, 0);
		k_swap_count<<<(C + 255) / 256, 256, 0, st>>>(s->swap_acc, C, 0, C, s->d.counters);
		k_swap_apply<<<(C * P + 255) / 256, 256, 0, st>>>(s->swap_src, C, P, s->d.pos, s->d.ll, s->d.lp, s->pos2, s->ll2, s->lp2);
		std::swap(s->d.pos, s->pos2);
		std::swap(s->d.ll, s->ll2);
		std::swap(s->d.lp, s->lp2);
		if (timed) SCUDA(ctx, cudaEventRecord(s->ev_sw1[s->n_sw_timed++], st));
		s->last_launches += 4;
		s->last_sweeps += 1;
	}
	s->sweep += 1;
	if (s->deferred)
		if (int rc = deferred_boundary(s)) return rc;
	if (s->nlanes > 1) {
		SCUDA(ctx, cudaEventRecord(s->ev_join, st));
		SCUDA(ctx, cudaStreamWaitEvent(s->st[1], s->ev_join, 0));
	}
	SCUDA(ctx, cudaGetLastError());
	return 0;
}

}  // namespace

extern "C" {

void gwat_b200_prior_init(gwat_b200_prior *p)
{
	if (!p) return;
	std::memset(p, 0, sizeof(*p));
	const double big = 1e300;
	auto open = [&](double *b) { b[0] = -big; b[1] = big; };
	open(p->mass1_prior); open(p->mass2_prior); open(p->spin1_prior); open(p->spin2_prior); open(p->a1_prior); open(p->a2_prior);
	open(p->ctheta1_prior); open(p->ctheta2_prior); open(p->phi1_prior); open(p->phi2_prior); open(p->tidal1_prior);
	open(p->tidal2_prior); open(p->tidal_s_prior); open(p->RA_bounds); open(p->sinDEC_bounds); open(p->DL_prior);
	for (int i = 0; i < GWAT_B200_MAX_MOD; i++) open(p->mod_priors[i]);
	p->T_merger = 0;
	p->tidal_love = 1;
}

void gwat_b200_sampler_options_init(gwat_b200_sampler_options *o)
{
	if (!o) return;
	std::memset(o, 0, sizeof(*o));
	o->swp_freq = 5;
	o->swap_rate = 1. / o->swp_freq;
	o->history_length = 1000;
	o->history_update = 10;
	o->fisher_exist = 1;
	o->fisher_update_number = 200;
	o->fisher_deriv_order = 4;
	o->check_stepsize_freq = 50;
	o->seed = 1;
	o->lanes = 2;
}

int gwat_b200_log_prior_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                              const gwat_b200_prior *prior, const double *params, double *logP)
{
	if (!ctx || !prior || W < 0 || (W > 0 && (!params || !logP))) return GWAT_B200_ERR_ARG;
	if (W == 0) return GWAT_B200_OK;
	PriorPlan pp;
	MethodDesc desc;
	bool af;
	int nm;
	if (int rc = make_prior_plan(method, mod, dimension, pp, desc, af, nm))
		return gwat_internal::set_error(ctx, rc == -1 ? GWAT_B200_ERR_METHOD : GWAT_B200_ERR_ARG, "log_prior_batch: unknown method or dimension too small for it");
	gwat_b200_prior pd = *prior;
	pd.tidal_love = pp.tidal_love;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	double *d_par = nullptr, *d_out = nullptr;
	SCUDA(ctx, dalloc(d_par, (size_t)W * dimension));
	SCUDA(ctx, dalloc(d_out, (size_t)W));
	SCUDA(ctx, cudaMemcpyAsync(d_par, params, sizeof(double) * W * dimension, cudaMemcpyHostToDevice, ctx->stream));
	k_prior<<<(W + 127) / 128, 128, 0, ctx->stream>>>(d_par, W, pd, pp, d_out);
	ctx->launches += 1;
	SCUDA(ctx, cudaMemcpyAsync(logP, d_out, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
	SCUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(d_par);
	cudaFree(d_out);
	return GWAT_B200_OK;
}

int gwat_b200_mcmc_fisher_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int order, int W,
                                const double *params, double gmst, double *fisher, double *eigenvalues, double *eigenvectors)
{
	if (!ctx || W < 0 || (W > 0 && !params)) return GWAT_B200_ERR_ARG;
	if (W == 0) return GWAT_B200_OK;
	PriorPlan pp;
	MethodDesc desc;
	bool af;
	int nm;
	if (int rc = make_prior_plan(method, mod, dimension, pp, desc, af, nm))
		return gwat_internal::set_error(ctx, rc == -1 ? GWAT_B200_ERR_METHOD : GWAT_B200_ERR_ARG, "mcmc_fisher_batch: unknown method or dimension too small for it");
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const int P = dimension;
	double *d_par = nullptr, *d_mat = nullptr, *d_vals = nullptr, *d_vecs = nullptr;
	SCUDA(ctx, dalloc(d_par, (size_t)W * P));
	SCUDA(ctx, dalloc(d_mat, (size_t)W * P * P));
	SCUDA(ctx, dalloc(d_vals, (size_t)W * P));
	SCUDA(ctx, dalloc(d_vecs, (size_t)W * P * P));
	cudaStream_t st = ctx->stream;
	SCUDA(ctx, cudaMemcpyAsync(d_par, params, sizeof(double) * W * P, cudaMemcpyHostToDevice, st));
	int rc = gwat_internal::fisher_mcmc_dev(ctx, method, mod, P, order, W, d_par, gmst, d_mat, st);
	if (rc == 0) {
		SCUDA(ctx, cudaMemsetAsync(d_vals, 0xff, sizeof(double) * W * P, st));  // NaN where the decomposition fails
		SCUDA(ctx, cudaMemsetAsync(d_vecs, 0xff, sizeof(double) * W * P * P, st));
		k_fisher_eigen<<<W, 32, 0, st>>>(d_mat, d_par, W, P, pp.pv2, af ? 1 : 0, nm, d_vals, d_vecs, nullptr);
		ctx->launches += 1;
		if (fisher) SCUDA(ctx, cudaMemcpyAsync(fisher, d_mat, sizeof(double) * W * P * P, cudaMemcpyDeviceToHost, st));
		if (eigenvalues) SCUDA(ctx, cudaMemcpyAsync(eigenvalues, d_vals, sizeof(double) * W * P, cudaMemcpyDeviceToHost, st));
		if (eigenvectors) SCUDA(ctx, cudaMemcpyAsync(eigenvectors, d_vecs, sizeof(double) * W * P * P, cudaMemcpyDeviceToHost, st));
		SCUDA(ctx, cudaStreamSynchronize(st));
	}
	cudaFree(d_par);
	cudaFree(d_mat);
	cudaFree(d_vals);
	cudaFree(d_vecs);
	return rc;
}

void gwat_b200_sampler_destroy(gwat_b200_sampler *s)
{
	if (!s) return;
	cudaSetDevice(s->device);  // (a sampler must be destroyed before its context; this at least does not read the context)
	for (int i = 0; i < 2; i++) {
		if (s->st[i]) cudaStreamSynchronize(s->st[i]);
		if (s->st_like[i]) cudaStreamSynchronize(s->st_like[i]);
	}
	DevState &d = s->d;
	void *ptrs[] = {d.pos, d.prop, d.ll, d.lp, d.llprop, d.lpprop, d.temps, d.hist, d.hist_pos, d.fvals, d.fvecs, d.widths, d.counters,
	                d.gauss_ct, d.type_last, d.info, d.cold_slot, d.cold, s->pos2, s->ll2, s->lp2, s->swap_thr, s->swap_kind, s->swap_src, s->swap_acc, s->swap_work};
	for (void *p : ptrs) cudaFree(p);
	if (s->st_fisher) cudaStreamSynchronize(s->st_fisher);
	for (gwat_b200_sampler::RefreshLane &r : s->rf) {
		for (int *p : r.d_idx) cudaFree(p);
		for (int *p : r.d_ok) cudaFree(p);
		for (auto *v : {&r.d_par, &r.d_mat, &r.d_vals, &r.d_vecs})
			for (double *p : *v) cudaFree(p);
		for (auto *v : {&r.ev_gathered, &r.ev_done, &r.ev_h2d})
			for (cudaEvent_t e : *v)
				if (e) cudaEventDestroy(e);
		if (r.h_idx) cudaFreeHost(r.h_idx);
	}
	for (int i = 0; i < 2; i++) {
		if (s->ev_lane[i]) cudaEventDestroy(s->ev_lane[i]);
		if (s->st[i]) cudaStreamDestroy(s->st[i]);
		if (s->st_like[i]) cudaStreamDestroy(s->st_like[i]);
	}
	if (s->st_fisher) cudaStreamDestroy(s->st_fisher);
	for (cudaEvent_t e : {s->ev_join, s->ev_t0, s->ev_t1, s->ev_g_gathered, s->ev_g_done})
		if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : s->gh_ev)
		if (e) cudaEventDestroy(e);
	for (void *p : {(void *)s->g_idx, (void *)s->g_ok, (void *)s->g_par, (void *)s->g_mat, (void *)s->g_vals, (void *)s->g_vecs}) cudaFree(p);
	if (s->gh_idx) cudaFreeHost(s->gh_idx);
	for (void *p : {(void *)s->x_send, (void *)s->x_recv, (void *)s->g_ll, (void *)s->g_temps, (void *)s->g_thr, (void *)s->g_kind, (void *)s->g_src,
	                (void *)s->g_acc, (void *)s->g_work})
		cudaFree(p);
	for (int i = 0; i < gwat_b200_sampler::NSW; i++) {
		if (s->ev_sw0[i]) cudaEventDestroy(s->ev_sw0[i]);
		if (s->ev_sw1[i]) cudaEventDestroy(s->ev_sw1[i]);
	}
	if (s->comm) nccl_api().CommDestroy(s->comm);
	delete s;
}

int gwat_b200_sampler_create(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, const gwat_b200_sampler_options *options,
                             const gwat_b200_prior *prior, const double *chain_temps, const double *initial_positions, double gmst,
                             double T_segment, gwat_b200_sampler **out)
{
	if (!ctx || !out || !options || !prior || !chain_temps || !initial_positions) return GWAT_B200_ERR_ARG;
	*out = nullptr;
	const gwat_b200_sampler_options &o = *options;
	if (o.chain_N < 1 || o.swp_freq < 1 || o.history_length < 2 || o.history_update < 1 || o.check_stepsize_freq < 1 ||
	    o.fisher_update_number < 1 || (o.fisher_deriv_order != 2 && o.fisher_deriv_order != 4))
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "sampler_create: invalid options");
	gwat_b200_sampler *s = new gwat_b200_sampler;
	s->ctx = ctx;
	s->device = ctx->device;
	s->method = method ? method : "";
	if (mod) s->mod = *mod;
	else gwat_b200_mod_init(&s->mod);
	s->opt = o;
	s->prior = *prior;
	if (int rc = make_prior_plan(method, &s->mod, o.dimension, s->pp, s->desc, s->alpha_fix, s->ppE_Nmod)) {
		delete s;
		return gwat_internal::set_error(ctx, rc == -1 ? GWAT_B200_ERR_METHOD : GWAT_B200_ERR_ARG,
		                                "sampler_create: unknown generation_method, or dimension does not fit it");
	}
	s->prior.tidal_love = s->pp.tidal_love;
	s->gmst = gmst;
	s->T_segment = T_segment;
	const int C = o.chain_N, P = o.dimension, H = o.history_length;
	s->k = StepConst{o.seed, C, P, H, o.history_update, o.check_stepsize_freq, o.fisher_exist, o.record_cold, o.chain_index_offset};
	s->nlanes = (o.lanes >= 2 && C >= 2) ? 2 : 1;
	s->lane_c0[0] = 0;
	s->lane_n[0] = s->nlanes == 2 ? C / 2 : C;
	s->lane_c0[1] = s->lane_n[0];
	s->lane_n[1] = C - s->lane_n[0];
	s->h_temps.assign(chain_temps, chain_temps + C);
	s->h_fisher_ct.assign(C, o.fisher_update_number);  // :2011
	std::vector<int> cold_slot(C, -1);
	for (int c = 0; c < C; c++)
		if (chain_temps[c] == 1.0) cold_slot[c] = s->n_cold++;

	std::unique_lock<std::mutex> lock(ctx->mu);
	auto bail = [&](int rc) {
		lock.unlock();
		gwat_b200_sampler_destroy(s);
		return rc;
	};
#define SC_TRY(expr)                                                                                                       \
	do {                                                                                                                     \
		cudaError_t e_ = (expr);                                                                                               \
		if (e_ != cudaSuccess)                                                                                                 \
			return bail(gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_))); \
	} while (0)
	SC_TRY(cudaSetDevice(ctx->device));
	DevState &d = s->d;
	SC_TRY(dalloc(d.pos, (size_t)C * P));
	SC_TRY(dalloc(d.prop, (size_t)C * P));
	SC_TRY(dalloc(s->pos2, (size_t)C * P));
	for (double **p : {&d.ll, &d.lp, &d.llprop, &d.lpprop, &d.temps, &s->ll2, &s->lp2, &s->swap_thr}) SC_TRY(dalloc(*p, (size_t)C));
	SC_TRY(dalloc(d.hist, (size_t)C * H * P));
	SC_TRY(dalloc(d.hist_pos, (size_t)C));
	SC_TRY(dalloc(d.fvals, (size_t)C * P));
	SC_TRY(dalloc(d.fvecs, (size_t)C * P * P));
	SC_TRY(dalloc(d.widths, (size_t)C * (P + 3)));
	SC_TRY(dalloc(d.counters, (size_t)C * NCT));
	SC_TRY(dalloc(d.gauss_ct, (size_t)C * P * 4));
	SC_TRY(dalloc(d.type_last, (size_t)C * 4));
	SC_TRY(dalloc(d.info, (size_t)C));
	SC_TRY(dalloc(d.cold_slot, (size_t)C));
	SC_TRY(dalloc(s->swap_kind, (size_t)C));
	SC_TRY(dalloc(s->swap_src, (size_t)C));
	SC_TRY(dalloc(s->swap_acc, (size_t)C));
	SC_TRY(dalloc(s->swap_work, (size_t)4 * C));
	int prio_lo = 0, prio_hi = 0;
	SC_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	s->lookahead = 0;
	s->deferred = o.fisher_exist && o.fisher_deferred != 0;
	if (o.fisher_exist) {
		// the group staging set: every chain's first matrix (below), and the deferred passes
		const size_t n = (size_t)C;
		SC_TRY(dalloc(s->g_idx, n));
		SC_TRY(dalloc(s->g_ok, n));
		SC_TRY(dalloc(s->g_par, n * P));
		SC_TRY(dalloc(s->g_mat, n * P * P));
		SC_TRY(dalloc(s->g_vals, n * P));
		SC_TRY(dalloc(s->g_vecs, n * P * P));
		SC_TRY(cudaEventCreateWithFlags(&s->ev_g_gathered, cudaEventDisableTiming));
		SC_TRY(cudaEventCreateWithFlags(&s->ev_g_done, cudaEventDisableTiming));
		for (int k = 0; k < gwat_b200_sampler::GNP; k++) SC_TRY(cudaEventCreateWithFlags(&s->gh_ev[k], cudaEventDisableTiming));
		SC_TRY(cudaMallocHost((void **)&s->gh_idx, (size_t)gwat_b200_sampler::GNP * n * sizeof(int)));
		s->due_mark.assign(C, 0);
	}
	for (int i = 0; i < s->nlanes; i++) {
		// short latency-bound kernels on a high-priority stream, the FP64-bound bin kernel on a low-priority one: the other
		// lane's setup is then scheduled into the SM slots the bin kernel frees instead of waiting for its last CTA
		SC_TRY(cudaStreamCreateWithPriority(&s->st[i], cudaStreamNonBlocking, prio_hi));
		SC_TRY(cudaStreamCreateWithPriority(&s->st_like[i], cudaStreamNonBlocking, prio_lo));
		SC_TRY(cudaEventCreateWithFlags(&s->ev_lane[i], cudaEventDisableTiming));
		if (!o.fisher_exist || s->deferred) continue;
		gwat_b200_sampler::RefreshLane &r = s->rf[i];
		const size_t n = (size_t)s->lane_n[i];
		r.ND = s->lookahead + 1;
		r.NP = s->lookahead + 33;
		r.n_in_slot.assign(r.ND, 0);
		r.h_used.assign(r.NP, 0);
		for (int k = 0; k < r.ND; k++) {
			int *pi = nullptr, *po = nullptr;
			double *a = nullptr, *b = nullptr, *c = nullptr, *e = nullptr;
			SC_TRY(dalloc(pi, n));
			r.d_idx.push_back(pi);
			SC_TRY(dalloc(po, n));
			r.d_ok.push_back(po);
			SC_TRY(dalloc(a, n * P));
			r.d_par.push_back(a);
			SC_TRY(dalloc(b, n * P * P));
			r.d_mat.push_back(b);
			SC_TRY(dalloc(c, n * P));
			r.d_vals.push_back(c);
			SC_TRY(dalloc(e, n * P * P));
			r.d_vecs.push_back(e);
			cudaEvent_t e1 = nullptr, e2 = nullptr;
			SC_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
			r.ev_gathered.push_back(e1);
			SC_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
			r.ev_done.push_back(e2);
		}
		for (int k = 0; k < r.NP; k++) {
			cudaEvent_t e1 = nullptr;
			SC_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
			r.ev_h2d.push_back(e1);
		}
		SC_TRY(cudaMallocHost((void **)&r.h_idx, std::max<size_t>(1, (size_t)r.NP * n) * sizeof(int)));
	}
	SC_TRY(cudaStreamCreateWithPriority(&s->st_fisher, cudaStreamNonBlocking, prio_hi));
	SC_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
	SC_TRY(cudaEventCreate(&s->ev_t0));
	SC_TRY(cudaEventCreate(&s->ev_t1));
	for (int i = 0; i < gwat_b200_sampler::NSW; i++) {
		SC_TRY(cudaEventCreate(&s->ev_sw0[i]));
		SC_TRY(cudaEventCreate(&s->ev_sw1[i]));
	}
	cudaStream_t st = s->st[0];
	SC_TRY(cudaMemcpyAsync(d.pos, initial_positions, sizeof(double) * C * P, cudaMemcpyHostToDevice, st));
	SC_TRY(cudaMemcpyAsync(d.temps, chain_temps, sizeof(double) * C, cudaMemcpyHostToDevice, st));
	SC_TRY(cudaMemcpyAsync(d.cold_slot, cold_slot.data(), sizeof(int) * C, cudaMemcpyHostToDevice, st));
	k_init<<<(C + 127) / 128, 128, 0, st>>>(d, s->k, s->prior, s->pp);
	if (int rc = gwat_internal::loglike_mcmc_lane(ctx, 1, s->method.c_str(), &s->mod, P, C, d.pos, gmst, T_segment, d.ll, st)) return bail(rc);
	SC_TRY(cudaStreamSynchronize(st));
	std::vector<double> lp(C), ll(C);
	SC_TRY(cudaMemcpy(lp.data(), d.lp, sizeof(double) * C, cudaMemcpyDeviceToHost));
	SC_TRY(cudaMemcpy(ll.data(), d.ll, sizeof(double) * C, cudaMemcpyDeviceToHost));
	for (int c = 0; c < C; c++)
		if (!(lp[c] > -INFINITY) || !(ll[c] == ll[c]))
			return bail(gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "sampler_create: initial position of chain " + std::to_string(c) +
			                                                                 " has zero prior or an undefined likelihood"));
	if (o.fisher_exist) {
		// every chain starts with the eigen-system of its initial position and a refresh counter of zero, as assign_initial_pos
		// leaves them (src/mcmc_sampler_internals.cpp:3182-3212)
		std::vector<int> all(C);
		for (int c = 0; c < C; c++) all[c] = c;
		if (int rc = group_launch(s, all, st)) return bail(rc);
		if (int rc = group_commit(s, st)) return bail(rc);
		SC_TRY(cudaStreamSynchronize(st));
		std::fill(s->h_fisher_ct.begin(), s->h_fisher_ct.end(), 0);
	}
#undef SC_TRY
	lock.unlock();
	*out = s;
	return GWAT_B200_OK;
}

int gwat_b200_sampler_run(gwat_b200_sampler *s, int n_steps)
{
	if (!s || n_steps < 0) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	if (int rc = ensure_cold(s, s->step + n_steps)) return rc;
	const int P = s->k.P;
	const long long launches0 = ctx->launches;
	s->last_launches = 0;
	s->n_sw_timed = 0;
	s->last_sweeps = 0;
	s->last_swap_ms = 0;
	SCUDA(ctx, cudaEventRecord(s->ev_t0, s->st[0]));
	if (s->nlanes > 1) {
		SCUDA(ctx, cudaEventRecord(s->ev_join, s->st[0]));
		SCUDA(ctx, cudaStreamWaitEvent(s->st[1], s->ev_join, 0));
	}
	std::vector<int> flagged;
	for (int it = 0; it < n_steps; it++) {
		const long long step = s->step;
		const long long cold_row = (s->opt.record_cold && s->n_cold) ? step * s->n_cold : -1;
		for (int ln = 0; ln < s->nlanes; ln++) {
			const int c0 = s->lane_c0[ln], n = s->lane_n[ln];
			if (n == 0) continue;
			cudaStream_t st = s->st[ln];
			if (s->opt.fisher_exist && s->deferred) {
				fisher_schedule(s, ln, step, flagged);
				for (int c : flagged)
					if (!s->due_mark[c]) {
						s->due_mark[c] = 1;
						s->due.push_back(c);
					}
			} else if (s->opt.fisher_exist) {
				// the reference's schedule: a chain's Fisher matrix is recomputed at the step that first uses it
				gwat_b200_sampler::RefreshLane &r = s->rf[ln];
				while (r.next_sched <= step + s->lookahead) {
					if (int rc = refresh_schedule(s, ln, r.next_sched, flagged)) return rc;
					r.next_sched += 1;
				}
				if (int rc = refresh_commit(s, ln, step)) return rc;
			}
			k_propose<<<(n + 63) / 64, 64, 0, st>>>(s->d, s->k, s->prior, s->pp, step, c0, n);
			if (int rc = gwat_internal::loglike_mcmc_lane(ctx, 1 + ln, s->method.c_str(), &s->mod, P, n, s->d.prop + (size_t)c0 * P, s->gmst,
			                                              s->T_segment, s->d.llprop + c0, st, s->st_like[ln]))
				return rc;
			k_accept<<<(n + 63) / 64, 64, 0, st>>>(s->d, s->k, step, c0, n, cold_row);
			s->last_launches += 2;
		}
		s->step += 1;
		s->since_swap += 1;
		if (s->since_swap == s->opt.swp_freq) {
			if (int rc = swap_sweep(s)) return rc;
			s->since_swap = 0;
		}
	}
	if (s->nlanes > 1) {
		SCUDA(ctx, cudaEventRecord(s->ev_lane[1], s->st[1]));
		SCUDA(ctx, cudaStreamWaitEvent(s->st[0], s->ev_lane[1], 0));
	}
	SCUDA(ctx, cudaEventRecord(s->ev_t1, s->st[0]));
	SCUDA(ctx, cudaStreamSynchronize(s->st[0]));
	// refreshes scheduled ahead for steps of the next call: let them finish, they use the context's Fisher scratch
	if (s->st_fisher) SCUDA(ctx, cudaStreamSynchronize(s->st_fisher));
	SCUDA(ctx, cudaGetLastError());
	float ms = 0;
	SCUDA(ctx, cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
	s->last_ms = ms;
	s->last_launches += ctx->launches - launches0;
	for (int i = 0; i < s->n_sw_timed; i++) {
		float t = 0;
		SCUDA(ctx, cudaEventElapsedTime(&t, s->ev_sw0[i], s->ev_sw1[i]));
		s->last_swap_ms += t;
	}
	if (s->n_sw_timed > 0) s->last_swap_ms /= s->n_sw_timed;
	return GWAT_B200_OK;
}

int gwat_b200_sampler_state(gwat_b200_sampler *s, double *positions, double *logL, double *logP)
{
	if (!s) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const int C = s->k.C, P = s->k.P;
	if (positions) SCUDA(ctx, cudaMemcpy(positions, s->d.pos, sizeof(double) * C * P, cudaMemcpyDeviceToHost));
	if (logL) SCUDA(ctx, cudaMemcpy(logL, s->d.ll, sizeof(double) * C, cudaMemcpyDeviceToHost));
	if (logP) SCUDA(ctx, cudaMemcpy(logP, s->d.lp, sizeof(double) * C, cudaMemcpyDeviceToHost));
	return GWAT_B200_OK;
}

int gwat_b200_sampler_set_state(gwat_b200_sampler *s, const double *positions, const double *logL, const double *logP)
{
	if (!s) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const int C = s->k.C, P = s->k.P;
	if (positions) SCUDA(ctx, cudaMemcpy(s->d.pos, positions, sizeof(double) * C * P, cudaMemcpyHostToDevice));
	if (logL) SCUDA(ctx, cudaMemcpy(s->d.ll, logL, sizeof(double) * C, cudaMemcpyHostToDevice));
	if (logP) SCUDA(ctx, cudaMemcpy(s->d.lp, logP, sizeof(double) * C, cudaMemcpyHostToDevice));
	return GWAT_B200_OK;
}

void gwat_b200_sampler_uniform(unsigned long long seed, unsigned long long step, unsigned chain, unsigned purpose, double *out2)
{
	uniform2(seed, step, chain, purpose, out2[0], out2[1]);
}

int gwat_b200_swap_sweep_host(int C, const double *logL, const double *temps, unsigned long long seed, long long sweep, int *src,
                              int *accepted)
{
	if (C < 1 || !logL || !temps || !src) return GWAT_B200_ERR_ARG;
	std::vector<double> thr(C);
	std::vector<int> kind(C), acc(C);
	for (int i = 0; i < C - 1; i++) {
		double alpha, unused;
		uniform2(seed, (uint64_t)sweep, (uint32_t)i, DRAW_SWAP, alpha, unused);
		swap_threshold(logL[i + 1], temps[i], temps[i + 1], alpha, kind[i], thr[i]);
	}
	swap_scan(logL, thr.data(), kind.data(), C, src, acc.data());
	if (accepted)
		for (int i = 0; i < C - 1; i++) accepted[i] = acc[i];
	return GWAT_B200_OK;
}

// The device sweep on caller-supplied inputs (thresholds + k_swap_scan), for tests against gwat_b200_swap_sweep_host: mode 0 = as the
// sampler runs it, 1 = the sequential fallback forced.
int gwat_b200_swap_sweep_device(gwat_b200_ctx *ctx, int C, const double *logL, const double *temps, unsigned long long seed, long long sweep,
                                int mode, int *src, int *accepted)
{
	if (!ctx || C < 2 || !logL || !temps || !src) return GWAT_B200_ERR_ARG;
	double *d_ll = nullptr, *d_t = nullptr, *d_thr = nullptr;
	int *d_kind = nullptr, *d_src = nullptr, *d_acc = nullptr, *d_work = nullptr;
	auto release = [&]() {
		for (void *p : {(void *)d_ll, (void *)d_t, (void *)d_thr, (void *)d_kind, (void *)d_src, (void *)d_acc, (void *)d_work}) cudaFree(p);
	};
#define SW_TRY(x)                                                                   \
	do {                                                                            \
		cudaError_t e_ = (x);                                                       \
		if (e_ != cudaSuccess) {                                                    \
			release();                                                              \
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string("swap_sweep_device: ") + cudaGetErrorString(e_)); \
		}                                                                           \
	} while (0)
	SW_TRY(cudaSetDevice(ctx->device));
	SW_TRY(cudaMalloc((void **)&d_ll, sizeof(double) * C));
	SW_TRY(cudaMalloc((void **)&d_t, sizeof(double) * C));
	SW_TRY(cudaMalloc((void **)&d_thr, sizeof(double) * C));
	SW_TRY(cudaMalloc((void **)&d_kind, sizeof(int) * C));
	SW_TRY(cudaMalloc((void **)&d_src, sizeof(int) * C));
	SW_TRY(cudaMalloc((void **)&d_acc, sizeof(int) * C));
	SW_TRY(cudaMalloc((void **)&d_work, sizeof(int) * 4 * (size_t)C));
	SW_TRY(cudaMemcpy(d_ll, logL, sizeof(double) * C, cudaMemcpyHostToDevice));
	SW_TRY(cudaMemcpy(d_t, temps, sizeof(double) * C, cudaMemcpyHostToDevice));
	SW_TRY(cudaMemset(d_src, 0xff, sizeof(int) * C));
	SW_TRY(cudaMemset(d_acc, 0xff, sizeof(int) * C));
	k_swap_prepare<<<(C + 255) / 256, 256>>>(d_ll, d_t, seed, sweep, C, 0, d_thr, d_kind);
	scan_smem_opt_in();
	k_swap_scan<<<1, kScanThreads, scan_smem_bytes(C)>>>(d_ll, d_thr, d_kind, C, d_src, d_acc, d_work, mode);
	SW_TRY(cudaGetLastError());
	SW_TRY(cudaDeviceSynchronize());
	SW_TRY(cudaMemcpy(src, d_src, sizeof(int) * C, cudaMemcpyDeviceToHost));
	if (accepted) SW_TRY(cudaMemcpy(accepted, d_acc, sizeof(int) * (C - 1), cudaMemcpyDeviceToHost));
#undef SW_TRY
	release();
	return GWAT_B200_OK;
}

int gwat_b200_sampler_counters(gwat_b200_sampler *s, long long *counters, double *widths)
{
	if (!s) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const int C = s->k.C, P = s->k.P;
	if (counters) SCUDA(ctx, cudaMemcpy(counters, s->d.counters, sizeof(long long) * C * NCT, cudaMemcpyDeviceToHost));
	if (widths) SCUDA(ctx, cudaMemcpy(widths, s->d.widths, sizeof(double) * C * (P + 3), cudaMemcpyDeviceToHost));
	return GWAT_B200_OK;
}

int gwat_b200_sampler_fisher_state(gwat_b200_sampler *s, double *eigenvalues, double *eigenvectors)
{
	if (!s) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const int C = s->k.C, P = s->k.P;
	if (eigenvalues) SCUDA(ctx, cudaMemcpy(eigenvalues, s->d.fvals, sizeof(double) * C * P, cudaMemcpyDeviceToHost));
	if (eigenvectors) SCUDA(ctx, cudaMemcpy(eigenvectors, s->d.fvecs, sizeof(double) * C * P * P, cudaMemcpyDeviceToHost));
	return GWAT_B200_OK;
}

int gwat_b200_sampler_cold(gwat_b200_sampler *s, long long first_step, int n, double *out, int *n_cold)
{
	if (!s) return GWAT_B200_ERR_ARG;
	if (n_cold) *n_cold = s->n_cold;
	if (!out) return GWAT_B200_OK;
	gwat_b200_ctx *ctx = s->ctx;
	if (!s->opt.record_cold) return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "sampler_cold: the sampler was created without record_cold");
	if (first_step < 0 || n < 0 || first_step + n > s->step) return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "sampler_cold: step range not recorded");
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	const size_t row = (size_t)s->n_cold * s->k.P;
	if (n > 0 && row > 0)
		SCUDA(ctx, cudaMemcpy(out, s->d.cold + (size_t)first_step * row, sizeof(double) * n * row, cudaMemcpyDeviceToHost));
	return GWAT_B200_OK;
}

// ---- dynamic temperature allocation (arXiv:1501.05823) --------------------------------------------------------------------------
// update_temperatures_full_ensemble, linear-swapping branch (src/mcmc_sampler_internals.cpp:3371-3413): chains at T = 1 and at the
// hottest temperature of an ensemble stay put, the others move so that neighbouring swap acceptances even out.
//   A[i] = 1 / 0: the last swap attempt between chains i - 1 and i was accepted / rejected (chain_swap, :1095-1112); A[0] is never written.
int gwat_b200_update_temperatures(int chain_N, double *chain_temps, const double *A, int t0, int nu, int t)
{
	if (chain_N < 1 || !chain_temps || !A || nu == 0) return GWAT_B200_ERR_ARG;
	const double thresh = 1e-10;  // DOUBLE_COMP_THRESH (include/gwat/util.h)
	std::vector<double> old_temps(chain_N, 0.0);
	double max_temp = 0;
	int ensemble_chain_number = 0;
	bool search = true;
	for (int i = 0; i < chain_N - 1; i++) {
		old_temps[i] = chain_temps[i];
		if (i != 0 && std::fabs(chain_temps[i] - 1) < thresh) {
			max_temp = chain_temps[i - 1];
			if (search) {
				ensemble_chain_number = i;
				search = false;
			}
		}
	}
	if (max_temp < thresh) {
		max_temp = chain_temps[chain_N - 1];
		ensemble_chain_number = chain_N;
	}
	(void)ensemble_chain_number;  // (only read by a branch the loop bounds below never reach, :3405-3408)
	const double kappa = (1. / nu) * (double)(t0) / (t + t0);  // PT_dynamical_timescale (:3224-3230)
	for (int i = 1; i < chain_N - 1; i++) {
		if (!(std::fabs(chain_temps[i] - 1) < thresh || std::fabs(chain_temps[i] - max_temp) < thresh)) {
			const double power = kappa * (A[i] - A[i + 1]);
			chain_temps[i] = chain_temps[i - 1] + (old_temps[i] - old_temps[i - 1]) * std::exp(power);
		}
	}
	return GWAT_B200_OK;
}

int gwat_b200_sampler_set_temperatures(gwat_b200_sampler *s, const double *chain_temps)
{
	if (!s || !chain_temps) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	if (s->comm) return gwat_internal::set_error(ctx, GWAT_B200_ERR_UNSUPPORTED, "sampler_set_temperatures: not for a sampler sharded over ranks");
	const int C = s->k.C;
	for (int c = 0; c < C; c++)
		if ((chain_temps[c] == 1.0) != (s->h_temps[c] == 1.0) || !(chain_temps[c] >= 1.0))
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "sampler_set_temperatures: the T = 1 chains must stay the T = 1 chains, and T >= 1");
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	SCUDA(ctx, cudaMemcpy(s->d.temps, chain_temps, sizeof(double) * C, cudaMemcpyHostToDevice));
	s->h_temps.assign(chain_temps, chain_temps + C);
	return GWAT_B200_OK;
}

int gwat_b200_sampler_temperatures(gwat_b200_sampler *s, double *chain_temps)
{
	if (!s || !chain_temps) return GWAT_B200_ERR_ARG;
	std::memcpy(chain_temps, s->h_temps.data(), sizeof(double) * s->h_temps.size());
	return GWAT_B200_OK;
}

int gwat_b200_sampler_last_swap_accepts(gwat_b200_sampler *s, int *accepted)
{
	if (!s || !accepted) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	if (s->comm) return gwat_internal::set_error(ctx, GWAT_B200_ERR_UNSUPPORTED, "sampler_last_swap_accepts: not for a sampler sharded over ranks");
	if (s->k.C < 2) return GWAT_B200_OK;
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	SCUDA(ctx, cudaMemcpy(accepted, s->swap_acc, sizeof(int) * (s->k.C - 1), cudaMemcpyDeviceToHost));
	return GWAT_B200_OK;
}

// dynamic_temperature_full_ensemble_internal with linear swapping (src/mcmc_sampler.cpp:453-545): blocks of swp_freq steps, one
// sweep over the whole ladder after each (the reference switches its probabilistic gate off and calls chain_swap itself; here
// the gate is held open for the duration), then the temperature update with t = steps taken so far.
int gwat_b200_sampler_dynamic_temperatures(gwat_b200_sampler *s, int N_steps, int nu, int t0, long long *sweeps_done)
{
	if (!s || N_steps < 0 || nu == 0) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	if (s->comm) return gwat_internal::set_error(ctx, GWAT_B200_ERR_UNSUPPORTED, "sampler_dynamic_temperatures: not for a sampler sharded over ranks");
	if (s->since_swap != 0) return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "sampler_dynamic_temperatures: start at a swap boundary (steps run so far must be a multiple of swp_freq)");
	const int C = s->k.C, steps = s->opt.swp_freq;
	const double swap_rate_saved = s->opt.swap_rate;
	s->opt.swap_rate = 2.0;  // every sweep happens
	std::vector<double> A(C + 1, 0.0), temps(C);
	std::vector<int> acc(C > 1 ? C - 1 : 1, 0);
	int t = 0, rc = GWAT_B200_OK;
	long long n = 0;
	while (t < N_steps - steps && rc == GWAT_B200_OK) {
		rc = gwat_b200_sampler_run(s, steps);
		t += steps;
		if (rc == GWAT_B200_OK) rc = gwat_b200_sampler_last_swap_accepts(s, acc.data());
		if (rc != GWAT_B200_OK) break;
		for (int i = 0; i + 1 < C; i++) A[i + 1] = acc[i] ? 1.0 : 0.0;
		temps = s->h_temps;
		gwat_b200_update_temperatures(C, temps.data(), A.data(), t0, nu, t);
		rc = gwat_b200_sampler_set_temperatures(s, temps.data());
		n++;
	}
	s->opt.swap_rate = swap_rate_saved;
	if (sweeps_done) *sweeps_done = n;
	return rc;
}

int gwat_b200_nccl_unique_id(unsigned char *id128)
{
	if (!id128) return GWAT_B200_ERR_ARG;
	NcclApi &api = nccl_api();
	if (!api.error.empty()) return gwat_internal::set_error(nullptr, GWAT_B200_ERR_UNSUPPORTED, api.error);
	static_assert(sizeof(ncclUniqueId) == GWAT_B200_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
	ncclUniqueId id;
	const ncclResult_t r = api.GetUniqueId(&id);
	if (r != ncclSuccess) return gwat_internal::set_error(nullptr, GWAT_B200_ERR_CUDA, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
	std::memcpy(id128, &id, sizeof(id));
	return GWAT_B200_OK;
}

int gwat_b200_sampler_attach_ranks(gwat_b200_sampler *s, const unsigned char *id128, int rank, int n_ranks)
{
	if (!s || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return GWAT_B200_ERR_ARG;
	gwat_b200_ctx *ctx = s->ctx;
	if (s->comm) return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "sampler_attach_ranks: already attached");
	if (s->step != 0) return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "sampler_attach_ranks: attach before the first step");
	if (s->k.chain_offset != rank * s->k.C)
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "sampler_attach_ranks: options.chain_index_offset must be rank * chain_N (equal shards)");
	NcclApi &api = nccl_api();
	if (!api.error.empty()) return gwat_internal::set_error(ctx, GWAT_B200_ERR_UNSUPPORTED, api.error);
	std::lock_guard<std::mutex> lock(ctx->mu);
	SCUDA(ctx, cudaSetDevice(ctx->device));
	ncclUniqueId id;
	std::memcpy(&id, id128, sizeof(id));
	ncclComm_t comm = nullptr;
	ncclResult_t r = api.CommInitRank(&comm, n_ranks, id, rank);
	if (r != ncclSuccess) return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
	const int C = s->k.C, P = s->k.P, Ct = C * n_ranks, R = P + 2;
	cudaStream_t st = s->st[0];
	auto fail_here = [&](const std::string &m) {
		api.CommDestroy(comm);
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, m);
	};
#define AT_TRY(expr)                                                                                  \
	do {                                                                                                \
		cudaError_t e_ = (expr);                                                                          \
		if (e_ != cudaSuccess) return fail_here(std::string(#expr) + ": " + cudaGetErrorString(e_));      \
	} while (0)
	AT_TRY(dalloc(s->x_send, (size_t)C * R));
	AT_TRY(dalloc(s->x_recv, (size_t)Ct * R));
	AT_TRY(dalloc(s->g_ll, (size_t)Ct));
	AT_TRY(dalloc(s->g_temps, (size_t)Ct));
	AT_TRY(dalloc(s->g_thr, (size_t)Ct));
	AT_TRY(dalloc(s->g_kind, (size_t)Ct));
	AT_TRY(dalloc(s->g_src, (size_t)Ct));
	AT_TRY(dalloc(s->g_acc, (size_t)Ct));
	AT_TRY(dalloc(s->g_work, (size_t)4 * Ct));
	// the whole ladder's temperatures, once
	r = api.AllGather(s->d.temps, s->g_temps, (size_t)C, ncclDouble, comm, st);
	if (r != ncclSuccess) return fail_here(std::string("ncclAllGather(temperatures): ") + api.GetErrorString(r));
	AT_TRY(cudaStreamSynchronize(st));
#undef AT_TRY
	s->comm = comm;
	s->rank = rank;
	s->n_ranks = n_ranks;
	return GWAT_B200_OK;
}

double gwat_b200_sampler_last_swap_ms(const gwat_b200_sampler *s) { return s ? s->last_swap_ms : 0; }
long long gwat_b200_sampler_last_sweeps(const gwat_b200_sampler *s) { return s ? s->last_sweeps : 0; }

double gwat_b200_sampler_last_ms(const gwat_b200_sampler *s) { return s ? s->last_ms : 0; }
long long gwat_b200_sampler_last_launches(const gwat_b200_sampler *s) { return s ? s->last_launches : 0; }

}  // extern "C"
