// tc/phic-maximised log-likelihoods of the reference's "intrinsic" samplers, batched.
//
//   maximized_Log_Likelihood_aligned_spin_internal     src/mcmc_gw.cpp:595-652     0.5 max_t |FFT(4 conj(d) r / S)|^2 df^2 / (r|r)
//   maximized_Log_Likelihood_unaligned_spin_internal   src/mcmc_gw.cpp:660-795     (arXiv:1603.02444) lambda(t) from rho_+, rho_x, I_+x
//   callers: MCMC_likelihood_wrapper, intrinsic branches src/mcmc_gw.cpp:2603-2722 (psi = 0, phiRef = 1, iota = 0, tc = 1,
//            f_ref = 10 (PhenomD) or 20 (PhenomP); horizon response with theta = phi = psi = 0, i.e. F+ = 1, Fx = 0)
//
// One pass = polarisations of a chunk of walkers (engine kernel) -> norms (k_max_norms) -> FFT inputs (k_max_fill) ->
// batched cuFFT Z2Z forward of length L, one transform per (walker, detector[, polarisation]) -> maximum over time
// (k_max_reduce).  cuFFT is a plain library FFT here, as FFTW is in the reference.
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "gwat_engine_internal.h"
#include "gwat_method.h"

using namespace gwat;

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double block_reduce(double v, bool take_max)
{
	__shared__ double sh[kT / 32];
	for (int o = 16; o > 0; o >>= 1) {
		const double other = __shfl_down_sync(0xffffffffu, v, o);
		v = take_max ? fmax(v, other) : v + other;
	}
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) sh[wid] = v;
	__syncthreads();
	if (wid == 0) {
		v = lane < kT / 32 ? sh[lane] : (take_max ? -INFINITY : 0.0);
		for (int o = 16; o > 0; o >>= 1) {
			const double other = __shfl_down_sync(0xffffffffu, v, o);
			v = take_max ? fmax(v, other) : v + other;
		}
	}
	return v;  // valid in thread 0
}

// norms[(w*D+d)*3 + {0,1,2}] = (h+|h+), (hx|hx), (h+|hx) with the Fisher/SNR convention of Simpson's rule (delta_f = f[1]-f[0])
__global__ void __launch_bounds__(kT) k_max_norms(const double *__restrict__ hp_re, const double *__restrict__ hp_im,
                                                 const double *__restrict__ hc_re, const double *__restrict__ hc_im,
                                                 const double *__restrict__ wq, int ld, int L, int D, double prefactor,
                                                 double *__restrict__ norms)
{
	const int w = blockIdx.x, d = blockIdx.y;
	const double *q = wq + (size_t)d * ld;
	const size_t base = (size_t)w * L;
	double pp = 0, cc = 0, pc = 0;
	for (int i = threadIdx.x; i < L; i += kT) {
		const double a = hp_re[base + i], b = hp_im[base + i], c = hc_re[base + i], e = hc_im[base + i];
		pp += q[i] * (a * a + b * b);
		cc += q[i] * (c * c + e * e);
		pc += q[i] * (a * c + b * e);
	}
	pp = block_reduce(pp, false);
	cc = block_reduce(cc, false);
	pc = block_reduce(pc, false);
	if (threadIdx.x == 0) {
		double *o = norms + ((size_t)w * D + d) * 3;
		o[0] = prefactor * pp;
		o[1] = prefactor * cc;
		o[2] = prefactor * pc;
	}
}

// in[(w*D+d)*npol + pol][i] = 4 conj(data_d[i]) h_pol[i] / S_d[i] / sqrt((h_pol|h_pol))   (normalisation only when npol == 2)
__global__ void __launch_bounds__(kT) k_max_fill(const double *__restrict__ hp_re, const double *__restrict__ hp_im,
                                                const double *__restrict__ hc_re, const double *__restrict__ hc_im,
                                                const double *__restrict__ wq, const double *__restrict__ dre,
                                                const double *__restrict__ dim, int ld, int L, int D, int npol,
                                                const double *__restrict__ norms, cufftDoubleComplex *__restrict__ in)
{
	const int i = blockIdx.x * kT + threadIdx.x;
	if (i >= L) return;
	const int w = blockIdx.y, d = blockIdx.z;
	const double coef = (i == 0 || i == L - 1) ? 1.0 : ((i % 2 == 0) ? 2.0 : 4.0);
	const double inv_psd = wq[(size_t)d * ld + i] / coef;  // exact: the table holds coef / S
	const double xr = dre[(size_t)d * ld + i], xi = dim[(size_t)d * ld + i];
	const double *n = norms + ((size_t)w * D + d) * 3;
	for (int pol = 0; pol < npol; pol++) {
		double hr = (pol == 0 ? hp_re : hc_re)[(size_t)w * L + i], hi = (pol == 0 ? hp_im : hc_im)[(size_t)w * L + i];
		if (npol == 2) {
			const double root = sqrt(n[pol]);
			hr = hr / root;
			hi = hi / root;
		}
		// 4 * conj(d) * h / S
		const double gr = 4. * (xr * hr + xi * hi) * inv_psd, gi = 4. * (xr * hi - xi * hr) * inv_psd;
		in[(((size_t)w * D + d) * npol + pol) * L + i] = cufftDoubleComplex{gr, gi};
	}
}

// out[w] = sum_d (aligned: 0.5 max_i |G_i|^2 df^2 / (r|r);  unaligned: 0.25 max_i lambda_i)
__global__ void __launch_bounds__(kT) k_max_reduce(const cufftDoubleComplex *__restrict__ G, int L, int D, int npol, double df,
                                                  const double *__restrict__ norms, double *__restrict__ out)
{
	const int w = blockIdx.x;
	double total = 0;
	for (int d = 0; d < D; d++) {
		const double *n = norms + ((size_t)w * D + d) * 3;
		const cufftDoubleComplex *gp = G + (((size_t)w * D + d) * npol) * L, *gc = gp + L;
		double best = -INFINITY;
		if (npol == 1) {
			for (int i = threadIdx.x; i < L; i += kT) best = fmax(best, gp[i].x * gp[i].x + gp[i].y * gp[i].y);
		} else {
			const double Ipc = n[2] / (sqrt(n[0]) * sqrt(n[1]));
			for (int i = threadIdx.x; i < L; i += kT) {
				const double rp2 = df * df * (gp[i].x * gp[i].x + gp[i].y * gp[i].y);
				const double rc2 = df * df * (gc[i].x * gc[i].x + gc[i].y * gc[i].y);
				const double gam = (df * gp[i].x) * (df * gc[i].x) + (df * gp[i].y) * (df * gc[i].y);
				const double lam = (rp2 + rc2 - 2 * gam * Ipc +
				                    sqrt((rp2 - rc2) * (rp2 - rc2) + 4. * (Ipc * rp2 - gam) * (Ipc * rc2 - gam))) /
				                   (1. - Ipc * Ipc);
				best = fmax(best, lam);
			}
		}
		best = block_reduce(best, true);
		if (threadIdx.x == 0) total += (npol == 1) ? .5 * (best * df * df) / n[0] : .25 * best;
	}
	if (threadIdx.x == 0) out[w] = total;
}

// ---- match (src/waveform_util.cpp:41-89) ------------------------------------------------------------------------------------
// in[i] = conj(d1[i]) d2[i] / S[i];  norms: 4 Simpson sums of |d1|^2 / S and |d2|^2 / S (data_snr, :108-127)
__global__ void __launch_bounds__(kT) k_match_fill(int L, const double *__restrict__ a_re, const double *__restrict__ a_im,
                                                  const double *__restrict__ b_re, const double *__restrict__ b_im,
                                                  const double *__restrict__ psd, cufftDoubleComplex *__restrict__ in,
                                                  double *__restrict__ norms)
{
	double n1 = 0, n2 = 0;
	for (int i = threadIdx.x; i < L; i += kT) {
		const double ar = a_re[i], ai = a_im[i], br = b_re[i], bi = b_im[i], s = psd[i];
		// conj(a) * b
		in[i] = cufftDoubleComplex{(ar * br + ai * bi) / s, (ar * bi - ai * br) / s};
		const double coef = (i == 0 || i == L - 1) ? 1.0 : ((i % 2 == 0) ? 2.0 : 4.0);
		n1 += coef * (4. * (ar * ar + ai * ai) / s);
		n2 += coef * (4. * (br * br + bi * bi) / s);
	}
	n1 = block_reduce(n1, false);
	n2 = block_reduce(n2, false);
	if (threadIdx.x == 0) {
		norms[0] = n1;
		norms[1] = n2;
	}
}
__global__ void __launch_bounds__(kT) k_match_max(int L, const cufftDoubleComplex *__restrict__ G, double *__restrict__ out)
{
	double best = -INFINITY;
	for (int i = threadIdx.x; i < L; i += kT) best = fmax(best, sqrt(G[i].x * G[i].x + G[i].y * G[i].y));
	best = block_reduce(best, true);
	if (threadIdx.x == 0) out[2] = best;
}

#define MCUDA(ctx, expr)                                                                                         \
	do {                                                                                                           \
		cudaError_t e_ = (expr);                                                                                     \
		if (e_ != cudaSuccess)                                                                                       \
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
	} while (0)

struct Scratch {
	cufftHandle plan = 0;
	bool have_plan = false;
	cufftDoubleComplex *buf = nullptr;
	double *norms = nullptr, *out = nullptr;
	~Scratch()
	{
		if (have_plan) cufftDestroy(plan);
		cudaFree(buf);
		cudaFree(norms);
		cudaFree(out);
	}
};

}  // namespace

extern "C" int gwat_b200_loglike_maximized_batch(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *sources,
                                                 double *logL)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (W < 0 || (W > 0 && (!sources || !logL))) return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "loglike_maximized_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc)
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_METHOD, std::string("unknown generation_method: ") + (method ? method : "(null)"));
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (ctx->L <= 0 || !ctx->have_data) return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "loglike_maximized_batch: set_network with data first");
	if (!ctx->uniform || ctx->gaussleg)
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_STATE, "loglike_maximized_batch: needs a uniform grid and Simpson's rule (the time axis is an FFT)");
	MCUDA(ctx, cudaSetDevice(ctx->device));
	const int L = ctx->L, D = ctx->D, npol = desc.pv2 ? 2 : 1;
	// the coalescence-frame source of the intrinsic branches (src/mcmc_gw.cpp:2606-2612, 2692-2698)
	std::vector<gwat_b200_source> local(sources, sources + W);
	for (gwat_b200_source &s : local) {
		s.psi = 0;
		s.phiRef = 1;
		s.f_ref = desc.pv2 ? 20 : 10;
		s.incl_angle = 0;
		s.tc = 1;
		s.theta = 0;
		s.phi = 0;
		s.horizon_coord = 0;
		s.equatorial_orientation = 0;
	}
	const size_t per_walker = (size_t)D * npol * L * sizeof(cufftDoubleComplex);
	// walkers per pass: a 1 GiB FFT buffer, and gridDim.y of k_max_fill (65535)
	const int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)W, 65535), ((size_t)1 << 30) / per_walker));
	Scratch sc;
	MCUDA(ctx, cudaMalloc((void **)&sc.buf, (size_t)chunk * per_walker));
	MCUDA(ctx, cudaMalloc((void **)&sc.norms, (size_t)chunk * D * 3 * sizeof(double)));
	MCUDA(ctx, cudaMalloc((void **)&sc.out, (size_t)chunk * sizeof(double)));
	cudaStream_t st = ctx->stream;
	const double *wq_f = ctx->d_net + 3 * (size_t)D * ctx->ld, *dre = ctx->d_net + (size_t)D * ctx->ld, *dim = ctx->d_net + 2 * (size_t)D * ctx->ld;
	const double df = ctx->h_f[1] - ctx->h_f[0];
	int planned_batch = -1;
	for (int w0 = 0; w0 < W; w0 += chunk) {
		const int nw = std::min(chunk, W - w0);
		double *pol = nullptr;
		if (int rc = gwat_internal::polarizations_dev(ctx, method, nw, local.data() + w0, &pol, st)) return rc;
		const size_t n = (size_t)nw * L;
		const double *hp_re = pol, *hp_im = pol + n, *hc_re = pol + 2 * n, *hc_im = pol + 3 * n;
		k_max_norms<<<dim3(nw, D), kT, 0, st>>>(hp_re, hp_im, hc_re, hc_im, wq_f, ctx->ld, L, D, ctx->pref_fisher, sc.norms);
		k_max_fill<<<dim3((L + kT - 1) / kT, nw, D), kT, 0, st>>>(hp_re, hp_im, hc_re, hc_im, wq_f, dre, dim, ctx->ld, L, D, npol, sc.norms, sc.buf);
		const int batch = nw * D * npol;
		if (batch != planned_batch) {
			if (sc.have_plan) cufftDestroy(sc.plan);
			sc.have_plan = false;
			int nfft[1] = {L};
			if (cufftPlanMany(&sc.plan, 1, nfft, nullptr, 1, L, nullptr, 1, L, CUFFT_Z2Z, batch) != CUFFT_SUCCESS)
				return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftPlanMany failed");
			sc.have_plan = true;
			planned_batch = batch;
			if (cufftSetStream(sc.plan, st) != CUFFT_SUCCESS) return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftSetStream failed");
		}
		if (cufftExecZ2Z(sc.plan, sc.buf, sc.buf, CUFFT_FORWARD) != CUFFT_SUCCESS)  // FFTW_FORWARD, src/util.cpp:968
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftExecZ2Z failed");
		k_max_reduce<<<nw, kT, 0, st>>>(sc.buf, L, D, npol, df, sc.norms, sc.out);
		ctx->launches += 3;
		MCUDA(ctx, cudaGetLastError());
		MCUDA(ctx, cudaMemcpyAsync(logL + w0, sc.out, sizeof(double) * nw, cudaMemcpyDeviceToHost, st));
		MCUDA(ctx, cudaStreamSynchronize(st));
	}
	return GWAT_B200_OK;
}

// The intrinsic branch of MCMC_likelihood_wrapper (src/mcmc_gw.cpp:2569-2722) for W sampling vectors of the intrinsic sets:
// MCMC_prep_params with mcmc_intrinsic -> repack_parameters("MCMC_" + method, sky_average) -> the maximised likelihood above.
extern "C" int gwat_b200_loglike_maximized_mcmc_batch(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                                                      const double *params, double gmst, double *logL)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (W < 0 || (W > 0 && (!params || !logL))) return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "loglike_maximized_mcmc_batch: NULL array");
	if (W == 0) return GWAT_B200_OK;
	std::vector<gwat_b200_source> src((size_t)W);
	if (int rc = gwat_b200_repack_mcmc_intrinsic_batch(ctx, method, mod, dimension, W, params, gmst, src.data())) return rc;
	return gwat_b200_loglike_maximized_batch(ctx, method, W, src.data(), logL);
}

// match(data1, data2, SN, frequencies, length) of the reference (src/waveform_util.cpp:41-89; gwatpy: match_py): the overlap of two
// frequency-domain series maximised over a relative time shift, 4 max_t |IFFT(conj(d1) d2 / S)| df / (||d1|| ||d2||), with the norms
// from data_snr (Simpson's rule, delta_f = f[1] - f[0]).  The inverse transform is cuFFT's (FFTW_BACKWARD in the reference).
extern "C" int gwat_b200_match(gwat_b200_ctx *ctx, int L, const double *frequencies, const double *psd, const double *data1_re,
                               const double *data1_im, const double *data2_re, const double *data2_im, double *match)
{
	if (!ctx || L < 4 || !frequencies || !psd || !data1_re || !data1_im || !data2_re || !data2_im || !match) return GWAT_B200_ERR_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	MCUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	double *d = nullptr;
	cufftDoubleComplex *buf = nullptr;
	MCUDA(ctx, cudaMalloc((void **)&d, sizeof(double) * ((size_t)5 * L + 4)));
	MCUDA(ctx, cudaMalloc((void **)&buf, sizeof(cufftDoubleComplex) * (size_t)L));
	auto cleanup = [&]() {
		cudaFree(d);
		cudaFree(buf);
	};
	const double *host[5] = {data1_re, data1_im, data2_re, data2_im, psd};
	for (int k = 0; k < 5; k++)
		if (cudaMemcpyAsync(d + (size_t)k * L, host[k], sizeof(double) * L, cudaMemcpyHostToDevice, st) != cudaSuccess) {
			cleanup();
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "match: upload failed");
		}
	double *norms = d + (size_t)5 * L;
	k_match_fill<<<1, kT, 0, st>>>(L, d, d + L, d + 2 * (size_t)L, d + 3 * (size_t)L, d + 4 * (size_t)L, buf, norms);
	cufftHandle plan;
	if (cufftPlan1d(&plan, L, CUFFT_Z2Z, 1) != CUFFT_SUCCESS || cufftSetStream(plan, st) != CUFFT_SUCCESS) {
		cleanup();
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "match: cufftPlan1d failed");
	}
	const bool ok = cufftExecZ2Z(plan, buf, buf, CUFFT_INVERSE) == CUFFT_SUCCESS;
	k_match_max<<<1, kT, 0, st>>>(L, buf, norms);
	ctx->launches += 2;
	double h[3] = {0, 0, 0};
	const cudaError_t e1 = cudaMemcpyAsync(h, norms, sizeof(h), cudaMemcpyDeviceToHost, st);
	const cudaError_t e2 = cudaStreamSynchronize(st);
	cufftDestroy(plan);
	cleanup();
	if (!ok || e1 != cudaSuccess || e2 != cudaSuccess) return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "match: transform failed");
	const double delta_f = frequencies[1] - frequencies[0];
	// simpsons_sum(delta_f, ...) = delta_f / 3 * sum coef_i g_i   (include/gwat/util.h:843-856)
	const double n1 = sqrt(delta_f / 3. * h[0]), n2 = sqrt(delta_f / 3. * h[1]);
	*match = 4. * (h[2] * delta_f) / (n1 * n2);
	return GWAT_B200_OK;
}
