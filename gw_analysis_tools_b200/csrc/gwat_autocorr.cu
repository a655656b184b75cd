// Autocorrelation lengths of the cold chains, as the reference's chain output computes them before thinning:
//
//   mcmc_sampler_output::calc_ac_vals            src/mcmc_io_util.cpp:434-520   one segment, cumulative, target 0.01
//   auto_corr_from_data_batch / _from_data       src/autocorrelation.cpp:152-332 every (chain, dimension) row is one job; rows longer than
//                                                                                MAX_SERIAL = 2 (:8) take the spectral job
//   auto_correlation_spectral_windowed           src/autocorrelation.cpp:401-462 emcee's estimator: x - mean zero-padded to
//                                                                                L = 2 * 2^ceil(log2 n), rho = IFFT(|FFT x|^2) / [0],
//                                                                                tau_i = 2 sum_{j<=i} rho_j - 1, window = first i > 5 tau_i
//   threaded_ac_spectral                         src/autocorrelation.cpp:334-347 lag = int(tau_window)
//
// The reference runs one FFTW plan pair per job on a thread pool (and has a CUDA file of its own for the brute-force variant,
// src/autocorrelation_cuda.cu, which this path never calls).  Here all rows of a pass go through two batched cuFFT transforms
// (plain library FFTs, as FFTW is in the reference) and one CTA per row forms the running sum and finds the window.
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "gwat_engine_internal.h"

namespace {

constexpr int kT = 256;

#define ACUDA(ctx, call)                                                                                     \
	do {                                                                                                       \
		cudaError_t e_ = (call);                                                                                 \
		if (e_ != cudaSuccess)                                                                                   \
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

__device__ __forceinline__ double block_sum(double v, double *sh)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) sh[wid] = v;
	__syncthreads();
	v = 0.0;
	if (wid == 0) {
		v = lane < kT / 32 ? sh[lane] : 0.0;
		for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	}
	return v;  // valid in thread 0
}

// row r = chain * dimension + dim of positions[chain][step][dim]: buf[r][i] = x_i - mean for i < n, 0 up to L (mean_list + the
// padding loop of auto_correlation_spectral_windowed, src/autocorrelation.cpp:419-430)
__global__ void __launch_bounds__(kT) k_ac_fill(const double *__restrict__ pos, long long steps, long long begin, int dimension, int n, int L,
                                               cufftDoubleComplex *__restrict__ buf)
{
	__shared__ double sh[kT / 32];
	__shared__ double mean_s;
	const int r = blockIdx.x, chain = r / dimension, dim = r % dimension;
	const double *x = pos + ((size_t)chain * steps + begin) * dimension + dim;
	double s = 0.0;
	for (int i = threadIdx.x; i < n; i += kT) s += x[(size_t)i * dimension];
	s = block_sum(s, sh);
	if (threadIdx.x == 0) mean_s = s / n;
	__syncthreads();
	const double mean = mean_s;
	cufftDoubleComplex *o = buf + (size_t)r * L;
	for (int i = threadIdx.x; i < L; i += kT) o[i] = cufftDoubleComplex{i < n ? x[(size_t)i * dimension] - mean : 0.0, 0.0};
}

__global__ void __launch_bounds__(kT) k_ac_power(cufftDoubleComplex *__restrict__ buf, size_t total)
{
	const size_t i = (size_t)blockIdx.x * kT + threadIdx.x;
	if (i >= total) return;
	const cufftDoubleComplex v = buf[i];
	buf[i] = cufftDoubleComplex{v.x * v.x + v.y * v.y, 0.0};
}

// One CTA per row: rho_i = buf[i].re / buf[0].re, tau_i = 2 sum_{j<=i} rho_j - 1, window = first i with i > 5 tau_i (else n - 1);
// tau[r] = tau_window, lag[r] = int(tau_window).  The running sum is formed chunk by chunk (256 elements: warp scans + carry).
__global__ void __launch_bounds__(kT) k_ac_window(const cufftDoubleComplex *__restrict__ buf, int n, int L, double *__restrict__ tau,
                                                 int *__restrict__ lag)
{
	__shared__ double warp_tot[kT / 32];
	__shared__ double carry_s;
	__shared__ int hit_s;
	__shared__ double tau_hit_s;
	const int r = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const cufftDoubleComplex *a = buf + (size_t)r * L;
	const double norm = a[0].x;
	if (threadIdx.x == 0) {
		carry_s = 0.0;
		hit_s = n;
		tau_hit_s = 0.0;
	}
	__syncthreads();
	const int c = 5;
	double last_tau = 0.0;
	for (int base = 0; base < n; base += kT) {
		const int i = base + threadIdx.x;
		double v = i < n ? a[i].x / norm : 0.0;
		for (int o = 1; o < 32; o <<= 1) {  // inclusive scan within the warp
			const double up = __shfl_up_sync(0xffffffffu, v, o);
			if (lane >= o) v += up;
		}
		if (lane == 31) warp_tot[wid] = v;
		__syncthreads();
		double before = carry_s;
		for (int w = 0; w < wid; w++) before += warp_tot[w];
		const double t = 2.0 * (before + v) - 1.0;
		if (i < n && (double)i > c * t) atomicMin(&hit_s, i);
		if (i == n - 1) last_tau = t;
		__syncthreads();
		const int hit = hit_s;
		if (hit < n) {
			if (i == hit) tau_hit_s = t;
			__syncthreads();
			break;
		}
		if (threadIdx.x == kT - 1) carry_s = before + v;
		__syncthreads();
	}
	if (hit_s >= n) {  // no window: the reference returns tau[n - 1]
		if (((n - 1) % kT) == (int)threadIdx.x) tau_hit_s = last_tau;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		tau[r] = tau_hit_s;
		lag[r] = (int)tau_hit_s;
	}
}

struct Scratch {
	cufftDoubleComplex *buf = nullptr;
	double *pos = nullptr, *tau = nullptr;
	int *lag = nullptr;
	cufftHandle plan = 0;
	bool have_plan = false;
	~Scratch()
	{
		if (have_plan) cufftDestroy(plan);
		cudaFree(buf);
		cudaFree(pos);
		cudaFree(tau);
		cudaFree(lag);
	}
};

}  // namespace

extern "C" int gwat_b200_autocorrelation_lengths(gwat_b200_ctx *ctx, int n_chains, int dimension, long long steps, const double *positions,
                                                 long long begin, int *ac_values, double *tau)
{
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (n_chains < 0 || dimension < 1 || steps < 1 || begin < 0 || begin >= steps || (n_chains > 0 && (!positions || !ac_values)))
		return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "autocorrelation_lengths: bad arguments");
	if (n_chains == 0) return GWAT_B200_OK;
	const long long n_ll = steps - begin;
	if (n_ll > (1LL << 28)) return gwat_internal::set_error(ctx, GWAT_B200_ERR_ARG, "autocorrelation_lengths: chains longer than 2^28 steps");
	const int n = (int)n_ll;
	const size_t rows = (size_t)n_chains * dimension;
	if (n <= 2) {
		// rows of at most MAX_SERIAL = 2 steps take auto_correlation_serial (src/autocorrelation.cpp:279, 360-393), whose loop leaves
		// at h = 2 because the next autocovariance is 0 / 0
		for (size_t r = 0; r < rows; r++) {
			ac_values[r] = 2;
			if (tau) tau[r] = NAN;
		}
		return GWAT_B200_OK;
	}
	const int L = 2 * (int)std::pow(2.0, std::ceil(std::log2((double)n)));  // (:406)
	std::lock_guard<std::mutex> lock(ctx->mu);
	ACUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	// chains per pass: a 1 GiB transform buffer
	const size_t per_chain = (size_t)dimension * L * sizeof(cufftDoubleComplex);
	const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_chains, ((size_t)1 << 30) / per_chain));
	Scratch sc;
	ACUDA(ctx, cudaMalloc((void **)&sc.buf, (size_t)chunk * per_chain));
	ACUDA(ctx, cudaMalloc((void **)&sc.pos, (size_t)chunk * steps * dimension * sizeof(double)));
	ACUDA(ctx, cudaMalloc((void **)&sc.tau, (size_t)chunk * dimension * sizeof(double)));
	ACUDA(ctx, cudaMalloc((void **)&sc.lag, (size_t)chunk * dimension * sizeof(int)));
	std::vector<double> h_tau((size_t)chunk * dimension);
	int planned = -1;
	for (int c0 = 0; c0 < n_chains; c0 += chunk) {
		const int nc = std::min(chunk, n_chains - c0), nr = nc * dimension;
		ACUDA(ctx, cudaMemcpyAsync(sc.pos, positions + (size_t)c0 * steps * dimension, sizeof(double) * (size_t)nc * steps * dimension,
		                           cudaMemcpyHostToDevice, st));
		if (nr != planned) {
			if (sc.have_plan) cufftDestroy(sc.plan);
			sc.have_plan = false;
			int nfft[1] = {L};
			if (cufftPlanMany(&sc.plan, 1, nfft, nullptr, 1, L, nullptr, 1, L, CUFFT_Z2Z, nr) != CUFFT_SUCCESS)
				return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftPlanMany failed");
			sc.have_plan = true;
			planned = nr;
			if (cufftSetStream(sc.plan, st) != CUFFT_SUCCESS) return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftSetStream failed");
		}
		k_ac_fill<<<nr, kT, 0, st>>>(sc.pos, steps, begin, dimension, n, L, sc.buf);
		if (cufftExecZ2Z(sc.plan, sc.buf, sc.buf, CUFFT_FORWARD) != CUFFT_SUCCESS)
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftExecZ2Z failed");
		const size_t total = (size_t)nr * L;
		k_ac_power<<<(unsigned)((total + kT - 1) / kT), kT, 0, st>>>(sc.buf, total);
		if (cufftExecZ2Z(sc.plan, sc.buf, sc.buf, CUFFT_INVERSE) != CUFFT_SUCCESS)  // FFTW_BACKWARD, unnormalised (allocate_FFTW_mem_reverse)
			return gwat_internal::set_error(ctx, GWAT_B200_ERR_CUDA, "cufftExecZ2Z failed");
		k_ac_window<<<nr, kT, 0, st>>>(sc.buf, n, L, sc.tau, sc.lag);
		ctx->launches += 3;
		ACUDA(ctx, cudaGetLastError());
		ACUDA(ctx, cudaMemcpyAsync(ac_values + (size_t)c0 * dimension, sc.lag, sizeof(int) * nr, cudaMemcpyDeviceToHost, st));
		ACUDA(ctx, cudaMemcpyAsync(h_tau.data(), sc.tau, sizeof(double) * nr, cudaMemcpyDeviceToHost, st));
		ACUDA(ctx, cudaStreamSynchronize(st));
		if (tau) std::memcpy(tau + (size_t)c0 * dimension, h_tau.data(), sizeof(double) * nr);
	}
	return GWAT_B200_OK;
}
