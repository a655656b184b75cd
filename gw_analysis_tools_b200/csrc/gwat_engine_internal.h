// Internal interface between the translation units of libgwat_b200.so (engine <-> ensemble sampler).  Not installed.
#ifndef GWAT_ENGINE_INTERNAL_H
#define GWAT_ENGINE_INTERNAL_H
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/gwat_b200.h"
#include "gwat_model.h"
#include "gwat_setup.h"

struct LikeLane {
	gwat::WalkerCoef *d_coef = nullptr;
	size_t cap_walkers = 0;
	double *d_partial = nullptr;
	size_t cap_partial = 0;
	unsigned long long *d_active = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaEvent_t ev_a = nullptr, ev_b = nullptr;  // hand-off between the light and the heavy stream of a lane
};

struct gwat_b200_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::string err;
	std::mutex mu;
	// network
	int D = 0, L = 0;
	int ld = 0;  // L padded to a whole number of 256-bin tiles: leading dimension of the per-detector tables
	bool have_data = false, gaussleg = false, log10F = false, uniform = false;
	double df = 0;
	gwat::Network net{};
	double pref_like = 0, pref_fisher = 0;
	double *d_grid = nullptr;  // f, sf_hi, sf_lo, logf : 4*L
	double *d_net = nullptr;   // wq, dre, dim, wq_fisher : 4*D*L
	std::vector<double> h_f;
	// scratch (grown on demand)
	size_t cap_walkers = 0, cap_partial = 0, cap_params = 0, cap_out = 0, cap_src = 0;
	gwat::WalkerCoef *d_coef = nullptr;
	double *d_partial = nullptr;
	double *d_params = nullptr;
	double *d_out = nullptr;
	gwat_b200_source *d_src = nullptr;
	unsigned long long *d_active = nullptr;
	unsigned long long *h_active = nullptr;  // pinned host word the active-bin count is copied into
	double *d_zero = nullptr;  // D*ld zeros: the strain seen by gwat_b200_snr_batch
	size_t cap_zero = 0;
	size_t cap_deriv = 0, cap_scale = 0, cap_fisher = 0, cap_bc = 0, cap_tcoef = 0;
	double *d_deriv = nullptr, *d_scale = nullptr, *d_fisher = nullptr, *d_tcoef = nullptr;
	int *d_bc = nullptr;
	int *d_binlim = nullptr;  // per source: bins from here on have exactly zero derivatives (Fisher passes)
	size_t cap_binlim = 0;
	// Fisher batches through host buffers: two pinned staging sets and two device source/result sets, so that the copies of
	// pass k+1 and k-1 overlap the kernels of pass k (the caller's arrays are pageable)
	struct FisherStage {
		gwat_b200_source *h_src = nullptr, *d_src = nullptr;
		double *h_out = nullptr, *d_out = nullptr;
		size_t cap_src = 0, cap_out = 0;
		cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
	} fstage[2];
	cudaStream_t copy_in = nullptr, copy_out = nullptr;
	// two more sets of likelihood scratch for callers that keep several batches in flight on their own streams (the
	// ensemble sampler); swapped in by LaneSwap while the caller holds `mu`
	LikeLane extra[2];
	// introspection
	long long launches = 0;
	bool kernel_timing = false;  // CUDA events around k_loglike (gwat_b200_set_kernel_timing)
	double last_ms = 0;
	long long last_active = 0;
};

namespace gwat_internal {
// All of these expect the caller to hold ctx->mu and to have made ctx->device current.
// lane 0 = the context's own scratch, 1..2 = ctx->extra[lane-1]
// st_heavy (optional): the bin kernel is launched there instead, chained to `st` by events, so that a caller can give the
// short latency-bound kernels (setup, finish, its own bookkeeping) a higher stream priority than the long FP64-bound one.
int loglike_mcmc_lane(gwat_b200_ctx *ctx, int lane, const char *method, const gwat_b200_mod *mod, int dimension, int W,
                      const double *d_params, double gmst, double T_segment, double *d_logL, cudaStream_t st,
                      cudaStream_t st_heavy = nullptr);
// Fisher matrices of MCMC_fisher_wrapper (src/mcmc_gw.cpp:2230-2300) before MCMC_fisher_transformations: sampling vectors on
// the device -> sum over the network's detectors of fisher_numerical("MCMC_"+method, detector d, reference detector 0).
// d_fisher: device [S][dimension][dimension].  Synchronous with respect to `st` only.
int fisher_mcmc_dev(gwat_b200_ctx *ctx, const char *method, const gwat_b200_mod *mod, int dimension, int order, int S,
                    const double *d_params, double gmst, double *d_fisher, cudaStream_t st);
// Polarisations of W sources (host records) on the context's grid, left on the device: *d_out = [hp_re | hp_im | hc_re | hc_im],
// each [W][L] (the context's output scratch: valid until the next call that produces outputs).
int polarizations_dev(gwat_b200_ctx *ctx, const char *method, int W, const gwat_b200_source *h_sources, double **d_out, cudaStream_t st);
int set_error(gwat_b200_ctx *ctx, int code, const std::string &msg);
}  // namespace gwat_internal
#endif
