/* gwat_b200_sampler.h -- the batched parallel-tempering Metropolis-Hastings step on the device (SURVEY.md 8f N1 + N3).
 *
 * What it replaces in the reference (one chain per CPU task, one likelihood per callback):
 *   mcmc_step                 src/mcmc_sampler_internals.cpp:35-146     proposal -> prior -> logL -> MH accept
 *   gaussian_step             :364-421      one random coordinate, per-dimension widths
 *   diff_ev_step              :846-977      differential evolution from the chain's own history
 *   fisher_step/update_fisher :424-629, 643-713   jump along an eigenvector of the Fisher matrix (MCMC_fisher_wrapper,
 *                                           src/mcmc_gw.cpp:2230-2300, + MCMC_fisher_transformations :2136-2189)
 *   assign_probabilities      :1196-1364    step-type probabilities from temperature / Fisher / history state (non-RJ)
 *   update_step_widths        :1623-1703    x0.9 / x1.1 width tuning towards the target acceptance band
 *   update_history            :2198-2219    ring buffer feeding differential evolution
 *   chain_swap/single_chain_swap :1086-1184 sequential sweep over adjacent chains, exp((l1-l2)/T2 - (l1-l2)/T1)
 *   PTMCMC_MH_step_incremental src/mcmc_sampler.cpp:4571-4660 (non-pool loop): swp_freq steps per chain, then one swap sweep
 *                                           with probability swap_rate
 *   logPriorStandard_{D,P,D_NRT,P_NRT}[_mod]::eval   src/standardPriorLibrary.cpp:321-526
 *
 * Everything runs on the device: positions, likelihoods, histories, Fisher eigen-systems and counters live in HBM; one call
 * advances every chain by n steps without a host round-trip per step.  The chains are cut into two halves that run on two
 * CUDA streams, so the latency-bound per-walker setup of one half overlaps the FP64-bound bin kernel of the other.
 *
 * Random numbers: Philox4x32-10 keyed by `seed`, counter = (step, chain, purpose) -- every draw is a pure function of
 * those three, so a run is reproducible for any lane split or GPU count, and tests can replay the draws on the CPU.
 * (The reference seeds one gsl_rng per chain with chain+1, :1939; bit-identical streams are neither possible nor needed.)
 *
 * Not built (outside this path): RJMCMC / nested models, KDE and MMALA proposals, block sampling, dynamic temperature
 * allocation, the thread-pool scheduler, checkpoint files.
 */
#ifndef GWAT_B200_SAMPLER_H
#define GWAT_B200_SAMPLER_H

#include "gwat_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* priorData of include/gwat/standardPriorLibrary.h:8-35 (the fields the standard priors read), [min, max] pairs */
typedef struct gwat_b200_prior {
	double mass1_prior[2], mass2_prior[2];
	double spin1_prior[2], spin2_prior[2];         /* aligned-spin models */
	double a1_prior[2], a2_prior[2];               /* precessing models: magnitudes, cos tilts, azimuths */
	double ctheta1_prior[2], ctheta2_prior[2];
	double phi1_prior[2], phi2_prior[2];
	double tidal1_prior[2], tidal2_prior[2], tidal_s_prior[2];
	double RA_bounds[2], sinDEC_bounds[2];
	double DL_prior[2];
	double T_merger;                               /* tc must lie within +-0.1 s of it */
	double mod_priors[GWAT_B200_MAX_MOD][2];       /* ppE / gIMR / theory parameters */
	int tidal_love;
	int reserved_;
} gwat_b200_prior;

typedef struct gwat_b200_sampler_options {
	int chain_N;                /* total chains (all temperatures, all ensembles), index order = the reference's chain order */
	int dimension;
	int swp_freq;               /* steps between swap sweeps */
	double swap_rate;           /* probability that a sweep happens (the reference sets 1/swp_freq, src/mcmc_sampler.cpp:4316) */
	int history_length;         /* 1000 */
	int history_update;         /* 10 */
	int fisher_exist;           /* 0: Gaussian steps only (the reference never uses DE without a Fisher, :1222-1231) */
	int fisher_update_number;   /* 200 when tuning, else the run length */
	int fisher_deriv_order;     /* 4 (include/gwat/mcmc_gw.h:45) */
	int check_stepsize_freq;    /* 50 when tuning, else the run length */
	unsigned long long seed;
	int lanes;                  /* 1 or 2 concurrent halves */
	int record_cold;            /* keep the positions of the T=1 chains of every step in a device buffer (gwat_b200_sampler_cold) */
	int fisher_deferred;        /* 0: a chain's Fisher matrix is recomputed at the step that first needs it, at the position it has
	                               then (the reference's schedule, :434-437).  1: the chains that come due between two swap sweeps
	                               are refreshed together at the next sweep, on a side stream that overlaps the likelihood kernels,
	                               and the new eigen-systems are installed at the sweep after that: same cadence per chain, same
	                               use, matrices up to 2 swp_freq steps staler out of the ~400 steps each is used for; all chains get
	                               their first matrix at creation. */
	int chain_index_offset;     /* global index of this sampler's chain 0 when an ensemble is sharded over several GPUs: random draws
	                               are functions of the GLOBAL chain index, so N samplers reproduce what one would do */
	int reserved_;
} gwat_b200_sampler_options;

typedef struct gwat_b200_sampler gwat_b200_sampler;

void gwat_b200_prior_init(gwat_b200_prior *p);                    /* wide-open bounds, tidal_love = 1 */
void gwat_b200_sampler_options_init(gwat_b200_sampler_options *o); /* the reference's defaults listed above */

/* The network (grid, PSDs, data) must already be set on ctx.  chain_temps[chain_N]; initial_positions[chain_N][dimension].
 * Evaluates prior and likelihood of the initial positions (all must have a finite prior, as assign_initial_pos requires). */
int gwat_b200_sampler_create(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                             const gwat_b200_sampler_options *options, const gwat_b200_prior *prior, const double *chain_temps,
                             const double *initial_positions, double gmst, double T_segment, gwat_b200_sampler **out);
void gwat_b200_sampler_destroy(gwat_b200_sampler *s);   /* before gwat_b200_ctx_destroy of its context */

/* Advance every chain by n_steps (swap sweeps every swp_freq steps).  Returns when the device has finished. */
int gwat_b200_sampler_run(gwat_b200_sampler *s, int n_steps);

/* Current state, any pointer may be NULL: positions[chain_N][dimension], logL[chain_N], logP[chain_N] */
int gwat_b200_sampler_state(gwat_b200_sampler *s, double *positions, double *logL, double *logP);
/* Overwrite the current state (any pointer may be NULL); histories, widths and Fisher eigen-systems stay with their slots, as
 * they do in the reference's swaps.  For callers that exchange states between GPUs (below). */
int gwat_b200_sampler_set_state(gwat_b200_sampler *s, const double *positions, const double *logL, const double *logP);
/* Several GPUs: create every rank's sampler with swap_rate = 0 and chain_index_offset = its first global chain, run swp_freq
 * steps, gather logL of all chains, and let every rank compute the reference's swap sweep over the WHOLE ladder with the very
 * draws and arithmetic the single-GPU sweep uses (host code, no GPU needed): src[i] = global slot whose state moves to slot i.
 * gwat_b200_sampler_uniform exposes the counter-based draws (purpose 5 = the swap_rate gate of sweep `step`, chain 0). */
int gwat_b200_swap_sweep_host(int chain_N_total, const double *logL, const double *temps, unsigned long long seed, long long sweep,
                              int *src, int *accepted /* [chain_N_total - 1] or NULL */);
void gwat_b200_sampler_uniform(unsigned long long seed, unsigned long long step, unsigned chain, unsigned purpose, double *out2);
/* The same sweep (chain_swap, src/mcmc_sampler_internals.cpp:1086-1184) computed the way the sampler computes it on the device --
 * thresholds for all pairs at once, the starts of the carried runs by pointer doubling -- on caller-supplied inputs; mode 0: as the
 * sampler runs it, mode 1: its sequential fallback.  Exists so that tests can hold the device sweep against the host one on
 * adversarial ladders; identical decisions for any input. */
int gwat_b200_swap_sweep_device(gwat_b200_ctx *ctx, int chain_N_total, const double *logL, const double *temps, unsigned long long seed,
                                long long sweep, int mode, int *src, int *accepted /* [chain_N_total - 1] or NULL */);
/*
 * One ladder sharded over the GPUs of a box, one process (rank) per GPU -- the split BASELINE.json's north_star names: "only the
 * PT swap step exchanges per-walker log-likelihoods and positions via NCCL allgather over NVLink".
 *   rank 0:     gwat_b200_nccl_unique_id(id), then hands the 128 bytes to the other ranks (MPI_Bcast, torch.distributed, a file ...)
 *   every rank: gwat_b200_sampler_create(...) with its own chain_N chains (equal on all ranks), their temperatures and
 *               options.chain_index_offset = rank * chain_N, then gwat_b200_sampler_attach_ranks(s, id, rank, n_ranks)
 *   every rank: gwat_b200_sampler_run(s, n) with the same n.
 * At every swap sweep the ranks ncclAllGather one record [position | logL | logP] per chain (8 (dimension + 2) bytes) on the
 * sampler's stream, every rank runs the reference's sequential sweep (chain_swap / single_chain_swap,
 * src/mcmc_sampler_internals.cpp:1086-1184) over the WHOLE ladder from the same gathered logL and the same counter-based draws,
 * and keeps the records that land in its own slots.  Nothing goes through the host.  Because every draw is a function of the
 * global chain index, n_ranks samplers reproduce bit for bit what one sampler with all the chains does.
 * NCCL is loaded with dlopen("libnccl.so.2") at the first of these calls; without it they return GWAT_B200_ERR_UNSUPPORTED.
 */
#define GWAT_B200_NCCL_UNIQUE_ID_BYTES 128
int gwat_b200_nccl_unique_id(unsigned char *id128);
int gwat_b200_sampler_attach_ranks(gwat_b200_sampler *s, const unsigned char *id128, int rank, int n_ranks);
/* mean device time of the swap exchange (pack, all-gather, sweep, take) over the first 64 sweeps of the last run, in ms, and the
 * number of sweeps that run performed */
double gwat_b200_sampler_last_swap_ms(const gwat_b200_sampler *s);
long long gwat_b200_sampler_last_sweeps(const gwat_b200_sampler *s);

/* counters[chain_N][GWAT_B200_SAMPLER_NCOUNTERS] (see the enum), widths[chain_N][dimension + 3]: Gaussian widths per dimension,
 * then the DE, (unused) and Fisher widths */
enum {
	GWAT_B200_CT_STEP_ACCEPT = 0, GWAT_B200_CT_STEP_REJECT, GWAT_B200_CT_GAUSS_ACCEPT, GWAT_B200_CT_GAUSS_REJECT,
	GWAT_B200_CT_DE_ACCEPT, GWAT_B200_CT_DE_REJECT, GWAT_B200_CT_FISHER_ACCEPT, GWAT_B200_CT_FISHER_REJECT,
	GWAT_B200_CT_SWAP_ACCEPT, GWAT_B200_CT_SWAP_REJECT, GWAT_B200_CT_FISHER_UPDATES, GWAT_B200_CT_FISHER_NAN,
	GWAT_B200_SAMPLER_NCOUNTERS
};
int gwat_b200_sampler_counters(gwat_b200_sampler *s, long long *counters, double *widths);
/* Fisher eigen-systems the chains currently jump along (fisher_vals / fisher_vecs of the reference's sampler struct):
 * eigenvalues[chain_N][dimension], eigenvectors[chain_N][dimension][dimension], row i = eigenvector i */
int gwat_b200_sampler_fisher_state(gwat_b200_sampler *s, double *eigenvalues, double *eigenvectors);
/* With record_cold: positions of the T == 1 chains for steps [first_step, first_step + n) of the steps run so far:
 * out[n][n_cold][dimension]; returns the number of cold chains through *n_cold when out is NULL. */
int gwat_b200_sampler_cold(gwat_b200_sampler *s, long long first_step, int n, double *out, int *n_cold);
/* wall time on the device of the last gwat_b200_sampler_run (CUDA events), and kernels launched by it */
double gwat_b200_sampler_last_ms(const gwat_b200_sampler *s);
long long gwat_b200_sampler_last_launches(const gwat_b200_sampler *s);

/*
 * Dynamic temperature allocation (arXiv:1501.05823; dynamic_temperature_full_ensemble_internal src/mcmc_sampler.cpp:453-545 with
 * linear swapping, update_temperatures_full_ensemble src/mcmc_sampler_internals.cpp:3371-3413, PT_dynamical_timescale :3224-3230).
 *   gwat_b200_update_temperatures       the update rule alone (host arithmetic): A[i] = 1 / 0 says whether the last swap attempt
 *                                       between chains i-1 and i was accepted; chains at T = 1 and at the ensemble's hottest
 *                                       temperature stay where they are
 *   gwat_b200_sampler_dynamic_temperatures  blocks of swp_freq steps on the device, a sweep after each, the update on the host
 *                                       (C doubles back and forth per block), until N_steps - swp_freq steps are done
 *   gwat_b200_sampler_set_temperatures / _temperatures / _last_swap_accepts   the pieces, for callers with their own schedule
 * Single-rank samplers only.
 */
int gwat_b200_update_temperatures(int chain_N, double *chain_temps, const double *A /* [chain_N + 1] */, int t0, int nu, int t);
int gwat_b200_sampler_set_temperatures(gwat_b200_sampler *s, const double *chain_temps);
int gwat_b200_sampler_temperatures(gwat_b200_sampler *s, double *chain_temps);
int gwat_b200_sampler_last_swap_accepts(gwat_b200_sampler *s, int *accepted /* [chain_N - 1] */);
int gwat_b200_sampler_dynamic_temperatures(gwat_b200_sampler *s, int N_steps, int nu, int t0, long long *sweeps_done);

/*
 * Chain output (host code, csrc/gwat_chain_io.cpp): the data dump and the thinned, flattened sample file of the reference's
 * mcmc_sampler_output (create_data_dump src/mcmc_io_util.cpp:643-990, write_flat_thin_output :555-642, count_indep_samples :521-552)
 * with the same dataset paths, shapes and thinning rule.  The reference writes HDF5; HDF5 is not available where this library is
 * built, so the datasets go into a flat self-describing container (layout at the top of gwat_chain_io.cpp: magic, then per record
 * the HDF5-style path, dtype 0 = float64 / 1 = int32, rank, dims, row-major payload).
 *   positions [n_chains][steps][dimension]; logl_logp [n_chains][steps][2] or NULL; trim_lengths [n_chains] or NULL;
 *   ac_values [n_cold][dimension] autocorrelation lengths (NULL: dataset omitted), e.g. from gwat_b200_autocorrelation_lengths below.
 */
typedef struct gwat_b200_dump gwat_b200_dump;
int gwat_b200_dump_create(const char *path, gwat_b200_dump **out);
int gwat_b200_dump_write(gwat_b200_dump *d, const char *dataset_path, int dtype, int rank, const long long *dims, const void *data);
int gwat_b200_dump_close(gwat_b200_dump *d);
int gwat_b200_write_data_dump(const char *path, int n_chains, int dimension, long long steps, const int *chain_ids, const double *temperatures,
                              const double *positions, const double *logl_logp, const int *trim_lengths, int n_cold, const int *ac_values);
int gwat_b200_write_flat_thin_output(const char *path, int n_cold, int dimension, long long steps, const double *positions,
                                     const int *trim_lengths, const int *ac_values, long long *n_rows);

/*
 * The autocorrelation lengths the thinning needs, as mcmc_sampler_output::calc_ac_vals obtains them (src/mcmc_io_util.cpp:434-520:
 * auto_corr_from_data_batch with one cumulative segment, src/autocorrelation.cpp:152-332): for every (chain, dimension) row of
 * positions[n_chains][steps][dimension], from step `begin` (the trim) on, emcee's windowed estimator
 * (auto_correlation_spectral_windowed, :401-462: x - mean zero-padded to L = 2 * 2^ceil(log2 n), rho = IFFT(|FFT x|^2) / rho_0,
 * tau_i = 2 sum_{j<=i} rho_j - 1, window = first i > 5 tau_i) and lag = int(tau_window) (:334-347).  All rows of a pass go through
 * two batched cuFFT transforms on the context's device.  ac_values[n_chains][dimension]; tau (same shape, the estimator before
 * truncation) may be NULL.  Rows of one or two steps give 2, as the reference's brute-force branch does (MAX_SERIAL = 2, :8).
 * The reference passes ONE length for all chains -- that of the first cold chain minus the trim of the last (:474) -- hence one `begin`.
 */
int gwat_b200_autocorrelation_lengths(gwat_b200_ctx *ctx, int n_chains, int dimension, long long steps, const double *positions,
                                      long long begin, int *ac_values, double *tau);

/* Building blocks, exposed for tests and for callers that keep their own sampler loop:
 * log prior of W sampling vectors (host arrays) with the standard prior of the method's family */
int gwat_b200_log_prior_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod, int dimension, int W,
                              const gwat_b200_prior *prior, const double *params, double *logP);
/* Fisher matrices as MCMC_fisher_wrapper returns them (sum over detectors + MCMC_fisher_transformations), then their
 * eigen-systems: fisher[W][dim][dim], eigenvalues[W][dim] ascending, eigenvectors[W][dim][dim] (row i = i-th vector). */
int gwat_b200_mcmc_fisher_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod, int dimension,
                                int order, int W, const double *params, double gmst, double *fisher, double *eigenvalues,
                                double *eigenvectors);
/* MCMC_fisher_wrapper of an INTRINSIC run (mcmc_intrinsic: the sets of gwat_b200_loglike_maximized_mcmc_batch; src/mcmc_gw.cpp:2229-2330):
 * sum over the network's detectors of fisher_numerical("MCMC_" + method) on the sky-averaged record, then the intrinsic branch of
 * MCMC_fisher_transformations (:2163-2179) and the dCS / EdGB unit factor.  IMRPhenomD family (with ppE / gIMR modifications) and
 * IMRPhenomPv2; the NRT sets are refused with the reason (see gwat_b200_fisher_numerical_batch).  fisher[W][dim][dim]. */
int gwat_b200_mcmc_fisher_intrinsic_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod, int dimension,
                                          int order, int W, const double *params, double gmst, double *fisher);

#ifdef __cplusplus
}
#endif
#endif
