/*
 * gwat_b200.h -- C ABI of the B200-native batched likelihood engine.
 *
 * This is the drop-in boundary for ONE path of scottperkins/gw_analysis_tools (GWAT):
 *   frequency-domain waveform -> detector projection -> noise-weighted inner product -> log-likelihood,
 *   and the finite-difference Fisher stencil built from the same responses,
 * evaluated for whole ensembles of walkers at once on one B200.
 *
 * Plain C: pointers, sizes and PODs only.  No torch / CUDA types appear in any signature; `stream` arguments are
 * cudaStream_t passed as void* (NULL = the context's own stream).  Every entry point returns 0 on success or a
 * negative gwat_b200_status; nothing here ever calls exit() (the reference does, src/detector_util.cpp:1155-1158).
 * There is NO CPU fallback: if no CUDA device is usable, gwat_b200_ctx_create fails with GWAT_B200_ERR_CUDA.
 *
 * Each declaration cites the reference interface (path relative to the GWAT repository root) it replaces.
 */
#ifndef GWAT_B200_H
#define GWAT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWAT_B200_ABI_VERSION 1
#define GWAT_B200_MAX_DETECTORS 8
#define GWAT_B200_MAX_MOD 8 /* max ppE terms, and max gIMR modifications per coefficient family */
#define GWAT_B200_MAX_DIM 32 /* max sampling / Fisher dimension */

typedef enum gwat_b200_status {
	GWAT_B200_OK = 0,
	GWAT_B200_ERR_ARG = -1,       /* NULL pointer, bad size, unknown detector ... */
	GWAT_B200_ERR_METHOD = -2,    /* generation_method string not supported by this library */
	GWAT_B200_ERR_CUDA = -3,      /* CUDA runtime error (message in gwat_b200_last_error) */
	GWAT_B200_ERR_STATE = -4,     /* network/grid not set, or context built for another shape */
	GWAT_B200_ERR_UNSUPPORTED = -5 /* valid GWAT option that is outside this library's path (LISA, autodiff ...) */
} gwat_b200_status;

/*
 * Flattened gen_params_base<double>  (include/gwat/util.h:121-378).
 * Same field names and units (solar masses, Mpc, radians, seconds) as the reference class; pointer members of the
 * reference (betappe/bppe, delta_*, *i) become fixed-size arrays so the record is a POD that can live in device memory.
 * Defaults of the reference are reproduced by gwat_b200_source_init().
 */
typedef struct gwat_b200_source {
	double mass1, mass2;
	double Luminosity_Distance;
	double spin1[3], spin2[3];
	double tc, phiRef, f_ref;
	double psi, incl_angle;
	double RA, DEC, gmst;
	double theta, phi;               /* sky position in the detector's horizon frame, read when horizon_coord != 0 by the one-detector response */
	double theta_l, phi_l;           /* equatorial direction of L, read when equatorial_orientation != 0 by the one-detector entry points */
	double tidal1, tidal2, tidal_s, tidal_a, tidal_weighted, delta_tidal_weighted;
	double diss_tidal1, diss_tidal2, diss_tidal_s, diss_tidal_a, diss_tidal_weighted;
	double chip, phip;               /* reduced PhenomPv2 parameterisation, used when chip != -1 */
	double betappe[GWAT_B200_MAX_MOD];
	double bppe[GWAT_B200_MAX_MOD];
	double delta_phi[GWAT_B200_MAX_MOD];
	double delta_sigma[GWAT_B200_MAX_MOD];
	double delta_beta[GWAT_B200_MAX_MOD];
	double delta_alpha[GWAT_B200_MAX_MOD];
	int phii[GWAT_B200_MAX_MOD];
	int sigmai[GWAT_B200_MAX_MOD];
	int betai[GWAT_B200_MAX_MOD];
	int alphai[GWAT_B200_MAX_MOD];
	int Nmod, Nmod_phi, Nmod_sigma, Nmod_beta, Nmod_alpha;
	int PNorder;
	int shift_time, shift_phase, sky_average;
	int tidal_love, tidal_love_error;
	int NSflag1, NSflag2;
	int dep_postmerger;
	int equatorial_orientation, horizon_coord;
	int cosmology;                   /* gen_params::cosmology as its index in the reference's cosmos[] (include/gwat/D_Z_Config.h:13):
	                                    0 PLANCK15 (the default), 1 PLANCK13, 2 WMAP9, 3 WMAP7, 4 WMAP5, 5 TESTING_COSMOLOGY;
	                                    gwat_b200_cosmology_index() maps the name.  Read by the theory mappings only (Z_from_DL) */
	int reserved_;
} gwat_b200_source;

/*
 * Flattened MCMC_modification_struct  (include/gwat/mcmc_gw.h:50-88): which extra sampling dimensions exist and what
 * they mean.  GAUSS_QUAD / log10F / weights of the reference struct are properties of the network here
 * (gwat_b200_set_network), the fisher_* overrides are arguments of the Fisher calls.
 */
typedef struct gwat_b200_mod {
	int ppE_Nmod;
	double bppe[GWAT_B200_MAX_MOD];
	int gIMR_Nmod_phi, gIMR_Nmod_sigma, gIMR_Nmod_beta, gIMR_Nmod_alpha;
	int gIMR_phii[GWAT_B200_MAX_MOD];
	int gIMR_sigmai[GWAT_B200_MAX_MOD];
	int gIMR_betai[GWAT_B200_MAX_MOD];
	int gIMR_alphai[GWAT_B200_MAX_MOD];
	int NSflag1, NSflag2;
	int tidal_love, tidal_love_error;
} gwat_b200_mod;

typedef struct gwat_b200_ctx gwat_b200_ctx;

/* ---- library / context ------------------------------------------------------------------------------------------ */

int gwat_b200_abi_version(void);

/* Fill a source / modification record with the reference's member defaults (include/gwat/util.h:125-285,
 * include/gwat/mcmc_gw.h:52-67). */
void gwat_b200_source_init(gwat_b200_source *src);
/* transform_orientation_coords (src/waveform_util.cpp:1535-1595) for terrestrial detectors, in place: incl_angle and psi of every
 * source from the equatorial direction of L (theta_l, phi_l); for IMRPhenomPv2 methods psi follows the direction of J.  The
 * one-detector entry points below apply it themselves to sources with equatorial_orientation set (on a copy: the reference
 * overwrites the caller's object); it is exported for callers that want the derived angles. */
int gwat_b200_transform_orientation_coords(const char *generation_method, int n, gwat_b200_source *sources);
/* Index of a cosmology name as Z_from_DL reads it (case-insensitive, src/util.cpp:356-365); -1 for a name the reference does not know. */
int gwat_b200_cosmology_index(const char *name);
/* Site constants of a GWAT detector (include/gwat/detector_util.h:29-157; what gwatpy's get_detector_parameters returns,
 * src/gwatpy_wrapping.cpp:743-831): latitude and longitude [rad], vertex location[3] [m], response_tensor[9] row-major.  The name is
 * one set_network takes; GWAT_B200_ERR_ARG for an unknown name or a NULL pointer.  Host constants: no context, no GPU. */
int gwat_b200_detector_site(const char *detector, double *latitude, double *longitude, double *location, double *response_tensor);
void gwat_b200_mod_init(gwat_b200_mod *mod);

/* One context per GPU and per submitting thread group.  Owns the device copies of the frequency grid, PSDs, data and
 * all scratch; replaces the file-static globals mcmc_data/mcmc_noise/mcmc_frequencies/mcmc_detectors/...
 * (include/gwat/mcmc_gw.h:22-46).  `device` is the CUDA ordinal. */
int gwat_b200_ctx_create(gwat_b200_ctx **ctx, int device);
void gwat_b200_ctx_destroy(gwat_b200_ctx *ctx);
/* Human-readable description of the last failure on this context ("" if none); NULL ctx -> last create failure. */
const char *gwat_b200_last_error(const gwat_b200_ctx *ctx);

/*
 * Upload the detector network: D detectors sharing one frequency grid of L bins (the reference only supports the shared
 * grid too: create_coherent_GW_detection, src/waveform_util.cpp:140-145).
 *   detectors[d]         GWAT detector names ("Hanford", "Livingston", "Virgo", "Kagra", "Indigo", "CE", "ET1".."ET3";
 *                        src/detector_util.cpp:1083-1158)
 *   frequencies[L]       Hz
 *   psd[D*L]             detector-major, as the gwatpy flat layout (src/gwatpy_wrapping.cpp:125-130)
 *   data_re/im[D*L]      frequency-domain strain; may both be NULL when only waveforms/Fishers are wanted
 *   weights[L]           quadrature weights for "GAUSSLEG" (NULL for "SIMPSONS")
 *   integration_method   "SIMPSONS" | "GAUSSLEG"   (Log_Likelihood_internal, src/mcmc_gw.cpp:801-868)
 *   log10F               GAUSSLEG nodes are in log10(f)  (src/mcmc_gw.cpp:822-826)
 */
int gwat_b200_set_network(gwat_b200_ctx *ctx, int num_detectors, const char *const *detectors, int length,
                          const double *frequencies, const double *psd, const double *data_re, const double *data_im,
                          const double *weights, const char *integration_method, int log10F);

/* ---- log-likelihood --------------------------------------------------------------------------------------------- */

/*
 * W evaluations of MCMC_likelihood_wrapper (src/mcmc_gw.cpp:2569-2791, extrinsic branch) in one call:
 *   MCMC_prep_params (:2492) -> repack_parameters("MCMC_"+method) (src/fisher.cpp:2167) -> MCMC_likelihood_extrinsic (:2374).
 *   params[W*dimension]  row-major sampling vectors (RA, sin DEC, psi, cos iota, phiRef, tc, ln DL, ln Mc, eta, ... )
 *   gmst                 replaces the global mcmc_gmst
 *   T_segment            segment duration; the reference forms tc_ref = T - tc with T = 1/(frequencies[1]-frequencies[0])
 *                        taken on a double** (pointer arithmetic, src/mcmc_gw.cpp:2466) -- here T is explicit.
 *   logL[W]              out
 * Host pointers.  NaN/invalid parameter points give NaN (the samplers reject those, src/mcmc_sampler_internals.cpp:101).
 */
int gwat_b200_loglike_mcmc_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                                 int dimension, int W, const double *params, double gmst, double T_segment,
                                 double *logL);
/* Same, with params/logL already in device memory of ctx's GPU; asynchronous on `stream` (NULL = the context's own stream).
 * The per-walker coefficient scratch belongs to the context: calls on ONE context must be ordered on one stream (or separated
 * by stream synchronisation); independent streams need independent contexts.  The host-buffer entry points of this header
 * are safe to call concurrently from several threads on one context (each call is one critical section, held from the upload of its
 * inputs to the download of its results: concurrent callers are SERIALISED, not overlapped -- a pool of threads with one walker per
 * call gets its throughput from gwat_b200_queue_* below, which merges the calls in flight into one batch, or from one context per
 * thread). */
int gwat_b200_loglike_mcmc_batch_dev(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                                     int dimension, int W, const double *d_params, double gmst, double T_segment,
                                     double *d_logL, void *stream);

/*
 * W evaluations of the body of MCMC_likelihood_extrinsic below its tc_ref line (src/mcmc_gw.cpp:2473-2486):
 * create_coherent_GW_detection (src/waveform_util.cpp:129) + sum_d Log_Likelihood_internal (src/mcmc_gw.cpp:801),
 * from physical parameters; sources[w].tc is used as given (i.e. it is the reference's tc_ref).
 */
int gwat_b200_loglike_batch(gwat_b200_ctx *ctx, const char *generation_method, int W, const gwat_b200_source *sources,
                            double *logL);

/*
 * One-chain-per-call front of gwat_b200_loglike_mcmc_batch: keeps the shape of the reference's likelihood callback
 *   double ll(double *param, int *status, int model_status, mcmc_data_interface *interface, void *parameters)
 * (include/gwat/mcmc_sampler_internals.h:169-171, bound to MCMC_likelihood_wrapper, src/mcmc_gw.cpp:2569), which the
 * samplers call concurrently from the workers of a thread pool, one chain per call (src/mcmc_sampler.cpp:347-447).
 * Calls that are in flight at the same moment are merged into ONE batched launch: the first caller waits until
 * `expected_callers` calls have joined (the pool's thread count), the batch holds `max_batch` vectors, or `max_wait_us`
 * microseconds have passed, then evaluates the whole group; every caller gets its own chain's value back.  Grouping never
 * changes a value.  The queue replaces the file-static globals mcmc_generation_method, mcmc_mod_struct, mcmc_gmst
 * (include/gwat/mcmc_gw.h:22-46) for the calls made through it; the network comes from `ctx` (gwat_b200_set_network).
 */
typedef struct gwat_b200_queue gwat_b200_queue;
int gwat_b200_queue_create(gwat_b200_queue **queue, gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                           int dimension, double gmst, double T_segment, int max_batch, int expected_callers, double max_wait_us);
void gwat_b200_queue_destroy(gwat_b200_queue *queue);
/* Blocking; safe to call from any number of threads.  Returns the chain's log-likelihood, or NaN when the point is
 * unphysical or the batched call failed (`*status`, may be NULL, then holds the gwat_b200_status; 0 on success). */
double gwat_b200_queue_loglike(gwat_b200_queue *queue, const double *param, int *status);
/* How many one-chain calls were served, in how many batched launches, and the largest group so far (any may be NULL). */
int gwat_b200_queue_stats(gwat_b200_queue *queue, long long *calls, long long *batches, int *largest_batch);

/*
 * Matched-filter signal-to-noise ratios of W templates in the context's network,
 *   snr[w] = sqrt( sum_d 4 int |r_d(f)|^2 / S_d(f) df ),
 * by the network's quadrature rule.  With a one-detector network this is calculate_snr(sensitivity_curve, detector,
 * generation_method, params, frequencies, length, integration_method, weights, log10_freq) of the reference
 * (src/waveform_util.cpp:290-344 -> calculate_snr_internal :479-510) with psd = populate_noise(curve)^2.  Data, if the
 * network has any, are ignored.  Same kernels as the likelihood.
 */
int gwat_b200_snr_batch(gwat_b200_ctx *ctx, const char *generation_method, int W, const gwat_b200_source *sources, double *snr);

/*
 * populate_noise (src/detector_util.cpp:87-282): amplitude spectral density sqrt(S_n) of a named noise curve at the given
 * frequencies.  Analytic models "aLIGO_analytic", "Hanford_O1_fitted"; tabulated curves ("AdLIGODesign", "AdLIGOAPlus",
 * "CE1", "CE2", "AdVIRGOPlus2_opt", "KAGRA_opt", "ET-D", "AdLIGOVoyager", ... and their "_smoothed" variants) are read from
 * the reference's two-column CSV files in `noise_data_dir` (GWAT installs them as GWAT_SHARE_DIR/noise_data; in its source
 * tree: data/noise_data/currently_supported) and interpolated linearly like gsl_interp_linear.  Frequencies outside a
 * table give NaN and GWAT_B200_ERR_ARG (GSL aborts there); LISA curves are GWAT_B200_ERR_UNSUPPORTED.  Host code, run once
 * per analysis before gwat_b200_set_network; noise_data_dir may be NULL for the analytic models.
 */
int gwat_b200_populate_noise(const double *frequencies, const char *curve, const char *noise_data_dir, int length,
                             double *noise_root);

/*
 * LOSC/GWOSC text files -> the inputs of gwat_b200_set_network: allocate_LOSC_data (src/io_util.cpp:523-661).
 *   data_files[D]          one strain text file per detector (read_LOSC_data_file, :403-464: three header lines -- sampling rate
 *                          on the second, GPS start and duration on the third -- then the samples)
 *   psd_file               header line, then rows "f S_1 ... S_D" (read_LOSC_PSD_file, :466-500); its spacing df fixes the
 *                          observation time T_obs = 1/df, its range the bins that are kept
 *   trigger_time (GPS), post_merger_duration (s): the segment is (trigger - (T_obs - post), trigger + post]
 *   capacity               rows the output arrays can hold; *length returns the rows of the PSD file (call with capacity 0 to ask)
 *   frequencies[length], psd[D*length], data_re/im[D*length]   detector-major, as gwat_b200_set_network takes them
 * The segment is Tukey-windowed (alpha = 0.8/T_obs), transformed (batched cuFFT, forward, as the reference's FFTW plan) and
 * multiplied by dt on the GPU.  Unlike the reference, missing files and a trigger outside the file are error codes.
 */
int gwat_b200_losc_prepare(gwat_b200_ctx *ctx, int num_detectors, const char *const *data_files, const char *psd_file,
                           double trigger_time, double post_merger_duration, int capacity, int *length, double *frequencies,
                           double *psd, double *data_re, double *data_im);

/* ---- waveforms and detector responses ---------------------------------------------------------------------------- */

/* Gauss-Legendre frequency grid as the reference builds it for "GAUSSLEG" integration (gauleg, src/ortho_basis.cpp:14-48, used as
 * in src/waveform_util.cpp:3113-3116): n nodes and weights on [f_lower, f_upper], or -- log10F != 0 -- nodes uniform-in-rule over
 * log10 f, returned as frequencies, with the weights of the log10 f integral (pass the same log10F to gwat_b200_set_network).
 * Host code; no context needed. */
int gwat_b200_gauss_legendre_grid(double f_lower, double f_upper, int n, int log10F, double *frequencies, double *weights);

/* tc/phic-maximised log-likelihood of the reference's "intrinsic" samplers, summed over the network's detectors:
 *   maximized_Log_Likelihood_aligned_spin_internal (src/mcmc_gw.cpp:595-652) for the IMRPhenomD family,
 *   maximized_Log_Likelihood_unaligned_spin_internal (src/mcmc_gw.cpp:660-795) for the IMRPhenomPv2 family,
 * called as the intrinsic branches of MCMC_likelihood_wrapper do (src/mcmc_gw.cpp:2603-2722): the sources' psi, phiRef,
 * incl_angle, tc, f_ref are overridden (0, 1, 0, 1, 10 or 20) and the response is the + polarisation (F+ = 1, Fx = 0).
 * Needs a uniform grid, Simpson's rule and data (the time axis is an FFT of length L; cuFFT replaces FFTW).  Like the
 * reference's, the value is not normalised: constant (d|d) terms are left out. */
int gwat_b200_loglike_maximized_batch(gwat_b200_ctx *ctx, const char *generation_method, int W, const gwat_b200_source *sources,
                                      double *logL);

/* The same for W sampling vectors of the reference's INTRINSIC sets, as the intrinsic branch of MCMC_likelihood_wrapper evaluates
 * one of them (src/mcmc_gw.cpp:2569-2722): MCMC_prep_params with mcmc_intrinsic (sky_average = true, :2494), then
 * repack_parameters("MCMC_" + method) (src/fisher.cpp:2308-2376, 2420-2431).  The sets (PTMCMC_method_specific_prep,
 * src/mcmc_gw.cpp:1880-1985): ln chirpmass, eta, chi1, chi2 for the IMRPhenomD family (+ ln tidal_s, or ln tidal1, ln tidal2, for
 * IMRPhenomD_NRT); ln chirpmass, eta, a1, a2, cos tilt1, cos tilt2, phi1, phi2 for IMRPhenomPv2; then the modifications. */
int gwat_b200_loglike_maximized_mcmc_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod, int dimension,
                                           int W, const double *params, double gmst, double *logL);

/*
 * W evaluations of fourier_waveform<double> (src/waveform_generator.cpp:104-294) on the context's grid.
 * Outputs are split real/imag like fourier_waveform_py (src/gwatpy_wrapping.cpp), shape [W*L] row-major; any may be NULL.
 */
int gwat_b200_fourier_waveform_batch(gwat_b200_ctx *ctx, const char *generation_method, int W,
                                     const gwat_b200_source *sources, double *hplus_re, double *hplus_im,
                                     double *hcross_re, double *hcross_im);

/*
 * W evaluations of fourier_amplitude<double> and fourier_phase<double> (src/waveform_generator.cpp:537-670, 672-814) for the
 * IMRPhenomD family (IMRPhenomD, ppE_IMRPhenomD_Inspiral/_IMR and the theories mapped onto them, gIMRPhenomD):
 * IMRPhenomD::construct_amplitude / construct_phase (src/IMRPhenomD.cpp:604-740).  amplitude/phase shape [W*L]; either may
 * be NULL.  The amplitude is zero above 0.2/M; the phase, like the reference's, has no cutoff.  For the theory-mapped
 * methods (dCS_, EdGB_ ...) the phase carries the mapped ppE term exactly as fourier_waveform does; the reference's deprecated
 * four-argument fourier_phase silently returns the GR phase there.  Other families: GWAT_B200_ERR_UNSUPPORTED.
 */
int gwat_b200_fourier_amplitude_phase_batch(gwat_b200_ctx *ctx, const char *generation_method, int W,
                                            const gwat_b200_source *sources, double *amplitude, double *phase);

/*
 * W evaluations of create_coherent_GW_detection_reuse_WF (src/waveform_util.cpp:153-184): responses of all D detectors
 * of the network including the inter-detector time-of-arrival phase.  resp_re/im shape [W*D*L].
 */
int gwat_b200_coherent_response_batch(gwat_b200_ctx *ctx, const char *generation_method, int W,
                                      const gwat_b200_source *sources, double *resp_re, double *resp_im);

/*
 * W evaluations of fourier_detector_response<double> (src/waveform_util.cpp:1070-1088) for ONE named detector, no
 * time-of-arrival shift.  resp_re/im shape [W*L].  Per source, as the reference's wrapper does: horizon_coord != 0 -> the
 * detector-frame patterns of (theta, phi, psi) (fourier_detector_response_horizon, :684-720); else the equatorial branch
 * (:936-985), where equatorial_orientation != 0 first derives incl_angle and psi from (theta_l, phi_l)
 * (transform_orientation_coords).  The coherent response above and the likelihoods read incl_angle / psi / RA / DEC as given,
 * as create_coherent_GW_detection does.
 */
int gwat_b200_fourier_detector_response_batch(gwat_b200_ctx *ctx, const char *generation_method, const char *detector,
                                              int W, const gwat_b200_source *sources, double *resp_re,
                                              double *resp_im);

/* ---- Fisher matrices --------------------------------------------------------------------------------------------- */

/*
 * S evaluations of fisher_numerical (src/fisher.cpp:81-145) for one detector of the network against a reference
 * detector, central differences of `order` 2 or 4 with the reference's epsilon = 1e-8 (src/fisher.cpp:361).
 *   generation_method    as the reference takes it, e.g. "IMRPhenomD" or "MCMC_IMRPhenomD"
 *   detector_index       which detector of the network supplies the antenna pattern and the PSD
 *   reference_index      the detector tc refers to (the reference's reference_detector)
 *   fisher[S*dim*dim]    out, row-major per source
 * With detector_index < 0 the matrices of all detectors are summed (MCMC_fisher_wrapper, src/mcmc_gw.cpp:2298-2312).
 * Sources with sky_average set take the sky-averaged branch of calculate_derivatives (src/fisher.cpp:183-338): "IMRPhenomD",
 * dimension 7 (ln A0, phic, tc, ln chirpmass, ln eta, chi_s, chi_a), derivatives of amplitude and phase, detector_index >= 0
 * naming the PSD; all sources of a batch must agree on the flag.  "MCMC_" + method with sky-averaged sources: the INTRINSIC sets of the
 * tc/phic-maximised samplers -- ln chirpmass, eta, chi1, chi2 (+ modifications) through the same amplitude / phase branch
 * (src/fisher.cpp:2000-2013, 2360-2376); "MCMC_IMRPhenomPv2", dimension 8 (ln chirpmass, eta, a1, a2, cos tilt1, cos tilt2, phi1, phi2)
 * through the response branch with the extrinsic members at the reference's constants (:1968-1990, 2308-2352; detector_index as for
 * pointed sources).  gwat_b200_repack_mcmc_intrinsic_batch makes such records from sampling vectors.
 */
int gwat_b200_fisher_numerical_batch(gwat_b200_ctx *ctx, const char *generation_method, int detector_index,
                                     int reference_index, int dimension, int order, int S,
                                     const gwat_b200_source *sources, double *fisher);

/* ---- helpers that mirror small reference utilities (evaluated on the GPU like everything else) ------------------- */

/* repack_parameters<double> for the "MCMC_"+method parameterisations (src/fisher.cpp:2167-2507), after the dCS/EdGB unit
 * change of MCMC_prep_params (src/mcmc_gw.cpp:2560-2565): sampling vectors -> physical records. */
int gwat_b200_repack_mcmc_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                                int dimension, int W, const double *params, double gmst, gwat_b200_source *sources);
/* ... and for the intrinsic sets (see gwat_b200_loglike_maximized_mcmc_batch): sky_average = 1 in the records, the members the set
 * does not hold at the constants of the reference (D_L = 1000 Mpc, or 100 Mpc and iota = pi/4 for IMRPhenomPv2; src/fisher.cpp:2308-2376). */
int gwat_b200_repack_mcmc_intrinsic_batch(gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                                          int dimension, int W, const double *params, double gmst, gwat_b200_source *sources);

/* Antenna patterns and time-of-arrival differences for W sky positions:
 * detector_response_functions_equatorial (src/detector_util.cpp:1037) and DTOA_DETECTOR (:677) for every detector of the
 * network relative to detector 0.  Outputs shape [W*D]. */
int gwat_b200_antenna_batch(gwat_b200_ctx *ctx, int W, const double *RA, const double *DEC, const double *psi,
                            double gmst, double *Fplus, double *Fcross, double *dtoa);

/* gps_to_GMST_radian (src/util.cpp:1793-1846): the `gmst` argument of the likelihood calls from a GPS time.  Host code. */
double gwat_b200_gps_to_gmst_radian(double gps_time);

/* ---- introspection used by bench.py / the tests ------------------------------------------------------------------ */

/*
 * Log_Likelihood_internal (include/gwat/mcmc_gw.h:276-284, src/mcmc_gw.cpp:801-868) for ONE detector and a response the caller
 * already holds in host memory: -1/2 ((r|r) - 2 (d|r)) with the reference's quadrature (SIMPSONS: 1,4,2,...,4,1 by index parity
 * and delta_f from the middle of the array; GAUSSLEG: weights, times f ln 10 when log10F).  Exists for callers that built the
 * response themselves (fourier_detector_response); the batched likelihoods above never materialise responses.  Does not use
 * or change the network set on the context.
 */
int gwat_b200_log_likelihood_internal(gwat_b200_ctx *ctx, int L, const double *frequencies, const double *psd, const double *data_re,
                                      const double *data_im, const double *weights, const char *integration_method, int log10F,
                                      const double *response_re, const double *response_im, double *logL);

/*
 * match(data1, data2, SN, frequencies, length) (include/gwat/waveform_util.h, src/waveform_util.cpp:41-89; gwatpy's match_py):
 * 4 max_t |IFFT(conj(d1) d2 / S)| delta_f / (||d1|| ||d2||), norms by Simpson's rule with delta_f = f[1] - f[0].  `psd` is S(f)
 * (the reference's argument `SN` is used as a power spectral density).  Uniform grid; one batched cuFFT where the reference
 * plans one FFTW transform per call.  Does not use or change the network set on the context.
 */
int gwat_b200_match(gwat_b200_ctx *ctx, int L, const double *frequencies, const double *psd, const double *data1_re, const double *data1_im,
                    const double *data2_re, const double *data2_im, double *match);

/*
 * How a generation_method string is read (the reference re-derives this with std::string::find in check_mod /
 * check_theory_support, src/ppE_utilities.cpp:65-134, and in MCMC_prep_params, src/mcmc_gw.cpp:2517-2565):
 *   ppe_like     the trailing sampling dimensions are ppE betas (ppE_* and every theory-mapped method)
 *   gimr         they are gIMR fractional deviations
 *   alpha_units  the first of them is sqrt(alpha) in km and is converted to alpha^2 in s^4 (dCS / EdGB family)
 * Returns GWAT_B200_ERR_METHOD for a string this library does not implement.  Any pointer may be NULL.
 */
int gwat_b200_method_info(const char *generation_method, int *ppe_like, int *gimr, int *alpha_units, int *pv2, int *nrt);

/* Measured FP64 FMA issue peak of ctx's GPU in TFLOP/s (a DFMA-chain microbenchmark; the roofline denominator of the
 * FP64-bound kernels -- the driver-written MEASURED_PEAKS.json carries no FP64 figure). */
int gwat_b200_measure_fp64_peak(gwat_b200_ctx *ctx, double *tflops);
/* Number of kernels this library has launched on ctx since creation (for the bench's gpu_launches). */
long long gwat_b200_launch_count(const gwat_b200_ctx *ctx);
/* Device-time (ms, CUDA events on the context's stream) of the hot kernel launches of the last *_batch call.  For the
 * likelihood entry points the events are recorded only after gwat_b200_set_kernel_timing(ctx, 1) (two events cost ~6 us per
 * call on the stream, 4 % of a 1024-walker pass); otherwise 0.  Fisher batches always report it. */
int gwat_b200_set_kernel_timing(gwat_b200_ctx *ctx, int on);
double gwat_b200_last_kernel_ms(const gwat_b200_ctx *ctx);
/* Active (walker,bin) pairs (f below the model's cutoff) evaluated by the last likelihood call. */
long long gwat_b200_last_active_bins(const gwat_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GWAT_B200_H */
