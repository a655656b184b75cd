// libgwat_b200_dropin.so -- the reference's OWN C++ entry points for T = double, defined on top of the C ABI of gwat_b200.h.
//
// This translation unit is compiled against GWAT's headers (-I$(REF)/include/gwat at build time; the third-party headers those
// include come from standins/ in this image, from the real ADOL-C / GSL / FFTW installations at a GWAT site) and DEFINES the
// symbols a GWAT program links against for the hot path, with the reference's exact signatures and therefore its exact mangled
// names:
//
//   fourier_waveform<double>(double*, int, waveform_polarizations<double>*, std::string, gen_params_base<double>*)
//                                                          include/gwat/waveform_generator.h:43-49   src/waveform_generator.cpp:104
//   fourier_waveform(double*, int, std::complex<double>*, std::string, gen_params*)   and the two split real/imag overloads  :51-81
//   fourier_detector_response<double>(double*, int, std::complex<double>*, std::string, std::string, gen_params_base<double>*, double*)
//                                                          include/gwat/waveform_util.h:191          src/waveform_util.cpp:1070
//   create_coherent_GW_detection<double>, create_coherent_GW_detection_reuse_WF<double>       waveform_util.h:16-33   :128-184
//   Log_Likelihood_internal(...)                           include/gwat/mcmc_gw.h:276                src/mcmc_gw.cpp:801
//   MCMC_likelihood_extrinsic(...)                         include/gwat/mcmc_gw.h:473                src/mcmc_gw.cpp:2374
//   MCMC_likelihood_wrapper(double*, mcmc_data_interface*, void*)   mcmc_gw.h:491                    src/mcmc_gw.cpp:2569
//   MCMC_fisher_wrapper(double*, double**, mcmc_data_interface*, void*)   mcmc_gw.h:487              src/mcmc_gw.cpp:2230
//   fisher_numerical(...)                                  include/gwat/fisher.h:35                  src/fisher.cpp:81
//
// A program written against GWAT (e.g. the reference's examples/waveform_output/src/waveform.cpp, which tests/ builds unmodified)
// that puts this library BEFORE libgwat on its link line gets these calls from the GPU and everything else from GWAT.
// Everything numerical happens behind the C ABI (CUDA kernels); this file only converts GWAT's argument types.
//
// Behaviour notes
//  * One process-wide context on device $GWAT_B200_DEVICE (default 0); calls are serialised by a mutex.  Throughput comes from the
//    batched entry points of gwat_b200.h / gwat_b200::CallbackQueue; these symbols exist so that existing callers keep working.
//  * The network (grid, PSDs, data) a call describes is uploaded once and reused while the caller passes the same arrays
//    (address, length and a 64-sample checksum per array); gwat_b200_dropin_invalidate() forgets it.
//  * MCMC_likelihood_extrinsic: the reference computes the segment duration as `1./(frequencies[1]-frequencies[0])` on a double**
//    (src/mcmc_gw.cpp:2466), i.e. from the distance between two ARRAYS.  By default this library does literally the same, so that
//    the same call gives the same number; gwat_b200_dropin_set_segment_duration(T > 0) replaces it by an explicit duration.
//  * MCMC_likelihood_wrapper / MCMC_fisher_wrapper read GWAT's file-static "globals" (static in include/gwat/mcmc_gw.h:22-46, so
//    every translation unit has its own copy; the reference's samplers set the copy of src/mcmc_gw.cpp).  This unit's copy is set
//    with gwat_b200_dropin_bind_mcmc(); a GWAT build that compiles this file in place of the wrappers' bodies needs no such call.
//  * Errors: status 0 / NaN, as gwat_b200_cxx.hpp; nothing calls exit().
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "fisher.h"
#include "mcmc_gw.h"
#include "mcmc_sampler_internals.h"
#include "util.h"
#include "waveform_generator.h"
#include "waveform_util.h"

#include "../../include/gwat_b200_cxx.hpp"
#include "../../include/gwat_b200_sampler.h"

namespace {

const double kNaN = std::numeric_limits<double>::quiet_NaN();

struct Session {
	std::mutex mu;
	gwat_b200::Engine *engine = nullptr;
	uint64_t net_key = 0;
	double T_override = -1;
};
Session &session()
{
	static Session s;
	return s;
}
gwat_b200::Engine *engine_locked(Session &S)
{
	if (!S.engine) {
		int dev = 0;
		if (const char *e = std::getenv("GWAT_B200_DEVICE")) dev = std::atoi(e);
		S.engine = new gwat_b200::Engine(dev);
	}
	return S.engine->ok() ? S.engine : nullptr;
}

uint64_t mix(uint64_t h, uint64_t w)
{
	h ^= w;
	h *= 1099511628211ULL;
	return h;
}
uint64_t array_key(uint64_t h, const void *p, size_t n_doubles)
{
	h = mix(h, (uint64_t)(uintptr_t)p);
	h = mix(h, (uint64_t)n_doubles);
	if (!p || n_doubles == 0) return h;
	const double *a = static_cast<const double *>(p);
	const size_t step = n_doubles > 64 ? n_doubles / 64 : 1;
	for (size_t i = 0; i < n_doubles; i += step) {
		uint64_t w;
		std::memcpy(&w, a + i, sizeof(w));
		h = mix(h, w);
	}
	uint64_t w;
	std::memcpy(&w, a + (n_doubles - 1), sizeof(w));
	return mix(h, w);
}

// Upload (or keep) the network described by GWAT's per-detector arrays.  data / psd / weights may be NULL.
bool ensure_network(Session &S, gwat_b200::Engine *e, const std::string *detectors, int D, int L, double **frequencies, double **psd,
                    std::complex<double> **data, double **weights, const std::string &integration_method, bool log10F)
{
	uint64_t key = 1469598103934665603ULL;
	key = mix(key, (uint64_t)D);
	key = mix(key, (uint64_t)L);
	for (int d = 0; d < D; d++) {
		for (char c : detectors[d]) key = mix(key, (uint64_t)(unsigned char)c);
		key = mix(key, 0xffu);
		key = array_key(key, frequencies[d], (size_t)L);
		key = array_key(key, psd ? psd[d] : nullptr, psd ? (size_t)L : 0);
		key = array_key(key, data ? data[d] : nullptr, data ? (size_t)2 * L : 0);
	}
	const bool gl = integration_method == "GAUSSLEG";
	key = array_key(key, (gl && weights) ? weights[0] : nullptr, (gl && weights) ? (size_t)L : 0);
	for (char c : integration_method) key = mix(key, (uint64_t)(unsigned char)c);
	key = mix(key, log10F ? 1u : 0u);
	if (key == S.net_key) return true;
	S.net_key = 0;
	if (e->set_network(detectors, D, L, frequencies, psd, data, weights, integration_method, log10F) != 0) return false;
	S.net_key = key;
	return true;
}

bool same_length(const int *lengths, int D)
{
	for (int d = 1; d < D; d++)
		if (lengths[d] != lengths[0]) return false;  // the reference supports the shared grid only, too (src/waveform_util.cpp:140-145)
	return true;
}

// the reference's `1./(frequencies[1]-frequencies[0])` on a double** (src/mcmc_gw.cpp:2466), or the explicit override
double segment_duration(Session &S, double **frequencies, int D)
{
	if (S.T_override > 0) return S.T_override;
	if (D < 2) return kNaN;  // (the reference reads frequencies[1] of a one-element array here: undefined)
	return 1. / (double)(frequencies[1] - frequencies[0]);
}

void to_mod(const MCMC_modification_struct *m, gwat_b200_mod &out, bool &ok)
{
	gwat_b200_mod_init(&out);
	ok = true;
	if (!m) return;
	const int counts[5] = {m->ppE_Nmod, m->gIMR_Nmod_phi, m->gIMR_Nmod_sigma, m->gIMR_Nmod_beta, m->gIMR_Nmod_alpha};
	for (int n : counts)
		if (n > GWAT_B200_MAX_MOD) ok = false;
	if (!ok) return;
	out.ppE_Nmod = m->ppE_Nmod;
	for (int i = 0; i < m->ppE_Nmod; i++) out.bppe[i] = m->bppe[i];
	out.gIMR_Nmod_phi = m->gIMR_Nmod_phi;
	out.gIMR_Nmod_sigma = m->gIMR_Nmod_sigma;
	out.gIMR_Nmod_beta = m->gIMR_Nmod_beta;
	out.gIMR_Nmod_alpha = m->gIMR_Nmod_alpha;
	for (int i = 0; i < m->gIMR_Nmod_phi; i++) out.gIMR_phii[i] = m->gIMR_phii[i];
	for (int i = 0; i < m->gIMR_Nmod_sigma; i++) out.gIMR_sigmai[i] = m->gIMR_sigmai[i];
	for (int i = 0; i < m->gIMR_Nmod_beta; i++) out.gIMR_betai[i] = m->gIMR_betai[i];
	for (int i = 0; i < m->gIMR_Nmod_alpha; i++) out.gIMR_alphai[i] = m->gIMR_alphai[i];
	out.NSflag1 = m->NSflag1;
	out.NSflag2 = m->NSflag2;
	out.tidal_love = m->tidal_love;
	out.tidal_love_error = m->tidal_love_error;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
// waveforms and responses
// ---------------------------------------------------------------------------------------------------------------------------

template <>
int fourier_waveform<double>(double *frequencies, int length, waveform_polarizations<double> *wp, std::string generation_method,
                             gen_params_base<double> *parameters)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e || !wp) return 0;
	S.net_key = 0;  // (fourier_waveform sets its own one-detector grid)
	return gwat_b200::fourier_waveform(*e, frequencies, length, wp->hplus, wp->hcross, generation_method, parameters);
}

int fourier_waveform(double *frequencies, int length, std::complex<double> *waveform, std::string generation_method, gen_params *parameters)
{
	// "legacy" overload (src/waveform_generator.cpp:406-494): construct_waveform of the aligned-spin models WITHOUT the inclination
	// factors, for four methods only; any other method leaves `waveform` untouched and returns 1, as the reference does.  The raw
	// (2,2) waveform is the plus polarisation seen face-on: cos(0) = 1 and .5 (1 + 1) = 1 exactly, so the product below is exact.
	if (generation_method != "IMRPhenomD" && generation_method != "ppE_IMRPhenomD_Inspiral" && generation_method != "ppE_IMRPhenomD_IMR" &&
	    generation_method != "IMRPhenomD_NRT")
		return 1;
	gen_params face_on = *parameters;
	face_on.incl_angle = 0;
	face_on.sky_average = false;
	waveform_polarizations<double> wp;
	std::vector<std::complex<double>> hc(length > 0 ? length : 0);
	wp.hplus = waveform;
	wp.hcross = hc.data();
	return fourier_waveform<double>(frequencies, length, &wp, generation_method, &face_on);
}

int fourier_waveform(double *frequencies, int length, double *waveform_real, double *waveform_imag, std::string generation_method,
                     gen_params *parameters)
{
	std::vector<std::complex<double>> h(length > 0 ? length : 0);
	const int st = fourier_waveform(frequencies, length, h.data(), generation_method, parameters);
	for (int i = 0; i < length; i++) {
		waveform_real[i] = h[i].real();
		waveform_imag[i] = h[i].imag();
	}
	return st;
}

int fourier_waveform(double *frequencies, int length, double *waveform_plus_real, double *waveform_plus_imag, double *waveform_cross_real,
                     double *waveform_cross_imag, double *waveform_x_real, double *waveform_x_imag, double *waveform_y_real,
                     double *waveform_y_imag, double *waveform_b_real, double *waveform_b_imag, double *waveform_l_real,
                     double *waveform_l_imag, std::string generation_method, gen_params *parameters)
{
	std::vector<std::complex<double>> hp(length > 0 ? length : 0), hc(length > 0 ? length : 0);
	waveform_polarizations<double> wp;
	wp.hplus = hp.data();
	wp.hcross = hc.data();
	const int st = fourier_waveform<double>(frequencies, length, &wp, generation_method, parameters);
	for (int i = 0; i < length; i++) {
		waveform_plus_real[i] = hp[i].real();
		waveform_plus_imag[i] = hp[i].imag();
		waveform_cross_real[i] = hc[i].real();
		waveform_cross_imag[i] = hc[i].imag();
	}
	// the vector and scalar polarisations belong to models outside this path (EA_IMRPhenomD_NRT): not touched, as the
	// reference leaves them for the tensor-only models (src/waveform_generator.cpp:352-377)
	(void)waveform_x_real; (void)waveform_x_imag; (void)waveform_y_real; (void)waveform_y_imag;
	(void)waveform_b_real; (void)waveform_b_imag; (void)waveform_l_real; (void)waveform_l_imag;
	return st;
}

template <>
int fourier_detector_response<double>(double *frequencies, int length, std::complex<double> *response, std::string detector,
                                      std::string generation_method, gen_params_base<double> *parameters, double *times)
{
	(void)times;  // space detectors only (src/waveform_util.cpp:1081-1086)
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return 0;
	S.net_key = 0;
	const int st = gwat_b200::fourier_detector_response(*e, frequencies, length, response, detector, generation_method, parameters);
	if (parameters->equatorial_orientation && !parameters->horizon_coord) {
		// the reference derives incl_angle and psi from (theta_l, phi_l) IN the caller's object (transform_orientation_coords,
		// src/waveform_util.cpp:947-949); callers read them back afterwards
		gwat_b200_source s;
		if (gwat_b200::flatten(*parameters, s) && gwat_b200_transform_orientation_coords(generation_method.c_str(), 1, &s) == 0) {
			parameters->incl_angle = s.incl_angle;
			parameters->psi = s.psi;
		}
	}
	return st;
}

// calculate_snr(sensitivity_curve, detector, generation_method, params, frequencies, length, integration_method, weights, log10_freq),
// src/waveform_util.cpp:290-344: sqrt(4 int |response|^2 / S_n) with S_n = populate_noise(curve)^2 (the curve's tables, if it is a tabulated
// one, from the directory GWAT_B200_NOISE_DIR names -- this library does not carry the reference's data files).  Like the reference, a
// source given by the equatorial direction of L has incl_angle and psi written into the caller's object.
double calculate_snr(std::string sensitivity_curve, std::string detector, std::string generation_method, gen_params_base<double> *params,
                     double *frequencies, int length, std::string integration_method, double *weights, bool log10_freq)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e || length < 1) return kNaN;
	std::vector<double> psd(length);
	if (gwat_b200_populate_noise(frequencies, sensitivity_curve.c_str(), std::getenv("GWAT_B200_NOISE_DIR"), length, psd.data()) != 0) return kNaN;
	for (double &v : psd) v *= v;
	S.net_key = 0;
	double *f1[1] = {frequencies}, *p1[1] = {psd.data()}, *w1[1] = {weights};
	if (e->set_network(&detector, 1, length, f1, p1, nullptr, w1, integration_method, log10_freq) != 0) return kNaN;
	gwat_b200_source s;
	if (!gwat_b200::flatten(*params, s)) return kNaN;
	double snr = kNaN;
	if (gwat_b200_snr_batch(e->ctx(), generation_method.c_str(), 1, &s, &snr) != 0) return kNaN;
	if (params->equatorial_orientation && !params->horizon_coord &&
	    gwat_b200_transform_orientation_coords(generation_method.c_str(), 1, &s) == 0) {
		params->incl_angle = s.incl_angle;
		params->psi = s.psi;
	}
	return snr;
}

template <>
void create_coherent_GW_detection_reuse_WF<double>(std::string *detectors, int detector_N, double *frequencies, int lengths,
                                                   gen_params_base<double> *gen_params, std::string generation_method,
                                                   std::complex<double> **responses)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return;
	std::vector<double *> f(detector_N, frequencies);
	std::vector<int> L(detector_N, lengths);
	S.net_key = 0;
	gwat_b200::create_coherent_GW_detection(*e, detectors, detector_N, f.data(), L.data(), true, gen_params, generation_method, responses);
}

template <>
void create_coherent_GW_detection<double>(std::string *detectors, int detector_N, double **frequencies, int *lengths, bool reuse_WF,
                                          gen_params_base<double> *gen_params, std::string generation_method,
                                          std::complex<double> **responses)
{
	(void)reuse_WF;  // one waveform evaluation serves all detectors either way
	if (!same_length(lengths, detector_N)) return;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return;
	S.net_key = 0;
	gwat_b200::create_coherent_GW_detection(*e, detectors, detector_N, frequencies, lengths, true, gen_params, generation_method, responses);
}

// ---------------------------------------------------------------------------------------------------------------------------
// likelihoods
// ---------------------------------------------------------------------------------------------------------------------------

// -1/2 ((h|h) - 2 (d|h)) for a response the CALLER supplies (src/mcmc_gw.cpp:801-868).  On this path responses never exist
// in host memory (MCMC_likelihood_extrinsic below evaluates everything on the device); this symbol serves callers that made
// one with fourier_detector_response.  The two weighted sums are formed by the library's inner-product kernel.
double Log_Likelihood_internal(std::complex<double> *data, double *psd, double *frequencies, double *weights,
                               std::complex<double> *detector_response, int length, bool log10F, std::string integration_method)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e || length < 4) return kNaN;
	std::vector<double> dr(length), di(length), rr(length), ri(length);
	for (int i = 0; i < length; i++) {
		dr[i] = data[i].real();
		di[i] = data[i].imag();
		rr[i] = detector_response[i].real();
		ri[i] = detector_response[i].imag();
	}
	double out = kNaN;
	if (gwat_b200_log_likelihood_internal(e->ctx(), length, frequencies, psd, dr.data(), di.data(), weights, integration_method.c_str(),
	                                      log10F ? 1 : 0, rr.data(), ri.data(), &out) != 0)
		return kNaN;
	S.net_key = 0;
	return out;
}

double MCMC_likelihood_extrinsic(bool save_waveform, gen_params_base<double> *parameters, std::string generation_method, int *data_length,
                                 double **frequencies, std::complex<double> **data, double **psd, double **weights,
                                 std::string integration_method, bool log10F, std::string *detectors, int num_detectors)
{
	(void)save_waveform;
	if (!same_length(data_length, num_detectors)) return kNaN;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return kNaN;
	if (!ensure_network(S, e, detectors, num_detectors, data_length[0], frequencies, psd, data, weights, integration_method, log10F)) return kNaN;
	const double T = segment_duration(S, frequencies, num_detectors);
	parameters->tc = T - parameters->tc;  // tc_ref, written back into the caller's object as the reference does (:2467, 2473)
	gwat_b200_source s;
	if (!gwat_b200::flatten(*parameters, s)) return kNaN;
	double ll = kNaN;
	if (gwat_b200_loglike_batch(e->ctx(), generation_method.c_str(), 1, &s, &ll) != 0) return kNaN;
	return ll;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fisher matrices
// ---------------------------------------------------------------------------------------------------------------------------

void fisher_numerical(double *frequency, int length, std::string generation_method, std::string detector, std::string reference_detector,
                      double **output, int dimension, gen_params_base<double> *parameters, int order, int *amp_tapes, int *phase_tapes,
                      double *noise)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return;
	std::vector<double> psd;
	if (!noise) {  // the reference falls back to Hanford_O1_fitted (src/fisher.cpp:98-103)
		psd.resize(length);
		if (gwat_b200_populate_noise(frequency, "Hanford_O1_fitted", nullptr, length, psd.data()) != 0) return;
		for (double &v : psd) v *= v;
		noise = psd.data();
	}
	S.net_key = 0;
	gwat_b200::fisher_numerical(*e, frequency, length, generation_method, detector, reference_detector, output, dimension, parameters, order,
	                            amp_tapes, phase_tapes, noise);
}

// ---------------------------------------------------------------------------------------------------------------------------
// the samplers' callbacks
// ---------------------------------------------------------------------------------------------------------------------------

extern "C" {

// Set this translation unit's copy of GWAT's file-static sampler state (include/gwat/mcmc_gw.h:22-46); see the header comment.
void gwat_b200_dropin_bind_mcmc(std::complex<double> **data, double **noise, double **frequencies, int *data_length, std::string *detectors,
                                int num_detectors, const char *generation_method, MCMC_modification_struct *mod_struct, double gmst,
                                int deriv_order)
{
	mcmc_data = data;
	mcmc_noise = noise;
	mcmc_frequencies = frequencies;
	mcmc_data_length = data_length;
	mcmc_detectors = detectors;
	mcmc_num_detectors = num_detectors;
	mcmc_generation_method = generation_method ? generation_method : "";
	mcmc_mod_struct = mod_struct;
	mcmc_gmst = gmst;
	mcmc_deriv_order = deriv_order;
	mcmc_intrinsic = false;
	mcmc_save_waveform = true;
	session().net_key = 0;
}

// mcmc_intrinsic of this unit (PTMCMC_method_specific_prep sets the reference's copy from the dimension, src/mcmc_gw.cpp:1880-1985):
// the wrappers then take the tc/phic-maximised likelihood and the sky-averaged Fisher of the 4 (+ modifications) / 8 parameter sets
void gwat_b200_dropin_set_intrinsic(int intrinsic) { mcmc_intrinsic = intrinsic != 0; }
void gwat_b200_dropin_set_segment_duration(double T) { session().T_override = T; }
void gwat_b200_dropin_invalidate(void) { session().net_key = 0; }
const char *gwat_b200_dropin_last_error(void)
{
	Session &S = session();
	static std::string msg;
	msg = S.engine ? S.engine->last_error() : std::string(gwat_b200_last_error(nullptr));
	return msg.c_str();
}

}  // extern "C"

double MCMC_likelihood_wrapper(double *param, mcmc_data_interface *interface, void *parameters)
{
	MCMC_user_param *user_param = (MCMC_user_param *)parameters;
	const int dimension = interface->max_dim;
	if (!mcmc_data || !same_length(mcmc_data_length, mcmc_num_detectors)) return kNaN;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return kNaN;
	const bool gl = user_param && user_param->GAUSS_QUAD;
	if (!ensure_network(S, e, mcmc_detectors, mcmc_num_detectors, mcmc_data_length[0], mcmc_frequencies, mcmc_noise, mcmc_data,
	                    user_param ? user_param->weights : nullptr, gl ? "GAUSSLEG" : "SIMPSONS", user_param ? user_param->log10F : false))
		return kNaN;
	gwat_b200_mod mod;
	bool ok;
	to_mod(mcmc_mod_struct, mod, ok);
	if (!ok) return kNaN;
	double ll = kNaN;
	if (mcmc_intrinsic) {
		// the tc/phic-maximised branches (src/mcmc_gw.cpp:2611-2722): the intrinsic sampling set, repacked with sky_average, through the
		// batched cuFFT likelihood (GAUSSLEG grids have no time axis: the C ABI refuses them, and so does this call)
		if (gwat_b200_loglike_maximized_mcmc_batch(e->ctx(), mcmc_generation_method.c_str(), &mod, dimension, 1, param, mcmc_gmst, &ll) != 0) return kNaN;
		return ll;
	}
	const double T = segment_duration(S, mcmc_frequencies, mcmc_num_detectors);
	if (gwat_b200_loglike_mcmc_batch(e->ctx(), mcmc_generation_method.c_str(), &mod, dimension, 1, param, mcmc_gmst, T, &ll) != 0) return kNaN;
	return ll;
}

void MCMC_fisher_wrapper(double *param, double **output, mcmc_data_interface *interface, void *parameters)
{
	const MCMC_user_param *user_param = (const MCMC_user_param *)parameters;
	const int dimension = interface->max_dim;
	for (int j = 0; j < dimension; j++)
		for (int k = 0; k < dimension; k++) output[j][k] = kNaN;
	// the Fisher matrices may have a grid of their own (user_param->fisher_freq / fisher_PSD / fisher_length, src/mcmc_gw.cpp:2257-2275:
	// typically a short Gauss-Legendre grid); fisher_numerical applies Simpson's rule to whatever grid it is given (src/fisher.cpp:128-131),
	// and so does the C ABI.  fisher_AD asks for ADOL-C derivatives, which are outside this path: the numerical stencil answers instead.
	double **local_freq = mcmc_frequencies, **local_noise = mcmc_noise;
	int *local_lengths = mcmc_data_length;
	if (user_param && user_param->fisher_freq) local_freq = user_param->fisher_freq;
	if (user_param && user_param->fisher_PSD) local_noise = user_param->fisher_PSD;
	if (user_param && user_param->fisher_length) local_lengths = user_param->fisher_length;
	if (!local_freq || !local_noise || !local_lengths || !same_length(local_lengths, mcmc_num_detectors)) return;
	const bool own_grid = local_freq != mcmc_frequencies || local_lengths != mcmc_data_length;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	gwat_b200::Engine *e = engine_locked(S);
	if (!e) return;
	if (!ensure_network(S, e, mcmc_detectors, mcmc_num_detectors, local_lengths[0], local_freq, local_noise, own_grid ? nullptr : mcmc_data, nullptr,
	                    "SIMPSONS", false))
		return;
	gwat_b200_mod mod;
	bool ok;
	to_mod(mcmc_mod_struct, mod, ok);
	if (!ok) return;
	std::vector<double> F((size_t)dimension * dimension), vals(dimension), vecs((size_t)dimension * dimension);
	if (mcmc_intrinsic) {  // sky-averaged records: amplitude / phase derivatives per detector PSD, intrinsic transformations (:2163-2179)
		if (gwat_b200_mcmc_fisher_intrinsic_batch(e->ctx(), mcmc_generation_method.c_str(), &mod, dimension, mcmc_deriv_order, 1, param, mcmc_gmst,
		                                          F.data()) != 0)
			return;
		for (int j = 0; j < dimension; j++)
			for (int k = 0; k < dimension; k++) output[j][k] = F[(size_t)j * dimension + k];
		return;
	}
	// sum over detectors of fisher_numerical("MCMC_" + method) + MCMC_fisher_transformations (:2298-2316)
	if (gwat_b200_mcmc_fisher_batch(e->ctx(), mcmc_generation_method.c_str(), &mod, dimension, mcmc_deriv_order, 1, param, mcmc_gmst, F.data(),
	                                vals.data(), vecs.data()) != 0)
		return;
	for (int j = 0; j < dimension; j++)
		for (int k = 0; k < dimension; k++) output[j][k] = F[(size_t)j * dimension + k];
}
