// gwat_b200_cxx.hpp -- header-only C++ forwarding layer with the reference's own entry-point names and argument lists
// for T = double, on top of the C ABI of gwat_b200.h.
//
// The reference's C++ API passes std::string, double**, std::complex<double>** and gen_params_base<double>* -- types that
// cannot cross a C ABI -- so this thin layer stays C++ and is meant to be compiled INSIDE GWAT (or a GWAT user's code),
// against GWAT's own include/gwat/util.h.  It is templated on the parameter-struct type, which only has to provide the
// reference's member names (mass1, spin1[3], betappe, bppe, Nmod, ...): gen_params_base<double> satisfies that, and so
// does the small mirror struct the GPU tests use (tests/dropin_types.hpp).
//
//   reference function                                            here
//   fourier_waveform<double>            include/gwat/waveform_generator.h:43     gwat_b200::fourier_waveform
//   fourier_detector_response<double>   include/gwat/waveform_util.h:191         gwat_b200::fourier_detector_response
//   create_coherent_GW_detection        src/waveform_util.cpp:129                gwat_b200::create_coherent_GW_detection
//   MCMC_likelihood_extrinsic           include/gwat/mcmc_gw.h:473               gwat_b200::MCMC_likelihood_extrinsic
//   Log_Likelihood_internal call path   src/mcmc_gw.cpp:801                      (inside MCMC_likelihood_extrinsic)
//   fisher_numerical                    include/gwat/fisher.h:35                 gwat_b200::fisher_numerical
//
// Error behaviour mirrors the reference: the waveform functions return int status 1 on success (the reference always
// returns 1, src/waveform_generator.cpp:113,293) and 0 on failure; likelihoods return NaN on failure, which the samplers
// reject (src/mcmc_sampler_internals.cpp:101).  Nothing calls exit().
#ifndef GWAT_B200_CXX_HPP
#define GWAT_B200_CXX_HPP

#include <complex>
#include <limits>
#include <string>
#include <vector>

#include "gwat_b200.h"

namespace gwat_b200 {

// gen_params_base<double>-like  ->  flat record (field for field; pointer members copied into the fixed arrays).
// Returns false -- and the caller reports failure (status 0 / NaN) -- when the source carries more modifications than the
// flat record holds (GWAT_B200_MAX_MOD per kind): nothing is ever evaluated with terms silently dropped.
template <class GenParams>
inline bool flatten(const GenParams &g, gwat_b200_source &s)
{
	gwat_b200_source_init(&s);
	s.mass1 = g.mass1;
	s.mass2 = g.mass2;
	s.Luminosity_Distance = g.Luminosity_Distance;
	for (int i = 0; i < 3; i++) {
		s.spin1[i] = g.spin1[i];
		s.spin2[i] = g.spin2[i];
	}
	s.tc = g.tc;
	s.phiRef = g.phiRef;
	s.f_ref = g.f_ref;
	s.psi = g.psi;
	s.incl_angle = g.incl_angle;
	s.RA = g.RA;
	s.DEC = g.DEC;
	s.gmst = g.gmst;
	s.tidal1 = g.tidal1;
	s.tidal2 = g.tidal2;
	s.tidal_s = g.tidal_s;
	s.tidal_a = g.tidal_a;
	s.tidal_weighted = g.tidal_weighted;
	s.delta_tidal_weighted = g.delta_tidal_weighted;
	s.diss_tidal1 = g.diss_tidal1;
	s.diss_tidal2 = g.diss_tidal2;
	s.diss_tidal_s = g.diss_tidal_s;
	s.diss_tidal_a = g.diss_tidal_a;
	s.diss_tidal_weighted = g.diss_tidal_weighted;
	s.chip = g.chip;
	s.phip = g.phip;
	s.PNorder = g.PNorder;
	s.shift_time = g.shift_time;
	s.shift_phase = g.shift_phase;
	s.sky_average = g.sky_average;
	s.tidal_love = g.tidal_love;
	s.tidal_love_error = g.tidal_love_error;
	s.NSflag1 = g.NSflag1;
	s.NSflag2 = g.NSflag2;
	s.dep_postmerger = g.dep_postmerger;
	s.equatorial_orientation = g.equatorial_orientation;
	s.horizon_coord = g.horizon_coord;
	s.cosmology = gwat_b200_cosmology_index(std::string(g.cosmology).c_str());
	if (s.cosmology < 0) return false;  // (the reference prints "Invalid Cosmology" and carries on with z = -1)
	if (g.equatorial_orientation) { s.theta_l = g.theta_l; s.phi_l = g.phi_l; }
	if (g.horizon_coord) { s.theta = g.theta; s.phi = g.phi; }
	const int counts[5] = {g.Nmod, g.Nmod_phi, g.Nmod_sigma, g.Nmod_beta, g.Nmod_alpha};
	for (int n : counts)
		if (n > GWAT_B200_MAX_MOD) return false;
	auto clampn = [](int n) { return n < 0 ? 0 : n; };
	s.Nmod = clampn(g.Nmod);
	for (int i = 0; i < s.Nmod; i++) {
		if (g.betappe) s.betappe[i] = g.betappe[i];
		if (g.bppe) s.bppe[i] = g.bppe[i];
	}
	s.Nmod_phi = clampn(g.Nmod_phi);
	s.Nmod_sigma = clampn(g.Nmod_sigma);
	s.Nmod_beta = clampn(g.Nmod_beta);
	s.Nmod_alpha = clampn(g.Nmod_alpha);
	for (int i = 0; i < s.Nmod_phi; i++) { s.phii[i] = g.phii[i]; s.delta_phi[i] = g.delta_phi[i]; }
	for (int i = 0; i < s.Nmod_sigma; i++) { s.sigmai[i] = g.sigmai[i]; s.delta_sigma[i] = g.delta_sigma[i]; }
	for (int i = 0; i < s.Nmod_beta; i++) { s.betai[i] = g.betai[i]; s.delta_beta[i] = g.delta_beta[i]; }
	for (int i = 0; i < s.Nmod_alpha; i++) { s.alphai[i] = g.alphai[i]; s.delta_alpha[i] = g.delta_alpha[i]; }
	return true;
}

// RAII handle on a context.  One per thread group that submits work; the uploaded network persists between calls.
class Engine {
public:
	explicit Engine(int device = 0) { ok_ = gwat_b200_ctx_create(&ctx_, device) == 0; }
	~Engine() { gwat_b200_ctx_destroy(ctx_); }
	Engine(const Engine &) = delete;
	Engine &operator=(const Engine &) = delete;
	bool ok() const { return ok_; }
	gwat_b200_ctx *ctx() const { return ctx_; }
	std::string last_error() const { return gwat_b200_last_error(ctx_); }

	// Shared grid for all detectors (the only layout the reference supports: src/waveform_util.cpp:140-145).
	// frequencies[d], psd[d], data[d], weights[d] are the reference's per-detector arrays; detector 0's grid is used.
	int set_network(const std::string *detectors, int num_detectors, int length, double **frequencies, double **psd,
	                std::complex<double> **data, double **weights, const std::string &integration_method, bool log10F)
	{
		std::vector<const char *> names(num_detectors);
		for (int d = 0; d < num_detectors; d++) names[d] = detectors[d].c_str();
		std::vector<double> p((size_t)num_detectors * length), re, im;
		for (int d = 0; d < num_detectors; d++)
			for (int i = 0; i < length; i++) p[(size_t)d * length + i] = psd ? psd[d][i] : 1.0;
		if (data) {
			re.resize(p.size());
			im.resize(p.size());
			for (int d = 0; d < num_detectors; d++)
				for (int i = 0; i < length; i++) {
					re[(size_t)d * length + i] = data[d][i].real();
					im[(size_t)d * length + i] = data[d][i].imag();
				}
		}
		const bool gl = integration_method == "GAUSSLEG";
		return gwat_b200_set_network(ctx_, num_detectors, names.data(), length, frequencies[0], p.data(), data ? re.data() : nullptr,
		                             data ? im.data() : nullptr, (gl && weights) ? weights[0] : nullptr, integration_method.c_str(),
		                             log10F ? 1 : 0);
	}

private:
	gwat_b200_ctx *ctx_ = nullptr;
	bool ok_ = false;
};

namespace detail {
inline int grid_only(Engine &e, const std::string &detector, double *frequencies, int length)
{
	const char *names[1] = {detector.c_str()};
	std::vector<double> ones(length, 1.0);
	return gwat_b200_set_network(e.ctx(), 1, names, length, frequencies, ones.data(), nullptr, nullptr, nullptr, "SIMPSONS", 0);
}
}  // namespace detail

// fourier_waveform<double>(frequencies, length, &wp, generation_method, parameters): hplus/hcross are caller-allocated
// std::complex<double>[length] (the reference's waveform_polarizations members).
template <class GenParams>
inline int fourier_waveform(Engine &e, double *frequencies, int length, std::complex<double> *hplus, std::complex<double> *hcross,
                            const std::string &generation_method, GenParams *parameters)
{
	gwat_b200_source s;
	if (!e.ok() || !flatten(*parameters, s) || detail::grid_only(e, "Hanford", frequencies, length) != 0) return 0;
	std::vector<double> b((size_t)4 * length);
	if (gwat_b200_fourier_waveform_batch(e.ctx(), generation_method.c_str(), 1, &s, &b[0], &b[length], &b[2 * (size_t)length],
	                                     &b[3 * (size_t)length]) != 0)
		return 0;
	for (int i = 0; i < length; i++) {
		if (hplus) hplus[i] = std::complex<double>(b[i], b[length + i]);
		if (hcross) hcross[i] = std::complex<double>(b[2 * (size_t)length + i], b[3 * (size_t)length + i]);
	}
	return 1;
}

// fourier_detector_response<double>(frequencies, length, response, detector, generation_method, parameters, times = NULL)
template <class GenParams>
inline int fourier_detector_response(Engine &e, double *frequencies, int length, std::complex<double> *response,
                                     const std::string &detector, const std::string &generation_method, GenParams *parameters)
{
	gwat_b200_source s;
	if (!e.ok() || !flatten(*parameters, s) || detail::grid_only(e, detector, frequencies, length) != 0) return 0;
	std::vector<double> re(length), im(length);
	if (gwat_b200_fourier_detector_response_batch(e.ctx(), generation_method.c_str(), detector.c_str(), 1, &s, re.data(), im.data()) != 0)
		return 0;
	for (int i = 0; i < length; i++) response[i] = std::complex<double>(re[i], im[i]);
	return 1;
}

// create_coherent_GW_detection(detectors, detector_N, frequencies, lengths, reuse_WF, gen_params, generation_method, responses)
template <class GenParams>
inline void create_coherent_GW_detection(Engine &e, std::string *detectors, int detector_N, double **frequencies, int *lengths,
                                         bool /*reuse_WF*/, GenParams *gen_params, const std::string &generation_method,
                                         std::complex<double> **responses)
{
	const int L = lengths[0];
	gwat_b200_source s;
	if (!e.ok() || !flatten(*gen_params, s) || e.set_network(detectors, detector_N, L, frequencies, nullptr, nullptr, nullptr, "SIMPSONS", false) != 0) return;
	std::vector<double> re((size_t)detector_N * L), im((size_t)detector_N * L);
	if (gwat_b200_coherent_response_batch(e.ctx(), generation_method.c_str(), 1, &s, re.data(), im.data()) != 0) return;
	for (int d = 0; d < detector_N; d++)
		for (int i = 0; i < L; i++) responses[d][i] = std::complex<double>(re[(size_t)d * L + i], im[(size_t)d * L + i]);
}

// MCMC_likelihood_extrinsic(save_waveform, parameters, generation_method, data_length, frequencies, data, psd, weights,
//                           integration_method, log10F, detectors, num_detectors)
// T_segment replaces the reference's `1./(frequencies[1]-frequencies[0])` on a double** (src/mcmc_gw.cpp:2466); pass <= 0 to
// get 1/(frequencies[0][1]-frequencies[0][0]), the evident intent.  The network is (re)uploaded when `reload` is true; a
// sampler sets it once and passes false afterwards.
template <class GenParams>
inline double MCMC_likelihood_extrinsic(Engine &e, bool /*save_waveform*/, GenParams *parameters, const std::string &generation_method,
                                        int *data_length, double **frequencies, std::complex<double> **data, double **psd,
                                        double **weights, const std::string &integration_method, bool log10F,
                                        std::string *detectors, int num_detectors, double T_segment = -1, bool reload = true)
{
	const double nan = std::numeric_limits<double>::quiet_NaN();
	if (!e.ok()) return nan;
	const int L = data_length[0];
	if (reload && e.set_network(detectors, num_detectors, L, frequencies, psd, data, weights, integration_method, log10F) != 0) return nan;
	gwat_b200_source s;
	if (!flatten(*parameters, s)) return nan;
	const double T = T_segment > 0 ? T_segment : 1. / (frequencies[0][1] - frequencies[0][0]);
	s.tc = T - s.tc;  // tc_ref (src/mcmc_gw.cpp:2467,2473)
	double ll = nan;
	if (gwat_b200_loglike_batch(e.ctx(), generation_method.c_str(), 1, &s, &ll) != 0) return nan;
	return ll;
}

// The samplers' likelihood callback, one chain per call from many pool threads
//   std::function<double(double *param, int *status, int model_status, mcmc_data_interface *interface, void *parameters)>
// (include/gwat/mcmc_sampler_internals.h:169-171; MCMC_likelihood_wrapper, src/mcmc_gw.cpp:2569).  `CallbackQueue` owns a
// gwat_b200_queue that merges the calls in flight into batched launches; bind it where the reference binds its wrapper:
//   gwat_b200::CallbackQueue q(engine, mcmc_generation_method, mod, dimension, mcmc_gmst, T, numThreads);
//   sampler.ll = [&q](double *p, int *, int, mcmc_data_interface *, void *) { return q(p); };
// The callback's `status` / `model_status` arguments are INPUTS in the reference (which dimensions an RJMCMC model has
// switched on); the fixed-dimension samplers on this path ignore them, and so does this.  Unphysical points and failed
// launches return NaN, which mcmc_step rejects (src/mcmc_sampler_internals.cpp:101-103).
class CallbackQueue {
public:
	CallbackQueue(Engine &e, const std::string &generation_method, const gwat_b200_mod *mod, int dimension, double gmst, double T_segment,
	              int pool_threads, int max_batch = 4096, double max_wait_us = 200.0)
	{
		if (e.ok())
			gwat_b200_queue_create(&q_, e.ctx(), generation_method.c_str(), mod, dimension, gmst, T_segment, max_batch, pool_threads, max_wait_us);
	}
	~CallbackQueue() { gwat_b200_queue_destroy(q_); }
	CallbackQueue(const CallbackQueue &) = delete;
	CallbackQueue &operator=(const CallbackQueue &) = delete;
	bool ok() const { return q_ != nullptr; }
	// `rc` (optional) receives the gwat_b200_status of the batched launch that served this call.
	double operator()(const double *param, int *rc = nullptr) const { return gwat_b200_queue_loglike(q_, param, rc); }
	gwat_b200_queue *handle() const { return q_; }

private:
	gwat_b200_queue *q_ = nullptr;
};

// fisher_numerical(frequency, length, generation_method, detector, reference_detector, output, dimension, parameters, order,
//                  amp_tapes = NULL, phase_tapes = NULL, noise)     -- `noise` (the PSD) is required here.
template <class GenParams>
inline void fisher_numerical(Engine &e, double *frequency, int length, const std::string &generation_method, const std::string &detector,
                             const std::string &reference_detector, double **output, int dimension, GenParams *parameters, int order,
                             int * /*amp_tapes*/, int * /*phase_tapes*/, double *noise)
{
	if (!e.ok() || !noise) return;
	std::string dets[2] = {reference_detector, detector};
	const int D = detector == reference_detector ? 1 : 2;
	double *f2[2] = {frequency, frequency};
	double *p2[2] = {noise, noise};
	gwat_b200_source s;
	if (!flatten(*parameters, s) || e.set_network(dets, D, length, f2, p2, nullptr, nullptr, "SIMPSONS", false) != 0) return;
	std::vector<double> flat((size_t)dimension * dimension);
	if (gwat_b200_fisher_numerical_batch(e.ctx(), generation_method.c_str(), D - 1, 0, dimension, order, 1, &s, flat.data()) != 0) return;
	for (int i = 0; i < dimension; i++)
		for (int j = 0; j < dimension; j++) output[i][j] = flat[(size_t)i * dimension + j];
}

}  // namespace gwat_b200
#endif
