// Host-side mirror of the gwatpy-facing C API of GWAT for the accelerated path -> libgwat_b200_gwatpy.so
//
// gwatpy loads libgwat.so with ctypes and calls `extern "C"` functions declared in include/gwat/gwatpy_wrapping.h
// (implemented in src/gwatpy_wrapping.cpp).  This file exports the on-path subset of those functions under the SAME names
// with the SAME argument lists, implemented on top of the C ABI of include/gwat_b200.h, so gwatpy (or any ctypes user)
// can point at this library for them.  The objects the reference hands to Python as opaque pointers
// (gen_params_base<double>*, MCMC_modification_struct*) are opaque here too; their layout is private to this file.
//
// Everything numerical runs in the CUDA kernels behind the C ABI.  The only host arithmetic in this file is the handful of
// scalar mass conversions gwatpy exposes (calculate_*_py), which are not part of the hot path.
//
// Deliberate deviations from the reference's wrappers (all documented in INTEGRATION.md):
//  * MCMC_likelihood_extrinsic computes the segment duration from pointer arithmetic on a double** (src/mcmc_gw.cpp:2466);
//    here T = 1/(frequencies[1]-frequencies[0]) -- the evident intent.  *_T variants take T explicitly.
//  * the detector letters are mapped letter by letter ("H","L","V","K","C"); the reference compares the REST of the string
//    (src/gwatpy_wrapping.cpp:131-140), which only works for the default order "HLV".
//  * unknown detectors / methods return NaN (or a negative status) instead of calling exit(1).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gwat_b200.h"

namespace {

struct GenParams {  // what gen_params_base_py returns
	gwat_b200_source s;
	std::string cosmology;
};
struct ModStruct {  // what MCMC_modification_struct_py returns
	gwat_b200_mod m;
};
struct DataInterface {  // what mcmc_data_interface_py returns: the fields of mcmc_data_interface (include/gwat/mcmc_sampler_internals.h:20-30)
	int min_dim, chain_id, max_dim, nested_model_number, chain_number;
	double RJ_step_width;
	bool burn_phase;
};

// One process-wide context; the uploaded network is cached and only re-uploaded when the caller's arrays change.
struct Session {
	std::mutex mu;
	gwat_b200_ctx *ctx = nullptr;
	uint64_t net_key = 0;
	std::string last_error;
};
Session &session()
{
	static Session s;
	return s;
}

// Identity of an input array for the network cache: its address, its length and a word-wise FNV-1a over 64 evenly spaced
// samples plus both ends.  A caller that overwrites an array in place between calls (same address, same length) with values
// that differ only between the samples must call gwat_b200_gwatpy_invalidate_network(); the reference's wrappers have no such
// cache because they re-read everything on every call, which is exactly the cost this avoids (cfg5: 100 MB per call).
uint64_t fnv_word(uint64_t h, uint64_t w)
{
	h ^= w;
	h *= 1099511628211ULL;
	return h;
}
uint64_t fnv(uint64_t h, const void *p, size_t n)
{
	const unsigned char *b = static_cast<const unsigned char *>(p);
	for (size_t i = 0; i < n; i++) h = fnv_word(h, b[i]);
	return h;
}
uint64_t array_key(uint64_t h, const double *a, size_t n)
{
	h = fnv_word(h, (uint64_t)(uintptr_t)a);
	h = fnv_word(h, (uint64_t)n);
	if (!a || n == 0) return h;
	const size_t step = n > 64 ? n / 64 : 1;
	for (size_t i = 0; i < n; i += step) {
		uint64_t w;
		std::memcpy(&w, a + i, sizeof(w));
		h = fnv_word(h, w);
	}
	uint64_t w;
	std::memcpy(&w, a + (n - 1), sizeof(w));
	return fnv_word(h, w);
}

const char *detector_from_letter(char c)
{
	switch (c) {
	case 'H': return "Hanford";
	case 'L': return "Livingston";
	case 'V': return "Virgo";
	case 'K': return "Kagra";
	case 'C': return "CE";
	default: return nullptr;
	}
}

int ensure_ctx(Session &S)
{
	if (S.ctx) return 0;
	int dev = 0;
	if (const char *e = std::getenv("GWAT_B200_DEVICE")) dev = std::atoi(e);
	const int rc = gwat_b200_ctx_create(&S.ctx, dev);
	if (rc != 0) {
		S.last_error = gwat_b200_last_error(nullptr);
		std::fprintf(stderr, "gwat_b200: %s\n", S.last_error.c_str());
	}
	return rc;
}

// Upload (or reuse) the network described by gwatpy's flat detector-major arrays.
int ensure_network(Session &S, int D, const char *const *dets, int L, const double *f, const double *psd, const double *dre,
                   const double *dim, const double *weights, const char *integ, bool log10F)
{
	if (int rc = ensure_ctx(S)) return rc;
	uint64_t key = 1469598103934665603ULL;
	key = fnv(key, &D, sizeof(D));
	key = fnv(key, &L, sizeof(L));
	for (int d = 0; d < D; d++) key = fnv(key, dets[d], std::strlen(dets[d]));
	key = array_key(key, f, (size_t)L);
	key = array_key(key, psd, (size_t)D * L);
	key = array_key(key, dre, dre ? (size_t)D * L : 0);
	key = array_key(key, dim, dim ? (size_t)D * L : 0);
	const bool gl = integ && std::string(integ) == "GAUSSLEG";
	key = array_key(key, (gl && weights) ? weights : nullptr, (gl && weights) ? (size_t)L : 0);
	key = fnv(key, integ ? integ : "", integ ? std::strlen(integ) : 0);
	key = fnv(key, &log10F, sizeof(log10F));
	if (key == S.net_key) return 0;
	const int rc = gwat_b200_set_network(S.ctx, D, dets, L, f, psd, dre, dim, gl ? weights : nullptr, integ, log10F ? 1 : 0);
	if (rc != 0) {
		S.last_error = gwat_b200_last_error(S.ctx);
		std::fprintf(stderr, "gwat_b200: %s\n", S.last_error.c_str());
		S.net_key = 0;
		return rc;
	}
	S.net_key = key;
	return 0;
}

int report(Session &S, int rc)
{
	if (rc != 0) {
		S.last_error = gwat_b200_last_error(S.ctx);
		std::fprintf(stderr, "gwat_b200: %s\n", S.last_error.c_str());
	}
	return rc;
}

// a throw-away one-detector network on the caller's grid, for the waveform / single-response entry points
int ensure_grid_only(Session &S, const char *detector, int L, const double *f)
{
	std::vector<double> ones(L, 1.0);
	const char *dets[1] = {detector};
	return ensure_network(S, 1, dets, L, f, ones.data(), nullptr, nullptr, nullptr, "SIMPSONS", false);
}

const double NaN = std::numeric_limits<double>::quiet_NaN();

double likelihood_common(bool mcmc_vector, const double *parameters, const gwat_b200_mod *mod, int dimension,
                         const gwat_b200_source *src, int W, const char *generation_method, const int *data_length,
                         const double *frequencies, const double *dataREAL, const double *dataIMAG, const double *psd,
                         const double *weights, const char *integration_method, bool log10F, const char *detectors,
                         int num_detectors, double gmst, double T_segment, double *out)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (num_detectors < 1 || num_detectors > GWAT_B200_MAX_DETECTORS || !detectors || !data_length) return NaN;
	const int L = data_length[0];
	for (int d = 1; d < num_detectors; d++)
		if (data_length[d] != L) return NaN;  // the reference only supports a shared grid too (src/waveform_util.cpp:140-145)
	const char *dets[GWAT_B200_MAX_DETECTORS];
	for (int d = 0; d < num_detectors; d++) {
		dets[d] = detector_from_letter(detectors[d]);
		if (!dets[d]) return NaN;
	}
	if (ensure_network(S, num_detectors, dets, L, frequencies, psd, dataREAL, dataIMAG, weights, integration_method, log10F)) return NaN;
	int rc;
	if (mcmc_vector) {
		const double T = T_segment > 0 ? T_segment : 1. / (frequencies[1] - frequencies[0]);
		rc = gwat_b200_loglike_mcmc_batch(S.ctx, generation_method, mod, dimension, W, parameters, gmst, T, out);
	} else {
		// MCMC_likelihood_extrinsic: tc_ref = T - tc  (src/mcmc_gw.cpp:2467,2473)
		std::vector<gwat_b200_source> tmp(src, src + W);
		const double T = T_segment > 0 ? T_segment : 1. / (frequencies[1] - frequencies[0]);
		for (auto &s : tmp) s.tc = T - s.tc;
		rc = gwat_b200_loglike_batch(S.ctx, generation_method, W, tmp.data(), out);
	}
	if (report(S, rc)) return NaN;
	return out[0];
}

}  // namespace

extern "C" {

// ---- object constructors / destructors (src/gwatpy_wrapping.cpp:244-420) ------------------------------------------------
void *gen_params_base_py(double mass1, double mass2, double *spin1, double *spin2, double Luminosity_Distance, double incl_angle,
                         double RA, double DEC, double psi, double gmst, double tc, double phiRef, double f_ref, double theta_l,
                         double phi_l, double theta, double phi, char *cosmology, bool equatorial_orientation, bool horizon_coord,
                         bool NSflag1, bool NSflag2, bool dep_postmerger, bool shift_time, bool shift_phase, bool sky_average,
                         double LISA_alpha0, double LISA_phi0, int Nmod_phi, int Nmod_sigma, int Nmod_beta, int Nmod_alpha,
                         int *phii, int *sigmai, int *betai, int *alphai, double *delta_phi, double *delta_sigma,
                         double *delta_beta, double *delta_alpha, int Nmod, double *bppe, double *betappe)
{
	(void)LISA_alpha0;
	(void)LISA_phi0;
	for (int n : {Nmod_phi, Nmod_sigma, Nmod_beta, Nmod_alpha, Nmod})
		if (n > GWAT_B200_MAX_MOD) {  // never evaluated with terms silently dropped
			std::fprintf(stderr, "gwat_b200: gen_params_base_py: more than %d modifications of one kind are not supported\n", GWAT_B200_MAX_MOD);
			return nullptr;
		}
	GenParams *p = new GenParams;
	gwat_b200_source &s = p->s;
	gwat_b200_source_init(&s);
	s.mass1 = mass1;
	s.mass2 = mass2;
	for (int i = 0; i < 3; i++) {
		s.spin1[i] = spin1 ? spin1[i] : 0;
		s.spin2[i] = spin2 ? spin2[i] : 0;
	}
	s.Luminosity_Distance = Luminosity_Distance;
	s.incl_angle = incl_angle;
	s.RA = RA;
	s.DEC = DEC;
	s.psi = psi;
	s.gmst = gmst;
	s.tc = tc;
	s.phiRef = phiRef;
	s.f_ref = f_ref;
	s.theta_l = theta_l;
	s.phi_l = phi_l;
	s.theta = theta;
	s.phi = phi;
	p->cosmology = cosmology ? cosmology : "PLANCK15";
	s.cosmology = gwat_b200_cosmology_index(p->cosmology.c_str());
	if (s.cosmology < 0) {
		std::fprintf(stderr, "gwat_b200: gen_params_base_py: unknown cosmology '%s'\n", p->cosmology.c_str());
		delete p;
		return nullptr;
	}
	s.equatorial_orientation = equatorial_orientation;
	s.horizon_coord = horizon_coord;
	s.NSflag1 = NSflag1;
	s.NSflag2 = NSflag2;
	s.dep_postmerger = dep_postmerger;
	s.shift_time = shift_time;
	s.shift_phase = shift_phase;
	s.sky_average = sky_average;
	auto clampn = [](int n) { return n < 0 ? 0 : n; };
	s.Nmod_phi = clampn(Nmod_phi);
	s.Nmod_sigma = clampn(Nmod_sigma);
	s.Nmod_beta = clampn(Nmod_beta);
	s.Nmod_alpha = clampn(Nmod_alpha);
	s.Nmod = clampn(Nmod);
	for (int i = 0; i < s.Nmod_phi; i++) { s.phii[i] = phii[i]; s.delta_phi[i] = delta_phi[i]; }
	for (int i = 0; i < s.Nmod_sigma; i++) { s.sigmai[i] = sigmai[i]; s.delta_sigma[i] = delta_sigma[i]; }
	for (int i = 0; i < s.Nmod_beta; i++) { s.betai[i] = betai[i]; s.delta_beta[i] = delta_beta[i]; }
	for (int i = 0; i < s.Nmod_alpha; i++) { s.alphai[i] = alphai[i]; s.delta_alpha[i] = delta_alpha[i]; }
	for (int i = 0; i < s.Nmod; i++) { s.bppe[i] = bppe[i]; s.betappe[i] = betappe[i]; }
	return p;
}
void gen_params_base_py_destructor(void *p) { delete static_cast<GenParams *>(p); }

// Accessor the reference does not have: tidal fields are plain members there and gwatpy never sets them through the
// constructor; exposed so NRT waveforms are reachable from Python.
void gen_params_base_set_tidal_py(void *p, double tidal1, double tidal2, double tidal_s, double tidal_weighted, bool tidal_love)
{
	gwat_b200_source &s = static_cast<GenParams *>(p)->s;
	s.tidal1 = tidal1;
	s.tidal2 = tidal2;
	s.tidal_s = tidal_s;
	s.tidal_weighted = tidal_weighted;
	s.tidal_love = tidal_love;
}
void gen_params_base_set_chip_py(void *p, double chip, double phip)
{
	static_cast<GenParams *>(p)->s.chip = chip;
	static_cast<GenParams *>(p)->s.phip = phip;
}
// copy of the flat record, for tests and for users who want to go to the C ABI directly
void gen_params_base_get_flat_py(void *p, gwat_b200_source *out) { *out = static_cast<GenParams *>(p)->s; }

void *MCMC_modification_struct_py(int ppE_Nmod, double *bppe, int gIMR_Nmod_phi, int *gIMR_phii, int gIMR_Nmod_sigma,
                                  int *gIMR_sigmai, int gIMR_Nmod_beta, int *gIMR_betai, int gIMR_Nmod_alpha, int *gIMR_alphai,
                                  bool NSflag1, bool NSflag2)
{
	for (int n : {ppE_Nmod, gIMR_Nmod_phi, gIMR_Nmod_sigma, gIMR_Nmod_beta, gIMR_Nmod_alpha})
		if (n > GWAT_B200_MAX_MOD) {
			std::fprintf(stderr, "gwat_b200: MCMC_modification_struct_py: more than %d modifications of one kind are not supported\n", GWAT_B200_MAX_MOD);
			return nullptr;
		}
	ModStruct *m = new ModStruct;
	gwat_b200_mod_init(&m->m);
	auto clampn = [](int n) { return n < 0 ? 0 : n; };
	m->m.ppE_Nmod = clampn(ppE_Nmod);
	for (int i = 0; i < m->m.ppE_Nmod; i++) m->m.bppe[i] = bppe[i];
	m->m.gIMR_Nmod_phi = clampn(gIMR_Nmod_phi);
	m->m.gIMR_Nmod_sigma = clampn(gIMR_Nmod_sigma);
	m->m.gIMR_Nmod_beta = clampn(gIMR_Nmod_beta);
	m->m.gIMR_Nmod_alpha = clampn(gIMR_Nmod_alpha);
	for (int i = 0; i < m->m.gIMR_Nmod_phi; i++) m->m.gIMR_phii[i] = gIMR_phii[i];
	for (int i = 0; i < m->m.gIMR_Nmod_sigma; i++) m->m.gIMR_sigmai[i] = gIMR_sigmai[i];
	for (int i = 0; i < m->m.gIMR_Nmod_beta; i++) m->m.gIMR_betai[i] = gIMR_betai[i];
	for (int i = 0; i < m->m.gIMR_Nmod_alpha; i++) m->m.gIMR_alphai[i] = gIMR_alphai[i];
	m->m.NSflag1 = NSflag1;
	m->m.NSflag2 = NSflag2;
	return m;
}
void MCMC_modification_struct_py_destructor(void *m) { delete static_cast<ModStruct *>(m); }

// ---- waveforms and responses (src/gwatpy_wrapping.cpp: fourier_waveform_py, fourier_detector_response_py) -------------------
int fourier_waveform_py(double *frequencies, int length, double *wf_plus_real, double *wf_plus_imaginary, double *wf_cross_real,
                        double *wf_cross_imaginary, char *generation_method, void *parameters)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, "Hanford", length, frequencies)) return 0;
	const int rc = gwat_b200_fourier_waveform_batch(S.ctx, generation_method, 1, &static_cast<GenParams *>(parameters)->s,
	                                                wf_plus_real, wf_plus_imaginary, wf_cross_real, wf_cross_imaginary);
	return report(S, rc) == 0 ? 1 : 0;  // the reference returns status 1 on success (src/waveform_generator.cpp:113,293)
}

int fourier_detector_response_py(double *frequencies, int length, double *response_real, double *response_imaginary,
                                 char *detector, char *generation_method, void *parameters)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, detector, length, frequencies)) return 0;
	const int rc = gwat_b200_fourier_detector_response_batch(S.ctx, generation_method, detector, 1,
	                                                         &static_cast<GenParams *>(parameters)->s, response_real,
	                                                         response_imaginary);
	return report(S, rc) == 0 ? 1 : 0;
}

// fourier_waveform_full_py (src/gwatpy_wrapping.cpp:492-546): all six polarisation arrays.  The models on this path are GR
// or phase-modified GR -- plus and cross only (assign_polarizations) -- so the vector x/y and scalar b/l outputs are zero.
int fourier_waveform_full_py(double *frequencies, int length, double *wf_plus_real, double *wf_plus_imaginary, double *wf_cross_real,
                             double *wf_cross_imaginary, double *wf_x_real, double *wf_x_imaginary, double *wf_y_real,
                             double *wf_y_imaginary, double *wf_b_real, double *wf_b_imaginary, double *wf_l_real,
                             double *wf_l_imaginary, char *generation_method, void *parameters)
{
	const int st = fourier_waveform_py(frequencies, length, wf_plus_real, wf_plus_imaginary, wf_cross_real, wf_cross_imaginary,
	                                   generation_method, parameters);
	double *extra[8] = {wf_x_real, wf_x_imaginary, wf_y_real, wf_y_imaginary, wf_b_real, wf_b_imaginary, wf_l_real, wf_l_imaginary};
	for (double *a : extra)
		if (a) std::memset(a, 0, sizeof(double) * (size_t)(length > 0 ? length : 0));
	return st;
}

double gps_to_GMST_radian_py(double gps) { return gwat_b200_gps_to_gmst_radian(gps); }

// batched versions (new): W parameter objects at once, outputs [W][length]
int fourier_waveform_batch_py(double *frequencies, int length, int W, void **parameters, char *generation_method,
                              double *wf_plus_real, double *wf_plus_imaginary, double *wf_cross_real, double *wf_cross_imaginary)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, "Hanford", length, frequencies)) return 0;
	std::vector<gwat_b200_source> src(W);
	for (int w = 0; w < W; w++) src[w] = static_cast<GenParams *>(parameters[w])->s;
	const int rc = gwat_b200_fourier_waveform_batch(S.ctx, generation_method, W, src.data(), wf_plus_real, wf_plus_imaginary,
	                                                wf_cross_real, wf_cross_imaginary);
	return report(S, rc) == 0 ? 1 : 0;
}

// ---- the plain-C waveform API (include/gwat/waveform_generator_C.h, src/waveform_generator_C.cpp) -----------------------------
// Argument ORDER follows the definitions in src/waveform_generator_C.cpp:8-33 (phiRef, tc, f_ref) -- the header lists
// (tc, f_ref, phiRef), but the compiled symbol is what callers bind to.  Every other gen_params member keeps its default.
namespace {
gwat_b200_source c_api_source(double mass1, double mass2, double DL, double s1x, double s1y, double s1z, double s2x, double s2y,
                              double s2z, double incl_angle, double theta, double phi)
{
	gwat_b200_source s;
	gwat_b200_source_init(&s);
	s.mass1 = mass1;
	s.mass2 = mass2;
	s.Luminosity_Distance = DL;
	s.spin1[0] = s1x; s.spin1[1] = s1y; s.spin1[2] = s1z;
	s.spin2[0] = s2x; s.spin2[1] = s2y; s.spin2[2] = s2z;
	s.incl_angle = incl_angle;
	s.theta = theta;
	s.phi = phi;
	s.NSflag1 = s.NSflag2 = 0;
	s.sky_average = 0;
	return s;
}
// false: more ppE terms than the flat record holds -> the caller returns status 0, nothing is evaluated
bool c_api_ppe(gwat_b200_source &s, const double *beta, const double *b, int Nmod)
{
	if (Nmod > GWAT_B200_MAX_MOD) return false;
	s.Nmod = Nmod < 0 ? 0 : Nmod;
	for (int i = 0; i < s.Nmod; i++) {
		if (beta) s.betappe[i] = beta[i];
		if (b) s.bppe[i] = b[i];
	}
	return true;
}
}  // namespace

int fourier_waveformC(double *frequencies, int length, double *waveform_plus_real, double *waveform_plus_imag,
                      double *waveform_cross_real, double *waveform_cross_imag, char *generation_method, double mass1, double mass2,
                      double DL, double spin1x, double spin1y, double spin1z, double spin2x, double spin2y, double spin2z,
                      double phiRef, double tc, double f_ref, double *ppE_beta, double *ppE_b, int Nmod, double incl_angle,
                      double theta, double phi)
{
	gwat_b200_source s = c_api_source(mass1, mass2, DL, spin1x, spin1y, spin1z, spin2x, spin2y, spin2z, incl_angle, theta, phi);
	s.tc = tc;
	s.phiRef = phiRef;
	s.f_ref = f_ref;
	if (!c_api_ppe(s, ppE_beta, ppE_b, Nmod)) return 0;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, "Hanford", length, frequencies)) return 0;
	const int rc = gwat_b200_fourier_waveform_batch(S.ctx, generation_method, 1, &s, waveform_plus_real, waveform_plus_imag,
	                                                waveform_cross_real, waveform_cross_imag);
	return report(S, rc) == 0 ? 1 : 0;
}

int fourier_amplitudeC(double *frequencies, int length, double *amplitude, char *generation_method, double mass1, double mass2,
                       double DL, double spin1x, double spin1y, double spin1z, double spin2x, double spin2y, double spin2z,
                       double incl_angle, double theta, double phi)
{
	gwat_b200_source s = c_api_source(mass1, mass2, DL, spin1x, spin1y, spin1z, spin2x, spin2y, spin2z, incl_angle, theta, phi);
	// the amplitude does not involve the time/phase reference (construct_amplitude never evaluates it); keep the record's
	// phase fields out of the way so that an unset f_ref = 0 cannot invalidate the coefficient block
	s.shift_time = 0;
	s.shift_phase = 0;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, "Hanford", length, frequencies)) return 0;
	return report(S, gwat_b200_fourier_amplitude_phase_batch(S.ctx, generation_method, 1, &s, amplitude, nullptr)) == 0 ? 1 : 0;
}

int fourier_phaseC(double *frequencies, int length, double *phase, char *generation_method, double mass1, double mass2, double DL,
                   double spin1x, double spin1y, double spin1z, double spin2x, double spin2y, double spin2z, double tc, double f_ref,
                   double phiRef, double *ppE_beta, double *ppE_b, int Nmod, double incl_angle, double theta, double phi)
{
	gwat_b200_source s = c_api_source(mass1, mass2, DL, spin1x, spin1y, spin1z, spin2x, spin2y, spin2z, incl_angle, theta, phi);
	s.tc = tc;
	s.phiRef = phiRef;
	s.f_ref = f_ref;
	if (!c_api_ppe(s, ppE_beta, ppE_b, Nmod)) return 0;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_grid_only(S, "Hanford", length, frequencies)) return 0;
	return report(S, gwat_b200_fourier_amplitude_phase_batch(S.ctx, generation_method, 1, &s, nullptr, phase)) == 0 ? 1 : 0;
}

// ---- noise curves and SNR (src/gwatpy_wrapping.cpp:59-69, 886-889) ----------------------------------------------------------
// The tabulated curves are read from $GWAT_B200_NOISE_DIR (the directory GWAT installs as GWAT_SHARE_DIR/noise_data; in its
// source tree data/noise_data/currently_supported).  integration_time only matters for the LISA confusion noise: unused.
void populate_noise_py(double *frequencies, char *detector, double *noise_root, int length, double integration_time)
{
	(void)integration_time;
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	const int rc = gwat_b200_populate_noise(frequencies, detector, std::getenv("GWAT_B200_NOISE_DIR"), length, noise_root);
	if (rc != 0) {
		S.last_error = std::string("populate_noise: curve '") + (detector ? detector : "(null)") +
		               "' unknown, unreadable (set GWAT_B200_NOISE_DIR) or evaluated outside its table";
		std::fprintf(stderr, "gwat_b200: %s\n", S.last_error.c_str());
	}
}

// calculate_snr(sensitivity_curve, detector, generation_method, params, frequencies, length, integration_method, weights,
// log10_freq), src/waveform_util.cpp:290-344: sqrt(4 int |response|^2 / S) with S = populate_noise(curve)^2.
double calculate_snr_py(char *sensitivity_curve, char *detector, char *generation_method, void *params, double *frequencies, int length,
                        char *integration_method, double *weights, bool log10_freq)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	std::vector<double> psd(length > 0 ? length : 0);
	if (gwat_b200_populate_noise(frequencies, sensitivity_curve, std::getenv("GWAT_B200_NOISE_DIR"), length, psd.data()) != 0) {
		S.last_error = "calculate_snr: noise curve unavailable";
		return NaN;
	}
	for (double &v : psd) v *= v;
	const char *dets[1] = {detector};
	if (ensure_network(S, 1, dets, length, frequencies, psd.data(), nullptr, nullptr, weights, integration_method, log10_freq)) return NaN;
	double snr = NaN;
	if (report(S, gwat_b200_snr_batch(S.ctx, generation_method, 1, &static_cast<GenParams *>(params)->s, &snr)) != 0) return NaN;
	return snr;
}

// ---- likelihoods (src/gwatpy_wrapping.cpp:96-241) ---------------------------------------------------------------------------
double MCMC_likelihood_extrinsic_py(bool save_waveform, void *parameters, char *generation_method, int *data_length,
                                    double *frequencies, double *dataREAL, double *dataIMAG, double *psd, double *weights,
                                    char *integration_method, bool log10F, char *detectors, int num_detectors)
{
	(void)save_waveform;
	double out = NaN;
	const gwat_b200_source &s = static_cast<GenParams *>(parameters)->s;
	return likelihood_common(false, nullptr, nullptr, 0, &s, 1, generation_method, data_length, frequencies, dataREAL, dataIMAG,
	                         psd, weights, integration_method, log10F, detectors, num_detectors, s.gmst, -1, &out);
}

// The reference's pyv2 wrapper runs MCMC_prep_params in a translation unit whose static mcmc_gmst is never set, i.e. with
// gmst = 0 (include/gwat/mcmc_gw.h:35); the same-named function keeps that, the *_gmst variant takes it explicitly.
double MCMC_likelihood_extrinsic_pyv2(bool save_waveform, double *parameters, void *mod_struct, int dimension,
                                      char *generation_method, int *data_length, double *frequencies, double *dataREAL,
                                      double *dataIMAG, double *psd, double *weights, char *integration_method, bool log10F,
                                      char *detectors, int num_detectors)
{
	(void)save_waveform;
	double out = NaN;
	return likelihood_common(true, parameters, mod_struct ? &static_cast<ModStruct *>(mod_struct)->m : nullptr, dimension, nullptr, 1,
	                         generation_method, data_length, frequencies, dataREAL, dataIMAG, psd, weights, integration_method,
	                         log10F, detectors, num_detectors, 0.0, -1, &out);
}

// NEW: W sampling vectors in one call; parameters[W][dimension] row-major, out[W].
// gmst and T_segment are explicit (T_segment <= 0: 1/(f[1]-f[0])).  Returns 0 on success.
int MCMC_likelihood_extrinsic_batch_py(double *parameters, int W, void *mod_struct, int dimension, char *generation_method,
                                       int *data_length, double *frequencies, double *dataREAL, double *dataIMAG, double *psd,
                                       double *weights, char *integration_method, bool log10F, char *detectors,
                                       int num_detectors, double gmst, double T_segment, double *out)
{
	if (W <= 0) return 0;
	const double r = likelihood_common(true, parameters, mod_struct ? &static_cast<ModStruct *>(mod_struct)->m : nullptr, dimension,
	                                   nullptr, W, generation_method, data_length, frequencies, dataREAL, dataIMAG, psd, weights,
	                                   integration_method, log10F, detectors, num_detectors, gmst, T_segment, out);
	(void)r;
	return session().last_error.empty() ? 0 : -1;
}

// repack_parameters_py (src/gwatpy_wrapping.cpp): sampling vector -> the gen_params object, "MCMC_"+method layout.
// The non-parameter options already stored in the object (gmst, flags, modification layout) are kept.
void repack_parameters_py(double *parameters, void *gen_param, char *generation_method, int dim)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_ctx(S)) return;
	GenParams *g = static_cast<GenParams *>(gen_param);
	gwat_b200_mod mod;
	gwat_b200_mod_init(&mod);
	mod.ppE_Nmod = g->s.Nmod;
	for (int i = 0; i < GWAT_B200_MAX_MOD; i++) mod.bppe[i] = g->s.bppe[i];
	mod.gIMR_Nmod_phi = g->s.Nmod_phi;
	mod.gIMR_Nmod_sigma = g->s.Nmod_sigma;
	mod.gIMR_Nmod_beta = g->s.Nmod_beta;
	mod.gIMR_Nmod_alpha = g->s.Nmod_alpha;
	for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
		mod.gIMR_phii[i] = g->s.phii[i];
		mod.gIMR_sigmai[i] = g->s.sigmai[i];
		mod.gIMR_betai[i] = g->s.betai[i];
		mod.gIMR_alphai[i] = g->s.alphai[i];
	}
	mod.NSflag1 = g->s.NSflag1;
	mod.NSflag2 = g->s.NSflag2;
	mod.tidal_love = g->s.tidal_love;
	std::string m(generation_method ? generation_method : "");
	if (m.compare(0, 5, "MCMC_") == 0) m.erase(0, 5);
	gwat_b200_source out;
	if (report(S, gwat_b200_repack_mcmc_batch(S.ctx, m.c_str(), &mod, dim, 1, parameters, g->s.gmst, &out))) return;
	// only the sampled quantities move; flags that MCMC_prep_params (not repack_parameters) would set stay as they were
	gwat_b200_source keep = g->s;
	g->s = out;
	g->s.f_ref = keep.f_ref;
	g->s.shift_time = keep.shift_time;
	g->s.shift_phase = keep.shift_phase;
	g->s.sky_average = keep.sky_average;
	g->s.gmst = keep.gmst;
}

// ---- sampler-side helpers gwatpy exposes (src/gwatpy_wrapping.cpp:244-372) -------------------------------------------------
void *mcmc_data_interface_py(int min_dim, int max_dim, int chain_id, int nested_model_number, int chain_number, double RJ_step_width,
                             bool burn_phase)
{
	(void)RJ_step_width;
	DataInterface *d = new DataInterface;
	d->min_dim = min_dim;
	d->max_dim = max_dim;
	d->chain_id = chain_id;
	d->nested_model_number = nested_model_number;
	d->chain_number = chain_number;
	d->RJ_step_width = chain_number;  // as the reference has it (src/gwatpy_wrapping.cpp:260)
	d->burn_phase = burn_phase;
	return d;
}
void mcmc_data_interface_destructor_py(void *d) { delete static_cast<DataInterface *>(d); }
// read-back for tests and ctypes users (the reference's struct is plain data that Python cannot see either)
void mcmc_data_interface_get_py(void *p, int *ints5, double *RJ_step_width, bool *burn_phase)
{
	const DataInterface *d = static_cast<DataInterface *>(p);
	ints5[0] = d->min_dim;
	ints5[1] = d->max_dim;
	ints5[2] = d->chain_id;
	ints5[3] = d->nested_model_number;
	ints5[4] = d->chain_number;
	*RJ_step_width = d->RJ_step_width;
	*burn_phase = d->burn_phase;
}

// MCMC_prep_params (src/mcmc_gw.cpp:2492-2568): the flags every sampler likelihood call sets on the gen_params object, the copy
// of the sampling vector into temp_params, the modification layout, and the dCS/EdGB unit change of the coupling.  Returns the
// generation method in a string the caller owns (new char[], as the reference's wrapper does).  The reference's wrapper runs in a
// translation unit whose static mcmc_gmst is never set, so gmst becomes 0 unless save_gmst (src/gwatpy_wrapping.cpp:357-381).
char *MCMC_prep_params_py(double *param, double *temp_params, void *gen_params, int dimension, char *generation_method, void *mod_struct,
                          bool save_gmst)
{
	GenParams *g = static_cast<GenParams *>(gen_params);
	const gwat_b200_mod &m = static_cast<ModStruct *>(mod_struct)->m;
	gwat_b200_source &s = g->s;
	const double gmst = s.gmst;
	s.sky_average = 0;
	s.tidal_love = m.tidal_love;
	s.tidal_love_error = m.tidal_love_error;
	s.f_ref = 20;
	s.shift_time = 1;
	s.shift_phase = 1;
	s.gmst = save_gmst ? gmst : 0.0;
	s.equatorial_orientation = 0;
	s.horizon_coord = 0;
	s.NSflag1 = m.NSflag1;
	s.NSflag2 = m.NSflag2;
	for (int i = 0; i < dimension; i++) temp_params[i] = param[i];
	int ppe_like = 0, gimr = 0, alpha_units = 0;
	std::string method(generation_method ? generation_method : "");
	std::string bare = method;
	if (bare.compare(0, 5, "MCMC_") == 0) bare.erase(0, 5);
	if (gwat_b200_method_info(bare.c_str(), &ppe_like, &gimr, &alpha_units, nullptr, nullptr) == 0) {
		int base = dimension;
		if (ppe_like) {
			s.Nmod = m.ppE_Nmod;
			for (int i = 0; i < m.ppE_Nmod; i++) s.bppe[i] = m.bppe[i];
			base = dimension - m.ppE_Nmod;
		} else if (gimr) {
			s.Nmod_phi = m.gIMR_Nmod_phi;
			s.Nmod_sigma = m.gIMR_Nmod_sigma;
			s.Nmod_beta = m.gIMR_Nmod_beta;
			s.Nmod_alpha = m.gIMR_Nmod_alpha;
			for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
				s.phii[i] = m.gIMR_phii[i];
				s.sigmai[i] = m.gIMR_sigmai[i];
				s.betai[i] = m.gIMR_betai[i];
				s.alphai[i] = m.gIMR_alphai[i];
			}
			base = dimension - m.gIMR_Nmod_phi - m.gIMR_Nmod_sigma - m.gIMR_Nmod_beta - m.gIMR_Nmod_alpha;
		}
		if (alpha_units && base >= 0 && base < dimension) {
			const double x = temp_params[base] / (299792458. / 1000.);  // pow_int(sqrt(alpha)[km] / (c/1000), 4):
			temp_params[base] = ((x * x) * x) * x;                       // a sequential product (src/util.cpp:1585-1597)
		}
	}
	char *out = new char[method.size() + 1];
	std::memcpy(out, method.c_str(), method.size() + 1);
	return out;
}

// pack_local_mod_structure (src/mcmc_gw.cpp:3401-3476): the gIMR modifications that are switched on in an RJMCMC status vector.
void pack_local_mod_structure_py(void *interface, double *param, int *status, char *waveform_extended, void *full_struct, void *local_struct)
{
	(void)param;
	const DataInterface *di = static_cast<DataInterface *>(interface);
	const gwat_b200_mod &full = static_cast<ModStruct *>(full_struct)->m;
	gwat_b200_mod &loc = static_cast<ModStruct *>(local_struct)->m;
	if (!waveform_extended || std::string(waveform_extended).find("gIMR") == std::string::npos) return;
	const int b_phi = full.gIMR_Nmod_phi + di->min_dim, b_sigma = full.gIMR_Nmod_sigma + b_phi, b_beta = full.gIMR_Nmod_beta + b_sigma,
	          b_alpha = full.gIMR_Nmod_alpha + b_beta;
	int dimct = 0;
	for (int i = 0; i < di->max_dim; i++) {
		if (status[i] == 1) dimct++;
		if (i >= di->min_dim && status[i] == 1) {
			if (i < b_phi) loc.gIMR_Nmod_phi++;
			else if (i < b_sigma) loc.gIMR_Nmod_sigma++;
			else if (i < b_beta) loc.gIMR_Nmod_beta++;
			else if (i < b_alpha) loc.gIMR_Nmod_alpha++;
		}
	}
	if (dimct == di->min_dim) return;
	int c_phi = 0, c_sigma = 0, c_beta = 0, c_alpha = 0;
	for (int i = di->min_dim; i < di->max_dim; i++) {
		if (status[i] != 1) continue;
		if (i < b_phi) { if (c_phi < GWAT_B200_MAX_MOD) loc.gIMR_phii[c_phi++] = full.gIMR_phii[i - b_phi + full.gIMR_Nmod_phi]; }
		else if (i < b_sigma) { if (c_sigma < GWAT_B200_MAX_MOD) loc.gIMR_sigmai[c_sigma++] = full.gIMR_sigmai[i - b_sigma + full.gIMR_Nmod_sigma]; }
		else if (i < b_beta) { if (c_beta < GWAT_B200_MAX_MOD) loc.gIMR_betai[c_beta++] = full.gIMR_betai[i - b_beta + full.gIMR_Nmod_beta]; }
		else if (i < b_alpha) { if (c_alpha < GWAT_B200_MAX_MOD) loc.gIMR_alphai[c_alpha++] = full.gIMR_alphai[i - b_alpha + full.gIMR_Nmod_alpha]; }
	}
}
// read-back of a modification struct (tests / ctypes users)
void MCMC_modification_struct_get_py(void *m, gwat_b200_mod *out) { *out = static_cast<ModStruct *>(m)->m; }

// match_py (src/gwatpy_wrapping.cpp:43-57 -> match, src/waveform_util.cpp:41-89)
double match_py(double *data1_real, double *data1_imag, double *data2_real, double *data2_imag, double *SN, double *frequencies, int length)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_ctx(S)) return NaN;
	double out = NaN;
	if (report(S, gwat_b200_match(S.ctx, length, frequencies, SN, data1_real, data1_imag, data2_real, data2_imag, &out))) return NaN;
	return out;
}

// ---- detector helpers ---------------------------------------------------------------------------------------------------
double DTOA_DETECTOR_py(double RA, double DEC, double GMST_rad, char *det1, char *det2)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	if (ensure_ctx(S)) return NaN;
	const char *dets[2] = {det1, det2};
	const double f[4] = {1, 2, 3, 4}, ones[8] = {1, 1, 1, 1, 1, 1, 1, 1};
	S.net_key = 0;
	if (report(S, gwat_b200_set_network(S.ctx, 2, dets, 4, f, ones, nullptr, nullptr, nullptr, "SIMPSONS", 0))) return NaN;
	double psi = 0, fp[2], fc[2], dt[2];
	if (report(S, gwat_b200_antenna_batch(S.ctx, 1, &RA, &DEC, &psi, GMST_rad, fp, fc, dt))) return NaN;
	return dt[1];
}

void detector_response_equatorial_py(char *detector, double ra, double dec, double psi, double gmst, bool *active_polarizations,
                                     double *response_functions)
{
	(void)active_polarizations;  // only the tensor polarisations are on this path
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	response_functions[0] = response_functions[1] = NaN;
	if (ensure_ctx(S)) return;
	const char *dets[1] = {detector};
	const double f[4] = {1, 2, 3, 4}, ones[4] = {1, 1, 1, 1};
	S.net_key = 0;
	if (report(S, gwat_b200_set_network(S.ctx, 1, dets, 4, f, ones, nullptr, nullptr, nullptr, "SIMPSONS", 0))) return;
	double fp, fc, dt;
	if (report(S, gwat_b200_antenna_batch(S.ctx, 1, &ra, &dec, &psi, gmst, &fp, &fc, &dt))) return;
	response_functions[0] = fp;
	response_functions[1] = fc;
}

// get_detector_parameters (src/gwatpy_wrapping.cpp:743-831): site constants by name.  The reference matches SUBSTRINGS, in this order,
// and knows the sites up to ET1 only (ET2 / ET3 fall through to "Unsupported detector", -1): kept, so the two libraries answer alike.
int get_detector_parameters(char *detector, double *LAT, double *LON, double *location, double *response_tensor)
{
	static const struct { const char *a, *b, *c, *canonical; } sites[] = {
		{"Hanford", "hanford", nullptr, "Hanford"},   {"Livingston", "livingston", nullptr, "Livingston"},
		{"Virgo", "virgo", nullptr, "Virgo"},         {"Kagra", "kagra", nullptr, "Kagra"},
		{"Indigo", "indigo", nullptr, "Indigo"},      {"CosmicExplorer", "cosmicexplorer", "CE", "CE"},
		{"Einstein Telescope 1", "einstein telescope 1", "ET1", "ET1"},
	};
	const std::string name(detector ? detector : "");
	for (const auto &s : sites)
		if (name.find(s.a) != std::string::npos || name.find(s.b) != std::string::npos || (s.c && name.find(s.c) != std::string::npos))
			return gwat_b200_detector_site(s.canonical, LAT, LON, location, response_tensor) == GWAT_B200_OK ? 0 : -1;
	std::fprintf(stderr, "Unsupported detector\n");
	return -1;
}

// ---- time-domain entry points (src/gwatpy_wrapping.cpp:389-407, 428-480) -----------------------------------------------------------
// The reference's time_waveform (src/waveform_generator.cpp:31-71) knows one model, TaylorT2, which is outside this library's path
// (DESIGN section 7): refused with status -1 and NaN outputs.  For every other method name the reference computes nothing -- status 1
// and the zero-initialised arrays -- and so does this, so that gwatpy.waveform_generator, which binds both symbols when it is imported,
// loads against this library and behaves alike wherever the reference has a defined answer.
static int time_domain_status(const char *generation_method, double fill[1])
{
	const bool taylor = generation_method && std::strstr(generation_method, "Taylor") != nullptr;
	fill[0] = taylor ? NaN : 0.0;
	if (taylor) std::fprintf(stderr, "gwat_b200: time-domain TaylorT2 waveforms are not part of this library\n");
	return taylor ? -1 : 1;
}
int time_waveform_full_py(double *times, int length, double *wf_plus_real, double *wf_plus_imaginary, double *wf_cross_real,
                          double *wf_cross_imaginary, double *wf_x_real, double *wf_x_imaginary, double *wf_y_real, double *wf_y_imaginary,
                          double *wf_b_real, double *wf_b_imaginary, double *wf_l_real, double *wf_l_imaginary, char *generation_method,
                          void *parameters)
{
	(void)times;
	(void)parameters;
	double fill;
	const int status = time_domain_status(generation_method, &fill);
	double *const outs[12] = {wf_plus_real, wf_plus_imaginary, wf_cross_real, wf_cross_imaginary, wf_x_real, wf_x_imaginary,
	                          wf_y_real,    wf_y_imaginary,    wf_b_real,     wf_b_imaginary,     wf_l_real, wf_l_imaginary};
	for (double *o : outs)
		if (o) std::fill(o, o + (length > 0 ? length : 0), fill);
	return status;
}
int time_detector_response_py(double *times, int length, double *response_real, double *response_imaginary, char *detector,
                              char *generation_method, void *parameters)
{
	(void)times;
	(void)detector;
	(void)parameters;
	double fill;
	const int status = time_domain_status(generation_method, &fill);
	for (double *o : {response_real, response_imaginary})
		if (o) std::fill(o, o + (length > 0 ? length : 0), fill);
	return status;
}

// ---- scalar conveniences gwatpy exposes (not on the hot path; plain host arithmetic, src/util.cpp:1492-1540) --------------------
int calculate_chirpmass_py(double mass1, double mass2, double *out)
{
	*out = std::pow(mass1 * mass2, 3. / 5) / std::pow(mass1 + mass2, 1. / 5);
	return 0;
}
int calculate_eta_py(double mass1, double mass2, double *out)
{
	*out = (mass1 * mass2) / std::pow(mass1 + mass2, 2);
	return 0;
}
int calculate_mass1_py(double chirpmass, double eta, double *out)
{
	const double etapow = std::pow(eta, 3. / 5);
	*out = 1. / 2 * (chirpmass / etapow + std::sqrt(1. - 4 * eta) * chirpmass / etapow);
	return 0;
}
int calculate_mass2_py(double chirpmass, double eta, double *out)
{
	const double etapow = std::pow(eta, 3. / 5);
	*out = 1. / 2 * (chirpmass / etapow - std::sqrt(1. - 4 * eta) * chirpmass / etapow);
	return 0;
}
int calculate_chirpmass_vectorized_py(double *mass1, double *mass2, double *out, int length)
{
	for (int i = 0; i < length; i++) calculate_chirpmass_py(mass1[i], mass2[i], out + i);
	return 0;
}
int calculate_eta_vectorized_py(double *mass1, double *mass2, double *out, int length)
{
	for (int i = 0; i < length; i++) calculate_eta_py(mass1[i], mass2[i], out + i);
	return 0;
}
int calculate_mass1_vectorized_py(double *chirpmass, double *eta, double *out, int length)
{
	for (int i = 0; i < length; i++) calculate_mass1_py(chirpmass[i], eta[i], out + i);
	return 0;
}
int calculate_mass2_vectorized_py(double *chirpmass, double *eta, double *out, int length)
{
	for (int i = 0; i < length; i++) calculate_mass2_py(chirpmass[i], eta[i], out + i);
	return 0;
}

// Small closed-form helpers of the gwatpy API (host arithmetic with the reference's expressions; src/gwatpy_wrapping.cpp:77-84, 832-836)
// f_0PN / t_0PN (src/pn_waveform_util.cpp:36-52): frequency at a time before merger at leading order, and its inverse
double f_0PN_py(double t, double chirpmass)
{
	const double factor = 0.07275685064;  // 5^(3/8) / (8 pi)
	return factor * std::pow(chirpmass, -5. / 8.) * std::pow(t, -3. / 8);
}
double t_0PN_py(double f, double chirpmass)
{
	const double factor = 0.07275685064;
	return std::pow(f / factor * std::pow(chirpmass, 5. / 8.), -8. / 3.);
}
// DL_from_Z (src/util.cpp:422-450): luminosity distance in Mpc from the redshift, the piecewise half-power series of D_Z_Config.h;
// -1 for an unknown cosmology or a redshift outside the tables, as the reference returns it
#include "gwat_tables_zd.inc"
int DL_from_Z_py(double z, char *COSMOLOGY, double *out)
{
	*out = -1;
	const int c = gwat_b200_cosmology_index(COSMOLOGY);
	if (c < 0) return 0;
	for (int i = 0; i < 3; i++)
		if (z < gwat_zd_boundaries[c][i + 1]) {
			const double *k = gwat_zd_coeffs[c][i];
			const double rootx = std::sqrt(z);
			double sum = k[0];
			for (int j = 1; j < 12; j++) {  // cosmology_interpolation_function with pow_int's sequential product (src/util.cpp:1585-1597)
				double prod = 1;
				for (int m = 0; m < j; m++) prod = prod * rootx;
				sum += k[j] * prod;
			}
			*out = sum;
			return 0;
		}
	return 0;
}

// Forget the cached network: the next call re-uploads its arrays.  Needed only by callers that overwrite their frequency /
// PSD / data arrays IN PLACE between calls (the cache key samples each array, see array_key).
void gwat_b200_gwatpy_invalidate_network(void)
{
	Session &S = session();
	std::lock_guard<std::mutex> lock(S.mu);
	S.net_key = 0;
}

// Why the last call failed ("" if it did not).
const char *gwat_b200_gwatpy_last_error(void) { return session().last_error.c_str(); }

}  // extern "C"
