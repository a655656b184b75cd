// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product library.
//
// C-ABI driver around the REAL reference implementation (scottperkins/gw_analysis_tools): this file is compiled together
// with the reference's own, unmodified .cpp files read from /root/reference at build time (see oracle/Makefile) into
// oracle/_ref/libgwat_ref.so.  It is the oracle the CUDA path is checked against (tests/, __graft_entry__.smoke()) and
// the CPU baseline bench.py times (`--impl reference`, `cpu_baseline.kind = "reference"`).
//
// It only converts the flat PODs of include/gwat_b200.h into the reference's own structs and calls the reference's
// public functions:
//   fourier_waveform<double>                 src/waveform_generator.cpp:104
//   fourier_detector_response<double>        src/waveform_util.cpp:1070
//   create_coherent_GW_detection_reuse_WF    src/waveform_util.cpp:153
//   Log_Likelihood_internal                  src/mcmc_gw.cpp:801
//   MCMC_prep_params / repack_parameters     src/mcmc_gw.cpp:2492, src/fisher.cpp:2167
//   fisher_numerical                         src/fisher.cpp:81
//   detector_response_functions_equatorial   src/detector_util.cpp:1019, DTOA_DETECTOR :677
// The one reference line that is deliberately NOT executed is `double T = 1./(frequencies[1]-frequencies[0])` on a
// double** (src/mcmc_gw.cpp:2466, pointer arithmetic); T is an explicit argument instead.
#include <complex>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>

#include "util.h"
#include "waveform_generator.h"
#include "waveform_util.h"
#include "detector_util.h"
#include "io_util.h"
#include "mcmc_gw.h"
#include "autocorrelation.h"
#include "fisher.h"
#include "pn_waveform_util.h"
#include "ortho_basis.h"
#include "ppE_utilities.h"
#include "IMRPhenomD.h"
#include "IMRPhenomP.h"

#include "gwat_b200.h"

// not declared in a reference header that this driver can include without pulling the samplers in
double Log_Likelihood_internal(std::complex<double> *data, double *psd, double *frequencies, double *weights,
                               std::complex<double> *detector_response, int length, bool log10F,
                               std::string integration_method);

namespace {

// Holds the heap arrays a gen_params_base borrows, so they die with the scope.
struct ParamBox {
	gen_params_base<double> gp;
	std::vector<double> betappe, bppe, dphi, dsigma, dbeta, dalpha;
	std::vector<int> phii, sigmai, betai, alphai;
};

void to_gen_params(const gwat_b200_source &s, ParamBox &b)
{
	gen_params_base<double> &g = b.gp;
	g.mass1 = s.mass1;
	g.mass2 = s.mass2;
	g.Luminosity_Distance = s.Luminosity_Distance;
	for (int i = 0; i < 3; i++) {
		g.spin1[i] = s.spin1[i];
		g.spin2[i] = s.spin2[i];
	}
	{
		static const char *const names[6] = {"PLANCK15", "PLANCK13", "WMAP9", "WMAP7", "WMAP5", "TESTING_COSMOLOGY"};  // cosmos[], D_Z_Config.h:13
		g.cosmology = names[(s.cosmology >= 0 && s.cosmology < 6) ? s.cosmology : 0];
	}
	g.tc = s.tc;
	g.phiRef = s.phiRef;
	g.f_ref = s.f_ref;
	g.psi = s.psi;
	g.incl_angle = s.incl_angle;
	g.RA = s.RA;
	g.DEC = s.DEC;
	g.gmst = s.gmst;
	g.theta = s.theta;
	g.phi = s.phi;
	g.theta_l = s.theta_l;
	g.phi_l = s.phi_l;
	g.tidal1 = s.tidal1;
	g.tidal2 = s.tidal2;
	g.tidal_s = s.tidal_s;
	g.tidal_a = s.tidal_a;
	g.tidal_weighted = s.tidal_weighted;
	g.delta_tidal_weighted = s.delta_tidal_weighted;
	g.diss_tidal1 = s.diss_tidal1;
	g.diss_tidal2 = s.diss_tidal2;
	g.diss_tidal_s = s.diss_tidal_s;
	g.diss_tidal_a = s.diss_tidal_a;
	g.diss_tidal_weighted = s.diss_tidal_weighted;
	g.chip = s.chip;
	g.phip = s.phip;
	g.PNorder = s.PNorder;
	g.shift_time = s.shift_time != 0;
	g.shift_phase = s.shift_phase != 0;
	g.sky_average = s.sky_average != 0;
	g.tidal_love = s.tidal_love != 0;
	g.tidal_love_error = s.tidal_love_error != 0;
	g.NSflag1 = s.NSflag1 != 0;
	g.NSflag2 = s.NSflag2 != 0;
	g.dep_postmerger = s.dep_postmerger != 0;
	g.equatorial_orientation = s.equatorial_orientation != 0;
	g.horizon_coord = s.horizon_coord != 0;
	g.Nmod = s.Nmod;
	b.betappe.assign(s.betappe, s.betappe + GWAT_B200_MAX_MOD);
	b.bppe.assign(s.bppe, s.bppe + GWAT_B200_MAX_MOD);
	g.betappe = b.betappe.data();
	g.bppe = b.bppe.data();
	g.Nmod_phi = s.Nmod_phi;
	g.Nmod_sigma = s.Nmod_sigma;
	g.Nmod_beta = s.Nmod_beta;
	g.Nmod_alpha = s.Nmod_alpha;
	b.dphi.assign(s.delta_phi, s.delta_phi + GWAT_B200_MAX_MOD);
	b.dsigma.assign(s.delta_sigma, s.delta_sigma + GWAT_B200_MAX_MOD);
	b.dbeta.assign(s.delta_beta, s.delta_beta + GWAT_B200_MAX_MOD);
	b.dalpha.assign(s.delta_alpha, s.delta_alpha + GWAT_B200_MAX_MOD);
	b.phii.assign(s.phii, s.phii + GWAT_B200_MAX_MOD);
	b.sigmai.assign(s.sigmai, s.sigmai + GWAT_B200_MAX_MOD);
	b.betai.assign(s.betai, s.betai + GWAT_B200_MAX_MOD);
	b.alphai.assign(s.alphai, s.alphai + GWAT_B200_MAX_MOD);
	g.delta_phi = b.dphi.data();
	g.delta_sigma = b.dsigma.data();
	g.delta_beta = b.dbeta.data();
	g.delta_alpha = b.dalpha.data();
	g.phii = b.phii.data();
	g.sigmai = b.sigmai.data();
	g.betai = b.betai.data();
	g.alphai = b.alphai.data();
}

void from_gen_params(const gen_params_base<double> &g, gwat_b200_source &s)
{
	std::memset(&s, 0, sizeof(s));
	s.theta = g.theta;
	s.phi = g.phi;
	s.theta_l = g.theta_l;
	s.phi_l = g.phi_l;
	s.delta_tidal_weighted = g.delta_tidal_weighted;
	s.diss_tidal1 = g.diss_tidal1;
	s.diss_tidal2 = g.diss_tidal2;
	s.diss_tidal_s = g.diss_tidal_s;
	s.diss_tidal_a = g.diss_tidal_a;
	s.diss_tidal_weighted = g.diss_tidal_weighted;
	s.PNorder = g.PNorder;
	s.dep_postmerger = g.dep_postmerger;
	s.equatorial_orientation = g.equatorial_orientation;
	s.horizon_coord = g.horizon_coord;
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
	s.chip = g.chip;
	s.phip = g.phip;
	s.shift_time = g.shift_time;
	s.shift_phase = g.shift_phase;
	s.sky_average = g.sky_average;
	s.tidal_love = g.tidal_love;
	s.tidal_love_error = g.tidal_love_error;
	s.NSflag1 = g.NSflag1;
	s.NSflag2 = g.NSflag2;
	s.Nmod = g.Nmod;
	for (int i = 0; i < g.Nmod && i < GWAT_B200_MAX_MOD; i++) {
		s.betappe[i] = g.betappe[i];
		s.bppe[i] = g.bppe[i];
	}
	s.Nmod_phi = g.Nmod_phi;
	s.Nmod_sigma = g.Nmod_sigma;
	s.Nmod_beta = g.Nmod_beta;
	s.Nmod_alpha = g.Nmod_alpha;
	for (int i = 0; i < g.Nmod_phi && i < GWAT_B200_MAX_MOD; i++) { s.delta_phi[i] = g.delta_phi[i]; s.phii[i] = g.phii[i]; }
	for (int i = 0; i < g.Nmod_sigma && i < GWAT_B200_MAX_MOD; i++) { s.delta_sigma[i] = g.delta_sigma[i]; s.sigmai[i] = g.sigmai[i]; }
	for (int i = 0; i < g.Nmod_beta && i < GWAT_B200_MAX_MOD; i++) { s.delta_beta[i] = g.delta_beta[i]; s.betai[i] = g.betai[i]; }
	for (int i = 0; i < g.Nmod_alpha && i < GWAT_B200_MAX_MOD; i++) { s.delta_alpha[i] = g.delta_alpha[i]; s.alphai[i] = g.alphai[i]; }
}

struct ModBox {
	MCMC_modification_struct m;
	std::vector<double> bppe;
	std::vector<int> phii, sigmai, betai, alphai;
};

void to_mod_struct(const gwat_b200_mod *in, ModBox &b)
{
	MCMC_modification_struct &m = b.m;
	if (!in) return;
	m.ppE_Nmod = in->ppE_Nmod;
	b.bppe.assign(in->bppe, in->bppe + GWAT_B200_MAX_MOD);
	m.bppe = b.bppe.data();
	m.gIMR_Nmod_phi = in->gIMR_Nmod_phi;
	m.gIMR_Nmod_sigma = in->gIMR_Nmod_sigma;
	m.gIMR_Nmod_beta = in->gIMR_Nmod_beta;
	m.gIMR_Nmod_alpha = in->gIMR_Nmod_alpha;
	b.phii.assign(in->gIMR_phii, in->gIMR_phii + GWAT_B200_MAX_MOD);
	b.sigmai.assign(in->gIMR_sigmai, in->gIMR_sigmai + GWAT_B200_MAX_MOD);
	b.betai.assign(in->gIMR_betai, in->gIMR_betai + GWAT_B200_MAX_MOD);
	b.alphai.assign(in->gIMR_alphai, in->gIMR_alphai + GWAT_B200_MAX_MOD);
	m.gIMR_phii = b.phii.data();
	m.gIMR_sigmai = b.sigmai.data();
	m.gIMR_betai = b.betai.data();
	m.gIMR_alphai = b.alphai.data();
	m.NSflag1 = in->NSflag1 != 0;
	m.NSflag2 = in->NSflag2 != 0;
	m.tidal_love = in->tidal_love != 0;
	m.tidal_love_error = in->tidal_love_error != 0;
}

// MCMC_prep_params allocates gen_params arrays with new[] (src/mcmc_gw.cpp:2529-2553); MCMC_likelihood_wrapper frees them
// the same way (:2753-2777).
void free_prepped(gen_params_base<double> &gp, const std::string &method, const MCMC_modification_struct &m)
{
	if (!check_mod(method)) return;
	if (method.find("ppE") != std::string::npos || check_theory_support(method)) {
		delete[] gp.betappe;
	} else if (method.find("gIMR") != std::string::npos) {
		if (m.gIMR_Nmod_phi != 0) delete[] gp.delta_phi;
		if (m.gIMR_Nmod_sigma != 0) delete[] gp.delta_sigma;
		if (m.gIMR_Nmod_beta != 0) delete[] gp.delta_beta;
		if (m.gIMR_Nmod_alpha != 0) delete[] gp.delta_alpha;
	}
}

// The body of MCMC_likelihood_extrinsic below its first line (src/mcmc_gw.cpp:2467-2486) for one shared grid.
double extrinsic_ll(gen_params_base<double> *gp, const std::string &method, std::string *dets, int D, double *f, int L,
                    std::complex<double> **data, double **psd, double *weights, const std::string &integ, bool log10F)
{
	std::vector<std::vector<std::complex<double>>> store(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> resp(D);
	for (int d = 0; d < D; d++) resp[d] = store[d].data();
	create_coherent_GW_detection_reuse_WF(dets, D, f, L, gp, method, resp.data());
	double ll = 0;
	for (int d = 0; d < D; d++) ll += Log_Likelihood_internal(data[d], psd[d], f, weights, resp[d], L, log10F, integ);
	return ll;
}

}  // namespace

extern "C" {

int oracle_ref_abi_version(void) { return GWAT_B200_ABI_VERSION; }
int oracle_ref_max_threads(void) { return omp_get_max_threads(); }
size_t oracle_ref_sizeof_source(void) { return sizeof(gwat_b200_source); }
size_t oracle_ref_sizeof_mod(void) { return sizeof(gwat_b200_mod); }

// fourier_waveform<double>, src/waveform_generator.cpp:104
int oracle_ref_fourier_waveform(const char *method, const gwat_b200_source *src, const double *f, int L, double *hp_re,
                                double *hp_im, double *hc_re, double *hc_im)
{
	ParamBox b;
	to_gen_params(*src, b);
	waveform_polarizations<double> wp;
	assign_polarizations(std::string(method), &wp);
	wp.allocate_memory(L);
	for (int i = 0; i < L; i++) {
		wp.hplus[i] = 0;
		wp.hcross[i] = 0;
	}
	int st = fourier_waveform(const_cast<double *>(f), L, &wp, std::string(method), &b.gp);
	for (int i = 0; i < L; i++) {
		if (hp_re) hp_re[i] = wp.hplus[i].real();
		if (hp_im) hp_im[i] = wp.hplus[i].imag();
		if (hc_re) hc_re[i] = wp.hcross[i].real();
		if (hc_im) hc_im[i] = wp.hcross[i].imag();
	}
	wp.deallocate_memory();
	return st;
}

// fourier_detector_response<double>, src/waveform_util.cpp:1070
int oracle_ref_fourier_detector_response(const char *method, const char *detector, const gwat_b200_source *src,
                                         const double *f, int L, double *re, double *im)
{
	ParamBox b;
	to_gen_params(*src, b);
	std::vector<std::complex<double>> r(L);
	int st = fourier_detector_response(const_cast<double *>(f), L, r.data(), std::string(detector), std::string(method),
	                                   &b.gp, (double *)NULL);
	for (int i = 0; i < L; i++) {
		re[i] = r[i].real();
		im[i] = r[i].imag();
	}
	return st;
}

// create_coherent_GW_detection_reuse_WF, src/waveform_util.cpp:153.  re/im shape [D][L].
int oracle_ref_coherent_response(const char *method, const gwat_b200_source *src, int D, const char *const *detectors,
                                 const double *f, int L, double *re, double *im)
{
	ParamBox b;
	to_gen_params(*src, b);
	std::vector<std::string> dets(D);
	for (int d = 0; d < D; d++) dets[d] = detectors[d];
	std::vector<std::vector<std::complex<double>>> store(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> resp(D);
	for (int d = 0; d < D; d++) resp[d] = store[d].data();
	create_coherent_GW_detection_reuse_WF(dets.data(), D, const_cast<double *>(f), L, &b.gp, std::string(method),
	                                      resp.data());
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) {
			re[(size_t)d * L + i] = store[d][i].real();
			im[(size_t)d * L + i] = store[d][i].imag();
		}
	return 1;
}

// Log_Likelihood_internal, src/mcmc_gw.cpp:801
double oracle_ref_log_likelihood_internal(const double *data_re, const double *data_im, const double *psd,
                                          const double *f, const double *weights, const double *resp_re,
                                          const double *resp_im, int L, int log10F, const char *integ)
{
	std::vector<std::complex<double>> d(L), r(L);
	for (int i = 0; i < L; i++) {
		d[i] = std::complex<double>(data_re[i], data_im[i]);
		r[i] = std::complex<double>(resp_re[i], resp_im[i]);
	}
	return Log_Likelihood_internal(d.data(), const_cast<double *>(psd), const_cast<double *>(f),
	                               const_cast<double *>(weights), r.data(), L, log10F != 0, std::string(integ));
}

// W x { create_coherent_GW_detection + sum_d Log_Likelihood_internal } from physical parameters (src/mcmc_gw.cpp:2473-2486),
// OpenMP over walkers = the reference's one-chain-per-thread pool (src/mcmc_sampler.cpp:347-447).
int oracle_ref_loglike_batch(const char *method, int W, const gwat_b200_source *sources, int D,
                             const char *const *detectors, const double *f, int L, const double *psd,
                             const double *data_re, const double *data_im, const double *weights, const char *integ,
                             int log10F, int nthreads, double *logL)
{
	std::vector<std::string> dets(D);
	for (int d = 0; d < D; d++) dets[d] = detectors[d];
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> datap(D);
	std::vector<double *> psdp(D);
	for (int d = 0; d < D; d++) {
		for (int i = 0; i < L; i++) data[d][i] = std::complex<double>(data_re[(size_t)d * L + i], data_im[(size_t)d * L + i]);
		datap[d] = data[d].data();
		psdp[d] = const_cast<double *>(psd) + (size_t)d * L;
	}
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	const std::string m(method), im(integ);
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
	for (int w = 0; w < W; w++) {
		ParamBox b;
		to_gen_params(sources[w], b);
		std::vector<std::string> ldets(dets);
		logL[w] = extrinsic_ll(&b.gp, m, ldets.data(), D, const_cast<double *>(f), L, datap.data(), psdp.data(),
		                       const_cast<double *>(weights), im, log10F != 0);
	}
	return 0;
}

// W x the intrinsic branches of MCMC_likelihood_wrapper (src/mcmc_gw.cpp:2603-2722): coalescence-frame source, then per
// detector maximized_Log_Likelihood_aligned_spin_internal (PhenomD family: horizon response with theta = phi = psi = 0) or
// maximized_Log_Likelihood_unaligned_spin_internal (PhenomP family: both polarisations).
int oracle_ref_loglike_maximized_batch(const char *method, int W, const gwat_b200_source *sources, int D,
                                       const char *const *detectors, const double *f, int L, const double *psd,
                                       const double *data_re, const double *data_im, int nthreads, double *logL)
{
	std::vector<std::string> dets(D);
	for (int d = 0; d < D; d++) dets[d] = detectors[d];
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(L));
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) data[d][i] = std::complex<double>(data_re[(size_t)d * L + i], data_im[(size_t)d * L + i]);
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	const std::string m(method);
	const bool precessing = m.find("IMRPhenomP") != std::string::npos;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
	for (int w = 0; w < W; w++) {
		ParamBox b;
		to_gen_params(sources[w], b);
		gen_params_base<double> &gp = b.gp;
		gp.theta = 0;
		gp.phi = 0;
		gp.psi = 0;
		gp.phiRef = 1;
		gp.f_ref = precessing ? 20 : 10;
		gp.incl_angle = 0;
		gp.tc = 1;
		fftw_outline plan;
		allocate_FFTW_mem_forward(&plan, L);
		double ll = 0;
		double *fp = const_cast<double *>(f);
		if (!precessing) {
			std::vector<std::complex<double>> response(L);
			for (int d = 0; d < D; d++) {
				fourier_detector_response_horizon(fp, L, response.data(), dets[d], m, &gp);
				ll += maximized_Log_Likelihood_aligned_spin_internal(data[d].data(), const_cast<double *>(psd) + (size_t)d * L, fp,
				                                                     response.data(), (size_t)L, &plan);
			}
		} else {
			waveform_polarizations<double> wp;
			assign_polarizations(m, &wp);
			wp.allocate_memory(L);
			fourier_waveform(fp, L, &wp, m, &gp);
			for (int d = 0; d < D; d++)
				ll += maximized_Log_Likelihood_unaligned_spin_internal(data[d].data(), const_cast<double *>(psd) + (size_t)d * L, fp, wp.hplus,
				                                                       wp.hcross, (size_t)L, &plan);
			wp.deallocate_memory();
		}
		deallocate_FFTW_mem(&plan);
		logL[w] = ll;
	}
	return 0;
}

// W x MCMC_likelihood_wrapper, extrinsic branch (src/mcmc_gw.cpp:2569-2791) with the globals passed explicitly:
//   MCMC_prep_params (:2492) -> repack_parameters("MCMC_"+method) (src/fisher.cpp:2167) -> tc_ref = T - tc (:2467) ->
//   create_coherent_GW_detection (:2474) -> sum_d Log_Likelihood_internal (:2476).
// If `sources_out` is not NULL the repacked physical parameters (with tc = tc_ref) are also returned.
int oracle_ref_loglike_mcmc_batch(const char *method, const gwat_b200_mod *mod, int dimension, int W,
                                  const double *params, double gmst, double T_segment, int D,
                                  const char *const *detectors, const double *f, int L, const double *psd,
                                  const double *data_re, const double *data_im, const double *weights,
                                  const char *integ, int log10F, int nthreads, double *logL,
                                  gwat_b200_source *sources_out)
{
	std::vector<std::string> dets(D);
	for (int d = 0; d < D; d++) dets[d] = detectors[d];
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> datap(D);
	std::vector<double *> psdp(D);
	const bool have_data = data_re && data_im && psd && logL;
	for (int d = 0; d < D && have_data; d++) {
		for (int i = 0; i < L; i++) data[d][i] = std::complex<double>(data_re[(size_t)d * L + i], data_im[(size_t)d * L + i]);
		datap[d] = data[d].data();
		psdp[d] = const_cast<double *>(psd) + (size_t)d * L;
	}
	ModBox mb;
	to_mod_struct(mod, mb);
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	const std::string m(method), im(integ ? integ : "SIMPSONS");
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
	for (int w = 0; w < W; w++) {
		std::vector<double> temp(dimension);
		gen_params_base<double> gp;
		std::string local_gen = MCMC_prep_params(const_cast<double *>(params) + (size_t)w * dimension, temp.data(), &gp,
		                                         dimension, m, &mb.m);
		gp.gmst = gmst;  // mcmc_gmst is a per-TU static (include/gwat/mcmc_gw.h:35); pass it explicitly
		repack_parameters(temp.data(), &gp, "MCMC_" + m, dimension, (gen_params_base<double> *)NULL);
		gp.tc = T_segment - gp.tc;  // src/mcmc_gw.cpp:2467,2473 with T explicit
		if (sources_out) {
			from_gen_params(gp, sources_out[w]);
		}
		if (have_data) {
			std::vector<std::string> ldets(dets);
			logL[w] = extrinsic_ll(&gp, local_gen, ldets.data(), D, const_cast<double *>(f), L, datap.data(),
			                       psdp.data(), const_cast<double *>(weights), im, log10F != 0);
		}
		free_prepped(gp, local_gen, mb.m);
	}
	return 0;
}

// fisher_numerical, src/fisher.cpp:81.  detector_index < 0: sum over detectors (MCMC_fisher_wrapper, src/mcmc_gw.cpp:2298-2312).
// psd shape [D][L]; fisher shape [S][dim][dim].
int oracle_ref_fisher_numerical_batch(const char *method, int detector_index, int reference_index, int dimension,
                                      int order, int S, const gwat_b200_source *sources, int D,
                                      const char *const *detectors, const double *f, int L, const double *psd,
                                      int nthreads, double *fisher)
{
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	const std::string m(method);
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
	for (int s = 0; s < S; s++) {
		std::vector<double> flat((size_t)dimension * dimension);
		std::vector<double *> rows(dimension);
		for (int i = 0; i < dimension; i++) rows[i] = flat.data() + (size_t)i * dimension;
		double *out = fisher + (size_t)s * dimension * dimension;
		for (int i = 0; i < dimension * dimension; i++) out[i] = 0;
		const int d0 = detector_index < 0 ? 0 : detector_index;
		const int d1 = detector_index < 0 ? D : detector_index + 1;
		for (int d = d0; d < d1; d++) {
			ParamBox b;
			to_gen_params(sources[s], b);
			fisher_numerical(const_cast<double *>(f), L, m, std::string(detectors[d]), std::string(detectors[reference_index]),
			                 rows.data(), dimension, &b.gp, order, (int *)NULL, (int *)NULL,
			                 const_cast<double *>(psd) + (size_t)d * L);
			for (int i = 0; i < dimension * dimension; i++) out[i] += flat[i];
		}
	}
	return 0;
}

// detector_response_functions_equatorial (src/detector_util.cpp:1019) and DTOA_DETECTOR (:677) vs detector 0.  Outputs [W][D].
int oracle_ref_antenna_batch(int W, const double *RA, const double *DEC, const double *psi, double gmst, int D,
                             const char *const *detectors, double *Fplus, double *Fcross, double *dtoa)
{
	bool active[6] = {true, true, false, false, false, false};
	for (int w = 0; w < W; w++)
		for (int d = 0; d < D; d++) {
			double fp, fc;
			det_res_pat<double> r;
			r.Fplus = &fp;
			r.Fcross = &fc;
			r.active_polarizations = active;
			detector_response_functions_equatorial(std::string(detectors[d]), RA[w], DEC[w], psi[w], gmst, &r);
			Fplus[(size_t)w * D + d] = fp;
			Fcross[(size_t)w * D + d] = fc;
			dtoa[(size_t)w * D + d] = DTOA_DETECTOR(RA[w], DEC[w], gmst, std::string(detectors[0]), std::string(detectors[d]));
		}
	return 0;
}

// populate_noise (src/detector_util.cpp:87): amplitude spectral density of a named analytic curve; psd = asd^2.
// gauleg (src/ortho_basis.cpp:14-48) and the pow(10, .) step of its callers (src/waveform_util.cpp:3113-3116)
int oracle_ref_gauleg_grid(double f_lower, double f_upper, int n, int log10F, double *freqs, double *weights)
{
	if (log10F) {
		gauleg(log10(f_lower), log10(f_upper), freqs, weights, n);
		for (int i = 0; i < n; i++) freqs[i] = pow(10, freqs[i]);
	} else {
		gauleg(f_lower, f_upper, freqs, weights, n);
	}
	return 0;
}

int oracle_ref_populate_noise(const double *f, const char *curve, double *asd, int L)
{
	populate_noise(const_cast<double *>(f), std::string(curve), asd, L);
	return 0;
}

// allocate_LOSC_data (src/io_util.cpp:523-661).  Outputs detector-major [D][psd_length]; freqs[psd_length].
int oracle_ref_losc(int D, const char *const *data_files, const char *psd_file, double trigger_time, double post_merger_duration,
                    int psd_length, int data_file_length, double *freqs, double *psds, double *data_re, double *data_im)
{
	std::vector<std::string> files(D);
	for (int d = 0; d < D; d++) files[d] = data_files[d];
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(psd_length));
	std::vector<std::vector<double>> p(D, std::vector<double>(psd_length)), f(D, std::vector<double>(psd_length));
	std::vector<std::complex<double> *> dptr(D);
	std::vector<double *> pptr(D), fptr(D);
	for (int d = 0; d < D; d++) { dptr[d] = data[d].data(); pptr[d] = p[d].data(); fptr[d] = f[d].data(); }
	allocate_LOSC_data(files.data(), std::string(psd_file), D, psd_length, data_file_length, trigger_time, post_merger_duration,
	                   dptr.data(), pptr.data(), fptr.data());
	for (int d = 0; d < D; d++)
		for (int i = 0; i < psd_length; i++) {
			psds[(size_t)d * psd_length + i] = p[d][i];
			data_re[(size_t)d * psd_length + i] = data[d][i].real();
			data_im[(size_t)d * psd_length + i] = data[d][i].imag();
		}
	for (int i = 0; i < psd_length; i++) freqs[i] = f[0][i];
	return 0;
}

// fourier_amplitude<double> / fourier_phase<double> (src/waveform_generator.cpp:537-670, 672-814)
int oracle_ref_fourier_amplitude_phase(const char *method, const gwat_b200_source *src, const double *f, int L, double *amplitude,
                                       double *phase)
{
	ParamBox b;
	to_gen_params(*src, b);
	if (amplitude) fourier_amplitude(const_cast<double *>(f), L, amplitude, std::string(method), &b.gp);
	ParamBox b2;
	to_gen_params(*src, b2);
	if (phase) fourier_phase(const_cast<double *>(f), L, phase, std::string(method), &b2.gp);
	return 0;
}

double oracle_ref_gps_to_gmst_radian(double gps) { return gps_to_GMST_radian(gps); }

// calculate_snr (src/waveform_util.cpp:290-344): SNR of one template in one detector against a named noise curve.
double oracle_ref_calculate_snr(const char *curve, const char *detector, const char *method, const gwat_b200_source *src,
                                const double *f, int L, const char *integration_method, const double *weights, int log10F)
{
	ParamBox b;
	to_gen_params(*src, b);
	return calculate_snr(std::string(curve), std::string(detector), std::string(method), &b.gp, const_cast<double *>(f), L,
	                     std::string(integration_method), const_cast<double *>(weights), log10F != 0);
}

// transform_orientation_coords (src/waveform_util.cpp:1535-1595) for a named (terrestrial) detector: incl_angle and psi it derives
int oracle_ref_transform_orientation_coords(const char *method, const char *detector, const gwat_b200_source *src, double *incl_angle, double *psi)
{
	ParamBox b;
	to_gen_params(*src, b);
	transform_orientation_coords(&b.gp, std::string(method), std::string(detector));
	*incl_angle = b.gp.incl_angle;
	*psi = b.gp.psi;
	return 0;
}

// match (src/waveform_util.cpp:41-89): overlap maximised over a relative time shift (FFTW_BACKWARD = the stand-in's DFT here).
double oracle_ref_match(const double *a_re, const double *a_im, const double *b_re, const double *b_im, const double *psd, const double *f,
                        int L)
{
	std::vector<std::complex<double>> a(L), b(L);
	for (int i = 0; i < L; i++) {
		a[i] = std::complex<double>(a_re[i], a_im[i]);
		b[i] = std::complex<double>(b_re[i], b_im[i]);
	}
	return match(a.data(), b.data(), const_cast<double *>(psd), const_cast<double *>(f), L);
}

// MCMC_prep_params (src/mcmc_gw.cpp:2492-2568) on a source that arrives with the fields `src` carries: returns temp_params, the
// flags it set on the gen_params object (as a gwat_b200_source) and the modification layout it attached.
int oracle_ref_mcmc_prep_params(const char *method, const gwat_b200_mod *mod, int dimension, const double *param, const gwat_b200_source *src,
                                double *temp_params, gwat_b200_source *out)
{
	ParamBox b;
	to_gen_params(*src, b);
	ModBox mb;
	to_mod_struct(mod, mb);
	gen_params_base<double> gp = b.gp;
	std::string local = MCMC_prep_params(const_cast<double *>(param), temp_params, &gp, dimension, std::string(method), &mb.m);
	gwat_b200_source s = *src;
	s.sky_average = gp.sky_average;
	s.tidal_love = gp.tidal_love;
	s.tidal_love_error = gp.tidal_love_error;
	s.f_ref = gp.f_ref;
	s.shift_time = gp.shift_time;
	s.shift_phase = gp.shift_phase;
	s.gmst = gp.gmst;
	s.NSflag1 = gp.NSflag1;
	s.NSflag2 = gp.NSflag2;
	s.Nmod = gp.Nmod;
	s.Nmod_phi = gp.Nmod_phi;
	s.Nmod_sigma = gp.Nmod_sigma;
	s.Nmod_beta = gp.Nmod_beta;
	s.Nmod_alpha = gp.Nmod_alpha;
	if (check_mod(local) && (local.find("ppE") != std::string::npos || check_theory_support(local)))
		for (int i = 0; i < gp.Nmod && i < GWAT_B200_MAX_MOD; i++) s.bppe[i] = gp.bppe[i];
	*out = s;
	free_prepped(gp, local, mb.m);
	return (int)local.size();
}

// The parameter handling of MCMC_likelihood_wrapper / MCMC_fisher_wrapper in an INTRINSIC run (src/mcmc_gw.cpp:2576-2581, 2236-2241):
// MCMC_prep_params with mcmc_intrinsic set (sky_average = true, :2494; the flag is a static of the reference's translation unit, so it
// is applied here on the object MCMC_prep_params returns) -> repack_parameters("MCMC_" + method) (src/fisher.cpp:2308-2376, 2420-2431).
int oracle_ref_repack_mcmc_intrinsic(const char *method, const gwat_b200_mod *mod, int dimension, int W, const double *params, double gmst,
                                     gwat_b200_source *sources_out)
{
	ModBox mb;
	to_mod_struct(mod, mb);
	const std::string m(method);
	for (int w = 0; w < W; w++) {
		std::vector<double> temp(dimension);
		gen_params_base<double> gp;
		std::string local_gen = MCMC_prep_params(const_cast<double *>(params) + (size_t)w * dimension, temp.data(), &gp, dimension, m, &mb.m);
		gp.sky_average = true;
		gp.gmst = gmst;
		repack_parameters(temp.data(), &gp, "MCMC_" + m, dimension, (gen_params_base<double> *)NULL);
		from_gen_params(gp, sources_out[w]);
		free_prepped(gp, local_gen, mb.m);
	}
	return 0;
}

// The autocorrelation lengths of calc_ac_vals (src/mcmc_io_util.cpp:434-520): auto_corr_from_data_batch with one cumulative segment
// (src/autocorrelation.cpp:152-217), positions[n_chains][steps][dimension] from step `begin` on; ac[n_chains][dimension].  tau (may be
// NULL) receives the estimator before truncation, from auto_correlation_spectral_windowed itself (:401-462) with the plan pair of :292-297.
int oracle_ref_autocorrelation_lengths(int n_chains, int dimension, int steps, const double *positions, int begin, double target, int nthreads,
                                       int *ac, double *tau)
{
	const int n = steps - begin;
	std::vector<std::vector<double *>> rows(n_chains, std::vector<double *>(n));
	std::vector<double **> data(n_chains);
	for (int c = 0; c < n_chains; c++) {
		for (int i = 0; i < n; i++) rows[c][i] = const_cast<double *>(positions) + ((size_t)c * steps + begin + i) * dimension;
		data[c] = rows[c].data();
	}
	std::vector<std::vector<int>> seg((size_t)n_chains * dimension, std::vector<int>(1));
	std::vector<std::vector<int *>> outp(n_chains, std::vector<int *>(dimension));
	std::vector<int **> out(n_chains);
	for (int c = 0; c < n_chains; c++) {
		for (int d = 0; d < dimension; d++) outp[c][d] = seg[(size_t)c * dimension + d].data();
		out[c] = outp[c].data();
	}
	auto_corr_from_data_batch(data.data(), n, dimension, n_chains, out.data(), 1, target, nthreads > 0 ? nthreads : omp_get_max_threads(), true);
	for (int c = 0; c < n_chains; c++)
		for (int d = 0; d < dimension; d++) ac[(size_t)c * dimension + d] = seg[(size_t)c * dimension + d][0];
	if (tau && n > 2) {
		const int L = 2 * std::pow(2, std::ceil(std::log2(n)));
		fftw_outline pf, pr;
		allocate_FFTW_mem_forward(&pf, L);
		allocate_FFTW_mem_reverse(&pr, L);
		std::vector<double> chain(n);
		for (int c = 0; c < n_chains; c++)
			for (int d = 0; d < dimension; d++) {
				for (int i = 0; i < n; i++) chain[i] = rows[c][i][d];
				auto_correlation_spectral_windowed(chain.data(), n, 0, &tau[(size_t)c * dimension + d], &pf, &pr);
			}
		deallocate_FFTW_mem(&pf);
		deallocate_FFTW_mem(&pr);
	}
	return 0;
}

// small helpers behind gwatpy's DL_from_Z_py, t_0PN_py, f_0PN_py (src/util.cpp:422, src/pn_waveform_util.cpp:36-52)
double oracle_ref_dl_from_z(double z, const char *cosmology) { return DL_from_Z(z, std::string(cosmology)); }
double oracle_ref_t_0pn(double f, double chirpmass) { return t_0PN<double>(f, chirpmass); }
double oracle_ref_f_0pn(double t, double chirpmass) { return f_0PN<double>(t, chirpmass); }

// The reference's own site constants (include/gwat/detector_util.h) in the order get_detector_parameters knows them
// (src/gwatpy_wrapping.cpp:743-831): 0 Hanford, 1 Livingston, 2 Virgo, 3 Kagra, 4 Indigo, 5 Cosmic Explorer, 6 ET1.
int oracle_ref_detector_site(int which, double *lat, double *lon, double *location, double *response_tensor)
{
	const double lats[7] = {H_LAT, L_LAT, V_LAT, K_LAT, I_LAT, CE_LAT, ET1_LAT};
	const double lons[7] = {H_LONG, L_LONG, V_LONG, K_LONG, I_LONG, CE_LONG, ET1_LONG};
	const double *locs[7] = {H_location, L_location, V_location, K_location, I_location, CE_location, ET1_location};
	const double(*tens[7])[3] = {Hanford_D, Livingston_D, Virgo_D, Kagra_D, Indigo_D, CE_D, ET1_D};
	if (which < 0 || which > 6) return -1;
	*lat = lats[which];
	*lon = lons[which];
	for (int i = 0; i < 3; i++) {
		location[i] = locs[which][i];
		for (int j = 0; j < 3; j++) response_tensor[3 * i + j] = tens[which][i][j];
	}
	return 0;
}

// time_waveform (src/waveform_generator.cpp:31-71) the way time_waveform_full_py calls it (src/gwatpy_wrapping.cpp:428-480): zeroed arrays in,
// status out; hplus/hcross receive whatever the reference wrote.
int oracle_ref_time_waveform(const char *method, const gwat_b200_source *src, const double *times, int length, double *hp_re, double *hp_im,
                             double *hc_re, double *hc_im)
{
	ParamBox b;
	to_gen_params(*src, b);
	std::vector<std::complex<double>> hp(length, 0.0), hc(length, 0.0);
	waveform_polarizations<double> wp;
	wp.hplus = hp.data();
	wp.hcross = hc.data();
	const int status = time_waveform<double>(const_cast<double *>(times), length, &wp, std::string(method), &b.gp);
	for (int i = 0; i < length; i++) {
		hp_re[i] = hp[i].real();
		hp_im[i] = hp[i].imag();
		hc_re[i] = hc[i].real();
		hc_im[i] = hc[i].imag();
	}
	return status;
}

// pack_local_mod_structure (src/mcmc_gw.cpp:3401-3476); counts[4] and idx[4][GWAT_B200_MAX_MOD] receive the local structure.
int oracle_ref_pack_local_mod_structure(int min_dim, int max_dim, const int *status, const char *waveform_extended, const gwat_b200_mod *full,
                                        int *counts, int *idx)
{
	mcmc_data_interface iface;
	iface.min_dim = min_dim;
	iface.max_dim = max_dim;
	ModBox fb;
	to_mod_struct(full, fb);
	MCMC_modification_struct loc;
	loc.gIMR_Nmod_phi = loc.gIMR_Nmod_sigma = loc.gIMR_Nmod_beta = loc.gIMR_Nmod_alpha = 0;
	loc.gIMR_phii = loc.gIMR_sigmai = loc.gIMR_betai = loc.gIMR_alphai = NULL;
	pack_local_mod_structure(&iface, (double *)NULL, const_cast<int *>(status), std::string(waveform_extended), (void *)NULL, &fb.m, &loc);
	counts[0] = loc.gIMR_Nmod_phi;
	counts[1] = loc.gIMR_Nmod_sigma;
	counts[2] = loc.gIMR_Nmod_beta;
	counts[3] = loc.gIMR_Nmod_alpha;
	int *arrs[4] = {loc.gIMR_phii, loc.gIMR_sigmai, loc.gIMR_betai, loc.gIMR_alphai};
	for (int k = 0; k < 4; k++)
		for (int i = 0; i < GWAT_B200_MAX_MOD; i++) idx[k * GWAT_B200_MAX_MOD + i] = (arrs[k] && i < counts[k]) ? arrs[k][i] : 0;
	for (int k = 0; k < 4; k++) delete[] arrs[k];
	return 0;
}

// Intermediate per-walker quantities of IMRPhenomD's setup, for unit-testing the GPU setup kernel
// (src/IMRPhenomD.cpp:414-466).  out[0..] = M, eta, chirpmass, chi_pn, A0, fRD, fdamp, f1, f3, f1_phase, f2_phase
int oracle_ref_phenomd_intermediates(const gwat_b200_source *src, double *out)
{
	ParamBox b;
	to_gen_params(*src, b);
	source_parameters<double> sp;
	prep_source_parameters(&sp, &b.gp, std::string("IMRPhenomD"));
	IMRPhenomD<double> model;
	lambda_parameters<double> lambda;
	model.assign_lambda_param(&sp, &lambda);
	model.post_merger_variables(&sp);
	sp.f1_phase = 0.018 / sp.M;
	sp.f2_phase = sp.fRD / 2.;
	sp.f1 = 0.014 / sp.M;
	sp.f3 = model.fpeak(&sp, &lambda);
	out[0] = sp.M;
	out[1] = sp.eta;
	out[2] = sp.chirpmass;
	out[3] = sp.chi_pn;
	out[4] = sp.A0;
	out[5] = sp.fRD;
	out[6] = sp.fdamp;
	out[7] = sp.f1;
	out[8] = sp.f3;
	out[9] = sp.f1_phase;
	out[10] = sp.f2_phase;
	return 0;
}

}  // extern "C"
