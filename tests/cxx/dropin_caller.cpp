// TEST ONLY.  A caller written the way GWAT's own programs are (examples/*/src/*.cpp, tests/src/test_mcmc.cpp:1026-1109,
// tests/src/test_fishers.cpp:425-548): it includes GWAT's headers and calls GWAT's C++ API with GWAT's types.  It is compiled ONCE
// against the reference's headers and linked twice (tests/test_dropin_link.py):
//   dropin_caller_ref    -> oracle/_ref/libgwat_ref.so alone: every call runs the reference's CPU code
//   dropin_caller_b200   -> libgwat_b200_dropin.so first, then libgwat_ref.so: the hot-path symbols resolve to the GPU library, the
//                           rest (populate_noise, MCMC_prep_params, repack_parameters, allocate_2D_array ...) to the reference
// Each run prints "name value" lines with 17 significant digits; the test compares the two outputs.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include <gwat/detector_util.h>
#include <gwat/fisher.h>
#include <gwat/mcmc_gw.h>
#include <gwat/ortho_basis.h>
#include <gwat/util.h>
#include <gwat/waveform_generator.h>
#include <gwat/waveform_util.h>

// only the b200 variant defines these (weak: NULL in the reference-only link)
extern "C" void gwat_b200_dropin_bind_mcmc(std::complex<double> **data, double **noise, double **frequencies, int *data_length,
                                           std::string *detectors, int num_detectors, const char *generation_method,
                                           MCMC_modification_struct *mod_struct, double gmst, int deriv_order) __attribute__((weak));

extern "C" void gwat_b200_dropin_set_intrinsic(int intrinsic) __attribute__((weak));

static void put(const char *name, double v) { std::printf("%s %.17g\n", name, v); }
static void put(const std::string &name, int i, double v) { std::printf("%s[%d] %.17g\n", name.c_str(), i, v); }

int main()
{
	gen_params gp;
	gp.mass1 = 36.4;
	gp.mass2 = 29.3;
	gp.Luminosity_Distance = 500;
	gp.spin1[0] = gp.spin1[1] = 0;
	gp.spin2[0] = gp.spin2[1] = 0;
	gp.spin1[2] = .3;
	gp.spin2[2] = .2;
	gp.RA = .275;
	gp.DEC = -.44;
	gp.psi = .2;
	gp.incl_angle = .51;
	gp.gmst = 2.1;
	gp.f_ref = 20;
	gp.phiRef = 2.;
	gp.tc = 3.;
	gp.shift_time = true;
	gp.shift_phase = true;
	gp.equatorial_orientation = false;
	gp.horizon_coord = false;
	gp.sky_average = false;

	const int D = 3, L = 2048;
	std::string detectors[3] = {"Hanford", "Livingston", "Virgo"};
	// one block per kind of array, detector after detector: the reference's `frequencies[1]-frequencies[0]` (a double** difference,
	// src/mcmc_gw.cpp:2466) is then exactly L, whatever the allocator does
	std::vector<double> fblock((size_t)D * L), pblock((size_t)D * L);
	std::vector<std::complex<double>> dblock((size_t)D * L);
	double *freq[3], *psd[3];
	std::complex<double> *data[3];
	int lengths[3] = {L, L, L};
	double *weights[3] = {NULL, NULL, NULL};  // (MCMC_likelihood_extrinsic indexes the array even for Simpson's rule, src/mcmc_gw.cpp:2476)
	for (int d = 0; d < D; d++) {
		freq[d] = fblock.data() + (size_t)d * L;
		psd[d] = pblock.data() + (size_t)d * L;
		data[d] = dblock.data() + (size_t)d * L;
		for (int i = 0; i < L; i++) freq[d][i] = 20. + 0.5 * i;
		populate_noise(freq[d], "aLIGO_analytic", psd[d], L, 48);
		for (int i = 0; i < L; i++) psd[d][i] *= psd[d][i];
	}

	// ---- fourier_waveform (new-style and the three legacy overloads) ---------------------------------------------------------
	for (const char *method : {"IMRPhenomD", "IMRPhenomPv2"}) {
		gen_params g = gp;
		if (std::string(method) == "IMRPhenomPv2") {
			g.spin1[0] = .3;
			g.spin1[1] = .1;
			g.spin2[1] = -.2;
		}
		std::vector<std::complex<double>> hp(L), hc(L);
		waveform_polarizations<double> wp;
		wp.hplus = hp.data();
		wp.hcross = hc.data();
		int st = fourier_waveform(freq[0], L, &wp, std::string(method), &g);
		put((std::string("status_wf_") + method).c_str(), st);
		for (int i : {0, 7, 100, 777, 1500}) {
			put(std::string("hp_re_") + method, i, hp[i].real());
			put(std::string("hp_im_") + method, i, hp[i].imag());
			put(std::string("hc_re_") + method, i, hc[i].real());
			put(std::string("hc_im_") + method, i, hc[i].imag());
		}
		// the legacy overloads know the aligned-spin models only (src/waveform_generator.cpp:406-494); for anything else they leave
		// the (here zero-initialised) output untouched and return 1
		std::vector<std::complex<double>> h1(L);
		put((std::string("status_legacy_") + method).c_str(), fourier_waveform(freq[0], L, h1.data(), std::string(method), &g));
		std::vector<double> re(L), im(L);
		fourier_waveform(freq[0], L, re.data(), im.data(), std::string(method), &g);
		for (int i : {0, 100, 777}) {
			put(std::string("legacy_complex_re_") + method, i, h1[i].real());
			put(std::string("legacy_complex_im_") + method, i, h1[i].imag());
			if (std::string(method) == "IMRPhenomD") {  // (the split overload copies out of an uninitialised malloc otherwise)
				put(std::string("legacy_split_re_") + method, i, re[i]);
				put(std::string("legacy_split_im_") + method, i, im[i]);
			}
		}
	}

	// ---- calculate_snr (noise curve by name; Simpson on the data grid, and Gauss-Legendre in log10 f with its own weights) ------
	for (const char *method : {"IMRPhenomD", "IMRPhenomPv2"}) {
		gen_params g = gp;
		if (std::string(method) == "IMRPhenomPv2") {
			g.spin1[0] = .3;
			g.spin1[1] = .1;
			g.spin2[1] = -.2;
		}
		put((std::string("calculate_snr_") + method).c_str(),
		    calculate_snr("aLIGO_analytic", "Hanford", std::string(method), &g, freq[0], L, "SIMPSONS", (double *)NULL, false));
		const int NG = 400;
		std::vector<double> fg(NG), wg(NG);
		gauleg(std::log10(20.), std::log10(900.), fg.data(), wg.data(), NG);
		for (int i = 0; i < NG; i++) fg[i] = std::pow(10., fg[i]);
		put((std::string("calculate_snr_gl_") + method).c_str(),
		    calculate_snr("Hanford_O1_fitted", "Virgo", std::string(method), &g, fg.data(), NG, "GAUSSLEG", wg.data(), true));
	}

	// ---- responses -------------------------------------------------------------------------------------------------------
	std::vector<std::complex<double>> resp(L);
	fourier_detector_response(freq[0], L, resp.data(), std::string("Livingston"), std::string("IMRPhenomD"), &gp, (double *)NULL);
	for (int i : {0, 100, 777}) {
		put("single_L_re", i, resp[i].real());
		put("single_L_im", i, resp[i].imag());
	}
	create_coherent_GW_detection(detectors, D, freq, lengths, true, &gp, std::string("IMRPhenomD"), data);
	for (int d = 0; d < D; d++)
		for (int i : {0, 100, 777}) {
			put("coherent_re_" + detectors[d], i, data[d][i].real());
			put("coherent_im_" + detectors[d], i, data[d][i].imag());
		}

	// ---- likelihoods -----------------------------------------------------------------------------------------------------
	// (data = the injection's responses scaled and rotated, so that the template below is mismatched)
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) data[d][i] *= std::complex<double>(0.9 * std::cos(0.3), 0.9 * std::sin(0.3));
	put("Log_Likelihood_internal", Log_Likelihood_internal(data[1], psd[1], freq[1], (double *)NULL, resp.data(), L, false, "SIMPSONS"));
	{
		gen_params t = gp;
		t.mass1 = 36.4 * 1.0005;
		// T = 1/(frequencies[1]-frequencies[0]) = 1/L with this layout (see above): tc_ref = 1/L - tc
		put("MCMC_likelihood_extrinsic_D", MCMC_likelihood_extrinsic(true, &t, "IMRPhenomD", lengths, freq, data, psd, weights, "SIMPSONS",
		                                                            false, detectors, D));
		put("tc_after_extrinsic", t.tc);  // the call replaces parameters->tc by T - tc (src/mcmc_gw.cpp:2473): the next one flips it back
		t.spin1[0] = .3;
		t.spin1[1] = .1;
		t.spin2[1] = -.2;
		put("MCMC_likelihood_extrinsic_Pv2", MCMC_likelihood_extrinsic(true, &t, "IMRPhenomPv2", lengths, freq, data, psd, weights,
		                                                              "SIMPSONS", false, detectors, D));
	}

	// ---- the samplers' callback: MCMC_prep_params + repack_parameters + MCMC_likelihood_extrinsic, and MCMC_likelihood_wrapper -----
	{
		const int dim = 11;
		double chirp = calculate_chirpmass(36.4, 29.3), eta = calculate_eta(36.4, 29.3);
		// RA, sin DEC, psi, cos iota, phiRef, tc, ln DL, ln Mc, eta, chi1, chi2
		double param[dim] = {.275, std::sin(-.44), .2, std::cos(.51), 2., 3., std::log(500.), std::log(chirp * 1.0003), eta, .3, .2};
		MCMC_modification_struct mod;
		double temp[dim];
		gen_params_base<double> g;
		std::string local = MCMC_prep_params(param, temp, &g, dim, "IMRPhenomD", &mod);
		g.gmst = 2.1;
		repack_parameters(temp, &g, "MCMC_" + std::string("IMRPhenomD"), dim, (gen_params_base<double> *)NULL);
		put("callback_chain", MCMC_likelihood_extrinsic(true, &g, local, lengths, freq, data, psd, weights, "SIMPSONS", false, detectors, D));
		if (gwat_b200_dropin_bind_mcmc) {
			gwat_b200_dropin_bind_mcmc(data, psd, freq, lengths, detectors, D, "IMRPhenomD", &mod, 2.1, 4);
			mcmc_data_interface iface;
			iface.min_dim = iface.max_dim = dim;
			iface.chain_id = 0;
			iface.chain_number = 1;
			iface.nested_model_number = 0;
			MCMC_user_param up;
			put("MCMC_likelihood_wrapper", MCMC_likelihood_wrapper(param, &iface, (void *)&up));
			double **F = allocate_2D_array(dim, dim);
			MCMC_fisher_wrapper(param, F, &iface, (void *)&up);
			for (int i = 0; i < dim; i++) put("MCMC_fisher_wrapper_diag", i, F[i][i]);
			deallocate_2D_array(F, dim, dim);
		}
		// what MCMC_fisher_wrapper computes, from the reference's pieces: sum_d fisher_numerical("MCMC_"+method) [+ transformations]
		double **F = allocate_2D_array(dim, dim), **Ft = allocate_2D_array(dim, dim);
		for (int i = 0; i < dim; i++)
			for (int j = 0; j < dim; j++) F[i][j] = 0;
		gen_params_base<double> g2;
		MCMC_prep_params(param, temp, &g2, dim, "IMRPhenomD", &mod);
		g2.gmst = 2.1;
		repack_parameters(temp, &g2, "MCMC_" + std::string("IMRPhenomD"), dim, (gen_params_base<double> *)NULL);
		for (int d = 0; d < D; d++) {
			fisher_numerical(freq[d], L, "MCMC_IMRPhenomD", detectors[d], detectors[0], Ft, dim, &g2, 4, NULL, NULL, psd[d]);
			for (int i = 0; i < dim; i++)
				for (int j = 0; j < dim; j++) F[i][j] += Ft[i][j];
		}
		{
			mcmc_data_interface iface2;
			iface2.min_dim = iface2.max_dim = dim;
			MCMC_fisher_transformations(temp, F, dim, "IMRPhenomD", false, &iface2, &mod, NULL);  // (the reference's, in both links)
		}
		for (int i = 0; i < dim; i++) put("fisher_sum_diag", i, F[i][i]);
		put("fisher_sum_offdiag_7_8", F[7][8]);
		deallocate_2D_array(F, dim, dim);
		deallocate_2D_array(Ft, dim, dim);
	}

	// ---- the Fisher matrices on a grid of their own: user_param->fisher_freq / fisher_PSD / fisher_length (src/mcmc_gw.cpp:2257-2275) --
	{
		const int dim = 11, Lf = 384;
		double chirp = calculate_chirpmass(36.4, 29.3), eta = calculate_eta(36.4, 29.3);
		double param[dim] = {.275, std::sin(-.44), .2, std::cos(.51), 2., 3., std::log(500.), std::log(chirp * 1.0003), eta, .3, .2};
		MCMC_modification_struct mod;
		double temp[dim];
		std::vector<double> ffb((size_t)D * Lf), fpb((size_t)D * Lf);
		double *ff[3], *fp[3];
		int fl[3] = {Lf, Lf, Lf};
		for (int d = 0; d < D; d++) {
			ff[d] = ffb.data() + (size_t)d * Lf;
			fp[d] = fpb.data() + (size_t)d * Lf;
			for (int i = 0; i < Lf; i++) ff[d][i] = 20. * std::pow(50., (double)i / (Lf - 1));  // 20 Hz ... 1000 Hz, geometric: not uniform
			populate_noise(ff[d], "aLIGO_analytic", fp[d], Lf, 48);
			for (int i = 0; i < Lf; i++) fp[d][i] *= fp[d][i] * (1. + .5 * d);
		}
		double **F = allocate_2D_array(dim, dim), **Ft = allocate_2D_array(dim, dim);
		for (int i = 0; i < dim; i++)
			for (int j = 0; j < dim; j++) F[i][j] = 0;
		gen_params_base<double> g2;
		MCMC_prep_params(param, temp, &g2, dim, "IMRPhenomD", &mod);
		g2.gmst = 2.1;
		repack_parameters(temp, &g2, "MCMC_" + std::string("IMRPhenomD"), dim, (gen_params_base<double> *)NULL);
		for (int d = 0; d < D; d++) {
			fisher_numerical(ff[d], Lf, "MCMC_IMRPhenomD", detectors[d], detectors[0], Ft, dim, &g2, 4, NULL, NULL, fp[d]);
			for (int i = 0; i < dim; i++)
				for (int j = 0; j < dim; j++) F[i][j] += Ft[i][j];
		}
		{
			mcmc_data_interface iface2;
			iface2.min_dim = iface2.max_dim = dim;
			MCMC_fisher_transformations(temp, F, dim, "IMRPhenomD", false, &iface2, &mod, NULL);
		}
		for (int i = 0; i < dim; i++) put("fisher_owngrid_sum_diag", i, F[i][i]);
		deallocate_2D_array(F, dim, dim);
		deallocate_2D_array(Ft, dim, dim);
		if (gwat_b200_dropin_bind_mcmc) {
			gwat_b200_dropin_bind_mcmc(data, psd, freq, lengths, detectors, D, "IMRPhenomD", &mod, 2.1, 4);
			mcmc_data_interface iface;
			iface.min_dim = iface.max_dim = dim;
			iface.chain_id = 0;
			iface.chain_number = 1;
			iface.nested_model_number = 0;
			MCMC_user_param up;
			up.fisher_freq = ff;
			up.fisher_PSD = fp;
			up.fisher_length = fl;
			double **Fw = allocate_2D_array(dim, dim);
			MCMC_fisher_wrapper(param, Fw, &iface, (void *)&up);
			for (int i = 0; i < dim; i++) put("MCMC_fisher_wrapper_owngrid_diag", i, Fw[i][i]);
			deallocate_2D_array(Fw, dim, dim);
			// ... and the likelihood afterwards is the data grid's again
			put("MCMC_likelihood_wrapper_after_owngrid", MCMC_likelihood_wrapper(param, &iface, (void *)&up));
		}
	}

	// ---- an INTRINSIC run (ln Mc, eta, chi1, chi2): the wrappers' maximised likelihood and sky-averaged Fisher ------------------
	{
		const int dim = 4;
		double chirp = calculate_chirpmass(36.4, 29.3), eta = calculate_eta(36.4, 29.3);
		double param[dim] = {std::log(chirp * 1.0003), eta, .3, .2};
		MCMC_modification_struct mod;
		double temp[dim];
		// the chain of reference calls MCMC_likelihood_wrapper stands for when mcmc_intrinsic is set (src/mcmc_gw.cpp:2576-2640, with
		// mcmc_save_waveform: one response for all detectors); the flag is a static of the reference's own translation unit, so its
		// one effect on the record -- sky_average (:2494) -- is applied here
		gen_params_base<double> g;
		std::string local = MCMC_prep_params(param, temp, &g, dim, "IMRPhenomD", &mod);
		g.sky_average = true;
		g.gmst = 2.1;
		repack_parameters(temp, &g, "MCMC_" + std::string("IMRPhenomD"), dim, (gen_params_base<double> *)NULL);
		gen_params_base<double> gl = g;
		gl.theta = 0;
		gl.phi = 0;
		gl.psi = 0;
		gl.phiRef = 1;
		gl.f_ref = 10;
		gl.incl_angle = 0;
		gl.tc = 1;
		{
			fftw_outline plan;
			allocate_FFTW_mem_forward(&plan, L);
			std::vector<std::complex<double>> response(L);
			fourier_detector_response_horizon(freq[0], L, response.data(), detectors[0], local, &gl);
			double ll = 0;
			for (int d = 0; d < D; d++) ll += maximized_Log_Likelihood_aligned_spin_internal(data[d], psd[d], freq[d], response.data(), (size_t)L, &plan);
			deallocate_FFTW_mem(&plan);
			put("intrinsic_callback_chain", ll);
		}
		double **F = allocate_2D_array(dim, dim), **Ft = allocate_2D_array(dim, dim);
		for (int i = 0; i < dim; i++)
			for (int j = 0; j < dim; j++) F[i][j] = 0;
		for (int d = 0; d < D; d++) {
			fisher_numerical(freq[d], L, "MCMC_IMRPhenomD", detectors[d], detectors[0], Ft, dim, &g, 4, NULL, NULL, psd[d]);
			for (int i = 0; i < dim; i++)
				for (int j = 0; j < dim; j++) F[i][j] += Ft[i][j];
		}
		{
			mcmc_data_interface iface2;
			iface2.min_dim = iface2.max_dim = dim;
			MCMC_fisher_transformations(temp, F, dim, "IMRPhenomD", true, &iface2, &mod, NULL);
		}
		for (int i = 0; i < dim; i++) put("intrinsic_fisher_sum_diag", i, F[i][i]);
		put("intrinsic_fisher_sum_offdiag_0_1", F[0][1]);
		deallocate_2D_array(F, dim, dim);
		deallocate_2D_array(Ft, dim, dim);
		if (gwat_b200_dropin_bind_mcmc && gwat_b200_dropin_set_intrinsic) {
			gwat_b200_dropin_bind_mcmc(data, psd, freq, lengths, detectors, D, "IMRPhenomD", &mod, 2.1, 4);
			gwat_b200_dropin_set_intrinsic(1);
			mcmc_data_interface iface;
			iface.min_dim = iface.max_dim = dim;
			iface.chain_id = 0;
			iface.chain_number = 1;
			iface.nested_model_number = 0;
			MCMC_user_param up;
			put("MCMC_likelihood_wrapper_intrinsic", MCMC_likelihood_wrapper(param, &iface, (void *)&up));
			double **Fw = allocate_2D_array(dim, dim);
			MCMC_fisher_wrapper(param, Fw, &iface, (void *)&up);
			for (int i = 0; i < dim; i++) put("MCMC_fisher_wrapper_intrinsic_diag", i, Fw[i][i]);
			put("MCMC_fisher_wrapper_intrinsic_offdiag_0_1", Fw[0][1]);
			deallocate_2D_array(Fw, dim, dim);
			gwat_b200_dropin_set_intrinsic(0);
		}
	}

	// ---- fisher_numerical, physical parameterisation, one detector --------------------------------------------------------
	{
		const int dim = 11;
		double **F = allocate_2D_array(dim, dim);
		fisher_numerical(freq[0], L, "IMRPhenomD", "Livingston", "Hanford", F, dim, &gp, 2, NULL, NULL, psd[1]);
		for (int i = 0; i < dim; i++) put("fisher_numerical_o2_diag", i, F[i][i]);
		put("fisher_numerical_o2_offdiag_6_7", F[6][7]);
		deallocate_2D_array(F, dim, dim);
	}
	return 0;
}
