// TEST ONLY.  Calls include/gwat_b200_cxx.hpp the way GWAT code would -- reference-style argument lists (std::string,
// double**, std::complex<double>**, a gen_params_base<double>-like struct) -- and exposes the results through a few
// extern "C" functions so that tests/test_cxx_adapter.py can compare them with the oracle.  The flat gwat_b200_source
// coming from Python is only the transport for the test inputs: it is unpacked into the C++ struct first.
#include <complex>
#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "dropin_types.hpp"
#include "gwat_b200_cxx.hpp"

namespace {
struct Unpacked {
	test_gen_params g;
	std::vector<double> betappe, bppe, dphi, dsigma, dbeta, dalpha;
	std::vector<int> phii, sigmai, betai, alphai;
};
void unpack(const gwat_b200_source &s, Unpacked &u)
{
	test_gen_params &g = u.g;
	g.mass1 = s.mass1; g.mass2 = s.mass2; g.Luminosity_Distance = s.Luminosity_Distance;
	for (int i = 0; i < 3; i++) { g.spin1[i] = s.spin1[i]; g.spin2[i] = s.spin2[i]; }
	g.tc = s.tc; g.phiRef = s.phiRef; g.f_ref = s.f_ref; g.psi = s.psi; g.incl_angle = s.incl_angle;
	g.RA = s.RA; g.DEC = s.DEC; g.gmst = s.gmst;
	g.tidal1 = s.tidal1; g.tidal2 = s.tidal2; g.tidal_s = s.tidal_s; g.tidal_a = s.tidal_a;
	g.tidal_weighted = s.tidal_weighted; g.delta_tidal_weighted = s.delta_tidal_weighted;
	g.diss_tidal1 = s.diss_tidal1; g.diss_tidal2 = s.diss_tidal2; g.diss_tidal_weighted = s.diss_tidal_weighted;
	g.chip = s.chip; g.phip = s.phip; g.PNorder = s.PNorder;
	g.shift_time = s.shift_time; g.shift_phase = s.shift_phase; g.sky_average = s.sky_average;
	g.tidal_love = s.tidal_love; g.tidal_love_error = s.tidal_love_error; g.NSflag1 = s.NSflag1; g.NSflag2 = s.NSflag2;
	g.dep_postmerger = s.dep_postmerger; g.equatorial_orientation = s.equatorial_orientation; g.horizon_coord = s.horizon_coord;
	u.betappe.assign(s.betappe, s.betappe + GWAT_B200_MAX_MOD);
	u.bppe.assign(s.bppe, s.bppe + GWAT_B200_MAX_MOD);
	u.dphi.assign(s.delta_phi, s.delta_phi + GWAT_B200_MAX_MOD);
	u.dsigma.assign(s.delta_sigma, s.delta_sigma + GWAT_B200_MAX_MOD);
	u.dbeta.assign(s.delta_beta, s.delta_beta + GWAT_B200_MAX_MOD);
	u.dalpha.assign(s.delta_alpha, s.delta_alpha + GWAT_B200_MAX_MOD);
	u.phii.assign(s.phii, s.phii + GWAT_B200_MAX_MOD);
	u.sigmai.assign(s.sigmai, s.sigmai + GWAT_B200_MAX_MOD);
	u.betai.assign(s.betai, s.betai + GWAT_B200_MAX_MOD);
	u.alphai.assign(s.alphai, s.alphai + GWAT_B200_MAX_MOD);
	g.Nmod = s.Nmod; g.betappe = u.betappe.data(); g.bppe = u.bppe.data();
	g.Nmod_phi = s.Nmod_phi; g.phii = u.phii.data(); g.delta_phi = u.dphi.data();
	g.Nmod_sigma = s.Nmod_sigma; g.sigmai = u.sigmai.data(); g.delta_sigma = u.dsigma.data();
	g.Nmod_beta = s.Nmod_beta; g.betai = u.betai.data(); g.delta_beta = u.dbeta.data();
	g.Nmod_alpha = s.Nmod_alpha; g.alphai = u.alphai.data(); g.delta_alpha = u.dalpha.data();
}
std::vector<std::string> split_names(const char *csv)
{
	std::vector<std::string> out;
	std::string cur;
	for (const char *p = csv; *p; p++) {
		if (*p == ',') { out.push_back(cur); cur.clear(); }
		else cur.push_back(*p);
	}
	out.push_back(cur);
	return out;
}
}  // namespace

extern "C" {

int cxa_fourier_waveform(const gwat_b200_source *src, const char *method, double *f, int L, double *out /* [4][L] */)
{
	gwat_b200::Engine e(0);
	Unpacked u;
	unpack(*src, u);
	std::vector<std::complex<double>> hp(L), hc(L);
	int st = gwat_b200::fourier_waveform(e, f, L, hp.data(), hc.data(), std::string(method), &u.g);
	for (int i = 0; i < L; i++) {
		out[i] = hp[i].real(); out[L + i] = hp[i].imag();
		out[2 * L + i] = hc[i].real(); out[3 * L + i] = hc[i].imag();
	}
	return st;
}

int cxa_fourier_detector_response(const gwat_b200_source *src, const char *method, const char *detector, double *f, int L,
                                  double *out /* [2][L] */)
{
	gwat_b200::Engine e(0);
	Unpacked u;
	unpack(*src, u);
	std::vector<std::complex<double>> r(L);
	int st = gwat_b200::fourier_detector_response(e, f, L, r.data(), std::string(detector), std::string(method), &u.g);
	for (int i = 0; i < L; i++) { out[i] = r[i].real(); out[L + i] = r[i].imag(); }
	return st;
}

int cxa_coherent_response(const gwat_b200_source *src, const char *method, const char *detectors_csv, double *f, int L,
                          double *out_re /* [D][L] */, double *out_im)
{
	std::vector<std::string> dets = split_names(detectors_csv);
	const int D = (int)dets.size();
	gwat_b200::Engine e(0);
	Unpacked u;
	unpack(*src, u);
	std::vector<std::vector<std::complex<double>>> resp(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> rptr(D);
	std::vector<double *> fptr(D);
	std::vector<int> lens(D, L);
	for (int d = 0; d < D; d++) { rptr[d] = resp[d].data(); fptr[d] = f; }
	gwat_b200::create_coherent_GW_detection(e, dets.data(), D, fptr.data(), lens.data(), true, &u.g, std::string(method), rptr.data());
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) { out_re[(size_t)d * L + i] = resp[d][i].real(); out_im[(size_t)d * L + i] = resp[d][i].imag(); }
	return e.ok() ? 1 : 0;
}

double cxa_loglike(const gwat_b200_source *trial, const char *method, const char *detectors_csv, double *f, double *psd /* [D][L] */,
                   const double *data_re, const double *data_im, int L, double T_segment)
{
	std::vector<std::string> dets = split_names(detectors_csv);
	const int D = (int)dets.size();
	gwat_b200::Engine e(0);
	Unpacked ut;
	unpack(*trial, ut);
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> dptr(D);
	std::vector<double *> fptr(D), pptr(D);
	std::vector<int> lens(D, L);
	for (int d = 0; d < D; d++) {
		for (int i = 0; i < L; i++) data[d][i] = std::complex<double>(data_re[(size_t)d * L + i], data_im[(size_t)d * L + i]);
		dptr[d] = data[d].data(); fptr[d] = f; pptr[d] = psd + (size_t)d * L;
	}
	return gwat_b200::MCMC_likelihood_extrinsic(e, false, &ut.g, std::string(method), lens.data(), fptr.data(), dptr.data(), pptr.data(),
	                                            (double **)nullptr, std::string("SIMPSONS"), false, dets.data(), D, T_segment);
}

int cxa_fisher_numerical(const gwat_b200_source *src, const char *method, const char *detector, double *f, double *psd, int L,
                         int dimension, int order, double *out /* [dim][dim] */)
{
	gwat_b200::Engine e(0);
	Unpacked u;
	unpack(*src, u);
	std::vector<double *> rows(dimension);
	for (int i = 0; i < dimension; i++) rows[i] = out + (size_t)i * dimension;
	for (int i = 0; i < dimension * dimension; i++) out[i] = 0;
	gwat_b200::fisher_numerical(e, f, L, std::string(method), std::string(detector), std::string(detector), rows.data(), dimension,
	                            &u.g, order, (int *)nullptr, (int *)nullptr, psd);
	return e.ok() ? 1 : 0;
}

// The reference's sampler shape: `threads` pool workers pull chains off a shared counter and call the likelihood callback
// one chain at a time (src/mcmc_sampler.cpp:347-447).  The callback is a gwat_b200::CallbackQueue bound the way
// INTEGRATION.md shows.  stats[0..2] = calls, batched launches, largest group.
int cxa_pool_loglike(const char *method, const gwat_b200_mod *mod, int dimension, int W, const double *params, double gmst,
                     double T_segment, const char *detectors_csv, double *f, double *psd, const double *data_re, const double *data_im,
                     int L, int threads, double *logL, long long *stats)
{
	std::vector<std::string> dets = split_names(detectors_csv);
	const int D = (int)dets.size();
	gwat_b200::Engine e(0);
	if (!e.ok()) return 0;
	std::vector<std::vector<std::complex<double>>> data(D, std::vector<std::complex<double>>(L));
	std::vector<std::complex<double> *> dptr(D);
	std::vector<double *> fptr(D), pptr(D);
	for (int d = 0; d < D; d++) {
		for (int i = 0; i < L; i++) data[d][i] = std::complex<double>(data_re[(size_t)d * L + i], data_im[(size_t)d * L + i]);
		dptr[d] = data[d].data(); fptr[d] = f; pptr[d] = psd + (size_t)d * L;
	}
	if (e.set_network(dets.data(), D, L, fptr.data(), pptr.data(), dptr.data(), nullptr, "SIMPSONS", false) != 0) return 0;
	gwat_b200::CallbackQueue q(e, std::string(method), mod, dimension, gmst, T_segment, threads);
	if (!q.ok()) return 0;
	std::function<double(double *, int *, int, void *, void *)> ll = [&q](double *p, int *, int, void *, void *) { return q(p); };
	std::atomic<int> next(0);
	auto worker = [&]() {
		for (int w = next++; w < W; w = next++) logL[w] = ll(const_cast<double *>(params) + (size_t)w * dimension, nullptr, 0, nullptr, nullptr);
	};
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(worker);
	for (auto &t : pool) t.join();
	int largest = 0;
	gwat_b200_queue_stats(q.handle(), &stats[0], &stats[1], &largest);
	stats[2] = largest;
	return 1;
}
}
