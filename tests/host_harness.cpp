// TEST INFRASTRUCTURE ONLY.  Compiles the kernels' GWAT_HD mathematics (gw_analysis_tools_b200/csrc/*.h) as plain C++ so
// the CPU-only test tier (`pytest -m "not gpu"`) can check the per-walker setup and the per-bin code against the oracle
// without a GPU.  This file is NOT part of the product: the C ABI (libgwat_b200.so) has no CPU path and never links this.
#include <cstring>
#include <string>
#include <vector>

#define GWAT_TABLE_QUALIFIER static const
#include "../gw_analysis_tools_b200/csrc/gwat_tables.inc"
#include "../gw_analysis_tools_b200/csrc/gwat_bins.h"
#include "../gw_analysis_tools_b200/csrc/gwat_grid.h"
#include "../gw_analysis_tools_b200/csrc/gwat_method.h"

using namespace gwat;

namespace {

Tables host_tables() { return Tables{gwat_phenomd_fit, gwat_qnm_knots, GWAT_QNM_N}; }

struct Grid {
	std::vector<double> f, hi, lo, lg;
};
Grid make_grid(const double *f, int L)
{
	Grid g;
	g.f.assign(f, f + L);
	build_frequency_tables(f, L, g.hi, g.lo, g.lg);
	return g;
}

template <class Fam>
void waveform_t(const gwat_b200_source *src, const Network &net, const Grid &g, double *hp_re, double *hp_im,
                double *hc_re, double *hc_im)
{
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), w);
	for (size_t i = 0; i < g.f.size(); i++) {
		cplx hp, hc;
		polarizations_bin<Fam>(w, g.f[i], g.hi[i], g.lo[i], g.lg[i], hp, hc);
		hp_re[i] = hp.re;
		hp_im[i] = hp.im;
		hc_re[i] = hc.re;
		hc_im[i] = hc.im;
	}
}

template <class Fam>
void response_t(const gwat_b200_source *src, const Network &net, const Grid &g, bool shift, double *re, double *im)
{
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), w);
	const size_t L = g.f.size();
	for (size_t i = 0; i < L; i++) {
		cplx hp, hc;
		polarizations_bin<Fam>(w, g.f[i], g.hi[i], g.lo[i], g.lg[i], hp, hc);
		for (int d = 0; d < net.D; d++) {
			const cplx r = project_bin(w.det[d], hp, hc, g.f[i], shift);
			re[d * L + i] = r.re;
			im[d * L + i] = r.im;
		}
	}
}

int make_network(int D, const char *const *dets, Network &net)
{
	net.D = D;
	for (int d = 0; d < D; d++) {
		const int id = detector_index(dets[d]);
		if (id < 0) return -1;
		std::memcpy(net.row[d], gwat_detector_table[id], sizeof(double) * 13);
	}
	return 0;
}

}  // namespace

#define DISPATCH_FAMILY(desc, CALL)                                                                     \
	switch (desc.family_id) {                                                                             \
	case FAM_D: { typedef Family<BASE_D, PPE_NONE, false, false> Fam; CALL; break; }                      \
	case FAM_D_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, false> Fam; CALL; break; }          \
	case FAM_D_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, false> Fam; CALL; break; }               \
	case FAM_D_GIMR: { typedef Family<BASE_D, PPE_NONE, true, false> Fam; CALL; break; }                  \
	case FAM_D_NRT: { typedef Family<BASE_D, PPE_NONE, false, true> Fam; CALL; break; }                   \
	case FAM_D_NRT_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, true> Fam; CALL; break; }       \
	case FAM_D_NRT_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, true> Fam; CALL; break; }            \
	case FAM_P: { typedef Family<BASE_P, PPE_NONE, false, false> Fam; CALL; break; }                      \
	case FAM_P_PPE_INS: { typedef Family<BASE_P, PPE_INSPIRAL, false, false> Fam; CALL; break; }          \
	case FAM_P_PPE_IMR: { typedef Family<BASE_P, PPE_IMR, false, false> Fam; CALL; break; }               \
	case FAM_P_GIMR: { typedef Family<BASE_P, PPE_NONE, true, false> Fam; CALL; break; }                  \
	default: return -2;                                                                                   \
	}

extern "C" {

int hh_fourier_waveform(const char *method, const gwat_b200_source *src, const double *f, int L, double *hp_re,
                        double *hp_im, double *hc_re, double *hc_im)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	Network net;
	net.D = 0;
	const Grid g = make_grid(f, L);
	DISPATCH_FAMILY(desc, waveform_t<Fam>(src, net, g, hp_re, hp_im, hc_re, hc_im));
	return 0;
}

int hh_coherent_response(const char *method, const gwat_b200_source *src, int D, const char *const *dets, const double *f,
                         int L, int with_shift, double *re, double *im)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	Network net;
	if (make_network(D, dets, net) != 0) return -1;
	const Grid g = make_grid(f, L);
	DISPATCH_FAMILY(desc, response_t<Fam>(src, net, g, with_shift != 0, re, im));
	return 0;
}

// setup intermediates for unit tests: fRD, fdamp, f1, f3, f1p, f2p, A0, tc, phic, beta0, beta1, alpha0, alpha1
int hh_phenomd_setup_probe(const gwat_b200_source *src, double *out)
{
	typedef Family<BASE_D, PPE_NONE, false, false> Fam;
	Network net;
	net.D = 0;
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), w);
	const DCoef &c = w.d;
	double v[] = {c.fRD, c.fdamp, c.f1a, c.f3a, c.f1p, c.f2p, c.A0, c.tc, c.phic, c.beta0, c.beta1, c.alpha0, c.alpha1};
	std::memcpy(out, v, sizeof(v));
	return 0;
}

}  // extern "C"
extern "C" int hh_debug_dcoef(const gwat_b200_source *src, double *out)
{
	typedef Family<BASE_D, PPE_NONE, false, false> Fam;
	Network net;
	net.D = 0;
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), w);
	std::memcpy(out, &w.d, sizeof(DCoef));
	return (int)(sizeof(DCoef) / sizeof(double));
}
