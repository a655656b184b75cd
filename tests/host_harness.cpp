// TEST INFRASTRUCTURE ONLY.  Compiles the kernels' GWAT_HD mathematics (gw_analysis_tools_b200/csrc/*.h) as plain C++ so
// the CPU-only test tier (`pytest -m "not gpu"`) can check the per-walker setup and the per-bin code against the oracle
// without a GPU.  This file is NOT part of the product: the C ABI (libgwat_b200.so) has no CPU path and never links this.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#define GWAT_TABLE_QUALIFIER static const
#include "../gw_analysis_tools_b200/csrc/gwat_tables.inc"
#include "../gw_analysis_tools_b200/csrc/gwat_like.h"
#include "../gw_analysis_tools_b200/csrc/gwat_grid.h"
#include "../gw_analysis_tools_b200/csrc/gwat_method.h"
#include "../gw_analysis_tools_b200/csrc/gwat_repack.h"
#include "../gw_analysis_tools_b200/csrc/gwat_setup_coop.h"
#include "../gw_analysis_tools_b200/csrc/gwat_orient.h"

using namespace gwat;

namespace {

Tables host_tables() { return Tables{gwat_phenomd_fit, gwat_qnm_knots, GWAT_QNM_N, DzTable{gwat_dz_boundaries, gwat_dz_coeffs, GWAT_DZ_SEGMENTS, GWAT_NUM_COSMOLOGIES, gwat_md_alphas, gwat_md_boundaries_z, gwat_md_coeffs, GWAT_MD_ALPHAS}}; }

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
void waveform_t(int theory, const gwat_b200_source *src, const Network &net, const Grid &g, double *hp_re, double *hp_im,
                double *hc_re, double *hc_im)
{
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), theory, w);
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
void response_t(int theory, const gwat_b200_source *src, const Network &net, const Grid &g, bool shift, double *re, double *im)
{
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), theory, w);
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
	net.horizon_mode = 0;
	for (int d = 0; d < D; d++) {
		const int id = detector_index(dets[d]);
		if (id < 0) return -1;
		std::memcpy(net.row[d], gwat_detector_table[id], sizeof(double) * 13);
	}
	return 0;
}

}  // namespace

#define DISPATCH_FAMILY(desc, ...)                                                                      \
	switch (desc.family_id) {                                                                             \
	case FAM_D: { typedef Family<BASE_D, PPE_NONE, false, false> Fam; __VA_ARGS__; break; }                      \
	case FAM_D_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, false> Fam; __VA_ARGS__; break; }          \
	case FAM_D_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, false> Fam; __VA_ARGS__; break; }               \
	case FAM_D_GIMR: { typedef Family<BASE_D, PPE_NONE, true, false> Fam; __VA_ARGS__; break; }                  \
	case FAM_D_NRT: { typedef Family<BASE_D, PPE_NONE, false, true> Fam; __VA_ARGS__; break; }                   \
	case FAM_D_NRT_PPE_INS: { typedef Family<BASE_D, PPE_INSPIRAL, false, true> Fam; __VA_ARGS__; break; }       \
	case FAM_D_NRT_PPE_IMR: { typedef Family<BASE_D, PPE_IMR, false, true> Fam; __VA_ARGS__; break; }            \
	case FAM_P: { typedef Family<BASE_P, PPE_NONE, false, false> Fam; __VA_ARGS__; break; }                      \
	case FAM_P_PPE_INS: { typedef Family<BASE_P, PPE_INSPIRAL, false, false> Fam; __VA_ARGS__; break; }          \
	case FAM_P_PPE_IMR: { typedef Family<BASE_P, PPE_IMR, false, false> Fam; __VA_ARGS__; break; }               \
	case FAM_P_GIMR: { typedef Family<BASE_P, PPE_NONE, true, false> Fam; __VA_ARGS__; break; }                  \
	default: return -2;                                                                                   \
	}

extern "C" {

int hh_fourier_waveform(const char *method, const gwat_b200_source *src, const double *f, int L, double *hp_re,
                        double *hp_im, double *hc_re, double *hc_im)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	Network net{};
	net.D = 0;
	const Grid g = make_grid(f, L);
	DISPATCH_FAMILY(desc, waveform_t<Fam>(desc.theory, src, net, g, hp_re, hp_im, hc_re, hc_im));
	return 0;
}

int hh_coherent_response(const char *method, const gwat_b200_source *src, int D, const char *const *dets, const double *f,
                         int L, int with_shift, double *re, double *im)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	Network net{};
	if (make_network(D, dets, net) != 0) return -1;
	const Grid g = make_grid(f, L);
	DISPATCH_FAMILY(desc, response_t<Fam>(desc.theory, src, net, g, with_shift != 0, re, im));
	return 0;
}

// setup intermediates for unit tests: fRD, fdamp, f1, f3, f1p, f2p, A0, tc, phic, beta0, beta1, alpha0, alpha1
int hh_phenomd_setup_probe(const gwat_b200_source *src, double *out)
{
	typedef Family<BASE_D, PPE_NONE, false, false> Fam;
	const int theory = 0;
	Network net{};
	net.D = 0;
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), theory, w);
	const DCoef &c = w.d;
	double v[] = {c.fRD, c.fdamp, c.f1a, c.f3a, c.f1p, c.f2p, c.A0, c.tc, c.phic, c.beta0, c.beta1, c.alpha0, c.alpha1};
	std::memcpy(out, v, sizeof(v));
	return 0;
}

}  // extern "C"
extern "C" int hh_debug_dcoef(const gwat_b200_source *src, double *out)
{
	typedef Family<BASE_D, PPE_NONE, false, false> Fam;
	const int theory = 0;
	Network net{};
	net.D = 0;
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), theory, w);
	std::memcpy(out, &w.d, sizeof(DCoef));
	return (int)(sizeof(DCoef) / sizeof(double));
}

// The cooperative setup kernel's dataflow on the host (gwat_setup_coop.h): the steps run role by role on SEPARATE carry
// structures, everything a role does not write itself is poisoned with NaN bit patterns beforehand, and the record is merged
// exactly as k_setup merges it.  out_coop / out_seq receive the WalkerCoef words of this flow and of walker_setup: the test
// wants them identical bit for bit.
template <class Fam>
int setup_coop_t(int theory, const gwat_b200_source *src, const Network &net, double *out_coop, double *out_seq)
{
	const Tables t = host_tables();
	constexpr int kRoles = setup_roles<Fam>();
	SetupRec r;
	std::memset(&r, 0xff, sizeof(r));
	SetupCarry k[4];
	std::memset(k, 0xff, sizeof(k));
	// the kernel's roles run concurrently between barriers; any order must do -- run them backwards
	for (int role = kRoles - 1; role >= 0; role--) setup_step1<Fam>(role, *src, net, t, theory, k[role], r);
	for (int role = kRoles - 1; role >= 0; role--) setup_step2<Fam>(role, t, k[role], r);
	for (int role = kRoles - 1; role >= 0; role--) setup_step3<Fam>(role, net, r);
	for (int role = kRoles - 1; role >= 0; role--) setup_step4<Fam>(role, src->shift_time != 0, r);
	if (r.refused) r.w.d.A0 = NAN;
	r.w.valid = 1;
	std::memcpy(out_coop, &r.w, sizeof(WalkerCoef));
	WalkerCoef w;
	std::memset(&w, 0xff, sizeof(w));
	walker_setup<Fam>(*src, net, t, theory, w);
	w.pad_ = 0;
	std::memcpy(out_seq, &w, sizeof(WalkerCoef));
	return (int)(sizeof(WalkerCoef) / sizeof(double));
}
extern "C" int hh_setup_coop(const char *method, const gwat_b200_source *src, int D, const char *const *dets, double *out_coop, double *out_seq)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -1;
	Network net{};
	if (make_network(D, dets, net) != 0) return -3;
	int n = 0;
	DISPATCH_FAMILY(desc, n = setup_coop_t<Fam>(desc.theory, src, net, out_coop, out_seq));
	return n;
}
// transform_orientation_coords of gwat_orient.h, in place
extern "C" int hh_transform_orientation(const char *method, gwat_b200_source *src)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -1;
	transform_orientation_coords(*src, desc.pv2 != 0);
	return 0;
}

// sampling vector of an intrinsic set -> physical record (repack_mcmc_walker with plan.sky), as k_repack_only runs it
extern "C" int hh_repack_mcmc_intrinsic(const char *method, const gwat_b200_mod *mod, int dimension, const double *param, double gmst,
                                        gwat_b200_source *out)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0 || desc.mcmc) return -1;
	RepackPlan plan;
	std::memset(&plan, 0, sizeof(plan));
	plan.dimension = dimension;
	plan.pv2 = desc.pv2;
	plan.nrt = desc.nrt;
	plan.ppe = desc.ppe || desc.theory != THEORY_NONE;
	plan.gimr = desc.gimr && !plan.ppe;
	plan.alpha_unit_fix = theory_alpha_units(desc.theory);
	plan.mcmc = 1;
	plan.sky = 1;
	plan.mod = *mod;
	repack_mcmc_walker(param, plan, gmst, 0.0, *out);
	return 0;
}

// which words of WalkerCoef each family defines (the others stay poisoned in both flows): offsets for the test
extern "C" void hh_walkercoef_layout(int *out)
{
	out[0] = (int)(offsetof(WalkerCoef, d) / 8);
	out[1] = (int)(offsetof(WalkerCoef, p) / 8);
	out[2] = (int)(offsetof(WalkerCoef, pfac) / 8);
	out[3] = (int)(offsetof(WalkerCoef, det) / 8);
	out[4] = (int)(offsetof(WalkerCoef, valid) / 8);
	out[5] = (int)(sizeof(DetCoef) / 8);
}

// Fisher matrix of one source on the host, through the same GWAT_HD pieces the Fisher kernels use (unpack/repack of the
// stencil points, walker setup, per-bin response); the stencil combination and the Simpson assembly are restated here
// (they live in __global__ kernels in the product).  psd is for `detector`.
template <class Fam>
int fisher_t(int theory, bool mcmc, const MethodDesc &desc, const gwat_b200_source *src, const char *detector,
             const char *reference, int dim, int order, const Grid &g, const double *psd, double *out)
{
	RepackPlan plan;
	std::memset(&plan, 0, sizeof(plan));
	plan.dimension = dim;
	plan.pv2 = desc.pv2;
	plan.nrt = desc.nrt;
	plan.ppe = desc.ppe || desc.theory != THEORY_NONE;
	plan.gimr = desc.gimr && !plan.ppe;
	plan.mcmc = mcmc;
	const int idd = detector_index(detector), idr = detector_index(reference);
	if (idd < 0 || idr < 0) return -1;
	const double *det_row = gwat_detector_table[idd], *ref_row = gwat_detector_table[idr];
	const bool same = idd == idr;
	const int npts = order == 4 ? 4 : 2;
	const size_t L = g.f.size();
	const double eps = 1e-8;
	if (src->sky_average && Fam::base == BASE_P) {
		// the reference's amplitude / phase branch is for the IMRPhenomD family (src/fisher.cpp:183); a sky-averaged IMRPhenomPv2 record takes the
		// response branch below with the 8 intrinsic parameters (the "MCMC_" set; the physical one prints "not supported", :1993)
		if (!mcmc) return -5;
		plan.sky = 1;
	} else if (src->sky_average) {
		// sky-averaged branch (src/fisher.cpp:183-338): amplitude / phase derivatives in the 7-parameter set
		if (Fam::nrt) return -5;
		plan.sky = 1;
		double v0s[GWAT_B200_MAX_DIM];
		int logf_[GWAT_B200_MAX_DIM];
		unpack_fisher(*src, plan, v0s, logf_);
		Network net{};
		net.D = 1;
		std::memcpy(net.row[0], det_row, sizeof(double) * 13);
		WalkerCoef w0;
		walker_setup<Fam>(*src, net, host_tables(), theory, w0, true);
		std::vector<std::vector<cplx>> deriv(dim, std::vector<cplx>(L));
		for (int i = 0; i < dim; i++) {
			WalkerCoef wk[4];
			for (int k = 0; k < npts; k++) {
				double v[GWAT_B200_MAX_DIM];
				for (int j = 0; j < dim; j++) v[j] = v0s[j];
				const double step = (k == 0) ? eps : (k == 1) ? -eps : (k == 2) ? 2 * eps : -2 * eps;
				if (!(step > 0 && i == 8 && v0s[8] > .25 - eps)) v[i] = v0s[i] + step;  // (the one-sided rule of parameter 8, :218-232)
				gwat_b200_source sp;
				repack_fisher_point(v, *src, plan, sp);
				walker_setup<Fam>(sp, net, host_tables(), theory, wk[k], true);
			}
			const double sc = (logf_[i] ? v0s[i] : 1.0) * ((i == 8 && v0s[8] > .25 - eps) ? 2.0 : 1.0);
			for (size_t b = 0; b < L; b++) {
				double a[4], ph[4], a0, p0;
				amplitude_phase_bin<Fam>(w0, g.f[b], g.hi[b], g.lo[b], g.lg[b], a0, p0);
				for (int k = 0; k < npts; k++) amplitude_phase_bin<Fam>(wk[k], g.f[b], g.hi[b], g.lo[b], g.lg[b], a[k], ph[k]);
				const cplx d = sky_derivative_bin(npts, a, ph, a0);
				deriv[i][b] = cplx{d.re * sc, d.im * sc};
			}
		}
		const double pref = quadrature_prefactor((int)L, false, g.f.data(), true);
		for (int j = 0; j < dim; j++)
			for (int k = 0; k <= j; k++) {
				double acc = 0;
				for (size_t b = 0; b < L; b++)
					acc += quadrature_coefficient((int)b, (int)L, false, false, nullptr, g.f.data()) *
					       ((deriv[j][b].re * deriv[k][b].re + deriv[j][b].im * deriv[k][b].im) / psd[b]);
				out[j * dim + k] = out[k * dim + j] = pref * acc;
			}
		return 0;
	}
	double v0[GWAT_B200_MAX_DIM];
	int logfac[GWAT_B200_MAX_DIM];
	unpack_fisher(*src, plan, v0, logfac);
	std::vector<std::vector<cplx>> deriv(dim, std::vector<cplx>(L));
	for (int i = 0; i < dim; i++) {
		const bool bc = (i == 8 && v0[8] > .25 - eps);
		std::vector<std::vector<cplx>> r(npts, std::vector<cplx>(L));
		for (int k = 0; k < npts; k++) {
			double v[GWAT_B200_MAX_DIM];
			for (int j = 0; j < dim; j++) v[j] = v0[j];
			const double step = (k == 0) ? eps : (k == 1) ? -eps : (k == 2) ? 2 * eps : -2 * eps;
			if (!(step > 0 && bc)) v[i] = v0[i] + step;
			gwat_b200_source sp;
			repack_fisher_point(v, *src, plan, sp);
			if (sp.equatorial_orientation) transform_orientation_coords(sp, Fam::base == BASE_P);
			double tshift = 0;
			if (!same) {
				const double dtoa = dtoa_between(ref_row + 9, det_row + 9, sp.RA, sp.DEC, sp.gmst);
				if (k < 2) tshift = (-2 * GWAT_PI) * dtoa;
				else sp.tc -= dtoa;
			}
			Network net{};
			net.D = 1;
			std::memcpy(net.row[0], det_row, sizeof(double) * 13);
			WalkerCoef w;
			walker_setup<Fam>(sp, net, host_tables(), theory, w);
			w.det[0].tshift = tshift;
			for (size_t b = 0; b < L; b++) {
				cplx hp, hc;
				polarizations_bin<Fam>(w, g.f[b], g.hi[b], g.lo[b], g.lg[b], hp, hc);
				r[k][b] = project_bin(w.det[0], hp, hc, g.f[b], true);
			}
		}
		const double sc = logfac[i] ? v0[i] : 1.0;
		for (size_t b = 0; b < L; b++) {
			cplx d;
			if (npts == 2) {
				const double den = bc ? eps : 2. * eps;
				d = cplx{(r[0][b].re - r[1][b].re) / den, (r[0][b].im - r[1][b].im) / den};
			} else {
				const double den = bc ? 6. * eps : 12. * eps;
				d = cplx{(((-r[2][b].re + 8. * r[0][b].re) - 8. * r[1][b].re) + r[3][b].re) / den,
				         (((-r[2][b].im + 8. * r[0][b].im) - 8. * r[1][b].im) + r[3][b].im) / den};
			}
			deriv[i][b] = cplx{d.re * sc, d.im * sc};
		}
	}
	const double pref = quadrature_prefactor((int)L, false, g.f.data(), true);
	for (int j = 0; j < dim; j++)
		for (int k = 0; k <= j; k++) {
			double acc = 0;
			for (size_t b = 0; b < L; b++)
				acc += quadrature_coefficient((int)b, (int)L, false, false, nullptr, g.f.data()) *
				       ((deriv[j][b].re * deriv[k][b].re + deriv[j][b].im * deriv[k][b].im) / psd[b]);
			out[j * dim + k] = out[k * dim + j] = pref * acc;
		}
	return 0;
}

extern "C" int hh_fisher_numerical(const char *method, const char *detector, const char *reference, int dim, int order,
                                   const gwat_b200_source *src, const double *f, int L, const double *psd, double *out)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	const Grid g = make_grid(f, L);
	int rc = 0;
	DISPATCH_FAMILY(desc, rc = fisher_t<Fam>(desc.theory, desc.mcmc, desc, src, detector, reference, dim, order, g, psd, out));
	return rc;
}


// The fused likelihood inner loop (gwat_like.h) with the kernel's exact work partition: chunks of `bins_per_cta` bins,
// 256 "threads" striding through each chunk, partial sums combined afterwards.
template <class Fam, int D>
double loglike_t(int theory, const gwat_b200_source *src, const Network &net, const Grid &g, const std::vector<double> &wq,
                 const double *dre, const double *dim, bool uniform, double df, double prefactor, int bins_per_cta)
{
	WalkerCoef w;
	walker_setup<Fam>(*src, net, host_tables(), theory, w);
	LikeGrid lg;
	lg.f = g.f.data();
	lg.sf_hi = g.hi.data();
	lg.sf_lo = g.lo.data();
	lg.logf = g.lg.data();
	lg.wq = wq.data();
	lg.dre = dre;
	lg.dim = dim;
	lg.L = (int)g.f.size();
	lg.ld = lg.L;
	lg.uniform = uniform ? 1 : 0;
	lg.df = df;
	// the kernel's cut: units of `bins_per_cta` bins (<= 0: the library's own unit size for this grid), 256 "threads" per
	// unit, per-thread seeds (the device's per-CTA seed tables are the same numbers to rounding)
	if (bins_per_cta <= 0) bins_per_cta = unit_bins_for(lg.L);
	const double fmax = walker_fmax<Fam>(w);
	double total = 0, nact = 0;
	for (int begin = 0; begin < lg.L; begin += bins_per_cta) {
		const int end = std::min(lg.L, begin + bins_per_cta);
		for (int tid = 0; tid < 256; tid++) {
			double acc = 0;
			loglike_unit<Fam, D>(w, lg, begin + tid, end, 256, fmax, acc, nact);
			total += acc;
		}
	}
	return -0.5 * (prefactor * total);
}

extern "C" int hh_loglike(const char *method, int W, const gwat_b200_source *src, int D, const char *const *dets, const double *f,
                          int L, const double *psd, const double *dre, const double *dim, const double *weights, int gaussleg,
                          int log10F, int bins_per_cta, double *out)
{
	MethodDesc desc;
	if (parse_method(method, desc) != 0) return -2;
	Network net{};
	if (make_network(D, dets, net) != 0) return -1;
	if (D < 1 || D > 5) return -3;
	const Grid g = make_grid(f, L);
	std::vector<double> wq((size_t)D * L);
	for (int d = 0; d < D; d++)
		for (int i = 0; i < L; i++) wq[(size_t)d * L + i] = quadrature_coefficient(i, L, gaussleg != 0, log10F != 0, weights, f) / psd[(size_t)d * L + i];
	const double pref = quadrature_prefactor(L, gaussleg != 0, f, false);
	const double df = (f[L - 1] - f[0]) / (L - 1);
	bool uni = df > 0;
	for (int i = 0; i < L && uni; i++) uni = std::fabs(f[i] - (f[0] + i * df)) <= 1e-9 * df;
	for (int wi = 0; wi < W; wi++) {
		switch (D) {  // (the kernels are instantiated per detector count, and so is this)
		case 1: DISPATCH_FAMILY(desc, out[wi] = loglike_t<Fam, 1>(desc.theory, src + wi, net, g, wq, dre, dim, uni, df, pref, bins_per_cta)); break;
		case 2: DISPATCH_FAMILY(desc, out[wi] = loglike_t<Fam, 2>(desc.theory, src + wi, net, g, wq, dre, dim, uni, df, pref, bins_per_cta)); break;
		case 3: DISPATCH_FAMILY(desc, out[wi] = loglike_t<Fam, 3>(desc.theory, src + wi, net, g, wq, dre, dim, uni, df, pref, bins_per_cta)); break;
		case 4: DISPATCH_FAMILY(desc, out[wi] = loglike_t<Fam, 4>(desc.theory, src + wi, net, g, wq, dre, dim, uni, df, pref, bins_per_cta)); break;
		default: DISPATCH_FAMILY(desc, out[wi] = loglike_t<Fam, 5>(desc.theory, src + wi, net, g, wq, dre, dim, uni, df, pref, bins_per_cta)); break;
		}
	}
	return 0;
}

// ---- sampler mathematics (gwat_sampler_math.h) as plain C++ -----------------------------------------------------------------
#include "../gw_analysis_tools_b200/csrc/gwat_sampler_math.h"

extern "C" void hh_philox(const unsigned *counter, const unsigned *key, unsigned *out)
{
	uint32_t r[4];
	smp::philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1], r);
	for (int i = 0; i < 4; i++) out[i] = r[i];
}
extern "C" void hh_uniform2(unsigned long long seed, unsigned long long step, unsigned chain, unsigned purpose, double *out)
{
	smp::uniform2(seed, step, chain, purpose, out[0], out[1]);
}
extern "C" double hh_normal_from(double u0, double u1) { return smp::normal_from(u0, u1); }
extern "C" void hh_step_boundaries(double T, int fisher_exist, int primed, double *out) { smp::step_boundaries(T, fisher_exist, primed, out); }
extern "C" double hh_log_prior(const gwat_b200_prior *prior, int pv2, int nrt, int dimension, const double *pos)
{
	smp::PriorPlan pp;
	pp.pv2 = pv2;
	pp.nrt = nrt;
	pp.tidal_love = prior->tidal_love;
	pp.dimension = dimension;
	pp.first_mod = (pv2 ? 15 : 11) + (nrt ? (prior->tidal_love ? 1 : 2) : 0);
	return smp::standard_log_prior(*prior, pp, pos);
}
extern "C" int hh_jacobi(const double *A, int n, double *vals, double *vecs)
{
	std::vector<double> work(A, A + n * n);
	return smp::jacobi_eigen(work.data(), n, vals, vecs) ? 1 : 0;
}
extern "C" void hh_fisher_transformations(double *F, int dim, int pv2, int alpha_fix, int nmod, const double *param)
{
	smp::fisher_transformations(F, dim, pv2, alpha_fix, nmod, param);
}
extern "C" int hh_propose(int type, const double *cur, double *prop, int dim, const double *widths, double u_pick, double u_pick2,
                          double z, double beta, int H, const double *hist, const double *vals, const double *vecs, double T)
{
	if (type == smp::STEP_GAUSS) return smp::propose_gaussian(cur, prop, dim, widths, u_pick, z);
	if (type == smp::STEP_DE) {
		int i, j;
		smp::de_pick(H, u_pick, u_pick2, i, j);
		smp::propose_de(cur, prop, dim, hist + (size_t)i * dim, hist + (size_t)j * dim, beta, z, widths[dim]);
		return i * H + j;
	}
	smp::propose_fisher(cur, prop, dim, vals, vecs, T, u_pick, z, widths[dim + 2]);
	return 0;
}
extern "C" int hh_mh_accept(double cll, double pll, double clp, double plp, double T, double u) { return smp::mh_accept(cll, pll, clp, plp, T, u) ? 1 : 0; }
extern "C" int hh_swap_decision(double ll1, double ll2, double T1, double T2, double alpha) { return smp::swap_decision(ll1, ll2, T1, T2, alpha); }
extern "C" double hh_tuned_width(double w, long long acc, long long rej, double lo, double hi) { return smp::tuned_width(w, acc, rej, lo, hi); }
