// TEST INFRASTRUCTURE ONLY -- drives the reference's OWN parallel-tempering sampler code and its OWN standard priors
// (compiled unmodified from /root/reference by oracle/Makefile) so that tests/ can compare the CUDA sampler with them:
//
//   src/mcmc_sampler_internals.cpp   allocate_sampler_mem, assign_probabilities, assign_initial_pos, mcmc_step, gaussian_step,
//                                    diff_ev_step, fisher_step, update_fisher, update_history, update_step_widths,
//                                    chain_swap / single_chain_swap
//   src/mcmc_sampler.cpp             PTMCMC_MH_step_incremental (the non-pool loop, :4571-4668), set up as PTMCMC_MH_internal does (:4259-4370)
//   src/standardPriorLibrary.cpp     logPriorStandard_{D,P,D_NRT,P_NRT}[_mod]::eval
//   src/mcmc_gw.cpp, src/fisher.cpp  the likelihood of every proposal (through oracle_ref_loglike_mcmc_batch of ref_driver.cpp)
//
// Third-party pieces the reference's sampler needs and this image lacks are stand-ins under standins/:
//   gsl_rng        scripted: every uniform / normal the reference draws is played back from per-chain queues that the harness
//                  fills with the CUDA sampler's counter-based draws, in the order the reference consumes them
//   Eigen          a Jacobi SelfAdjointEigenSolver; a test that follows the device's eigen-system step by step installs it through
//                  eigen_standin::override_next (eigenvectors are defined up to sign)
//   BayesShip      the abstract probabilityFn / positionInfo the priors derive from
// Nothing under gw_analysis_tools_b200/ links or loads this.
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include <eigen3/Eigen/Eigen>

#include "../include/gwat_b200.h"
#include "../include/gwat_b200_sampler.h"
#include "mcmc_sampler.h"
#include "mcmc_sampler_internals.h"
#include "standardPriorLibrary.h"
#include "util.h"

extern "C" int oracle_ref_loglike_mcmc_batch(const char *method, const gwat_b200_mod *mod, int dimension, int W, const double *params,
                                             double gmst, double T_segment, int D, const char *const *detectors, const double *f, int L,
                                             const double *psd, const double *data_re, const double *data_im, const double *weights,
                                             const char *integ, int log10F, int nthreads, double *logL, gwat_b200_source *sources_out);

namespace {

struct FisherScript {
	std::vector<double> vals, vecs;  // vecs[i*P + j]: component j of eigenvector i
};

struct PriorBox {
	priorData pd;
	std::vector<double *> mod_rows;
	std::vector<double> mod_store;
	bayesship::probabilityFn *fn = nullptr;
	~PriorBox() { delete fn; }
};

void fill_prior(const gwat_b200_prior &p, bool pv2, bool nrt, int n_mod, PriorBox &b)
{
	priorData &d = b.pd;
	std::memset(&d, 0, sizeof(d));
	auto cp = [](double *dst, const double *src) {
		dst[0] = src[0];
		dst[1] = src[1];
	};
	cp(d.mass1_prior, p.mass1_prior);
	cp(d.mass2_prior, p.mass2_prior);
	cp(d.spin1_prior, p.spin1_prior);
	cp(d.spin2_prior, p.spin2_prior);
	cp(d.a1_prior, p.a1_prior);
	cp(d.a2_prior, p.a2_prior);
	cp(d.ctheta1_prior, p.ctheta1_prior);
	cp(d.ctheta2_prior, p.ctheta2_prior);
	cp(d.phi1_prior, p.phi1_prior);
	cp(d.phi2_prior, p.phi2_prior);
	cp(d.tidal1_prior, p.tidal1_prior);
	cp(d.tidal2_prior, p.tidal2_prior);
	cp(d.tidal_s_prior, p.tidal_s_prior);
	cp(d.RA_bounds, p.RA_bounds);
	cp(d.sinDEC_bounds, p.sinDEC_bounds);
	cp(d.DL_prior, p.DL_prior);
	d.T_merger = p.T_merger;
	d.tidal_love = p.tidal_love != 0;
	d.tidal_love_error = false;
	d.alpha_param = false;
	b.mod_store.resize(2 * GWAT_B200_MAX_MOD);
	b.mod_rows.resize(GWAT_B200_MAX_MOD);
	for (int i = 0; i < GWAT_B200_MAX_MOD; i++) {
		b.mod_store[2 * i] = p.mod_priors[i][0];
		b.mod_store[2 * i + 1] = p.mod_priors[i][1];
		b.mod_rows[i] = &b.mod_store[2 * i];
	}
	d.mod_priors = b.mod_rows.data();
	// the class the reference's front-end picks for the method and the modification count (src/mcmc_gw_extended.cpp: the
	// *_mod variants when extra dimensions are present)
	if (pv2 && nrt) b.fn = n_mod > 0 ? (bayesship::probabilityFn *)new logPriorStandard_P_NRT_mod(&d) : new logPriorStandard_P_NRT(&d);
	else if (pv2) b.fn = n_mod > 0 ? (bayesship::probabilityFn *)new logPriorStandard_P_mod(&d) : new logPriorStandard_P(&d);
	else if (nrt) b.fn = n_mod > 0 ? (bayesship::probabilityFn *)new logPriorStandard_D_NRT_mod(&d) : new logPriorStandard_D_NRT(&d);
	else b.fn = n_mod > 0 ? (bayesship::probabilityFn *)new logPriorStandard_D_mod(&d) : new logPriorStandard_D(&d);
}

double eval_prior(PriorBox &b, double *pos, int dim)
{
	bayesship::positionInfo pi;
	pi.dimension = dim;
	pi.parameters = pos;
	return b.fn->eval(&pi, 0);
}

struct RefSampler {
	sampler s;
	int C = 0, P = 0, N = 0;
	std::vector<double> temps;
	double ***output = nullptr;
	// likelihood inputs
	std::string method;
	gwat_b200_mod mod;
	bool have_mod = false;
	double gmst = 0, T_segment = 0;
	std::vector<std::string> det_names;
	std::vector<const char *> det_ptrs;
	std::vector<double> f, psd, dre, dim;
	PriorBox prior;
	std::vector<std::deque<FisherScript>> fisher_script;
	long fisher_script_underflow = 0, ll_calls = 0, fisher_calls = 0;
	bool ran = false;
};

}  // namespace

extern "C" {

// log prior of W sampling vectors by the reference's own classes (N3)
int oracle_ref_log_prior_batch(int pv2, int nrt, int n_mod, const gwat_b200_prior *prior, int dimension, int W, const double *params, double *out)
{
	PriorBox b;
	fill_prior(*prior, pv2 != 0, nrt != 0, n_mod, b);
	std::vector<double> pos(dimension);
	for (int w = 0; w < W; w++) {
		for (int i = 0; i < dimension; i++) pos[i] = params[(size_t)w * dimension + i];
		out[w] = eval_prior(b, pos.data(), dimension);
	}
	return 0;
}

// Eigen stand-in alone: eigenvalues[n] ascending, eigenvectors[n][n] row i = vector i (the layout update_fisher stores)
int oracle_ref_eigen_standin(int n, const double *A, double *vals, double *vecs)
{
	std::vector<double> a(A, A + (size_t)n * n);
	Eigen::Map<Eigen::MatrixXd> m(a.data(), n, n);
	Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> es(m);
	for (int i = 0; i < n; i++) {
		vals[i] = es.eigenvalues()(i);
		for (int j = 0; j < n; j++) vecs[(size_t)i * n + j] = es.eigenvectors().col(i)(j);
	}
	return 0;
}

void *oracle_sampler_create(const char *method, const gwat_b200_mod *mod, int pv2, int nrt, int n_mod, int C, int P, int N_steps,
                            const double *temps, const double *initial /*[C][P]*/, int swp_freq, double swap_rate, int history_length,
                            int history_update, int fisher_exist, int fisher_update_number, int check_stepsize_freq,
                            const gwat_b200_prior *prior, double gmst, double T_segment, int D, const char *const *detectors,
                            const double *f, int L, const double *psd, const double *data_re, const double *data_im,
                            const double *init_fvals /*[C][P] or NULL*/, const double *init_fvecs /*[C][P][P] or NULL*/)
{
	RefSampler *r = new RefSampler;
	r->C = C;
	r->P = P;
	r->N = N_steps;
	r->temps.assign(temps, temps + C);
	r->method = method;
	if (mod) {
		r->mod = *mod;
		r->have_mod = true;
	}
	r->gmst = gmst;
	r->T_segment = T_segment;
	for (int d = 0; d < D; d++) r->det_names.push_back(detectors[d]);
	for (int d = 0; d < D; d++) r->det_ptrs.push_back(r->det_names[d].c_str());
	r->f.assign(f, f + L);
	r->psd.assign(psd, psd + (size_t)D * L);
	r->dre.assign(data_re, data_re + (size_t)D * L);
	r->dim.assign(data_im, data_im + (size_t)D * L);
	fill_prior(*prior, pv2 != 0, nrt != 0, n_mod, r->prior);
	r->fisher_script.resize(C);

	// ---- as PTMCMC_MH_internal sets the struct up (src/mcmc_sampler.cpp:4284-4330) ------------------------------------------
	sampler *sp = &r->s;
	sp->tune = true;
	sp->burn_phase = false;
	sp->fisher_exist = fisher_exist != 0;
	sp->log_ll = true;
	sp->log_lp = true;
	sp->lp = [r](double *pos, int *, int, mcmc_data_interface *, void *) { return eval_prior(r->prior, pos, r->P); };
	sp->ll = [r](double *pos, int *, int, mcmc_data_interface *, void *) {
		double out = 0;
		r->ll_calls++;
		const int D = (int)r->det_ptrs.size(), L = (int)r->f.size();
		oracle_ref_loglike_mcmc_batch(r->method.c_str(), r->have_mod ? &r->mod : nullptr, r->P, 1, pos, r->gmst, r->T_segment, D, r->det_ptrs.data(),
		                              r->f.data(), L, r->psd.data(), r->dre.data(), r->dim.data(), nullptr, "SIMPSONS", 0, 1, &out, nullptr);
		return out;
	};
	if (fisher_exist)
		sp->fish = [r](double *, int *, int, double **fisher, mcmc_data_interface *iface, void *) {
			// The matrix the chain's next eigen-decomposition is scripted to come from (see the header): F = V^T diag(vals) V, and the
			// Eigen stand-in is told to return exactly (vals, V) for it.
			const int c = iface->chain_id, P = r->P;
			r->fisher_calls++;
			if (r->fisher_script[c].empty()) {
				r->fisher_script_underflow++;
				for (int i = 0; i < P; i++)
					for (int j = 0; j < P; j++) fisher[i][j] = (i == j) ? 1.0 : 0.0;
				return;
			}
			const FisherScript fs = r->fisher_script[c].front();
			r->fisher_script[c].pop_front();
			for (int i = 0; i < P; i++)
				for (int j = 0; j < P; j++) {
					double acc = 0;
					for (int k = 0; k < P; k++) acc += fs.vals[k] * fs.vecs[(size_t)k * P + i] * fs.vecs[(size_t)k * P + j];
					fisher[i][j] = acc;
				}
			eigen_standin::override_next(P, fs.vals.data(), fs.vecs.data());
		};
	sp->swp_freq = swp_freq;      // (PTMCMC_MH_internal hard-codes 2 here and puts the caller's value into swap_rate, :4315-4316)
	sp->swap_rate = swap_rate;
	sp->chain_temps = r->temps.data();
	sp->chain_N = C;
	sp->N_steps = N_steps;
	sp->dimension = sp->min_dim = sp->max_dim = P;
	sp->show_progress = false;
	sp->num_threads = 1;
	sp->numThreads = 1;
	sp->user_parameters = nullptr;
	sp->pool = false;
	sp->history_length = history_length;
	sp->history_update = history_update;
	sp->random_swaps = false;      // chain_swap (the sweep over adjacent chains) instead of full_random_swap
	r->output = allocate_3D_array(C, N_steps, P);
	sp->output = r->output;
	allocate_sampler_mem(sp);
	// allocate_sampler_mem fixes these two for tune == true (200 and 50, :1846, 1956); tests shorten them
	sp->fisher_update_number = fisher_update_number;
	for (int j = 0; j < C; j++) {
		sp->check_stepsize_freq[j] = check_stepsize_freq;
		sp->fisher_update_ct[j] = fisher_update_number;
		sp->rvec[j]->scripted = true;
	}
	for (int j = 0; j < C; j++) assign_probabilities(sp, j);
	std::vector<int> init_status(P, 1);
	std::vector<std::vector<double>> pos(C, std::vector<double>(P));
	std::vector<std::vector<int>> stat(C, std::vector<int>(P, 1));
	std::vector<double *> posp(C);
	std::vector<int *> statp(C);
	for (int j = 0; j < C; j++) {
		for (int i = 0; i < P; i++) pos[j][i] = initial[(size_t)j * P + i];
		posp[j] = pos[j].data();
		statp[j] = stat[j].data();
	}
	// assign_initial_pos evaluates prior and likelihood of every chain and, with a Fisher, every chain's first matrix (:3182-3212)
	if (fisher_exist && init_fvals && init_fvecs)
		for (int j = 0; j < C; j++) {
			FisherScript fs;
			fs.vals.assign(init_fvals + (size_t)j * P, init_fvals + (size_t)(j + 1) * P);
			fs.vecs.assign(init_fvecs + (size_t)j * P * P, init_fvecs + (size_t)(j + 1) * P * P);
			r->fisher_script[j].push_back(fs);
		}
	assign_initial_pos(sp, pos[0].data(), init_status.data(), 0, posp.data(), statp.data(), nullptr, nullptr);
	return r;
}

// queue draws for one chain: uniforms and unit normals, each consumed in order
int oracle_sampler_push(void *h, int chain, int n_u, const double *u, int n_n, const double *n)
{
	RefSampler *r = static_cast<RefSampler *>(h);
	if (chain < 0 || chain >= r->C) return -1;
	gsl_rng *g = r->s.rvec[chain];
	for (int i = 0; i < n_u; i++) g->script_u.push_back(u[i]);
	for (int i = 0; i < n_n; i++) g->script_n.push_back(n[i]);
	return 0;
}

// queue the eigen-system the chain's next update_fisher will obtain
int oracle_sampler_push_fisher(void *h, int chain, const double *vals, const double *vecs)
{
	RefSampler *r = static_cast<RefSampler *>(h);
	if (chain < 0 || chain >= r->C) return -1;
	FisherScript fs;
	fs.vals.assign(vals, vals + r->P);
	fs.vecs.assign(vecs, vecs + (size_t)r->P * r->P);
	r->fisher_script[chain].push_back(fs);
	return 0;
}

// the reference's loop over all N_steps (one call: its swap indexes `output` by the loop-local step count, :4657)
int oracle_sampler_run(void *h)
{
	RefSampler *r = static_cast<RefSampler *>(h);
	if (r->ran) return -1;
	r->ran = true;
	PTMCMC_MH_step_incremental(&r->s, r->N);
	return 0;
}

// trajectories output[C][N][P], ll_lp[C][N][2]; counters[C][12] in the order of GWAT_B200_CT_*; widths[C][P+3] (Gaussian per
// dimension, DE, MMALA, Fisher); diag[4] = {rng underflows, unconsumed uniforms, unconsumed normals, fisher-script underflows}
int oracle_sampler_results(void *h, double *output, double *ll_lp, long long *counters, double *widths, double *fvals, double *fvecs,
                           int *chain_pos, long long *diag)
{
	RefSampler *r = static_cast<RefSampler *>(h);
	sampler *sp = &r->s;
	const int C = r->C, P = r->P, N = r->N;
	long long under = 0, left_u = 0, left_n = 0;
	for (int j = 0; j < C; j++) {
		for (int l = 0; l < N; l++) {
			for (int i = 0; i < P; i++) output[((size_t)j * N + l) * P + i] = sp->output[j][l][i];
			ll_lp[((size_t)j * N + l) * 2 + 0] = sp->ll_lp_output[j][l][0];
			ll_lp[((size_t)j * N + l) * 2 + 1] = sp->ll_lp_output[j][l][1];
		}
		long long *ct = counters + (size_t)j * GWAT_B200_SAMPLER_NCOUNTERS;
		ct[GWAT_B200_CT_STEP_ACCEPT] = sp->step_accept_ct[j];
		ct[GWAT_B200_CT_STEP_REJECT] = sp->step_reject_ct[j];
		ct[GWAT_B200_CT_GAUSS_ACCEPT] = sp->gauss_accept_ct[j];
		ct[GWAT_B200_CT_GAUSS_REJECT] = sp->gauss_reject_ct[j];
		ct[GWAT_B200_CT_DE_ACCEPT] = sp->de_accept_ct[j];
		ct[GWAT_B200_CT_DE_REJECT] = sp->de_reject_ct[j];
		ct[GWAT_B200_CT_FISHER_ACCEPT] = sp->fish_accept_ct[j];
		ct[GWAT_B200_CT_FISHER_REJECT] = sp->fish_reject_ct[j];
		ct[GWAT_B200_CT_SWAP_ACCEPT] = sp->swap_accept_ct[j];
		ct[GWAT_B200_CT_SWAP_REJECT] = sp->swap_reject_ct[j];
		ct[GWAT_B200_CT_FISHER_UPDATES] = 0;
		ct[GWAT_B200_CT_FISHER_NAN] = sp->nan_counter[j];
		for (int i = 0; i < P; i++) widths[(size_t)j * (P + 3) + i] = sp->randgauss_width[j][0][i];
		widths[(size_t)j * (P + 3) + P + 0] = sp->randgauss_width[j][1][0];
		widths[(size_t)j * (P + 3) + P + 1] = sp->randgauss_width[j][2][0];
		widths[(size_t)j * (P + 3) + P + 2] = sp->randgauss_width[j][3][0];
		for (int i = 0; i < P; i++) {
			fvals[(size_t)j * P + i] = sp->fisher_vals[j][i];
			for (int k = 0; k < P; k++) fvecs[((size_t)j * P + i) * P + k] = sp->fisher_vecs[j][i][k];
		}
		chain_pos[j] = sp->chain_pos[j];
		under += sp->rvec[j]->underflow;
		left_u += (long long)sp->rvec[j]->script_u.size();
		left_n += (long long)sp->rvec[j]->script_n.size();
	}
	diag[0] = under;
	diag[1] = left_u;
	diag[2] = left_n;
	diag[3] = r->fisher_script_underflow;
	diag[4] = r->ll_calls;
	diag[5] = r->fisher_calls;
	return 0;
}

void oracle_sampler_destroy(void *h)
{
	RefSampler *r = static_cast<RefSampler *>(h);
	if (!r) return;
	// (deallocate_sampler_mem frees what allocate_sampler_mem made; `output` is ours)
	deallocate_sampler_mem(&r->s);
	deallocate_3D_array(r->output, r->C, r->N, r->P);
	delete r;
}

// update_temperatures_full_ensemble (src/mcmc_sampler_internals.cpp:3312-3413) with linear swapping, on a sampler struct that
// carries nothing but the ladder and the swap indicators A.
int oracle_ref_update_temperatures(int chain_N, double *chain_temps, int *A, int t0, int nu, int t)
{
	sampler s;
	s.chain_N = chain_N;
	s.chain_temps = chain_temps;
	s.A = A;
	s.linear_swapping = true;
	update_temperatures_full_ensemble(&s, t0, nu, t);
	return 0;
}

}  // extern "C"
