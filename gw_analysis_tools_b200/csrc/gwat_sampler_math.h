// Per-chain mathematics of the batched parallel-tempering step (host/device, see gwat_hd.h): counter-based random numbers,
// the standard priors, the three proposals, the Metropolis-Hastings and swap rules, width tuning and a Jacobi eigen-solver.
// Each function cites the reference code it follows; the kernels in gwat_sampler.cu only move data around these.
#ifndef GWAT_SAMPLER_MATH_H
#define GWAT_SAMPLER_MATH_H

#include <stdint.h>

#include "../../include/gwat_b200_sampler.h"
#include "gwat_hd.h"
#include "gwat_repack.h"

namespace gwat {
namespace smp {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (c0..c3), key = (k0, k1) ----------------------------------------------
GWAT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
	for (int r = 0; r < 10; r++) {
		const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
		const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += W0; k1 += W1;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// what a draw is for; part of the counter so that purposes never share numbers
enum Purpose { DRAW_TYPE_ACCEPT = 0, DRAW_PICK = 1, DRAW_NORMAL = 2, DRAW_DE_SCALE = 3, DRAW_SWAP = 4, DRAW_SWAP_GATE = 5 };

// two uniforms in [0,1) with 53 random bits each (the (a>>5, b>>6) construction)
GWAT_HD void uniform2(uint64_t seed, uint64_t step, uint32_t chain, uint32_t purpose, double &u0, double &u1)
{
	uint32_t r[4];
	philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), chain, purpose, (uint32_t)seed, (uint32_t)(seed >> 32), r);
	u0 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) * (1.0 / 9007199254740992.0);
	u1 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) * (1.0 / 9007199254740992.0);
}
// standard normal (Box-Muller; gsl_ran_gaussian is the polar variant -- same distribution)
GWAT_HD double normal_from(double u0, double u1) { return sqrt(-2.0 * log(1.0 - u0)) * cos(GWAT_TWOPI * u1); }

// ---- step-type probabilities, non-RJ branch of assign_probabilities (src/mcmc_sampler_internals.cpp:1196-1250,1357-1362) ----
// bounds[0..3]: cumulative boundaries for Gaussian, DE, MMALA (always empty), Fisher
GWAT_HD void step_boundaries(double T, bool fisher_exist, bool de_primed, double *bounds)
{
	double p[4] = {0, 0, 0, 0};
	if (!fisher_exist) {
		p[0] = 1.;  // with or without a primed history: the reference leaves DE off when there is no Fisher
	} else if (!de_primed) {
		p[3] = .1 + .8 / T;
		const double sum = p[1] + p[2] + p[3];
		p[0] = 1 - sum;
	} else {
		p[1] = .7 - .4 / T;
		p[3] = .2 + .5 / T;
		const double sum = p[1] + p[2] + p[3] + 0.;
		p[0] = 1 - sum;
	}
	bounds[0] = p[0];
	bounds[1] = p[1] + bounds[0];
	bounds[2] = p[2] + bounds[1];
	bounds[3] = p[3] + bounds[2];
}
enum StepType { STEP_GAUSS = 0, STEP_DE = 1, STEP_MMALA = 2, STEP_FISHER = 3 };
// mcmc_step's cascade (:46-82); anything past the Fisher boundary would be RJ/KDE, which have probability 0 here
GWAT_HD int step_type(double alpha, const double *bounds)
{
	if (alpha < bounds[0]) return STEP_GAUSS;
	if (alpha < bounds[1]) return STEP_DE;
	if (alpha < bounds[2]) return STEP_MMALA;
	return STEP_FISHER;
}

// ---- standard priors (src/standardPriorLibrary.cpp) ------------------------------------------------------------------------
struct PriorPlan {
	int pv2, nrt, tidal_love, dimension;
	int first_mod;  // initial_nongr_id of the *_mod variants; == dimension when the model has no modifications
};

GWAT_HD double chirpmass_eta_jac(double chirpmass, double eta)
{  // :10-19
	const double epsilon = 1e-12;
	double delta = sqrt(1. - 4. * eta);
	if (eta > .25 - epsilon) delta = sqrt(1. - 4. * (eta - epsilon));
	return chirpmass * chirpmass / (delta * pow(eta, 1.2));
}
GWAT_HD double aligned_spin_prior(double chi)
{  // :25-28
	const double a = 0.0039132, b = 3.95381;
	return a * exp(-b * fabs(chi));
}
GWAT_HD bool tidal_love_boundary_violation(double q, double lambda_s) { return q < 1.2321 - .124616 * log(lambda_s); }  // :30-38

GWAT_HD bool outside(double x, const double *b) { return x < b[0] || x > b[1]; }

// logPriorStandard_D::eval (:409-436) and logPriorStandard_P::eval (:438-469)
GWAT_HD double log_prior_base(const gwat_b200_prior &PD, bool pv2, const double *pos)
{
	const double a = -INFINITY;
	const double chirp = exp(pos[7]);
	const double eta = pos[8];
	if (eta < .0 || eta > .25) return a;
	double m1, m2;
	masses_of(chirp, eta, m1, m2);
	if (outside(m1, PD.mass1_prior)) return a;
	if (outside(m2, PD.mass2_prior)) return a;
	if (outside(pos[0], PD.RA_bounds)) return a;
	if (outside(pos[1], PD.sinDEC_bounds)) return a;
	if (pos[2] < 0 || pos[2] > GWAT_PI) return a;
	if (pos[3] < -1 || pos[3] > 1) return a;
	if (pos[4] < 0 || pos[4] > 2 * GWAT_PI) return a;
	if (pos[5] < (PD.T_merger - .1) || pos[5] > (PD.T_merger + .1)) return a;
	if (outside(exp(pos[6]), PD.DL_prior)) return a;
	if (!pv2) {
		if (outside(pos[9], PD.spin1_prior)) return a;
		if (outside(pos[10], PD.spin2_prior)) return a;
		return log(aligned_spin_prior(pos[9])) + log(aligned_spin_prior(pos[10])) + log(chirpmass_eta_jac(chirp, eta)) + 3 * pos[6];
	}
	if (outside(pos[9], PD.a1_prior)) return a;
	if (outside(pos[10], PD.a2_prior)) return a;
	if (outside(pos[11], PD.ctheta1_prior)) return a;
	if (outside(pos[12], PD.ctheta2_prior)) return a;
	if (outside(pos[13], PD.phi1_prior)) return a;
	if (outside(pos[14], PD.phi2_prior)) return a;
	return log(chirpmass_eta_jac(chirp, eta)) + 3 * pos[6];
}

// The whole family: [_mod] -> [_NRT] -> base, in the order the reference nests them (:321-335, 337-353, 383-407, 471-526)
GWAT_HD double standard_log_prior(const gwat_b200_prior &PD, const PriorPlan &pp, const double *pos)
{
	const double a = -INFINITY;
	for (int i = pp.first_mod; i < pp.dimension; i++)
		if (outside(pos[i], PD.mod_priors[i - pp.first_mod])) return a;
	double factor = 0;
	if (pp.nrt) {
		const double chirp = exp(pos[7]);
		double m1, m2;
		masses_of(chirp, pos[8], m1, m2);
		const double q = m2 / m1;
		const int t0 = pp.pv2 ? 15 : 11;
		if (PD.tidal_love) {
			if (outside(exp(pos[t0]), PD.tidal_s_prior)) return a;
			// the precessing variant tests and weights pos[11] here, not pos[15] (:499-500); kept as the reference has it
			if (tidal_love_boundary_violation(q, exp(pos[11]))) return a;
			factor += pos[11];
		} else {
			if (outside(exp(pos[t0]), PD.tidal1_prior)) return a;
			if (outside(exp(pos[t0 + 1]), PD.tidal2_prior)) return a;
			factor += pos[t0];
			factor += pos[t0 + 1];
		}
	}
	const double base = log_prior_base(PD, pp.pv2 != 0, pos);
	return pp.nrt ? base + factor : base;
}

// ---- proposals ---------------------------------------------------------------------------------------------------------------
// gaussian_step (:364-397) with every coordinate active: one coordinate moves by N(0, width[beta])
GWAT_HD int propose_gaussian(const double *cur, double *prop, int dim, const double *widths, double u_pick, double z)
{
	const int beta = (int)(u_pick * dim);
	for (int i = 0; i < dim; i++) prop[i] = cur[i];
	prop[beta] = z * widths[beta] + cur[beta];
	return beta;
}
// diff_ev_step, "regular PTMCMC" branch (:930-975).  The second index is drawn from the other H-1 slots directly instead of
// by rejection (same distribution as the reference's do-while).
GWAT_HD void de_pick(int H, double u_i, double u_j, int &i, int &j)
{
	i = (int)(H * u_i);
	j = (i + 1 + (int)((H - 1) * u_j)) % H;
}
GWAT_HD void propose_de(const double *cur, double *prop, int dim, const double *hist_i, const double *hist_j, double beta,
                        double z, double width)
{
	double alpha = 1;
	if (beta < .9) alpha = z * width;
	for (int k = 0; k < dim; k++) prop[k] = cur[k] + alpha * (hist_i[k] - hist_j[k]);
}
// fisher_step, plain branch (:491-518): along eigenvector beta, scaled by 1/sqrt(|eigenvalue|/T), floor 10
GWAT_HD void propose_fisher(const double *cur, double *prop, int dim, const double *vals, const double *vecs /*[dim][dim]*/,
                            double T, double u_pick, double z, double width)
{
	const int beta = (int)(dim * u_pick);
	const double alpha = z * width;
	double scaling;
	if (fabs(vals[beta]) < 10) scaling = 10.;
	else scaling = fabs(vals[beta]) / T;
	const double s = alpha / sqrt(scaling);
	for (int i = 0; i < dim; i++) prop[i] = cur[i] + s * vecs[beta * dim + i];
}

// ---- acceptance rules --------------------------------------------------------------------------------------------------------
// mcmc_step (:84-146): returns true when the proposal is accepted
GWAT_HD bool mh_accept(double current_ll, double proposed_ll, double current_lp, double proposed_lp, double T, double u)
{
	double MH_ratio;
	if (current_lp == -INFINITY || proposed_lp == -INFINITY) MH_ratio = -INFINITY;
	else if (proposed_ll != proposed_ll) MH_ratio = -INFINITY;
	else MH_ratio = (-current_ll + proposed_ll) / T - current_lp + proposed_lp;
	const double beta = log(u);
	return !(MH_ratio < beta);
}
// single_chain_swap (:1121-1184): -1 same temperature (chain_swap books it as a rejection, :1100-1117), 0 rejected, 1 accepted
GWAT_HD int swap_decision(double ll1, double ll2, double T1, double T2, double alpha)
{
	if (T1 == T2) return -1;
	const double pw = (ll1 - ll2) / T2 - (ll1 - ll2) / T1;
	const double MH_ratio = exp(pw);
	return (MH_ratio < alpha) ? 0 : 1;
}

// The same test as a threshold on ll1 for fixed ll2, T1, T2, alpha (the swap sweep carries ll1 along the ladder, so everything
// else can be prepared for all pairs at once).  kind 0: never (equal temperatures), 1: swap iff ll1 >= thr, 2: swap iff
// ll1 <= thr, 3: always.
GWAT_HD void swap_threshold(double ll2, double T1, double T2, double alpha, int &kind, double &thr)
{
	thr = 0;
	if (T1 == T2) {
		kind = 0;
		return;
	}
	const double g = 1. / T2 - 1. / T1;  // (ll1 - ll2) * g >= ln(alpha)
	const double la = log(alpha);        // alpha = 0 -> -inf -> always
	if (g == 0 || la == -INFINITY) {
		kind = (la <= 0) ? 3 : 0;
	} else {
		kind = g > 0 ? 1 : 2;
		thr = ll2 + la / g;
	}
}
// chain_swap's sweep (:1086-1118) over C slots: src[i] = slot whose state ends up in slot i; accepted[i] for pair (i, i+1)
GWAT_HD void swap_scan(const double *ll, const double *thr, const int *kind, int C, int *src, int *accepted)
{
	double carry = ll[0];
	int carry_src = 0;
	for (int i = 0; i < C - 1; i++) {
		const int kd = kind[i];
		const double th = thr[i];
		const bool sw = (kd == 3) || (kd == 1 && carry >= th) || (kd == 2 && carry <= th);
		if (sw) {
			src[i] = i + 1;  // slot i receives the untouched state of slot i+1; the carried state moves on to slot i+1
		} else {
			src[i] = carry_src;
			carry = ll[i + 1];
			carry_src = i + 1;
		}
		accepted[i] = sw ? 1 : 0;
	}
	src[C - 1] = carry_src;
}

// update_step_widths (:1623-1703): one width, its accept/reject counts since the last check
GWAT_HD double tuned_width(double width, long long acc, long long rej, double min_target, double max_target)
{
	const double frac = (double)acc / (double)(acc + rej);  // 0/0 -> NaN -> neither branch, as in the reference
	if (frac < min_target) return width * .9;
	if (frac > max_target) return width * 1.1;
	return width;
}

// MCMC_fisher_transformations (src/mcmc_gw.cpp:2136-2189), extrinsic (non-"intrinsic") branch
GWAT_HD void fisher_transformations(double *F, int dim, bool pv2, bool alpha_unit_fix, int ppE_Nmod, const double *param)
{
	const double pi2 = 4 * GWAT_PI * GWAT_PI;
	F[0 * dim + 0] += 1. / pi2;
	F[1 * dim + 1] += 1. / 4;
	F[2 * dim + 2] += 1. / pi2;
	F[3 * dim + 3] += 1. / (4);
	F[4 * dim + 4] += 1. / pi2;
	F[5 * dim + 5] += 1. / (.01);
	F[8 * dim + 8] += 1. / .25;
	F[9 * dim + 9] += 1. / 4;
	F[10 * dim + 10] += 1. / 4;
	if (pv2) {
		F[11 * dim + 11] += 1. / 4;
		F[12 * dim + 12] += 1. / 4;
		F[13 * dim + 13] += 1. / pi2;
		F[14 * dim + 14] += 1. / pi2;
	}
	if (alpha_unit_fix) {  // dCS / EdGB: sqrt(alpha) in km sampled, alpha^2 in s^4 differentiated
		const int base = dim - ppE_Nmod;
		double factor = 4 * pow(param[base], 3. / 4.);
		factor *= 1000 / GWAT_C_SI;
		for (int i = 0; i < dim; i++) {
			F[base * dim + i] *= factor;
			F[i * dim + base] *= factor;
		}
	}
}

// Symmetric eigen-decomposition by cyclic Jacobi rotations (replaces Eigen::SelfAdjointEigenSolver in update_fisher, :664-667).
// A[n][n] is destroyed; vals ascending, vecs[i*n + j] = component j of eigenvector i (the layout fisher_vecs uses, :687).
// Returns false when the result contains a NaN (the reference then keeps the old eigen-system, :676-711).
GWAT_HD bool jacobi_eigen(double *A, int n, double *vals, double *vecs)
{
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) vecs[i * n + j] = (i == j) ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 60; sweep++) {
		double off = 0, diag = 0;
		for (int i = 0; i < n; i++) {
			diag += A[i * n + i] * A[i * n + i];
			for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
		}
		if (!(off > 1e-32 * diag)) break;
		for (int p = 0; p < n - 1; p++) {
			for (int q = p + 1; q < n; q++) {
				const double apq = A[p * n + q];
				if (apq == 0.0) continue;
				const double app = A[p * n + p], aqq = A[q * n + q];
				const double theta = (aqq - app) / (2.0 * apq);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
				for (int k = 0; k < n; k++) {
					const double akp = A[k * n + p], akq = A[k * n + q];
					A[k * n + p] = c * akp - s * akq;
					A[k * n + q] = s * akp + c * akq;
				}
				for (int k = 0; k < n; k++) {
					const double apk = A[p * n + k], aqk = A[q * n + k];
					A[p * n + k] = c * apk - s * aqk;
					A[q * n + k] = s * apk + c * aqk;
				}
				for (int k = 0; k < n; k++) {
					const double vpk = vecs[p * n + k], vqk = vecs[q * n + k];
					vecs[p * n + k] = c * vpk - s * vqk;
					vecs[q * n + k] = s * vpk + c * vqk;
				}
			}
		}
	}
	for (int i = 0; i < n; i++) vals[i] = A[i * n + i];
	// ascending order (Eigen's convention), selection sort on the small n
	for (int i = 0; i < n - 1; i++) {
		int m = i;
		for (int j = i + 1; j < n; j++)
			if (vals[j] < vals[m]) m = j;
		if (m != i) {
			const double tv = vals[i];
			vals[i] = vals[m];
			vals[m] = tv;
			for (int k = 0; k < n; k++) {
				const double t = vecs[i * n + k];
				vecs[i * n + k] = vecs[m * n + k];
				vecs[m * n + k] = t;
			}
		}
	}
	bool ok = true;
	for (int i = 0; i < n; i++) {
		ok = ok && (vals[i] == vals[i]);
		for (int j = 0; j < n; j++) ok = ok && (vecs[i * n + j] == vecs[i * n + j]);
	}
	return ok;
}

}  // namespace smp
}  // namespace gwat
#endif
