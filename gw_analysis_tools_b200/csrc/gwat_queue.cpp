// Coalescing front of gwat_b200_loglike_mcmc_batch for callers that evaluate ONE chain per call from many threads.
//
// The reference's samplers hand one chain at a time to `ll(param, &status, model_status, interface, user)`
// (include/gwat/mcmc_sampler_internals.h:169-171 -> MCMC_likelihood_wrapper, src/mcmc_gw.cpp:2569) from the workers of a
// thread pool (src/mcmc_sampler.cpp:347-447).  A GPU evaluation of a single chain is launch-bound, so the unmodified
// sampler would see no speed-up from a drop-in that forwards each call on its own.  This queue keeps the one-chain
// signature and merges the calls that are in flight at the same moment into one batched launch:
//
//   * the first caller of a generation becomes its leader; it waits until `expected_callers` calls have joined, the
//     batch is full, or `max_wait_us` have passed since it arrived, whichever comes first;
//   * it then closes the generation (later arrivals open the next one and can fill up while the GPU is busy), runs ONE
//     gwat_b200_loglike_mcmc_batch over the collected vectors and publishes the results;
//   * followers sleep on a condition variable until their generation is published.
//
// Results do not depend on how calls were grouped: the per-walker reduction order of k_loglike/k_finish is fixed by the
// grid, not by the batch (tests/test_queue.py checks bit-equality with a direct batch call).
// Host code only; everything numerical happens behind the batched C ABI call.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gwat_b200.h"

namespace {

struct Generation {
	std::vector<double> params;  // [n][dimension]
	std::vector<double> logL;    // [n]
	int n = 0;
	int status = 0;
	bool closed = false;  // no more joiners
	bool done = false;    // results published
};

}  // namespace

struct gwat_b200_queue {
	gwat_b200_ctx *ctx = nullptr;
	std::string method;
	gwat_b200_mod mod;
	bool has_mod = false;
	int dimension = 0;
	double gmst = 0, T_segment = 0;
	int max_batch = 0, expected = 0;
	double max_wait_us = 0;

	std::mutex m;
	std::condition_variable joined;     // a caller joined the open generation (wakes its leader)
	std::condition_variable published;  // a generation's results are available (wakes its followers)
	std::shared_ptr<Generation> open;   // generation currently accepting joiners (null: none)
	std::mutex gpu;                     // generations run on the context one at a time, in closing order

	long long calls = 0, batches = 0;
	int largest = 0, last_status = 0;
};

extern "C" {

int gwat_b200_queue_create(gwat_b200_queue **queue, gwat_b200_ctx *ctx, const char *generation_method, const gwat_b200_mod *mod,
                           int dimension, double gmst, double T_segment, int max_batch, int expected_callers, double max_wait_us)
{
	if (!queue) return GWAT_B200_ERR_ARG;
	*queue = nullptr;
	if (!ctx || !generation_method || dimension <= 0 || dimension > GWAT_B200_MAX_DIM || max_batch <= 0 || expected_callers <= 0 ||
	    !(max_wait_us >= 0))
		return GWAT_B200_ERR_ARG;
	gwat_b200_queue *q = new gwat_b200_queue;
	q->ctx = ctx;
	q->method = generation_method;
	if (mod) {
		q->mod = *mod;
		q->has_mod = true;
	}
	q->dimension = dimension;
	q->gmst = gmst;
	q->T_segment = T_segment;
	q->max_batch = max_batch;
	q->expected = expected_callers < max_batch ? expected_callers : max_batch;
	q->max_wait_us = max_wait_us;
	*queue = q;
	return GWAT_B200_OK;
}

void gwat_b200_queue_destroy(gwat_b200_queue *q) { delete q; }

double gwat_b200_queue_loglike(gwat_b200_queue *q, const double *param, int *status)
{
	const double nan = std::numeric_limits<double>::quiet_NaN();
	if (status) *status = GWAT_B200_ERR_ARG;
	if (!q || !param) return nan;
	std::unique_lock<std::mutex> lk(q->m);
	q->calls++;
	bool leader = false;
	if (!q->open) {
		q->open = std::make_shared<Generation>();
		q->open->params.reserve((size_t)q->expected * q->dimension);
		leader = true;
	}
	std::shared_ptr<Generation> g = q->open;
	const int slot = g->n++;
	g->params.insert(g->params.end(), param, param + q->dimension);
	if (g->n >= q->max_batch) {  // full: nobody else may join, whoever arrives next opens a new generation
		g->closed = true;
		q->open.reset();
	}

	if (!leader) {
		q->joined.notify_all();
		q->published.wait(lk, [&] { return g->done; });
	} else {
		const auto deadline = std::chrono::steady_clock::now() + std::chrono::nanoseconds((long long)(q->max_wait_us * 1e3));
		q->joined.wait_until(lk, deadline, [&] { return g->closed || g->n >= q->expected; });
		if (!g->closed) {
			g->closed = true;
			q->open.reset();
		}
		const int n = g->n;
		g->logL.assign(n, nan);
		lk.unlock();
		int rc;
		{
			std::lock_guard<std::mutex> run(q->gpu);
			rc = gwat_b200_loglike_mcmc_batch(q->ctx, q->method.c_str(), q->has_mod ? &q->mod : nullptr, q->dimension, n,
			                                  g->params.data(), q->gmst, q->T_segment, g->logL.data());
		}
		lk.lock();
		g->status = rc;
		g->done = true;
		q->batches++;
		if (n > q->largest) q->largest = n;
		q->last_status = rc;
		q->published.notify_all();
	}
	if (status) *status = g->status;
	return g->status == 0 ? g->logL[slot] : nan;
}

int gwat_b200_queue_stats(gwat_b200_queue *q, long long *calls, long long *batches, int *largest_batch)
{
	if (!q) return GWAT_B200_ERR_ARG;
	std::lock_guard<std::mutex> lk(q->m);
	if (calls) *calls = q->calls;
	if (batches) *batches = q->batches;
	if (largest_batch) *largest_batch = q->largest;
	return GWAT_B200_OK;
}

}  // extern "C"
