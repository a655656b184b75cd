// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_rng.h>.
// Two modes per generator:
//   free-running (default)  a std::mt19937_64 -- random draws are off the likelihood path (only proposals and
//                           tidal_love_error, which the oracle keeps false);
//   scripted                the draws are played back from two queues (uniforms, unit normals) that a test harness fills
//                           (oracle/sampler_driver.cpp): this is how the reference's OWN sampler steps are driven with the
//                           counter-based draws the CUDA sampler uses.  An empty queue is recorded (underflow) and answered
//                           with 0.5 / 0 so that the harness can report it instead of crashing the test process.
#ifndef ORACLE_STUB_GSL_RNG_H
#define ORACLE_STUB_GSL_RNG_H
#include <deque>
#include <random>
struct gsl_rng_type { int dummy; };
struct gsl_rng {
	std::mt19937_64 eng;
	unsigned long seed = 0;
	bool scripted = false;
	std::deque<double> script_u, script_n;
	long underflow = 0, served_u = 0, served_n = 0;
};
static const gsl_rng_type oracle_gsl_rng_default_t = {0};
static const gsl_rng_type *gsl_rng_default = &oracle_gsl_rng_default_t;
static const gsl_rng_type *gsl_rng_mt19937 = &oracle_gsl_rng_default_t;
inline const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng; }
inline void gsl_rng_free(gsl_rng *r) { delete r; }
inline void gsl_rng_set(gsl_rng *r, unsigned long s)
{
	r->seed = s;
	r->eng.seed(s);
}
inline double gsl_rng_uniform(gsl_rng *r)
{
	if (r->scripted) {
		if (r->script_u.empty()) {
			r->underflow++;
			return 0.5;
		}
		const double v = r->script_u.front();
		r->script_u.pop_front();
		r->served_u++;
		return v;
	}
	return std::uniform_real_distribution<double>(0., 1.)(r->eng);
}
inline unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long n) { return std::uniform_int_distribution<unsigned long>(0, n - 1)(r->eng); }
// one unit normal of the script (gsl_ran_gaussian scales it by sigma)
inline double oracle_gsl_scripted_normal(gsl_rng *r)
{
	if (r->script_n.empty()) {
		r->underflow++;
		return 0.0;
	}
	const double v = r->script_n.front();
	r->script_n.pop_front();
	r->served_n++;
	return v;
}
#endif
