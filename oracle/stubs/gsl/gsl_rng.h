// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_rng.h>: any PRNG will do, random draws are off-path
// (only proposals and tidal_love_error, which the oracle keeps false).
#ifndef ORACLE_STUB_GSL_RNG_H
#define ORACLE_STUB_GSL_RNG_H
#include <random>
struct gsl_rng_type { int dummy; };
struct gsl_rng { std::mt19937_64 eng; };
static const gsl_rng_type oracle_gsl_rng_default_t = {0};
static const gsl_rng_type *gsl_rng_default = &oracle_gsl_rng_default_t;
static const gsl_rng_type *gsl_rng_mt19937 = &oracle_gsl_rng_default_t;
inline const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng; }
inline void gsl_rng_free(gsl_rng *r) { delete r; }
inline void gsl_rng_set(gsl_rng *r, unsigned long s) { r->eng.seed(s); }
inline double gsl_rng_uniform(gsl_rng *r) { return std::uniform_real_distribution<double>(0., 1.)(r->eng); }
inline unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long n) { return std::uniform_int_distribution<unsigned long>(0, n - 1)(r->eng); }
#endif
