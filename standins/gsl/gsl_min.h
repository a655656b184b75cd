// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_min.h> (off-path).
#ifndef ORACLE_STUB_GSL_MIN_H
#define ORACLE_STUB_GSL_MIN_H
#include <cstdio>
#include <cstdlib>
#include "gsl/gsl_math.h"
struct gsl_min_fminimizer_type { int dummy; };
struct gsl_min_fminimizer { int dummy; };
static const gsl_min_fminimizer_type oracle_gsl_min_brent_t = {0};
static const gsl_min_fminimizer_type *gsl_min_fminimizer_brent = &oracle_gsl_min_brent_t;
static const gsl_min_fminimizer_type *gsl_min_fminimizer_goldensection = &oracle_gsl_min_brent_t;
#define ORACLE_GSL_MIN_ABORT do { std::fprintf(stderr, "oracle stub: gsl_min called (off-path)\n"); std::abort(); } while (0)
inline gsl_min_fminimizer *gsl_min_fminimizer_alloc(const gsl_min_fminimizer_type *) { ORACLE_GSL_MIN_ABORT; return 0; }
inline void gsl_min_fminimizer_free(gsl_min_fminimizer *) {}
inline int gsl_min_fminimizer_set(gsl_min_fminimizer *, gsl_function *, double, double, double) { ORACLE_GSL_MIN_ABORT; return -1; }
inline int gsl_min_fminimizer_iterate(gsl_min_fminimizer *) { ORACLE_GSL_MIN_ABORT; return -1; }
inline double gsl_min_fminimizer_x_minimum(const gsl_min_fminimizer *) { ORACLE_GSL_MIN_ABORT; return 0; }
inline double gsl_min_fminimizer_x_upper(const gsl_min_fminimizer *) { ORACLE_GSL_MIN_ABORT; return 0; }
inline double gsl_min_fminimizer_x_lower(const gsl_min_fminimizer *) { ORACLE_GSL_MIN_ABORT; return 0; }
inline int gsl_min_test_interval(double, double, double, double) { ORACLE_GSL_MIN_ABORT; return -1; }
#endif
