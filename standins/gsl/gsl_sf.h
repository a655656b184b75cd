// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_sf.h>: elliptic functions are only used by IMRPhenomPv3 (off-path).
#ifndef ORACLE_STUB_GSL_SF_H
#define ORACLE_STUB_GSL_SF_H
#include <cstdio>
#include <cstdlib>
typedef unsigned int gsl_mode_t;
#define GSL_PREC_DOUBLE 0
#define GSL_PREC_SINGLE 1
#define GSL_PREC_APPROX 2
inline int gsl_sf_elljac_e(double, double, double *, double *, double *) { std::fprintf(stderr, "oracle stub: gsl_sf_elljac_e called (off-path)\n"); std::abort(); return -1; }
inline double gsl_sf_ellint_F(double, double, gsl_mode_t) { std::fprintf(stderr, "oracle stub: gsl_sf_ellint_F called (off-path)\n"); std::abort(); return 0; }
#endif
