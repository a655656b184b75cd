// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_sf_gamma.h>: included by the reference's tests/src/test_mcmc.cpp, whose
// only uses of it are commented out.
#ifndef ORACLE_STUB_GSL_SF_GAMMA_H
#define ORACLE_STUB_GSL_SF_GAMMA_H
#include <cmath>
inline double gsl_sf_gamma(double x) { return std::tgamma(x); }
#endif
