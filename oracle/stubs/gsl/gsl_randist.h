// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_randist.h> (off-path).
#ifndef ORACLE_STUB_GSL_RANDIST_H
#define ORACLE_STUB_GSL_RANDIST_H
#include "gsl/gsl_rng.h"
inline double gsl_ran_gaussian(gsl_rng *r, double sigma) { return std::normal_distribution<double>(0., sigma)(r->eng); }
inline double gsl_ran_flat(gsl_rng *r, double a, double b) { return std::uniform_real_distribution<double>(a, b)(r->eng); }
inline double gsl_ran_gaussian_pdf(double x, double sigma) { return std::exp(-x * x / (2 * sigma * sigma)) / (std::sqrt(2 * M_PI) * sigma); }
#endif
