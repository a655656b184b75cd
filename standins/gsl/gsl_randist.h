// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_randist.h> (see gsl_rng.h for the scripted mode).
#ifndef ORACLE_STUB_GSL_RANDIST_H
#define ORACLE_STUB_GSL_RANDIST_H
#include <cmath>
#include "gsl/gsl_rng.h"
inline double gsl_ran_gaussian(gsl_rng *r, double sigma)
{
	if (r->scripted) return sigma * oracle_gsl_scripted_normal(r);
	return std::normal_distribution<double>(0., sigma)(r->eng);
}
inline double gsl_ran_flat(gsl_rng *r, double a, double b) { return std::uniform_real_distribution<double>(a, b)(r->eng); }
inline double gsl_ran_gaussian_pdf(double x, double sigma) { return std::exp(-x * x / (2 * sigma * sigma)) / (std::sqrt(2 * M_PI) * sigma); }
#endif
