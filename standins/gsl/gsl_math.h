// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_math.h>.
#ifndef ORACLE_STUB_GSL_MATH_H
#define ORACLE_STUB_GSL_MATH_H
#include <cmath>
struct gsl_function { double (*function)(double x, void *params); void *params; };
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
inline int gsl_fcmp(const double x1, const double x2, const double epsilon)
{
	int exponent;
	const double max = (std::fabs(x1) > std::fabs(x2)) ? x1 : x2;
	std::frexp(max, &exponent);
	const double delta = std::ldexp(epsilon, exponent);
	const double difference = x1 - x2;
	if (difference > delta) return 1;
	if (difference < -delta) return -1;
	return 0;
}
#endif
