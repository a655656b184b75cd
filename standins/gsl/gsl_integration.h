// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_integration.h>: adaptive quadrature is off-path.
#ifndef ORACLE_STUB_GSL_INTEGRATION_H
#define ORACLE_STUB_GSL_INTEGRATION_H
#include <cstdio>
#include <cstdlib>
#include "gsl/gsl_math.h"
struct gsl_integration_workspace { size_t limit; };
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3, GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };
inline gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) { gsl_integration_workspace *w = new gsl_integration_workspace; w->limit = n; return w; }
inline void gsl_integration_workspace_free(gsl_integration_workspace *w) { delete w; }
inline int gsl_integration_qag(const gsl_function *, double, double, double, double, size_t, int, gsl_integration_workspace *, double *, double *)
{ std::fprintf(stderr, "oracle stub: gsl_integration_qag called (off-path)\n"); std::abort(); return -1; }
inline int gsl_integration_qags(const gsl_function *, double, double, double, double, size_t, gsl_integration_workspace *, double *, double *)
{ std::fprintf(stderr, "oracle stub: gsl_integration_qags called (off-path)\n"); std::abort(); return -1; }
#endif
