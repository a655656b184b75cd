// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for ADOL-C drivers: autodiff never runs in the oracle.
#ifndef ORACLE_STUB_ADOLC_DRIVERS_H
#define ORACLE_STUB_ADOLC_DRIVERS_H
#include <cstdlib>
#include <cstdio>
#include "adolc/taping.h"
#define ORACLE_ADOLC_ABORT(name) do { std::fprintf(stderr, "oracle stub: ADOL-C %s called (autodiff is off-path)\n", name); std::abort(); } while (0)
inline int gradient(short, int, const double *, double *) { ORACLE_ADOLC_ABORT("gradient"); return -1; }
inline int jacobian(short, int, int, const double *, double **) { ORACLE_ADOLC_ABORT("jacobian"); return -1; }
inline int hessian(short, int, const double *, double **) { ORACLE_ADOLC_ABORT("hessian"); return -1; }
inline int function(short, int, int, double *, double *) { ORACLE_ADOLC_ABORT("function"); return -1; }
inline int binomi(int, int) { ORACLE_ADOLC_ABORT("binomi"); return -1; }
inline int tensor_eval(short, int, int, int, int, double *, double **, double **) { ORACLE_ADOLC_ABORT("tensor_eval"); return -1; }
inline int tensor_address(int, int *) { ORACLE_ADOLC_ABORT("tensor_address"); return -1; }
#endif
