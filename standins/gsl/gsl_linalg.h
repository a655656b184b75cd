// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_linalg.h>: matrix inversion is off-path (Fisher proposals), abort if reached.
#ifndef ORACLE_STUB_GSL_LINALG_H
#define ORACLE_STUB_GSL_LINALG_H
#include <cstdio>
#include "gsl/gsl_matrix_double.h"
#define ORACLE_GSL_LA_ABORT(n) do { std::fprintf(stderr, "oracle stub: %s called (off-path)\n", n); std::abort(); } while (0)
inline int gsl_linalg_LU_decomp(gsl_matrix *, gsl_permutation *, int *) { ORACLE_GSL_LA_ABORT("gsl_linalg_LU_decomp"); return -1; }
inline int gsl_linalg_LU_invert(const gsl_matrix *, const gsl_permutation *, gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_LU_invert"); return -1; }
inline double gsl_linalg_LU_lndet(gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_LU_lndet"); return 0; }
inline int gsl_linalg_cholesky_decomp(gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_cholesky_decomp"); return -1; }
inline int gsl_linalg_cholesky_decomp1(gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_cholesky_decomp1"); return -1; }
inline int gsl_linalg_cholesky_invert(gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_cholesky_invert"); return -1; }
inline int gsl_linalg_pcholesky_decomp(gsl_matrix *, gsl_permutation *) { ORACLE_GSL_LA_ABORT("gsl_linalg_pcholesky_decomp"); return -1; }
inline int gsl_linalg_pcholesky_invert(const gsl_matrix *, const gsl_permutation *, gsl_matrix *) { ORACLE_GSL_LA_ABORT("gsl_linalg_pcholesky_invert"); return -1; }
#endif
