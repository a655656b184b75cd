// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_matrix_double.h> + permutation (off-path linear algebra).
#ifndef ORACLE_STUB_GSL_MATRIX_H
#define ORACLE_STUB_GSL_MATRIX_H
#include <cstdlib>
#include <vector>
struct gsl_matrix { size_t size1, size2; std::vector<double> data; };
struct gsl_permutation { size_t size; std::vector<size_t> data; };
struct gsl_vector { size_t size; std::vector<double> data; };
inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) { gsl_matrix *m = new gsl_matrix; m->size1 = n1; m->size2 = n2; m->data.assign(n1 * n2, 0.); return m; }
inline void gsl_matrix_free(gsl_matrix *m) { delete m; }
inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
inline gsl_permutation *gsl_permutation_alloc(size_t n) { gsl_permutation *p = new gsl_permutation; p->size = n; p->data.assign(n, 0); return p; }
inline void gsl_permutation_free(gsl_permutation *p) { delete p; }
#endif
