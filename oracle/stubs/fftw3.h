// TEST INFRASTRUCTURE ONLY (oracle build). Type-only stand-in for <fftw3.h>.
// FFTW is used by the reference only for tc/phic-maximised likelihoods, which are off the path.
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <cstdlib>
#include <cstdio>
typedef double fftw_complex[2];
typedef struct oracle_fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
inline void *fftw_malloc(size_t n) { return std::malloc(n); }
inline void fftw_free(void *p) { std::free(p); }
inline fftw_plan fftw_plan_dft_1d(int, fftw_complex *, fftw_complex *, int, unsigned) { return (fftw_plan)0; }
inline void fftw_execute(const fftw_plan) { std::fprintf(stderr, "oracle stub: fftw_execute called (off-path)\n"); std::abort(); }
inline void fftw_execute_dft(const fftw_plan, fftw_complex *, fftw_complex *) { std::fprintf(stderr, "oracle stub: fftw_execute_dft called (off-path)\n"); std::abort(); }
inline void fftw_destroy_plan(fftw_plan) {}
inline void fftw_cleanup(void) {}
#endif
