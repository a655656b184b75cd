// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for <fftw3.h>: the reference uses FFTW only for the tc/phic-maximised
// likelihoods (src/mcmc_gw.cpp:595-795), one complex 1-d transform of the grid length per call.  FFTW itself is not in this
// image, so the plan/execute entry points the reference calls are implemented here as a textbook transform (the discrete
// Fourier transform is defined by its formula, not by FFTW): iterative radix-2 Cooley-Tukey for power-of-two lengths, the
// O(n^2) sum otherwise.  sign = FFTW_FORWARD (-1): out[k] = sum_j in[j] exp(-2 pi i j k / n), unnormalised, like FFTW.
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef double fftw_complex[2];
struct oracle_fftw_plan_s {
	int n, sign;
	fftw_complex *in, *out;
};
typedef struct oracle_fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
inline void *fftw_malloc(size_t n) { return std::malloc(n); }
inline void fftw_free(void *p) { std::free(p); }
inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned)
{
	fftw_plan p = new oracle_fftw_plan_s;
	p->n = n;
	p->sign = sign;
	p->in = in;
	p->out = out;
	return p;
}
inline void oracle_dft(int n, int sign, fftw_complex *in, fftw_complex *out)
{
	typedef std::complex<long double> cld;
	const long double tau = 6.283185307179586476925286766559005768L;
	std::vector<cld> a(n);
	for (int i = 0; i < n; i++) a[i] = cld(in[i][0], in[i][1]);
	if (n > 0 && (n & (n - 1)) == 0) {
		for (int i = 1, j = 0; i < n; i++) {  // bit reversal
			int bit = n >> 1;
			for (; j & bit; bit >>= 1) j ^= bit;
			j ^= bit;
			if (i < j) std::swap(a[i], a[j]);
		}
		for (int len = 2; len <= n; len <<= 1) {
			for (int i = 0; i < n; i += len)
				for (int k = 0; k < len / 2; k++) {
					const long double ang = sign * tau * k / len;
					const cld w(cosl(ang), sinl(ang));
					const cld u = a[i + k], v = a[i + k + len / 2] * w;
					a[i + k] = u + v;
					a[i + k + len / 2] = u - v;
				}
		}
		for (int i = 0; i < n; i++) {
			out[i][0] = (double)a[i].real();
			out[i][1] = (double)a[i].imag();
		}
		return;
	}
	for (int k = 0; k < n; k++) {
		cld s = 0;
		for (int j = 0; j < n; j++) {
			const long double ang = sign * tau * (long double)(((long long)j * k) % n) / n;
			s += a[j] * cld(cosl(ang), sinl(ang));
		}
		out[k][0] = (double)s.real();
		out[k][1] = (double)s.imag();
	}
}
inline void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out) { oracle_dft(p->n, p->sign, in, out); }
inline void fftw_execute(const fftw_plan p) { oracle_dft(p->n, p->sign, p->in, p->out); }
inline void fftw_destroy_plan(fftw_plan p) { delete p; }
inline void fftw_cleanup(void) {}
#endif
