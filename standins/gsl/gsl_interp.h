// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for GSL's <gsl/gsl_interp.h> + <gsl/gsl_spline.h>
// (GSL is not installed in this image; the reference does not pin a version -- distro libgsl-dev).
// Unlike the other stubs this one is a REAL implementation, because the reference evaluates
// natural cubic splines on its hot path (src/IMRPhenomD.cpp:1190-1216, src/IMRPhenomP.cpp:595-602).
// It restates GSL's published algorithm (interpolation/cspline.c, linalg/tridiag.c, interpolation/linear.c):
//   * natural boundary conditions c[0]=c[n-1]=0,
//   * symmetric tridiagonal system diag=2(h_i+h_{i+1}), offdiag=h_{i+1}, rhs=3(dy_{i+1}/h_{i+1}-dy_i/h_i),
//     solved by the L.D.L^T recurrence in the same operation order,
//   * per-interval b=dy/dx-dx(c_{i+1}+2c_i)/3, d=(c_{i+1}-c_i)/(3dx), Horner evaluation.
#ifndef ORACLE_STUB_GSL_INTERP_H
#define ORACLE_STUB_GSL_INTERP_H
#include <cstdlib>
#include <cmath>
#include <limits>
#include <vector>

struct gsl_interp_type { int kind; };  // 0 linear, 1 cspline
static const gsl_interp_type oracle_gsl_interp_linear_t = {0};
static const gsl_interp_type oracle_gsl_interp_cspline_t = {1};
static const gsl_interp_type *const gsl_interp_linear = &oracle_gsl_interp_linear_t;
static const gsl_interp_type *const gsl_interp_cspline = &oracle_gsl_interp_cspline_t;

struct gsl_interp_accel { size_t cache; };
inline gsl_interp_accel *gsl_interp_accel_alloc(void) { gsl_interp_accel *a = new gsl_interp_accel; a->cache = 0; return a; }
inline void gsl_interp_accel_free(gsl_interp_accel *a) { delete a; }
inline int gsl_interp_accel_reset(gsl_interp_accel *a) { a->cache = 0; return 0; }

struct gsl_spline {
	int kind;
	size_t size;
	std::vector<double> x, y, c;
};

inline gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
	gsl_spline *s = new gsl_spline;
	s->kind = T->kind;
	s->size = size;
	return s;
}
inline void gsl_spline_free(gsl_spline *s) { delete s; }

inline int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size)
{
	s->size = size;
	s->x.assign(xa, xa + size);
	s->y.assign(ya, ya + size);
	s->c.assign(size, 0.0);
	if (s->kind == 0 || size < 3) return 0;
	const size_t max_index = size - 1;
	const size_t N = max_index - 1;
	std::vector<double> g(N), diag(N), offdiag(N);
	for (size_t i = 0; i < N; i++) {
		const double h_i = xa[i + 1] - xa[i];
		const double h_ip1 = xa[i + 2] - xa[i + 1];
		const double ydiff_i = ya[i + 1] - ya[i];
		const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
		const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
		const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
		offdiag[i] = h_ip1;
		diag[i] = 2.0 * (h_ip1 + h_i);
		g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
	}
	double *sol = &s->c[1];
	if (N == 1) { sol[0] = g[0] / diag[0]; return 0; }
	std::vector<double> gamma(N), alpha(N), cc(N), z(N);
	alpha[0] = diag[0];
	gamma[0] = offdiag[0] / alpha[0];
	for (size_t i = 1; i < N - 1; i++) {
		alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
		gamma[i] = offdiag[i] / alpha[i];
	}
	if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
	z[0] = g[0];
	for (size_t i = 1; i < N; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
	for (size_t i = 0; i < N; i++) cc[i] = z[i] / alpha[i];
	sol[N - 1] = cc[N - 1];
	if (N >= 2) {
		for (size_t i = N - 2, j = 0; j <= N - 2; j++, i--) sol[i] = cc[i] - gamma[i] * sol[i + 1];
	}
	return 0;
}

inline size_t oracle_gsl_bsearch(const double *xa, double x, size_t lo, size_t hi)
{
	size_t ilo = lo, ihi = hi;
	while (ihi > ilo + 1) {
		size_t i = (ihi + ilo) / 2;
		if (xa[i] > x) ihi = i; else ilo = i;
	}
	return ilo;
}

inline int oracle_gsl_eval(const gsl_spline *s, double x, int order, double *out)
{
	const size_t n = s->size;
	const double *xa = &s->x[0], *ya = &s->y[0];
	if (x < xa[0] || x > xa[n - 1]) { *out = std::numeric_limits<double>::quiet_NaN(); return 1; }
	const size_t idx = oracle_gsl_bsearch(xa, x, 0, n - 1);
	const double x_lo = xa[idx], x_hi = xa[idx + 1];
	const double dx = x_hi - x_lo;
	if (!(dx > 0.0)) { *out = 0.0; return 1; }
	const double y_lo = ya[idx], y_hi = ya[idx + 1];
	const double dy = y_hi - y_lo;
	if (s->kind == 0) {
		if (order == 0) *out = y_lo + (x - x_lo) / dx * dy;
		else if (order == 1) *out = dy / dx;
		else *out = 0.0;
		return 0;
	}
	const double delx = x - x_lo;
	const double c_i = s->c[idx], c_ip1 = s->c[idx + 1];
	const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
	const double d_i = (c_ip1 - c_i) / (3.0 * dx);
	if (order == 0) *out = y_lo + delx * (b_i + delx * (c_i + delx * d_i));
	else if (order == 1) *out = b_i + delx * (2.0 * c_i + 3.0 * d_i * delx);
	else *out = 2.0 * c_i + 6.0 * d_i * delx;
	return 0;
}
inline double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *) { double y; oracle_gsl_eval(s, x, 0, &y); return y; }
inline double gsl_spline_eval_deriv(const gsl_spline *s, double x, gsl_interp_accel *) { double y; oracle_gsl_eval(s, x, 1, &y); return y; }
inline double gsl_spline_eval_deriv2(const gsl_spline *s, double x, gsl_interp_accel *) { double y; oracle_gsl_eval(s, x, 2, &y); return y; }
inline int gsl_spline_eval_e(const gsl_spline *s, double x, gsl_interp_accel *, double *y) { return oracle_gsl_eval(s, x, 0, y); }
#endif
