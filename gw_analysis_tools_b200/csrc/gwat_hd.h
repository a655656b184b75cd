// Host/device portability layer and exact-arithmetic helpers for the gwat_b200 kernels.
//
// All waveform mathematics lives in headers of `GWAT_HD` functions so that the very same source is compiled
//   * by nvcc for sm_100a (the product: kernels in gwat_engine.cu), and
//   * by g++ as plain C++ for tests/host_harness.cpp (TEST ONLY: lets the CPU-only test tier check the kernels' logic
//     against the oracle without a GPU; it is not reachable from the C ABI).
#ifndef GWAT_HD_H
#define GWAT_HD_H

#include <math.h>

#if defined(__CUDACC__)
#define GWAT_HD __host__ __device__ __forceinline__
#define GWAT_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define GWAT_HD inline
#define GWAT_HD_NOINLINE inline
#endif

#define GWAT_PI 3.14159265358979323846
#define GWAT_TWOPI 6.283185307179586476925286766559005768

namespace gwat {

// ---- arithmetic that must not be contracted into FMAs ---------------------------------------------------------------
// The reference is built by g++ -O2 for baseline x86-64: every a*b+c there rounds twice.  Where the phase is large
// (1e3..1e6 rad) a fused multiply-add changes the result by up to half an ulp of the big term, so the phase chain uses
// these helpers, which nvcc is not allowed to contract.  Amplitudes use ordinary operators (FMA contraction welcome).
GWAT_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
GWAT_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
GWAT_HD double sub_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}
GWAT_HD double fma_rn(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
	return __fma_rn(a, b, c);
#else
	return fma(a, b, c);
#endif
}

// ---- libm for the per-walker setup code ---------------------------------------------------------------------------------
// The setup kernels are long straight-line code (IMRPhenomPv2: 19 k SASS instructions = 300 KB, executed once per warp) and
// were bound by instruction fetch, not arithmetic: ncu attributed 62 % of their stall samples to "no instruction".  Most
// of that footprint is CUDA's libm inlined at every call site (pow ~300 instructions, log/exp/sincos/cbrt/acos/atan2
// 80-150 each, ~150 call sites).  Setup code calls them through these out-of-line copies instead: one body per function,
// resident in the instruction cache after its first use.  The per-bin code keeps its inlined fast_* versions.
// For the same reason loops of the setup code whose bodies are large are kept rolled (GWAT_SETUP_LOOP before the `for`): the
// body is fetched once and then runs out of the instruction cache.  -DGWAT_SETUP_UNROLLED restores full unrolling (A/B runs).
#if defined(__CUDACC__) && !defined(GWAT_SETUP_UNROLLED)
#define GWAT_SETUP_LOOP _Pragma("unroll 1")
#else
#define GWAT_SETUP_LOOP
#endif
namespace sm {
GWAT_HD_NOINLINE double pow(double a, double b) { return ::pow(a, b); }
GWAT_HD_NOINLINE double log(double a) { return ::log(a); }
GWAT_HD_NOINLINE double exp(double a) { return ::exp(a); }
GWAT_HD_NOINLINE double cbrt(double a) { return ::cbrt(a); }
GWAT_HD_NOINLINE double sin(double a) { return ::sin(a); }
GWAT_HD_NOINLINE double cos(double a) { return ::cos(a); }
GWAT_HD_NOINLINE void sincos(double a, double *s, double *c) { ::sincos(a, s, c); }
GWAT_HD_NOINLINE double acos(double a) { return ::acos(a); }
GWAT_HD_NOINLINE double asin(double a) { return ::asin(a); }
GWAT_HD_NOINLINE double atan2(double y, double x) { return ::atan2(y, x); }
GWAT_HD_NOINLINE double tan(double a) { return ::tan(a); }
}  // namespace sm

// ---- double-double (unevaluated sum hi+lo) --------------------------------------------------------------------------
struct dd {
	double hi, lo;
};

GWAT_HD dd two_sum(double a, double b)
{
	double s = add_rn(a, b);
	double bb = sub_rn(s, a);
	double e = add_rn(sub_rn(a, sub_rn(s, bb)), sub_rn(b, bb));
	return dd{s, e};
}
GWAT_HD dd quick_two_sum(double a, double b)
{
	double s = add_rn(a, b);
	double e = sub_rn(b, sub_rn(s, a));
	return dd{s, e};
}
GWAT_HD dd two_prod(double a, double b)
{
	double p = mul_rn(a, b);
	double e = fma_rn(a, b, -p);
	return dd{p, e};
}
GWAT_HD dd dd_mul(dd a, dd b)
{
	dd p = two_prod(a.hi, b.hi);
	double e = add_rn(p.lo, add_rn(mul_rn(a.hi, b.lo), mul_rn(a.lo, b.hi)));
	return quick_two_sum(p.hi, e);
}
GWAT_HD dd dd_mul_d(dd a, double b)
{
	dd p = two_prod(a.hi, b);
	double e = add_rn(p.lo, mul_rn(a.lo, b));
	return quick_two_sum(p.hi, e);
}
GWAT_HD dd dd_add(dd a, dd b)
{
	dd s = two_sum(a.hi, b.hi);
	double e = add_rn(s.lo, add_rn(a.lo, b.lo));
	return quick_two_sum(s.hi, e);
}
// correctly rounded (to within a few 1e-32 relative) product of two double-doubles, returned as one double
GWAT_HD double dd_mul_to_double(double ahi, double alo, double bhi, double blo)
{
	double p = mul_rn(ahi, bhi);
	double e = fma_rn(ahi, bhi, -p);
	e = fma_rn(ahi, blo, e);
	e = fma_rn(alo, bhi, e);
	return add_rn(p, e);
}

// The reference raises (M*f) to the power 1./6. -- the DOUBLE nearest to 1/6, not 1/6 itself (src/IMRPhenomD.cpp:882).
// x^(fl(1/6)) = x^(1/6) * exp(-GWAT_SIXTH_DEFECT * ln x), a shift of up to ~0.7 ulp that has to be reproduced.
#define GWAT_SIXTH 0.16666666666666666
#define GWAT_SIXTH_DEFECT 9.2518585385429707e-18 /* 1/6 - fl(1/6) */

// x^(fl(1/6)) as a double-double, accurate to ~1e-30 relative, for x > 0.
GWAT_HD_NOINLINE dd pow_sixth_dd(double x)
{
	// Newton iteration on y^6 = x in double-double, from the double estimate y0 = sqrt(cbrt(x)).
	// y0 is good to ~2 ulp, one quadratic step in double-double takes it to ~1e-31
	double y0 = sqrt(sm::cbrt(x));
	dd y = dd{y0, 0.0};
	for (int it = 0; it < 1; it++) {
		dd y2 = dd_mul(y, y);
		dd y3 = dd_mul(y2, y);
		dd y6 = dd_mul(y3, y3);
		// residual r = y^6 - x  (x exact)
		dd r = dd_add(y6, dd{-x, 0.0});
		// y <- y - r / (6 y^5);   y^5 = y^6 / y  ~ y6.hi / y.hi  (double accuracy suffices for the correction)
		double corr = (r.hi + r.lo) / (6.0 * (y6.hi / y.hi));
		y = dd_add(y, dd{-corr, 0.0});
	}
	// exponent defect: multiply by (1 - defect*ln x)
	double shift = -GWAT_SIXTH_DEFECT * sm::log(x);
	dd t = dd_mul_d(y, shift);
	return dd_add(y, t);
}

// Reciprocal / reciprocal square root to ~1 ulp without the IEEE-exact routines' special-case paths.  Used for amplitudes
// and for phase terms of O(1..100) rad, where one ulp is < 1e-14 rad; the leading TaylorF2 term keeps the exact division.
GWAT_HD double fast_rcp(double x)
{
#if defined(__CUDA_ARCH__)
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	// the hardware seed is good to 2^-23; each Newton step squares the error
	double e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
#else
	return 1.0 / x;
#endif
}
GWAT_HD double fast_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
	double r;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	const double h = 0.5 * x;
	r = r * fma(-h * r, r, 1.5);
	r = r * fma(-h * r, r, 1.5);
	return r;
#else
	return 1.0 / sqrt(x);
#endif
}
GWAT_HD double fast_sqrt(double x) { return x * fast_rsqrt(x); }

// ---- trigonometry for the per-bin code -----------------------------------------------------------------------------------
// CUDA's sincos()/atan() spend half of their ~80 instructions materialising polynomial coefficients as immediates (two UMOVs
// per double) and guarding a slow path for |x| > 1e5.  The per-bin code evaluates two sincos and one atan per bin (a third of
// its instructions), so it uses these instead: the same algorithms -- three-term Cody-Waite reduction by pi/2 and the
// fdlibm kernel polynomials; reciprocal fold and a degree-20 polynomial in x^2 for atan -- with the coefficients in constant
// memory, where one LDCU.128 fetches two of them.  Accuracy ~1 ulp for |x| < 2^31 (sincos) and all x (atan); NaN in, NaN out.
// On the host (test harness) they are the libm functions.
#if defined(__CUDACC__)
static __constant__ double gwat_trig_k[40] = {
    /* 0 */ 0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,
    /* 4: sin */ -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
    -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    /* 10: cos */ 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
    2.08757232129817482790e-09, -1.13596475577881948265e-11,
    /* 16: atan(x)/x in x^2 on [0,1], Chebyshev-node interpolant, max error 2.6e-17 */ 1.0, -0.3333333333333286, 0.19999999999929946,
    -0.14285714281592693, 0.11111110982087126, -0.09090906605656898, 0.0769227555520563, -0.06666371187721098, 0.05880342002401543,
    -0.052527255573225747, 0.04719723992321112, -0.042125723963855326, 0.03651081352721035, -0.02970071773623423, 0.021740213830758134,
    -0.013674139288399478, 0.007038646202989813, -0.0028047655531701315, 0.0008033604181626027, -0.00014617088163625013,
    1.2631178430477426e-05,
    /* 37 */ 1.5707963267948966, 6.123233995736766e-17, 0.0};
#endif

GWAT_HD void fast_sincos(double x, double *s, double *c)
{
#if defined(__CUDA_ARCH__)
	const double *k = gwat_trig_k;
	const int n = __double2int_rn(x * k[0]);
	const double q = (double)n;
	double r = fma(-q, k[1], x);
	r = fma(-q, k[2], r);
	r = fma(-q, k[3], r);
	const double z = r * r;
	double ps = fma(z, k[9], k[8]);
	ps = fma(z, ps, k[7]);
	ps = fma(z, ps, k[6]);
	ps = fma(z, ps, k[5]);
	ps = fma(z, ps, k[4]);
	const double sn = fma(r * z, ps, r);
	double pc = fma(z, k[15], k[14]);
	pc = fma(z, pc, k[13]);
	pc = fma(z, pc, k[12]);
	pc = fma(z, pc, k[11]);
	pc = fma(z, pc, k[10]);
	const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
	const double a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;
	*s = (n & 2) ? -a : a;
	*c = ((n + 1) & 2) ? -b : b;
#else
	sincos(x, s, c);
#endif
}
GWAT_HD double fast_atan(double x)
{
#if defined(__CUDA_ARCH__)
	const double *k = gwat_trig_k;
	const double ax = fabs(x);
	const bool big = ax > 1.0;
	const double t = big ? fast_rcp(ax) : ax;
	const double z = t * t;
	double p = k[36];
#pragma unroll
	for (int j = 35; j >= 16; j--) p = fma(z, p, k[j]);
	double v = t * p;
	if (big) v = (k[37] - v) + k[38];
	return copysign(v, x);
#else
	return atan(x);
#endif
}

GWAT_HD double sq(double x) { return x * x; }
GWAT_HD double cube(double x) { return x * x * x; }

}  // namespace gwat
#endif
