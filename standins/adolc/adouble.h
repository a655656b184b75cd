// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for the third-party
// ADOL-C header <adolc/adouble.h>, which is not installed in this image.
// The reference templates its waveform code on T in {double, adouble}; the
// oracle only ever *executes* the T=double instantiations, but the adouble
// ones must still compile.  This is a passive wrapper around double: no
// taping, no derivatives.
#ifndef ORACLE_STUB_ADOUBLE_H
#define ORACLE_STUB_ADOUBLE_H
#include <cmath>
#include <math.h>  // must precede the adouble overloads below: its using-declarations would otherwise clash with them
#include <iostream>
#include <type_traits>
#include <omp.h>

class adouble {
public:
	double v;
	adouble() : v(0.) {}
	template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>
	adouble(U x) : v((double)x) {}
	double value() const { return v; }
	double getValue() const { return v; }
	void setValue(double x) { v = x; }
	adouble &operator<<=(double x) { v = x; return *this; }
	adouble &operator>>=(double &x) { x = v; return *this; }
	template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>
	adouble &operator=(U x) { v = (double)x; return *this; }
	adouble &operator+=(const adouble &o) { v += o.v; return *this; }
	adouble &operator-=(const adouble &o) { v -= o.v; return *this; }
	adouble &operator*=(const adouble &o) { v *= o.v; return *this; }
	adouble &operator/=(const adouble &o) { v /= o.v; return *this; }
	adouble operator-() const { return adouble(-v); }
	adouble operator+() const { return *this; }
	adouble operator++(int) { adouble t(*this); v += 1; return t; }
	adouble &operator++() { v += 1; return *this; }
	explicit operator double() const { return v; }
};

#define ORACLE_AD_ARITH typename std::enable_if<std::is_arithmetic<U>::value, int>::type = 0
#define ORACLE_AD_BINOP(op)                                                                      \
	inline adouble operator op(const adouble &a, const adouble &b) { return adouble(a.v op b.v); } \
	template <class U, ORACLE_AD_ARITH>                                                            \
	inline adouble operator op(const adouble &a, U b) { return adouble(a.v op(double) b); }        \
	template <class U, ORACLE_AD_ARITH>                                                            \
	inline adouble operator op(U a, const adouble &b) { return adouble((double)a op b.v); }
ORACLE_AD_BINOP(+)
ORACLE_AD_BINOP(-)
ORACLE_AD_BINOP(*)
ORACLE_AD_BINOP(/)
#undef ORACLE_AD_BINOP
#define ORACLE_AD_CMP(op)                                                              \
	inline bool operator op(const adouble &a, const adouble &b) { return a.v op b.v; } \
	template <class U, ORACLE_AD_ARITH>                                                  \
	inline bool operator op(const adouble &a, U b) { return a.v op(double) b; }          \
	template <class U, ORACLE_AD_ARITH>                                                  \
	inline bool operator op(U a, const adouble &b) { return (double)a op b.v; }
ORACLE_AD_CMP(<)
ORACLE_AD_CMP(>)
ORACLE_AD_CMP(<=)
ORACLE_AD_CMP(>=)
ORACLE_AD_CMP(==)
ORACLE_AD_CMP(!=)
#undef ORACLE_AD_CMP

#define ORACLE_AD_FN1(fn) \
	inline adouble fn(const adouble &a) { return adouble(std::fn(a.v)); }
ORACLE_AD_FN1(sqrt)
ORACLE_AD_FN1(exp)
ORACLE_AD_FN1(log)
ORACLE_AD_FN1(log10)
ORACLE_AD_FN1(sin)
ORACLE_AD_FN1(cos)
ORACLE_AD_FN1(tan)
ORACLE_AD_FN1(asin)
ORACLE_AD_FN1(acos)
ORACLE_AD_FN1(atan)
ORACLE_AD_FN1(sinh)
ORACLE_AD_FN1(cosh)
ORACLE_AD_FN1(tanh)
ORACLE_AD_FN1(fabs)
ORACLE_AD_FN1(floor)
ORACLE_AD_FN1(ceil)
ORACLE_AD_FN1(cbrt)
#undef ORACLE_AD_FN1
inline adouble abs(const adouble &a) { return adouble(std::fabs(a.v)); }
inline bool isnan(const adouble &a) { return std::isnan(a.v); }
inline bool isinf(const adouble &a) { return std::isinf(a.v); }
inline adouble pow(const adouble &a, const adouble &b) { return adouble(std::pow(a.v, b.v)); }
template <class U, ORACLE_AD_ARITH>
inline adouble pow(const adouble &a, U b) { return adouble(std::pow(a.v, (double)b)); }
template <class U, ORACLE_AD_ARITH>
inline adouble pow(U a, const adouble &b) { return adouble(std::pow((double)a, b.v)); }
inline adouble atan2(const adouble &a, const adouble &b) { return adouble(std::atan2(a.v, b.v)); }
template <class U, ORACLE_AD_ARITH>
inline adouble atan2(const adouble &a, U b) { return adouble(std::atan2(a.v, (double)b)); }
template <class U, ORACLE_AD_ARITH>
inline adouble atan2(U a, const adouble &b) { return adouble(std::atan2((double)a, b.v)); }
inline adouble fmax(const adouble &a, const adouble &b) { return adouble(std::fmax(a.v, b.v)); }
inline adouble fmin(const adouble &a, const adouble &b) { return adouble(std::fmin(a.v, b.v)); }
template <class U, ORACLE_AD_ARITH>
inline adouble fmax(const adouble &a, U b) { return adouble(std::fmax(a.v, (double)b)); }
template <class U, ORACLE_AD_ARITH>
inline adouble fmin(const adouble &a, U b) { return adouble(std::fmin(a.v, (double)b)); }
template <class U, ORACLE_AD_ARITH>
inline adouble fmax(U a, const adouble &b) { return adouble(std::fmax((double)a, b.v)); }
template <class U, ORACLE_AD_ARITH>
inline adouble fmin(U a, const adouble &b) { return adouble(std::fmin((double)a, b.v)); }
inline std::ostream &operator<<(std::ostream &o, const adouble &a) { return o << a.v; }
inline std::istream &operator>>(std::istream &i, adouble &a) { return i >> a.v; }
#undef ORACLE_AD_ARITH

namespace std {
inline bool isnan(const adouble &a) { return std::isnan(a.v); }
}  // namespace std
#endif
