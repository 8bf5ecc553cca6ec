// fpmath.h — one implementation of the non-IEEE-basic functions on the kino-A* path,
// compiled for BOTH host (g++ -ffp-contract=off) and device (nvcc -fmad=false).
//
// Why this file exists.  The reference search (kino_astar.cpp:339-372 `cubic`,
// :416-471 `computeShotTraj`, minimum_control.cpp:5-96 `pow`) calls glibc's cbrt / acos /
// cos / pow.  CUDA's libdevice versions of those differ from glibc in the last ulp, and a
// last-ulp difference in an f-cost can reorder a heap pop, which breaks "identical
// expanded-node sets".  IEEE-754 +,-,*,/,sqrt,floor are correctly rounded on both sides,
// so everything below is built ONLY from those (plus integer bit moves) and therefore
// gives bit-identical results on x86-64 and sm_100a.  Accuracy is ~1 ulp, i.e. the same
// class as glibc; tests/test_fpmath.py measures the ulp distance to the container's glibc.
//
// Polynomial coefficients are exact rationals (1/n!, (2n)!/(4^n n!^2 (2n+1))) rounded to
// nearest and written as hex floats, so no table was borrowed from any libm.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#else
#define FP_HD static inline
#endif

namespace fpm {

FP_HD uint64_t to_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
FP_HD double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}

// ---- error-free transforms (no FMA anywhere: Veltkamp/Dekker) ------------------------
FP_HD void split(double a, double& hi, double& lo) {
  const double C = 134217729.0;  // 2^27 + 1
  double t = C * a;
  hi = t - (t - a);
  lo = a - hi;
}
FP_HD void two_prod(double a, double b, double& p, double& e) {
  p = a * b;
  double ah, al, bh, bl;
  split(a, ah, al);
  split(b, bh, bl);
  e = ((ah * bh - p) + ah * bl + al * bh) + al * bl;
}
FP_HD void fast_two_sum(double a, double b, double& s, double& e) {  // |a| >= |b|
  s = a + b;
  e = b - (s - a);
}

// x^n for small non-negative integer n, correctly rounded except in astronomically rare
// half-way cases (double-double accumulation, ~2^-100 relative error before the final
// rounding).  Stands in for glibc pow(t, i) with integer i (kino_astar.cpp:443,
// minimum_control.cpp:10-17), which is itself correctly rounded in all but such cases.
FP_HD double powi(double x, int n) {
  if (n == 0) return 1.0;
  double hi = x, lo = 0.0;
  for (int k = 1; k < n; ++k) {
    double p, e;
    two_prod(hi, x, p, e);
    e = e + lo * x;
    fast_two_sum(p, e, hi, lo);
  }
  return hi;
}

// ---- cbrt -----------------------------------------------------------------------------
FP_HD double cbrt(double x) {
  if (x == 0.0 || x != x) return x;
  uint64_t ux = to_bits(x);
  uint64_t sign = ux & 0x8000000000000000ull;
  uint64_t ua = ux & 0x7fffffffffffffffull;
  if (ua >= 0x7ff0000000000000ull) return x;  // inf
  int eadj = 0;
  if (ua < 0x0010000000000000ull) {  // subnormal: scale by 2^54 (exact)
    double s = from_bits(ua) * 18014398509481984.0;
    ua = to_bits(s);
    eadj = -18;  // cbrt(2^54) = 2^18
  }
  int e = (int)(ua >> 52) - 1023;
  // e = 3q + r, r in {0,1,2}, floor division
  int q = (e >= 0) ? (e / 3) : -((2 - e) / 3);
  int r = e - 3 * q;
  // m in [1, 8)
  double m = from_bits((ua & 0x000fffffffffffffull) | ((uint64_t)(1023 + r) << 52));
  // chord initial guess, <= ~1.6 % off
  double y;
  if (r == 0)      y = 1.0 + (m - 1.0) * 0.2599210498948732;
  else if (r == 1) y = 1.2599210498948732 + (m - 2.0) * 0.16374002795240733;
  else             y = 1.5874010519681994 + (m - 4.0) * 0.10314973700795015;
  // Halley: y <- y (y^3 + 2m) / (2 y^3 + m), cubic convergence
  for (int it = 0; it < 3; ++it) {
    double y3 = y * y * y;
    y = y * ((y3 + (m + m)) / ((y3 + y3) + m));
  }
  // one Newton correction with the residual m - y^3 evaluated in double-double
  double y2h, y2l, y3h, y3l;
  two_prod(y, y, y2h, y2l);
  two_prod(y2h, y, y3h, y3l);
  y3l = y3l + y2l * y;
  double res = (m - y3h) - y3l;
  y = y + res / (3.0 * y2h);
  // scale by 2^(q + eadj): y in [1,2) so adding to the exponent field is exact
  uint64_t uy = to_bits(y);
  uy = uy + ((uint64_t)(int64_t)(q + eadj) << 52);
  return from_bits(uy | sign);
}

// ---- sin / cos kernels on |r| <= pi/4 (+ tiny), r = rh + rl ------------------------------
FP_HD double sin_kernel(double rh, double rl) {
  double z = rh * rh;
  double p = 0x1.71b8ef6dcf572p-66;             // 1/21!
  p = -0x1.2f49b46814157p-57 + z * p;            // -1/19!
  p = 0x1.952c77030ad4ap-49 + z * p;             // 1/17!
  p = -0x1.ae7f3e733b81fp-41 + z * p;            // -1/15!
  p = 0x1.6124613a86d09p-33 + z * p;             // 1/13!
  p = -0x1.ae64567f544e4p-26 + z * p;            // -1/11!
  p = 0x1.71de3a556c734p-19 + z * p;             // 1/9!
  p = -0x1.a01a01a01a01ap-13 + z * p;            // -1/7!
  p = 0x1.1111111111111p-7 + z * p;              // 1/5!
  p = -0x1.5555555555555p-3 + z * p;             // -1/3!
  double w = rh * z;
  return rh + (w * p + rl);
}
FP_HD double cos_kernel(double rh, double rl) {
  double z = rh * rh;
  double p = 0x1.f2cf01972f578p-80;             // 1/24!
  p = -0x1.0ce396db7f853p-70 + z * p;            // -1/22!
  p = 0x1.e542ba4020225p-62 + z * p;             // 1/20!
  p = -0x1.6827863b97d97p-53 + z * p;            // -1/18!
  p = 0x1.ae7f3e733b81fp-45 + z * p;             // 1/16!
  p = -0x1.93974a8c07c9dp-37 + z * p;            // -1/14!
  p = 0x1.1eed8eff8d898p-29 + z * p;             // 1/12!
  p = -0x1.27e4fb7789f5cp-22 + z * p;            // -1/10!
  p = 0x1.a01a01a01a01ap-16 + z * p;             // 1/8!
  p = -0x1.6c16c16c16c17p-10 + z * p;            // -1/6!
  p = 0x1.5555555555555p-5 + z * p;              // 1/4!
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + ((z * z) * p - rh * rl));
}

// cos(x) for |x| < ~1e6 (the search only feeds it (theta + 2k pi)/3, theta in [0, pi]).
FP_HD double cos(double x) {
  if (x != x) return x;
  const double two_over_pi = 0x1.45f306dc9c883p-1;
  const double p1 = 0x1.921fb54400000p+0;   // pi/2 split in 33-bit pieces: n*p1, n*p2 exact
  const double p2 = 0x1.0b4611a600000p-34;
  const double p3 = 0x1.3198a2e037073p-69;
  double fn = ::floor(x * two_over_pi + 0.5);
  double a = (x - fn * p1) - fn * p2;
  double w = fn * p3;
  double rh = a - w;
  double rl = (a - rh) - w;
  int n = (int)((long long)fn & 3);
  if (n == 0) return cos_kernel(rh, rl);
  if (n == 1) return -sin_kernel(rh, rl);
  if (n == 2) return -cos_kernel(rh, rl);
  return sin_kernel(rh, rl);
}

// ---- acos -----------------------------------------------------------------------------
// R(z) = (asin(s) - s) / (s z) with z = s^2 <= 1/4: Maclaurin series, 26 exact-rational terms.
FP_HD double asin_R(double z) {
  double p = 0x1.1052bc5fa960ap-9;
  p = 0x1.208d3570ae5a6p-9 + z * p;
  p = 0x1.3275586c5f2f0p-9 + z * p;
  p = 0x1.464c0950f7d47p-9 + z * p;
  p = 0x1.5c5f56efaaaabp-9 + z * p;
  p = 0x1.750de64d7d05fp-9 + z * p;
  p = 0x1.90cb77f60c7cep-9 + z * p;
  p = 0x1.b026f57b13b14p-9 + z * p;
  p = 0x1.d3d2a8e0dd67dp-9 + z * p;
  p = 0x1.fcaf8fb6db6dbp-9 + z * p;
  p = 0x1.15ee9d45d1746p-8 + z * p;
  p = 0x1.31683bdef7bdfp-8 + z * p;
  p = 0x1.51ba308d3dcb1p-8 + z * p;
  p = 0x1.782dda12f684cp-8 + z * p;
  p = 0x1.a6863d70a3d71p-8 + z * p;
  p = 0x1.df3bd37a6f4dfp-8 + z * p;
  p = 0x1.12ef3cf3cf3cfp-7 + z * p;
  p = 0x1.3fde50d79435ep-7 + z * p;
  p = 0x1.7a87878787878p-7 + z * p;
  p = 0x1.c99999999999ap-7 + z * p;
  p = 0x1.1c4ec4ec4ec4fp-6 + z * p;
  p = 0x1.6e8ba2e8ba2e9p-6 + z * p;
  p = 0x1.f1c71c71c71c7p-6 + z * p;
  p = 0x1.6db6db6db6db7p-5 + z * p;
  p = 0x1.3333333333333p-4 + z * p;
  p = 0x1.5555555555555p-3 + z * p;
  return p;
}

FP_HD double acos(double x) {
  const double pio2_hi = 0x1.921fb54442d18p+0, pio2_lo = 0x1.1a62633145c07p-54;
  const double pi_hi = 0x1.921fb54442d18p+1, pi_lo = 0x1.1a62633145c07p-53;
  if (x != x) return x;
  double ax = x < 0.0 ? -x : x;
  if (ax > 1.0) return from_bits(0x7ff8000000000000ull);  // NaN, as glibc
  if (ax == 1.0) return x > 0.0 ? 0.0 : pi_hi + pi_lo;
  if (ax < 0.5) {
    if (ax < 0x1p-57) return pio2_hi + pio2_lo;
    double z = x * x;
    double r = z * asin_R(z);
    return pio2_hi - (x - (pio2_lo - x * r));
  }
  if (x < 0.0) {
    double z = (1.0 + x) * 0.5;
    double s = ::sqrt(z);
    double r = z * asin_R(z);
    double w = r * s - pio2_lo;
    return pi_hi - 2.0 * (s + w);   // pi - 2 asin(s); 2*pio2_lo == pi_lo
  }
  double z = (1.0 - x) * 0.5;
  double s = ::sqrt(z);
  double df = from_bits(to_bits(s) & 0xffffffff00000000ull);
  double c = (z - df * df) / (s + df);
  double r = z * asin_R(z);
  double w = r * s + c;
  return 2.0 * (df + w);
}

}  // namespace fpm
