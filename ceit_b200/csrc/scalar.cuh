// Scalar helpers shared by every kernel: double and complex-double (interleaved re, im) with the
// same free-function vocabulary so kernels can be written once for the real path (zero base
// variable, K real SPD) and the complex-symmetric path (non-zero base variable).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#define CEIT_HD __host__ __device__ __forceinline__

namespace ceit {

struct cd {
  double x, y;
};

CEIT_HD cd mk(double x, double y) {
  cd r;
  r.x = x;
  r.y = y;
  return r;
}

template <class T>
struct Sc;
template <>
struct Sc<double> {
  static CEIT_HD double zero() { return 0.0; }
  static CEIT_HD double one() { return 1.0; }
  static CEIT_HD double from_real(double a) { return a; }
  static CEIT_HD double from_cd(cd a) { return a.x; }
  static constexpr bool is_complex = false;
};
template <>
struct Sc<cd> {
  static CEIT_HD cd zero() { return mk(0.0, 0.0); }
  static CEIT_HD cd one() { return mk(1.0, 0.0); }
  static CEIT_HD cd from_real(double a) { return mk(a, 0.0); }
  static CEIT_HD cd from_cd(cd a) { return a; }
  static constexpr bool is_complex = true;
};

CEIT_HD double add(double a, double b) { return a + b; }
CEIT_HD cd add(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
CEIT_HD double sub(double a, double b) { return a - b; }
CEIT_HD cd sub(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
CEIT_HD double neg(double a) { return -a; }
CEIT_HD cd neg(cd a) { return mk(-a.x, -a.y); }
CEIT_HD double mul(double a, double b) { return a * b; }
CEIT_HD cd mul(cd a, cd b) { return mk(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }
// c + a*b
CEIT_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
CEIT_HD cd fma_(cd a, cd b, cd c) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(a.y, b.x, c.y);
  return c;
}
CEIT_HD double scale(double s, double a) { return s * a; }
CEIT_HD cd scale(double s, cd a) { return mk(s * a.x, s * a.y); }
CEIT_HD cd to_cd(double a) { return mk(a, 0.0); }
CEIT_HD cd to_cd(cd a) { return a; }
CEIT_HD double abs_(double a) { return fabs(a); }
CEIT_HD double abs_(cd a) { return sqrt(fma(a.x, a.x, a.y * a.y)); }
CEIT_HD double inv_(double a) { return 1.0 / a; }
CEIT_HD cd inv_(cd a) {
  double d = fma(a.x, a.x, a.y * a.y);
  return mk(a.x / d, -a.y / d);
}
// reciprocal to ~1 ulp without the slow-path checks of 1.0/x (device); used on the serial critical path of the
// tile factorisation.  MUFU seed r0 with relative error e = 1 - a r0, |e| <= 2^-20; one cubic step
// r0 (1 + e + e^2) leaves e^3 (3 dependent FMAs), a final Newton step rounds it off.
__device__ __forceinline__ double rcp_fast(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  r = fma(fma(-a, r, 1.0), r, r);
  return r;
}
__device__ __forceinline__ cd rcp_fast(cd a) { return inv_(a); }
CEIT_HD double sqrt_(double a) { return sqrt(a); }
// principal square root (Re >= 0); pivots of the complex-symmetric Cholesky have Re > 0
CEIT_HD cd sqrt_(cd a) {
  double r = sqrt(fma(a.x, a.x, a.y * a.y));
  double sx = sqrt(0.5 * (r + fabs(a.x)));
  double sy = (sx > 0.0) ? 0.5 * fabs(a.y) / sx : 0.0;
  if (a.x >= 0.0) return mk(sx, a.y >= 0.0 ? sy : -sy);
  return mk(sy, a.y >= 0.0 ? sx : -sx);
}
CEIT_HD bool bad_pivot(double a) { return !(a > 0.0); }
CEIT_HD bool bad_pivot(cd a) { return !(fma(a.x, a.x, a.y * a.y) > 0.0); }

}  // namespace ceit
