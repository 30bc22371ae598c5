// Scalar abstraction of libhpddm_b200.so: every source file is compiled twice, once with
// K = double (C ABI hpddm_b200_*, include/hpddm_b200.h) and once with -DHB_COMPLEX,
// K = complex double (C ABI hpddm_b200z_*, include/hpddm_b200z.h) -- the same "one source, one
// object per scalar type" scheme as the reference's template parameter K (HPDDM::Schwarz<..., K>,
// include/HPDDM_schwarz.hpp:52) and as MUMPS' s/d/c/z builds.  Quantities that are real in the
// reference (underlying_type<K>: the partition of unity d_, norms, tolerances) stay double.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#ifdef HB_COMPLEX
#define hb hbz  // separate internal namespace per scalar type: both object sets link into one .so
#define HB_API(name) hpddm_b200z_##name
#define HB_PREFIX "hpddm_b200z"
#else
#define HB_API(name) hpddm_b200_##name
#define HB_PREFIX "hpddm_b200"
#endif

namespace hb {

#ifdef HB_COMPLEX
// layout-compatible with std::complex<double> / C99 double _Complex; 16-byte aligned so that one
// element is one 128-bit load
struct __align__(16) cplx {
  double re, im;
};
typedef cplx K;
constexpr bool IS_COMPLEX = true;
#define HB_HD __host__ __device__ __forceinline__
HB_HD K mk(double re, double im = 0.0) {
  K r;
  r.re = re;
  r.im = im;
  return r;
}
HB_HD K operator+(K a, K b) { return mk(a.re + b.re, a.im + b.im); }
HB_HD K operator-(K a, K b) { return mk(a.re - b.re, a.im - b.im); }
HB_HD K operator-(K a) { return mk(-a.re, -a.im); }
HB_HD K operator*(K a, K b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
HB_HD K operator*(double a, K b) { return mk(a * b.re, a * b.im); }
HB_HD K operator*(K a, double b) { return mk(a.re * b, a.im * b); }
HB_HD K operator/(K a, double b) { return mk(a.re / b, a.im / b); }
HB_HD K operator/(K a, K b) {  // Smith's algorithm: no spurious overflow
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, den = b.re + r * b.im;
    return mk((a.re + a.im * r) / den, (a.im - a.re * r) / den);
  }
  const double r = b.re / b.im, den = b.im + r * b.re;
  return mk((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}
HB_HD K &operator+=(K &a, K b) {
  a.re += b.re;
  a.im += b.im;
  return a;
}
HB_HD K &operator-=(K &a, K b) {
  a.re -= b.re;
  a.im -= b.im;
  return a;
}
HB_HD K &operator*=(K &a, K b) {
  a = a * b;
  return a;
}
HB_HD K &operator*=(K &a, double b) {
  a.re *= b;
  a.im *= b;
  return a;
}
HB_HD K &operator/=(K &a, K b) {
  a = a / b;
  return a;
}
HB_HD bool operator==(K a, K b) { return a.re == b.re && a.im == b.im; }
HB_HD bool operator!=(K a, K b) { return !(a == b); }
HB_HD K hb_conj(K a) { return mk(a.re, -a.im); }
HB_HD double hb_real(K a) { return a.re; }
HB_HD double hb_imag(K a) { return a.im; }
HB_HD double hb_abs(K a) { return hypot(a.re, a.im); }
HB_HD double hb_norm(K a) { return a.re * a.re + a.im * a.im; }  // |a|^2 (HPDDM::norm)
// c + a * b
HB_HD K hb_fma(K a, K b, K c) { return mk(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im))); }
HB_HD K hb_fma(double a, K b, K c) { return mk(fma(a, b.re, c.re), fma(a, b.im, c.im)); }
#else
typedef double K;
constexpr bool IS_COMPLEX = false;
#define HB_HD __host__ __device__ __forceinline__
HB_HD K mk(double re, double = 0.0) { return re; }
HB_HD K hb_conj(K a) { return a; }
HB_HD double hb_real(K a) { return a; }
HB_HD double hb_imag(K) { return 0.0; }
HB_HD double hb_abs(K a) { return fabs(a); }
HB_HD double hb_norm(K a) { return a * a; }
HB_HD K hb_fma(K a, K b, K c) { return fma(a, b, c); }
#endif

constexpr int VE = 16 / (int)sizeof(K);   // elements per 128-bit vector: 2 (real) or 1 (complex)
constexpr int KD = (int)sizeof(K) / 8;    // doubles per element (NCCL / copy counts)

#ifdef __CUDACC__
// ---- device helpers --------------------------------------------------------------------------
#ifdef HB_COMPLEX
__device__ __forceinline__ K hb_shfl_xor(K v, int m) { return mk(__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.im, m)); }
__device__ __forceinline__ K hb_shfl(K v, int lane) { return mk(__shfl_sync(0xffffffffu, v.re, lane), __shfl_sync(0xffffffffu, v.im, lane)); }
__device__ __forceinline__ K hb_shfl_down(K v, int o, int w) { return mk(__shfl_down_sync(0xffffffffu, v.re, o, w), __shfl_down_sync(0xffffffffu, v.im, o, w)); }
__device__ __forceinline__ void hb_atomic_add(K *p, K v) {
  atomicAdd(&p->re, v.re);
  atomicAdd(&p->im, v.im);
}
// one 128-bit vector = one complex element.  acc += t (panel) * b (vector)
__device__ __forceinline__ K hb_vdot(double2 t, double2 b, K acc) { return mk(fma(t.x, b.x, fma(-t.y, b.y, acc.re)), fma(t.x, b.y, fma(t.y, b.x, acc.im))); }
// acc (one complex accumulator held in a double2) += t * u
__device__ __forceinline__ void hb_vaxpy(double2 t, K u, double2 &acc) {
  acc.x = fma(t.x, u.re, fma(-t.y, u.im, acc.x));
  acc.y = fma(t.x, u.im, fma(t.y, u.re, acc.y));
}
// publish the VE elements of a vector accumulator into x[c .. c+VE), bounded by lim
__device__ __forceinline__ void hb_vpublish(K *x, int c, int lim, double2 acc) {
  if (c < lim) hb_atomic_add(x + c, mk(acc.x, acc.y));
}
#else
__device__ __forceinline__ K hb_shfl_xor(K v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ K hb_shfl(K v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
__device__ __forceinline__ K hb_shfl_down(K v, int o, int w) { return __shfl_down_sync(0xffffffffu, v, o, w); }
__device__ __forceinline__ void hb_atomic_add(K *p, K v) { atomicAdd(p, v); }
// one 128-bit vector = two consecutive real elements
__device__ __forceinline__ K hb_vdot(double2 t, double2 b, K acc) { return fma(t.x, b.x, fma(t.y, b.y, acc)); }
__device__ __forceinline__ void hb_vaxpy(double2 t, K u, double2 &acc) {
  acc.x = fma(t.x, u, acc.x);
  acc.y = fma(t.y, u, acc.y);
}
__device__ __forceinline__ void hb_vpublish(K *x, int c, int lim, double2 acc) {
  if (c < lim) atomicAdd(x + c, acc.x);
  if (c + 1 < lim) atomicAdd(x + c + 1, acc.y);
}
#endif
#endif  // __CUDACC__

}  // namespace hb
