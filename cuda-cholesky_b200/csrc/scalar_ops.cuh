// scalar_ops.cuh — uniform arithmetic over the four element types
//   float, double, float2 (= C99 float complex), double2 (= C99 double complex)
#pragma once
#include <cuda_runtime.h>

namespace chol {

template <typename T> struct ScalarOf;
template <> struct ScalarOf<float>   { using Real = float;  static constexpr bool cplx = false; };
template <> struct ScalarOf<double>  { using Real = double; static constexpr bool cplx = false; };
template <> struct ScalarOf<float2>  { using Real = float;  static constexpr bool cplx = true; };
template <> struct ScalarOf<double2> { using Real = double; static constexpr bool cplx = true; };

__host__ __device__ __forceinline__ float  mk(float, float re, float)    { return re; }
__host__ __device__ __forceinline__ double mk(double, double re, double) { return re; }
__host__ __device__ __forceinline__ float2  mk(float2, float re, float im)    { return make_float2(re, im); }
__host__ __device__ __forceinline__ double2 mk(double2, double re, double im) { return make_double2(re, im); }
template <typename T> __host__ __device__ __forceinline__ T zero_of() { return mk(T(), 0, 0); }
template <typename T> __host__ __device__ __forceinline__ T one_of() { return mk(T(), 1, 0); }
template <typename T> __host__ __device__ __forceinline__ T from_real(typename ScalarOf<T>::Real r) { return mk(T(), r, 0); }

__host__ __device__ __forceinline__ float  re_of(float a)   { return a; }
__host__ __device__ __forceinline__ double re_of(double a)  { return a; }
__host__ __device__ __forceinline__ float  re_of(float2 a)  { return a.x; }
__host__ __device__ __forceinline__ double re_of(double2 a) { return a.x; }

__host__ __device__ __forceinline__ float  conj_of(float a)   { return a; }
__host__ __device__ __forceinline__ double conj_of(double a)  { return a; }
__host__ __device__ __forceinline__ float2  conj_of(float2 a)  { return make_float2(a.x, -a.y); }
__host__ __device__ __forceinline__ double2 conj_of(double2 a) { return make_double2(a.x, -a.y); }

__host__ __device__ __forceinline__ float  add(float a, float b)   { return a + b; }
__host__ __device__ __forceinline__ double add(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ float2  add(float2 a, float2 b)   { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float  sub(float a, float b)   { return a - b; }
__host__ __device__ __forceinline__ double sub(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ float2  sub(float2 a, float2 b)   { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float  neg(float a)   { return -a; }
__host__ __device__ __forceinline__ double neg(double a)  { return -a; }
__host__ __device__ __forceinline__ float2  neg(float2 a)  { return make_float2(-a.x, -a.y); }
__host__ __device__ __forceinline__ double2 neg(double2 a) { return make_double2(-a.x, -a.y); }

__host__ __device__ __forceinline__ float  mul(float a, float b)   { return a * b; }
__host__ __device__ __forceinline__ double mul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ float2  mul(float2 a, float2 b)   { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ double2 mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a += b*c
__host__ __device__ __forceinline__ void fma_acc(float &a, float b, float c)    { a = fmaf(b, c, a); }
__host__ __device__ __forceinline__ void fma_acc(double &a, double b, double c) { a = fma(b, c, a); }
__host__ __device__ __forceinline__ void fma_acc(float2 &a, float2 b, float2 c) {
  a.x = fmaf(b.x, c.x, a.x); a.x = fmaf(-b.y, c.y, a.x); a.y = fmaf(b.x, c.y, a.y); a.y = fmaf(b.y, c.x, a.y);
}
__host__ __device__ __forceinline__ void fma_acc(double2 &a, double2 b, double2 c) {
  a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x); a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
// scale by a real
__host__ __device__ __forceinline__ float  rscale(float a, float r)    { return a * r; }
__host__ __device__ __forceinline__ double rscale(double a, double r)  { return a * r; }
__host__ __device__ __forceinline__ float2  rscale(float2 a, float r)   { return make_float2(a.x * r, a.y * r); }
__host__ __device__ __forceinline__ double2 rscale(double2 a, double r) { return make_double2(a.x * r, a.y * r); }
// reciprocal
__host__ __device__ __forceinline__ float  recip(float a)  { return 1.0f / a; }
__host__ __device__ __forceinline__ double recip(double a) { return 1.0 / a; }
__host__ __device__ __forceinline__ float2 recip(float2 a) {
  const float d = a.x * a.x + a.y * a.y; return make_float2(a.x / d, -a.y / d);
}
__host__ __device__ __forceinline__ double2 recip(double2 a) {
  const double d = a.x * a.x + a.y * a.y; return make_double2(a.x / d, -a.y / d);
}
__host__ __device__ __forceinline__ bool is_zero(float a)   { return a == 0.0f; }
__host__ __device__ __forceinline__ bool is_zero(double a)  { return a == 0.0; }
__host__ __device__ __forceinline__ bool is_zero(float2 a)  { return a.x == 0.0f && a.y == 0.0f; }
__host__ __device__ __forceinline__ bool is_zero(double2 a) { return a.x == 0.0 && a.y == 0.0; }

}  // namespace chol
