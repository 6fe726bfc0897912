// Shared device/host helpers for libamt_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amt_b200.h"

namespace amt {

// ---------------------------------------------------------------- error plumbing
void set_last_cuda_error(cudaError_t e);
void count_launch(int n = 1);

#define AMT_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) {                       \
      ::amt::set_last_cuda_error(_e);              \
      return AMT_ERR_CUDA;                         \
    }                                              \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define AMT_LAUNCH_CHECK()                         \
  do {                                             \
    ::amt::count_launch();                         \
    cudaError_t _e = cudaPeekAtLastError();        \
    if (_e != cudaSuccess) {                       \
      ::amt::set_last_cuda_error(_e);              \
      (void)cudaGetLastError();                    \
      return AMT_ERR_CUDA;                         \
    }                                              \
  } while (0)

#define AMT_TRY(expr)                              \
  do {                                             \
    int _s = (expr);                               \
    if (_s != AMT_OK) return _s;                   \
  } while (0)

static inline cudaStream_t as_stream(amt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------- order-preserving keys
__host__ __device__ __forceinline__ uint64_t f64_to_key(double x) {
#ifdef __CUDA_ARCH__
  uint64_t b = (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  __builtin_memcpy(&b, &x, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_to_f64(uint64_t k) {
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  __builtin_memcpy(&x, &b, 8);
  return x;
#endif
}

// 12-bit monotone bucket of a double (selection pre-binning written by the DoG's second pass):
// sign + the exponent and top 6 mantissa bits above 2^-20, straight from the high word, i.e. 64
// buckets per octave (1.5 % wide) for 2^-20 <= |x| < 2^12, which covers difference-of-Gaussians
// values of [0, 1] images with room to spare.  Truncation, shift and clamp are monotone, so the
// buckets are ordered like the values (and like f64_to_key: -0.0 sits below +0.0).
__device__ __forceinline__ uint32_t bucket12(double x) {
  const uint32_t hi = (uint32_t)__double2hiint(x);
  int m = (int)((hi & 0x7fffffffu) >> 14) - ((1023 - 20) << 6);
  m = m < 0 ? 0 : (m > 2047 ? 2047 : m);
  return (hi >> 31) ? (uint32_t)(2047 - m) : (uint32_t)(2048 + m);
}

// ---------------------------------------------------------------- exactly-rounded f64 helpers
// Separately rounded add / sub / mul: never contracted into FMA, whatever -fmad says.
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// numpy _lerp(a, b, t): a + (b-a)*t for t < 0.5, else b - (b-a)*(1-t)
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
  double diff = dsub(b, a);
  if (t >= 0.5) return dsub(b, dmul(diff, dsub(1.0, t)));
  return dadd(a, dmul(diff, t));
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit loads that do not pollute L1
__device__ __forceinline__ int4 ld_nc_int4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace amt
