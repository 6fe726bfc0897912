// The exact-order symmetric correlation inner loop shared by every Gaussian kernel.
#pragma once

#include "common.cuh"

namespace amt {

constexpr int GR = 8;  // outputs per thread along the filter axis

template <typename T>
__device__ __forceinline__ double convert_to_f64(T v, double scale);
template <>
__device__ __forceinline__ double convert_to_f64<double>(double v, double) {
  return v;
}
template <>
__device__ __forceinline__ double convert_to_f64<uint16_t>(uint16_t v, double scale) {
  return dmul((double)v, scale);  // img_as_float: multiply by 1/65535
}

// acc[o], o < R: correlation centred at sample `o` of the thread's window, scipy's order:
//     acc = x[0]*w[c];  for j = r .. 1:  acc += (x[-j] + x[+j]) * w[c-j]
// `at(k)` returns sample k relative to output 0's centre; hw[j] = weights[c-j] (shared memory).
// Two R-wide register windows (left taps, right taps) slide over the samples, so one tap
// step costs 2 sample loads + 1 broadcast weight load for 3*R DP instructions; the j-loop is
// unrolled by R with static register renaming (logical L_o lives in L[(o+u)%R], logical R_o in
// Rt[(o-u+R)%R]), so the window shift costs no MOVs.
template <int R, typename At>
__device__ __forceinline__ void conv_exact(At at, const double* __restrict__ hw, const int r, double (&acc)[R]) {
  double L[R], Rt[R];
  {
    const double w0 = hw[0];
#pragma unroll
    for (int o = 0; o < R; ++o) acc[o] = dmul(at(o), w0);
  }
  if (r == 0) return;
#pragma unroll
  for (int o = 0; o < R; ++o) {
    L[o] = at(o - r);
    Rt[o] = at(o + r);
  }
  int j = r;
  // r % R leading steps with an explicit window shift
  for (int t = r % R; t > 0; --t, --j) {
    const double wj = hw[j];
#pragma unroll
    for (int o = 0; o < R; ++o) acc[o] = dadd(acc[o], dmul(dadd(L[o], Rt[o]), wj));
#pragma unroll
    for (int o = 0; o < R - 1; ++o) L[o] = L[o + 1];
    L[R - 1] = at(R - j);
#pragma unroll
    for (int o = R - 1; o > 0; --o) Rt[o] = Rt[o - 1];
    Rt[0] = at(j - 1);
  }
  // Each tap step is written as three passes over the R outputs (pair sums, products,
  // accumulations) so that the R dependency chains are interleaved instead of serialised, and
  // the samples / weight of the NEXT step are fetched before this step's arithmetic.
  double wj = hw[j];
  for (; j >= R; j -= R) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int jj = j - u;
      double t[R];
#pragma unroll
      for (int o = 0; o < R; ++o) t[o] = dadd(L[(o + u) % R], Rt[(o - u + R) % R]);
      // the two window slots that just became dead take the next step's samples right away
      L[u] = at(R - jj);
      Rt[R - 1 - u] = at(jj - 1);
      const double wn = hw[jj - 1];
#pragma unroll
      for (int o = 0; o < R; ++o) t[o] = dmul(t[o], wj);
#pragma unroll
      for (int o = 0; o < R; ++o) acc[o] = dadd(acc[o], t[o]);
      wj = wn;
    }
  }
}

// generic per-axis launchers (gauss.cu), used as the fallback when a radius does not fit the
// persistent ring kernels of dog.cu
int dog_axis0_generic(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                      const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
                      cudaStream_t st);
int dog_axis1_generic(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                      const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, uint64_t* minmax, cudaStream_t st);

}  // namespace amt
