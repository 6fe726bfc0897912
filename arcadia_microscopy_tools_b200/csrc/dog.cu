// Fused 2-D difference of Gaussians: persistent ring-buffer kernels for the hot path.
//
// Reference path: operations.py:91 ski.filters.difference_of_gaussians [3p] (see gauss.cu for
// the arithmetic contract: float64, scipy's operation order, no FMA — bit-identical).
//
// These kernels are FP64-pipe-bound (3 DP instructions per tap pair, 2*(1+3*64)+2*(1+3*2) = 400
// per sample for sigma = 0.6 / 16).  On B200 a DP warp instruction holds the SMSP's issue port
// for two cycles, so every non-DP instruction in the loop costs DP throughput; the design goal
// is therefore "nothing but DADD/DMUL and three LDS per tap step":
//  * one kernel shape for both axes.  Pass 1 filters along axis 0 and writes its two results
//    TRANSPOSED; pass 2 filters the transposed planes along their axis 0 (= image axis 1) and
//    writes lo - hi transposed back.  Lanes always run along the contiguous axis (conflict-free
//    LDS, coalesced LDG) and every thread's 8 outputs are contiguous in the transposed layout
//    (four 16-byte stores).
//  * a CTA walks 512 samples along the filter axis and keeps the samples it needs in a
//    shared-memory ring [rows][32]; the next 64 rows are prefetched into registers while the
//    current 64 are computed (uint16 -> float64 conversion once per sample, on the way in), so
//    each sample is fetched once per strip and its latency is hidden.
//  * the first 2r ring rows are mirrored behind the ring, so a step's window is one contiguous
//    run and every shared-memory access in the unrolled inner loop is base + constant.
//  * the narrow (sigma_lo) operand of pass 2 needs only 2*r_lo+8 samples per thread; they come
//    straight from global memory (L1) instead of a second ring, which keeps two CTAs per SM.
// The inner loop is conv_exact (conv.cuh).

#include <cstdlib>

#include "conv.cuh"

namespace amt {

constexpr int PV_TH = 64;  // samples per step along the filter axis (8 thread rows * GR)
constexpr int PV_TW = 32;  // lanes along the contiguous axis
constexpr int SEG = 512;   // samples along the filter axis per CTA
constexpr size_t kSmemMax = 113 * 1024;  // two CTAs per SM

template <int R>
__device__ __forceinline__ void store_run(double* dst, const double (&v)[R], const int first, const int end, const int vec2) {
  if (vec2 && first + R <= end) {
#pragma unroll
    for (int o = 0; o < R; o += 2) *reinterpret_cast<double2*>(dst + o) = make_double2(v[o], v[o + 1]);
  } else {
#pragma unroll
    for (int o = 0; o < R; ++o)
      if (first + o < end) dst[o] = v[o];
  }
}

// grid (ceil(inner/32), ceil(n/SEG), planes); block (32, 8).
// FIRST pass : in = image (InT), out_a = G_hi^T, out_b = G_lo^T          (both inner x n)
// SECOND pass: in = G_hi^T (double), in_lo = G_lo^T, out_a = G_lo - G_hi  transposed back
template <typename InT, bool SECOND>
__global__ void __launch_bounds__(256, 2)
dog_pass_kernel(const InT* __restrict__ in, const double* __restrict__ in_lo, const double scale,
                double* __restrict__ out_a, double* __restrict__ out_b, const int n, const int inner,
                const double* __restrict__ hw_lo, const int r_lo, const double* __restrict__ hw_hi, const int r_hi,
                const int ring_rows, uint64_t* __restrict__ minmax, const int vec2) {
  extern __shared__ double smem[];
  __shared__ uint64_t s_mm[16];
  const int rmax = SECOND ? r_hi : (r_lo > r_hi ? r_lo : r_hi);
  const int mirror_rows = 2 * rmax;  // ring rows [0, mirror) are also kept at [ring_rows, ring_rows + mirror)
  double* ring = smem;
  double* wlo = ring + (ring_rows + mirror_rows) * PV_TW;
  double* whi = wlo + (r_lo + 1);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * PV_TW + tx;
  for (int i = tid; i <= r_lo; i += 256) wlo[i] = hw_lo[i];
  for (int i = tid; i <= r_hi; i += 256) whi[i] = hw_hi[i];

  const int mask = ring_rows - 1;
  auto ring_store = [&](int s, double v) {
    const int p = s & mask;
    ring[p * PV_TW + tx] = v;
    if (p < mirror_rows) ring[(p + ring_rows) * PV_TW + tx] = v;
  };
  const int x = blockIdx.x * PV_TW + tx;
  const bool xok = x < inner;
  const int y_begin = blockIdx.y * SEG;
  const int y_end = (y_begin + SEG < n) ? y_begin + SEG : n;
  const int64_t plane = (int64_t)blockIdx.z * n * inner;
  const InT* src = in + plane + (xok ? x : 0);
  const int base = y_begin - rmax;  // sample index of ring row 0
  const int pro_rows = PV_TH + 2 * rmax;

  for (int s0 = 0; s0 < pro_rows; s0 += 64) {  // prologue: 8 rows per thread in flight
    InT raw[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      int y = base + s0 + ty + 8 * m;
      y = y < 0 ? 0 : (y > n - 1 ? n - 1 : y);  // mode='nearest'
      raw[m] = __ldg(src + (int64_t)y * inner);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int s = s0 + ty + 8 * m;
      if (s < pro_rows) ring_store(s, convert_to_f64<InT>(raw[m], scale));
    }
  }
  __syncthreads();

  uint64_t kmin = ~0ull, kmax = 0ull;
  const int n_steps = (y_end - y_begin + PV_TH - 1) / PV_TH;
  for (int k = 0; k < n_steps; ++k) {
    const bool more = k + 1 < n_steps;
    InT pre[8];
    if (more) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        int y = base + pro_rows + k * PV_TH + ty + 8 * m;
        y = y > n - 1 ? n - 1 : y;
        pre[m] = __ldg(src + (int64_t)y * inner);
      }
    }
    // the step's window starts at a multiple of 64 inside the ring and runs on into the mirror
    // rows, so every access of the unrolled loop is col + constant
    const double* col = ring + (((k * PV_TH) & mask) + rmax + ty * GR) * PV_TW + tx;
    const int yb = y_begin + k * PV_TH + ty * GR;
    double acc[GR];
    conv_exact<GR>([&](int kk) -> double { return col[kk * PV_TW]; }, whi, r_hi, acc);
    double* dst_a = out_a + plane + (int64_t)x * n + yb;  // transposed: 8 contiguous outputs
    if (!SECOND) {
      // pass 1 keeps the image layout: one coalesced 256-byte row segment per store
      if (xok) {
#pragma unroll
        for (int o = 0; o < GR; ++o)
          if (yb + o < y_end) out_a[plane + (int64_t)(yb + o) * inner + x] = acc[o];
      }
      conv_exact<GR>([&](int kk) -> double { return col[kk * PV_TW]; }, wlo, r_lo, acc);
      if (xok) {
#pragma unroll
        for (int o = 0; o < GR; ++o)
          if (yb + o < y_end) out_b[plane + (int64_t)(yb + o) * inner + x] = acc[o];
      }
    } else {
      // narrow operand: 2*r_lo + 8 samples per thread, straight from global memory (L1)
      const double* lo_col = in_lo + plane + (xok ? x : 0);
      double acc_lo[GR];
      conv_exact<GR>(
          [&](int kk) -> double {
            int y = yb + kk;
            y = y < 0 ? 0 : (y > n - 1 ? n - 1 : y);
            return __ldg(lo_col + (int64_t)y * inner);
          },
          wlo, r_lo, acc_lo);
#pragma unroll
      for (int o = 0; o < GR; ++o) acc[o] = dsub(acc_lo[o], acc[o]);
      if (xok) {
        store_run<GR>(dst_a, acc, yb, y_end, vec2);
        if (minmax != nullptr) {
#pragma unroll
          for (int o = 0; o < GR; ++o) {
            if (yb + o < y_end) {
              const uint64_t key = f64_to_key(acc[o]);
              kmin = key < kmin ? key : kmin;
              kmax = key > kmax ? key : kmax;
            }
          }
        }
      }
    }
    if (more) {
      __syncthreads();  // every warp is done with the oldest 64 rows
#pragma unroll
      for (int m = 0; m < 8; ++m) ring_store(pro_rows + k * PV_TH + ty + 8 * m, convert_to_f64<InT>(pre[m], scale));
      __syncthreads();
    }
  }
  if (SECOND && minmax != nullptr) {
    kmin = warp_min_u64(kmin);
    kmax = warp_max_u64(kmax);
    if (tx == 0) {
      s_mm[ty] = kmin;
      s_mm[8 + ty] = kmax;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 1; i < 8; ++i) {
        kmin = s_mm[i] < kmin ? s_mm[i] : kmin;
        kmax = s_mm[8 + i] > kmax ? s_mm[8 + i] : kmax;
      }
      atomicMin((unsigned long long*)&minmax[2 * blockIdx.z], (unsigned long long)kmin);
      atomicMax((unsigned long long*)&minmax[2 * blockIdx.z + 1], (unsigned long long)kmax);
    }
  }
}

static int pow2_at_least(int v) {
  int p = 64;
  while (p < v) p <<= 1;
  return p;
}

int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st);  // gauss.cu

struct DogPlan {
  bool fast;
  int ring1, ring2;
  size_t smem1, smem2;
};

static DogPlan dog_plan(int in_dtype, int64_t n_img, int64_t h, int64_t w, int r_lo, int r_hi) {
  DogPlan p;
  const int rmax = r_lo > r_hi ? r_lo : r_hi;
  p.ring1 = pow2_at_least(PV_TH + 2 * rmax);
  p.ring2 = pow2_at_least(PV_TH + 2 * r_hi);
  const size_t wbytes = (size_t)((r_lo + 1) + (r_hi + 1)) * sizeof(double);
  p.smem1 = (size_t)(p.ring1 + 2 * rmax) * PV_TW * sizeof(double) + wbytes;
  p.smem2 = (size_t)(p.ring2 + 2 * r_hi) * PV_TW * sizeof(double) + wbytes;
  p.fast = p.smem1 <= kSmemMax && p.smem2 <= kSmemMax && n_img <= 65535 && ceil_div(h, SEG) <= 65535 &&
           ceil_div(w, SEG) <= 65535 && h * w < (1ll << 31) && (in_dtype == AMT_U16 || in_dtype == AMT_F64);
  return p;
}

// pass 1: image (h x w) -> tmp_hi, tmp_lo (same layout)
static int dog_axis0(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                     const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
                     cudaStream_t st) {
  const DogPlan p = dog_plan(in_dtype, n_img, h, w, r_lo, r_hi);
  if (!p.fast) return dog_axis0_generic(in, in_dtype, in_scale, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, tmp_lo, tmp_hi, st);
  const int vec2 = (h % 2 == 0) && (((uintptr_t)tmp_lo) % 16 == 0) && (((uintptr_t)tmp_hi) % 16 == 0);
  dim3 grid((unsigned)ceil_div(w, PV_TW), (unsigned)ceil_div(h, SEG), (unsigned)n_img), block(PV_TW, 8);
  if (in_dtype == AMT_U16) {
    AMT_CUDA_TRY(cudaFuncSetAttribute(dog_pass_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)p.smem1));
    dog_pass_kernel<uint16_t, false><<<grid, block, p.smem1, st>>>((const uint16_t*)in, nullptr, in_scale, tmp_hi, tmp_lo,
                                                                   (int)h, (int)w, hw_lo, r_lo, hw_hi, r_hi, p.ring1,
                                                                   nullptr, vec2);
  } else {
    AMT_CUDA_TRY(cudaFuncSetAttribute(dog_pass_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)p.smem1));
    dog_pass_kernel<double, false><<<grid, block, p.smem1, st>>>((const double*)in, nullptr, 1.0, tmp_hi, tmp_lo, (int)h,
                                                                 (int)w, hw_lo, r_lo, hw_hi, r_hi, p.ring1, nullptr, vec2);
  }
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

// pass 2: image-layout planes filtered along axis 1 by the tile kernel of gauss.cu (loads are
// transposed through shared memory, lo - hi leaves through a staged, coalesced store)
static int dog_axis1(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                     const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, uint64_t* minmax, cudaStream_t st) {
  return dog_axis1_generic(tmp_lo, tmp_hi, out, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, minmax, st);
}

int dog2d(const void* in, int in_dtype, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
          const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
          uint64_t* minmax, cudaStream_t st) {
  if (!in || !out || !tmp_lo || !tmp_hi || !hw_lo || !hw_hi) return AMT_ERR_INVALID;
  if (n_img <= 0 || h <= 0 || w <= 0 || r_lo < 0 || r_hi < 0) return AMT_ERR_INVALID;
  if (in_dtype != AMT_U16 && in_dtype != AMT_F64) return AMT_ERR_UNSUPPORTED;
  if (minmax) AMT_TRY(minmax_init(minmax, n_img, st));
  AMT_TRY(dog_axis0(in, in_dtype, in_scale, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, tmp_lo, tmp_hi, st));
  return dog_axis1(tmp_lo, tmp_hi, out, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, minmax, st);
}

}  // namespace amt

extern "C" {

int amt_dog2d(const void* in, int in_dtype, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
              const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, double* tmp_lo, double* tmp_hi,
              uint64_t* minmax_keys, amt_stream_t stream) {
  return amt::dog2d(in, in_dtype, in_scale, out, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, tmp_lo, tmp_hi,
                    minmax_keys, amt::as_stream(stream));
}

int amt_dog2d_axis0(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, double* tmp_lo, double* tmp_hi,
                    amt_stream_t stream) {
  using namespace amt;
  if (!in || !tmp_lo || !tmp_hi || !half_w_lo || !half_w_hi || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (in_dtype != AMT_U16 && in_dtype != AMT_F64) return AMT_ERR_UNSUPPORTED;
  return dog_axis0(in, in_dtype, in_scale, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, tmp_lo, tmp_hi,
                   as_stream(stream));
}

int amt_dog2d_axis1(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, uint64_t* minmax_keys,
                    amt_stream_t stream) {
  using namespace amt;
  if (!tmp_lo || !tmp_hi || !out || !half_w_lo || !half_w_hi || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (minmax_keys) AMT_TRY(minmax_init(minmax_keys, n_img, as_stream(stream)));
  return dog_axis1(tmp_lo, tmp_hi, out, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, minmax_keys, as_stream(stream));
}

}  // extern "C"
