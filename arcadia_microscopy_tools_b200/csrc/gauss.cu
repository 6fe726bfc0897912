// Separable float64 Gaussian / difference-of-Gaussians, bit-identical to scipy.ndimage.
//
// Reference path: operations.py:91 ski.filters.difference_of_gaussians -> [3p]
// scipy.ndimage.gaussian_filter(mode='nearest') -> correlate1d, whose symmetric-kernel loop is
//     acc = x[0]*w[c];  for j = r .. 1:  acc += (x[-j] + x[+j]) * w[c-j]
// with separately rounded add / mul / add.  The kernels below keep that
// exact operation order per output sample, so the result is bit-for-bit scipy's.
//
// Roofline: this is the one FP64-pipe-bound stage of the path (3 DP instructions per tap pair,
// 193 per output sample and axis at sigma=16); HBM traffic is ~10x below its bandwidth time.
// Design for the DP pipe: each thread owns R=8 consecutive outputs ALONG THE FILTER AXIS and
// slides two R-wide register windows (left taps / right taps) over the shared-memory tile, so
// one tap step costs 2 LDS.64 + 1 broadcast LDS.64 for 24 DP instructions (LDS:DP = 1:8, the
// shared-memory port would otherwise be the limiter at 2:3).  The j-loop is unrolled by R with
// static register renaming, so the window shift costs no MOVs.
//
//  * axis != last ("V pass"): tile [TH + 2r][32] in natural layout, lanes along the contiguous
//    axis -> conflict-free LDS, coalesced LDG/STG.
//  * last axis ("H pass"): tile stored TRANSPOSED [64 + 2r][33] so lanes are 32 different rows
//    and the window again slides along the slow smem axis; results are staged back through
//    shared memory for coalesced stores.  The fused DoG variant reads both axis-0 results,
//    writes lo - hi and reduces the plane min / max (order-preserving keys, one atomic pair
//    per block) for the percentile stage that follows.

#include "conv.cuh"

namespace amt {

template <typename T>
__device__ __forceinline__ double load_as_f64(const T* p, double scale) {
  return convert_to_f64<T>(__ldg(p), scale);
}

// acc[o] for the R outputs centred at col[o*stride]; hw[j] = weights[c-j].
template <int R>
__device__ __forceinline__ void conv_window(const double* __restrict__ col, const int stride,
                                            const double* __restrict__ hw, const int r, double (&acc)[R]) {
  conv_exact<R>([&](int k) -> double { return col[k * stride]; }, hw, r, acc);
}

// boundary extension of a sample index: 0 = scipy mode 'nearest' (the path of the reference's DoG),
// 1 = scipy mode 'reflect' (d c b a | a b c d | d c b a; threshold_local's default)
__device__ __forceinline__ int64_t extend_index(int64_t i, const int64_t n, const int mode) {
  if (mode == 0) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  const int64_t period = 2 * n;
  i %= period;
  i = i < 0 ? i + period : i;
  return i < n ? i : period - 1 - i;
}

// ------------------------------------------------------------------ V pass
constexpr int GV_TW = 32;
constexpr int GV_TH = 64;  // = 8 thread rows * GR
constexpr int GV_BATCH = 24;  // tile rows per thread when radius <= 64

template <typename InT, bool DUAL>
__global__ void __launch_bounds__(256, 3)
gauss_v_kernel(const InT* __restrict__ in, const double scale, double* __restrict__ out_a,
               double* __restrict__ out_b, const int64_t n, const int64_t inner,
               const double* __restrict__ hw_a, const int r_a, const double* __restrict__ hw_b, const int r_b,
               const int mode) {
  extern __shared__ double smem[];
  const int rmax = DUAL ? (r_a > r_b ? r_a : r_b) : r_a;
  const int rows = GV_TH + 2 * rmax;
  double* tile = smem;
  double* wa = tile + rows * GV_TW;
  double* wb = wa + (r_a + 1);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * GV_TW + tx;
  for (int i = tid; i <= r_a; i += 256) wa[i] = hw_a[i];
  if (DUAL)
    for (int i = tid; i <= r_b; i += 256) wb[i] = hw_b[i];

  const int64_t x = (int64_t)blockIdx.x * GV_TW + tx;
  const int64_t y0 = (int64_t)blockIdx.y * GV_TH;
  const int64_t plane = (int64_t)blockIdx.z * n * inner;
  const InT* src = in + plane;
  const bool xok = x < inner;
  if (rows <= 8 * GV_BATCH) {
    // all global loads of the tile in flight at once (one exposed latency, not rows/8 of them)
    InT raw[GV_BATCH];
#pragma unroll
    for (int i = 0; i < GV_BATCH; ++i) {
      const int s = ty + 8 * i;
      int64_t y = y0 - rmax + s;
      y = extend_index(y, n, mode);
      raw[i] = (xok && s < rows) ? __ldg(src + y * inner + x) : InT(0);
    }
#pragma unroll
    for (int i = 0; i < GV_BATCH; ++i) {
      const int s = ty + 8 * i;
      if (s < rows) tile[s * GV_TW + tx] = convert_to_f64<InT>(raw[i], scale);
    }
  } else {
    for (int s = ty; s < rows; s += 8) {
      int64_t y = y0 - rmax + s;
      y = extend_index(y, n, mode);
      tile[s * GV_TW + tx] = xok ? load_as_f64<InT>(src + y * inner + x, scale) : 0.0;
    }
  }
  __syncthreads();

  const double* col = tile + (rmax + ty * GR) * GV_TW + tx;
  double acc[GR];
  conv_window<GR>(col, GV_TW, wa, r_a, acc);
  const int64_t yb = y0 + ty * GR;
  if (xok) {
#pragma unroll
    for (int o = 0; o < GR; ++o)
      if (yb + o < n) out_a[plane + (yb + o) * inner + x] = acc[o];
  }
  if (DUAL) {
    conv_window<GR>(col, GV_TW, wb, r_b, acc);
    if (xok) {
#pragma unroll
      for (int o = 0; o < GR; ++o)
        if (yb + o < n) out_b[plane + (yb + o) * inner + x] = acc[o];
    }
  }
}

// ------------------------------------------------------------------ H pass
constexpr int GH_ROWS = 32;
constexpr int GH_TX = 64;  // = 8 warps * GR
constexpr int GH_PITCH = 33;
constexpr int GH_STAGE_PITCH = GH_TX + 1;
constexpr int GH_BATCH = 6;  // column groups per thread and row when radius <= 64

template <typename InT>
__device__ __forceinline__ void gh_load_tile(double* tile, const InT* __restrict__ src, const double scale,
                                             const int64_t row0, const int64_t nrows, const int64_t n,
                                             const int64_t x0, const int r, const int warp, const int lane,
                                             const int mode) {
  const int width = GH_TX + 2 * r;
  if (width <= 32 * GH_BATCH) {
    // 4 rows x GH_BATCH column groups per thread, every load issued before the first store
    InT raw[4][GH_BATCH];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t row = row0 + warp + 8 * k;
      const bool rok = row < nrows;
      const InT* p = src + (rok ? row : 0) * n;
#pragma unroll
      for (int i = 0; i < GH_BATCH; ++i) {
        const int xx = lane + 32 * i;
        int64_t gx = x0 - r + xx;
        gx = extend_index(gx, n, mode);
        raw[k][i] = (rok && xx < width) ? __ldg(p + gx) : InT(0);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int i = 0; i < GH_BATCH; ++i) {
        const int xx = lane + 32 * i;
        if (xx < width) tile[xx * GH_PITCH + warp + 8 * k] = convert_to_f64<InT>(raw[k][i], scale);
      }
    }
    return;
  }
  for (int rr = warp; rr < GH_ROWS; rr += 8) {
    const int64_t row = row0 + rr;
    const bool rok = row < nrows;
    const InT* p = src + (rok ? row : 0) * n;
    for (int xx = lane; xx < width; xx += 32) {
      int64_t gx = x0 - r + xx;
      gx = extend_index(gx, n, mode);
      tile[xx * GH_PITCH + rr] = rok ? load_as_f64<InT>(p + gx, scale) : 0.0;
    }
  }
}

// grid: (row blocks, x tiles, planes).  DUAL: out = conv(in_a, hw_a) - conv(in_b, hw_b) and
// per-plane min/max keys (minmax may be null).
template <typename InT, bool DUAL>
__global__ void __launch_bounds__(256, 3)
gauss_h_kernel(const InT* __restrict__ in_a, const double* __restrict__ in_b, const double scale,
               double* __restrict__ out, const int64_t nrows, const int64_t n,
               const double* __restrict__ hw_a, const int r_a, const double* __restrict__ hw_b, const int r_b,
               uint64_t* __restrict__ minmax, const int mode) {
  extern __shared__ double smem[];
  double* tile_a = smem;
  double* tile_b = tile_a + (GH_TX + 2 * r_a) * GH_PITCH;
  double* wa = tile_b + (DUAL ? (GH_TX + 2 * r_b) * GH_PITCH : 0);
  double* wb = wa + (r_a + 1);
  __shared__ uint64_t s_mm[16];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i <= r_a; i += 256) wa[i] = hw_a[i];
  if (DUAL)
    for (int i = tid; i <= r_b; i += 256) wb[i] = hw_b[i];

  const int64_t row0 = (int64_t)blockIdx.x * GH_ROWS;
  const int64_t x0 = (int64_t)blockIdx.y * GH_TX;
  const int64_t plane = (int64_t)blockIdx.z * nrows * n;
  gh_load_tile<InT>(tile_a, in_a + plane, scale, row0, nrows, n, x0, r_a, warp, lane, mode);
  if (DUAL) gh_load_tile<double>(tile_b, in_b + plane, 1.0, row0, nrows, n, x0, r_b, warp, lane, mode);
  __syncthreads();

  double acc[GR];
  conv_window<GR>(tile_a + (r_a + warp * GR) * GH_PITCH + lane, GH_PITCH, wa, r_a, acc);
  if (DUAL) {
    double acc_b[GR];
    conv_window<GR>(tile_b + (r_b + warp * GR) * GH_PITCH + lane, GH_PITCH, wb, r_b, acc_b);
#pragma unroll
    for (int o = 0; o < GR; ++o) acc[o] = dsub(acc[o], acc_b[o]);
  }
  __syncthreads();  // everyone is done reading tile_a: reuse it as the output stage
  double* stage = tile_a;
#pragma unroll
  for (int o = 0; o < GR; ++o) stage[lane * GH_STAGE_PITCH + warp * GR + o] = acc[o];

  if (DUAL && minmax != nullptr) {
    uint64_t kmin = ~0ull, kmax = 0ull;
    const bool rok = row0 + lane < nrows;
#pragma unroll
    for (int o = 0; o < GR; ++o) {
      if (rok && x0 + warp * GR + o < n) {
        const uint64_t k = f64_to_key(acc[o]);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
      }
    }
    kmin = warp_min_u64(kmin);
    kmax = warp_max_u64(kmax);
    if (lane == 0) {
      s_mm[warp] = kmin;
      s_mm[8 + warp] = kmax;
    }
  }
  __syncthreads();
  for (int rr = warp; rr < GH_ROWS; rr += 8) {
    const int64_t row = row0 + rr;
    if (row >= nrows) break;
    for (int xx = lane; xx < GH_TX; xx += 32) {
      const int64_t gx = x0 + xx;
      if (gx < n) out[plane + row * n + gx] = stage[rr * GH_STAGE_PITCH + xx];
    }
  }
  if (DUAL && minmax != nullptr && tid == 0) {
    uint64_t kmin = s_mm[0], kmax = s_mm[8];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      kmin = s_mm[i] < kmin ? s_mm[i] : kmin;
      kmax = s_mm[8 + i] > kmax ? s_mm[8 + i] : kmax;
    }
    atomicMin((unsigned long long*)&minmax[2 * blockIdx.z], (unsigned long long)kmin);
    atomicMax((unsigned long long*)&minmax[2 * blockIdx.z + 1], (unsigned long long)kmax);
  }
}

__global__ void minmax_init_kernel(uint64_t* mm, int64_t n_img) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_img) {
    mm[2 * i] = ~0ull;
    mm[2 * i + 1] = 0ull;
  }
}

__global__ void sub_f64_kernel(const double* __restrict__ a, const double* __restrict__ b,
                               double* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) out[i] = dsub(a[i], b[i]);
}

constexpr size_t kMaxSmem = 227 * 1024;
int g_stream_pad_kb = 0;
int g_stream_ctas = 16;  // grid cap (CTAs per SM) of the grid-stride streaming kernels (amt_tune "stream_ctas")

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > kMaxSmem) return AMT_ERR_CAPACITY;
  if (bytes > 48 * 1024) {
    AMT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return AMT_OK;
}

template <typename InT, bool DUAL>
static int launch_v(const InT* in, double scale, double* out_a, double* out_b, int64_t outer, int64_t n,
                    int64_t inner, const double* hw_a, int r_a, const double* hw_b, int r_b, cudaStream_t st,
                    int mode = 0) {
  const int rmax = DUAL ? (r_a > r_b ? r_a : r_b) : r_a;
  const size_t smem = ((size_t)(GV_TH + 2 * rmax) * GV_TW + (r_a + 1) + (DUAL ? r_b + 1 : 0)) * sizeof(double);
  AMT_TRY(set_smem(gauss_v_kernel<InT, DUAL>, smem));
  const int64_t gy = ceil_div(n, GV_TH);
  if (gy > 65535 || outer > 65535) return AMT_ERR_CAPACITY;
  dim3 grid((unsigned)ceil_div(inner, GV_TW), (unsigned)gy, (unsigned)outer), block(GV_TW, 8);
  gauss_v_kernel<InT, DUAL><<<grid, block, smem, st>>>(in, scale, out_a, out_b, n, inner, hw_a, r_a, hw_b, r_b, mode);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

template <typename InT, bool DUAL>
static int launch_h(const InT* in_a, const double* in_b, double scale, double* out, int64_t planes,
                    int64_t nrows, int64_t n, const double* hw_a, int r_a, const double* hw_b, int r_b,
                    uint64_t* minmax, cudaStream_t st, int mode = 0) {
  const size_t smem = ((size_t)(GH_TX + 2 * r_a) * GH_PITCH + (DUAL ? (size_t)(GH_TX + 2 * r_b) * GH_PITCH : 0) +
                       (r_a + 1) + (DUAL ? r_b + 1 : 0)) * sizeof(double);
  AMT_TRY(set_smem(gauss_h_kernel<InT, DUAL>, smem));
  const int64_t gy = ceil_div(n, GH_TX);
  if (gy > 65535 || planes > 65535) return AMT_ERR_CAPACITY;
  dim3 grid((unsigned)ceil_div(nrows, GH_ROWS), (unsigned)gy, (unsigned)planes), block(256);
  gauss_h_kernel<InT, DUAL><<<grid, block, smem, st>>>(in_a, in_b, scale, out, nrows, n, hw_a, r_a, hw_b, r_b, minmax, mode);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st) {
  minmax_init_kernel<<<(unsigned)ceil_div(n_img, 256), 256, 0, st>>>(mm, n_img);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int dog_axis0_generic(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                      const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
                      cudaStream_t st) {
  if (in_dtype == AMT_U16)
    return launch_v<uint16_t, true>((const uint16_t*)in, in_scale, tmp_lo, tmp_hi, n_img, h, w, hw_lo, r_lo, hw_hi,
                                    r_hi, st);
  if (in_dtype == AMT_F64)
    return launch_v<double, true>((const double*)in, 1.0, tmp_lo, tmp_hi, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, st);
  return AMT_ERR_UNSUPPORTED;
}

int dog_axis1_generic(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                      const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, uint64_t* minmax, cudaStream_t st) {
  return launch_h<double, true>(tmp_lo, tmp_hi, 1.0, out, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, minmax, st);
}

}  // namespace amt

extern "C" {

int amt_gaussian_axis_mode(const void* in, int in_dtype, double in_scale, double* out, int64_t outer, int64_t n,
                           int64_t inner, const double* half_w, int radius, int mode, amt_stream_t stream) {
  using namespace amt;
  if (!in || !out || !half_w || outer <= 0 || n <= 0 || inner <= 0 || radius < 0) return AMT_ERR_INVALID;
  if (mode != AMT_EXTEND_NEAREST && mode != AMT_EXTEND_REFLECT) return AMT_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (inner == 1) {
    // rows = outer, filter along the contiguous axis; fold rows into (planes, nrows) to fit grid limits
    if (in_dtype == AMT_U16)
      return launch_h<uint16_t, false>((const uint16_t*)in, nullptr, in_scale, out, 1, outer, n, half_w, radius,
                                       nullptr, 0, nullptr, st, mode);
    if (in_dtype == AMT_F64)
      return launch_h<double, false>((const double*)in, nullptr, 1.0, out, 1, outer, n, half_w, radius, nullptr, 0,
                                     nullptr, st, mode);
    return AMT_ERR_UNSUPPORTED;
  }
  if (in_dtype == AMT_U16)
    return launch_v<uint16_t, false>((const uint16_t*)in, in_scale, out, nullptr, outer, n, inner, half_w, radius,
                                     nullptr, 0, st, mode);
  if (in_dtype == AMT_F64)
    return launch_v<double, false>((const double*)in, 1.0, out, nullptr, outer, n, inner, half_w, radius, nullptr,
                                   0, st, mode);
  return AMT_ERR_UNSUPPORTED;
}

int amt_gaussian_axis(const void* in, int in_dtype, double in_scale, double* out, int64_t outer, int64_t n,
                      int64_t inner, const double* half_w, int radius, amt_stream_t stream) {
  return amt_gaussian_axis_mode(in, in_dtype, in_scale, out, outer, n, inner, half_w, radius, AMT_EXTEND_NEAREST, stream);
}

int amt_sub_f64(const double* a, const double* b, double* out, int64_t n, amt_stream_t stream) {
  using namespace amt;
  if (!a || !b || !out || n < 0) return AMT_ERR_INVALID;
  if (n == 0) return AMT_OK;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * g_stream_ctas) blocks = kNumSMs * g_stream_ctas;
  const size_t pad = (size_t)g_stream_pad_kb * 1024;  // probe knob: dynamic shared memory the kernel never touches
  if (pad > 48 * 1024) AMT_CUDA_TRY(cudaFuncSetAttribute(sub_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
  sub_f64_kernel<<<(unsigned)blocks, 256, pad, as_stream(stream)>>>(a, b, out, n);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
