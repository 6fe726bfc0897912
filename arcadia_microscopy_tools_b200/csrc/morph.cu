// Flat grey-scale erosion / dilation along one axis and the white top-hat built from them.
//
// north_star names a "top-hat / rolling-background filter"; the reference ships only the DoG
// (operations.py:57-97), so this is an extension with no reference call site.  It is pinned to
// scipy.ndimage.white_tophat(x, size=s) (importable in the test container):
//     tmp = grey_erosion(x, size=s)   = minimum_filter1d along every axis, window
//                                       [i - s//2, i - s//2 + s), mode='reflect'
//     tmp = grey_dilation(tmp, size=s) = maximum_filter1d along every axis, window shifted by
//                                       one for even s (scipy negates the origin and subtracts 1)
//     out = x - tmp                    (input dtype; never negative: an opening is <= x)
// min / max are exact, so any evaluation order gives scipy's bits.
//
// HBM-bound streaming: 2 B (uint16) or 8 B (float64) read + written per sample and pass.  The
// window is evaluated with a doubling table in shared memory: m_k[i] = min(x[i .. i+2^k)) for
// k = 0..floor(log2 s) costs one min per sample and level, then out = min(m_K[i], m_K[i+s-2^K]):
// log2(s)+1 operations per sample whatever the window (a direct scan would be s).
//  * strided axis (inner > 1): tile [seg + s - 1][32], lanes along the contiguous axis.
//  * contiguous axis (inner == 1): 32 rows per CTA, tile loaded transposed ([position][row],
//    pitch 33) so the same doubling runs with lanes = rows; results leave through a transposed,
//    coalesced store.

#include "common.cuh"

namespace amt {

constexpr int MF_LANES = 32;
constexpr int MF_SEG = 128;      // outputs along the filter axis per CTA
constexpr int MF_THREADS = 256;  // 8 warps

template <typename T>
__device__ __forceinline__ T mf_op(T a, T b, bool is_max);
template <>
__device__ __forceinline__ int mf_op<int>(int a, int b, bool is_max) {
  return is_max ? max(a, b) : min(a, b);
}
template <>
__device__ __forceinline__ double mf_op<double>(double a, double b, bool is_max) {
  return is_max ? fmax(a, b) : fmin(a, b);
}

__device__ __forceinline__ int reflect_index(int i, const int n) {
  // scipy mode='reflect': d c b a | a b c d | d c b a
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

// grid (ceil(inner/32), ceil(n/SEG), outer); block 256.  ST = storage type (uint16_t / double),
// CT = compute type held in shared memory (int / double).  Window of output i: [i-left, i-left+size).
// minuend (optional): out = minuend - result (the top-hat's final subtraction, fused).
template <typename ST, typename CT>
__global__ void __launch_bounds__(MF_THREADS)
minmax_axis_kernel(const ST* __restrict__ in, ST* __restrict__ out, const ST* __restrict__ minuend, const int64_t n,
                   const int64_t inner, const int size, const int left, const int levels, const int is_max) {
  extern __shared__ __align__(16) unsigned char mf_smem[];
  CT* tile = reinterpret_cast<CT*>(mf_smem);
  const int rows = MF_SEG + size - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t x = (int64_t)blockIdx.x * MF_LANES + lane;
  const bool xok = x < inner;
  const int64_t y0 = (int64_t)blockIdx.y * MF_SEG;
  const int64_t base = (int64_t)blockIdx.z * n * inner;
  for (int r = warp; r < rows; r += MF_THREADS / 32) {
    const int y = reflect_index((int)(y0 - left) + r, (int)n);
    tile[r * MF_LANES + lane] = xok ? (CT)in[base + (int64_t)y * inner + x] : (CT)0;
  }
  __syncthreads();
  // doubling, ping-pong between the two halves of the buffer: after level k, cur[r] = op over
  // [r, r + 2^(k+1)) (valid while r + 2^(k+1) <= rows)
  CT* cur = tile;
  CT* nxt = tile + rows * MF_LANES;
  int span = 1;
  for (int k = 0; k < levels; ++k) {
    for (int r = warp; r + 2 * span <= rows; r += MF_THREADS / 32)
      nxt[r * MF_LANES + lane] = mf_op<CT>(cur[r * MF_LANES + lane], cur[(r + span) * MF_LANES + lane], is_max != 0);
    __syncthreads();
    CT* t = cur;
    cur = nxt;
    nxt = t;
    span *= 2;
  }
  // span = 2^levels <= size < 2 * span
  for (int o = warp; o < MF_SEG; o += MF_THREADS / 32) {
    const int64_t y = y0 + o;
    if (y >= n || !xok) continue;
    const CT r = mf_op<CT>(cur[o * MF_LANES + lane], cur[(o + size - span) * MF_LANES + lane], is_max != 0);
    const int64_t idx = base + y * inner + x;
    out[idx] = minuend ? (ST)((CT)minuend[idx] - r) : (ST)r;
  }
}

// contiguous axis: rows of length n; a CTA takes 32 rows x SEG outputs and loads its tile
// transposed ([position][row], pitch 33 for the float64 transposition) so that the same doubling
// code runs with lanes = rows.
template <typename ST, typename CT>
__global__ void __launch_bounds__(MF_THREADS)
minmax_last_kernel(const ST* __restrict__ in, ST* __restrict__ out, const ST* __restrict__ minuend, const int64_t nrows,
                   const int64_t n, const int size, const int left, const int levels, const int is_max) {
  extern __shared__ __align__(16) unsigned char mf_smem[];
  CT* tile = reinterpret_cast<CT*>(mf_smem);
  constexpr int P = MF_LANES + 1;
  const int cols = MF_SEG + size - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * MF_LANES;
  const int64_t x0 = (int64_t)blockIdx.y * MF_SEG;
  // load: lanes along the contiguous axis (coalesced), one row per warp iteration
  for (int rr = warp; rr < MF_LANES; rr += MF_THREADS / 32) {
    const int64_t row = row0 + rr;
    for (int c = lane; c < cols; c += 32) {
      const int xx = reflect_index((int)(x0 - left) + c, (int)n);
      tile[c * P + rr] = row < nrows ? (CT)in[row * n + xx] : (CT)0;
    }
  }
  __syncthreads();
  CT* cur = tile;
  CT* nxt = tile + cols * P;
  int span = 1;
  for (int k = 0; k < levels; ++k) {
    for (int c = warp; c + 2 * span <= cols; c += MF_THREADS / 32)
      nxt[c * P + lane] = mf_op<CT>(cur[c * P + lane], cur[(c + span) * P + lane], is_max != 0);
    __syncthreads();
    CT* t = cur;
    cur = nxt;
    nxt = t;
    span *= 2;
  }
  // combine into the other half, then a transposed, coalesced store
  for (int o = warp; o < MF_SEG; o += MF_THREADS / 32)
    nxt[o * P + lane] = mf_op<CT>(cur[o * P + lane], cur[(o + size - span) * P + lane], is_max != 0);
  __syncthreads();
  for (int rr = warp; rr < MF_LANES; rr += MF_THREADS / 32) {
    const int64_t row = row0 + rr;
    if (row >= nrows) break;
    for (int o = lane; o < MF_SEG; o += 32) {
      const int64_t xx = x0 + o;
      if (xx >= n) continue;
      const CT r = nxt[o * P + rr];
      const int64_t idx = row * n + xx;
      out[idx] = minuend ? (ST)((CT)minuend[idx] - r) : (ST)r;
    }
  }
}

template <typename ST, typename CT>
static int minmax_axis_launch(const ST* in, ST* out, const ST* minuend, int64_t outer, int64_t n, int64_t inner, int size,
                              int left, int is_max, cudaStream_t st) {
  int levels = 0;
  while ((2 << levels) <= size) ++levels;  // 2^levels <= size
  if (inner == 1) {
    const size_t smem = 2 * (size_t)(MF_SEG + size - 1) * (MF_LANES + 1) * sizeof(CT);
    if (smem > 220 * 1024) return AMT_ERR_CAPACITY;
    auto k = minmax_last_kernel<ST, CT>;
    AMT_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t gx = ceil_div(outer, MF_LANES), gy = ceil_div(n, MF_SEG);
    if (gy > 65535) return AMT_ERR_CAPACITY;
    k<<<dim3((unsigned)gx, (unsigned)gy), MF_THREADS, smem, st>>>(in, out, minuend, outer, n, size, left, levels, is_max);
  } else {
    const size_t smem = 2 * (size_t)(MF_SEG + size - 1) * MF_LANES * sizeof(CT);
    if (smem > 220 * 1024) return AMT_ERR_CAPACITY;
    auto k = minmax_axis_kernel<ST, CT>;
    AMT_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t gy = ceil_div(n, MF_SEG);
    if (gy > 65535 || outer > 65535) return AMT_ERR_CAPACITY;
    k<<<dim3((unsigned)ceil_div(inner, MF_LANES), (unsigned)gy, (unsigned)outer), MF_THREADS, smem, st>>>(
        in, out, minuend, n, inner, size, left, levels, is_max);
  }
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt

extern "C" {

int amt_minmax_filter_axis(const void* in, int dtype, void* out, const void* minuend, int64_t outer, int64_t n,
                           int64_t inner, int size, int left, int is_max, amt_stream_t stream) {
  using namespace amt;
  if (!in || !out || outer <= 0 || n <= 0 || inner <= 0 || size < 1 || size > 256 || left < 0 || left >= size)
    return AMT_ERR_INVALID;
  if (n >= (1ll << 30)) return AMT_ERR_CAPACITY;
  if (dtype == AMT_U16)
    return minmax_axis_launch<uint16_t, int>((const uint16_t*)in, (uint16_t*)out, (const uint16_t*)minuend, outer, n,
                                             inner, size, left, is_max, as_stream(stream));
  if (dtype == AMT_F64)
    return minmax_axis_launch<double, double>((const double*)in, (double*)out, (const double*)minuend, outer, n, inner,
                                              size, left, is_max, as_stream(stream));
  return AMT_ERR_UNSUPPORTED;
}

}  // extern "C"
