// Float-image route of threshold_li (ref: operations.py:186 -> [3p] ski.filters.threshold_li): scikit-image iterates
//   t <- (mean_back - mean_fore) / (log mean_back - log mean_fore)
// on image - image.min(), where the two means are np.mean(image[image > t]) and np.mean(image[~(image > t)]) of the
// PIXELS (no histogram for float images), and stops when t moves by less than np.min(np.diff(np.unique(image))) / 2.
// The per-pixel work is built here; the scalar recurrence stays with the caller:
//  * amt_li_shift_f64: image - min, rounded per element like NumPy's in-place subtraction;
//  * amt_li_min_gap_f64: the smallest positive difference of neighbouring values of the sorted plane (= min of
//    np.diff(np.unique(.))): a bitonic sort of a padded copy, then one pass over neighbours;
//  * amt_li_split_f64: boolean-mask indexing, i.e. a STABLE split of the plane into the samples above t and the rest,
//    both in raster order, so that amt_pairwise_sum_f64 of either part is np.sum of NumPy's compacted array bit for bit.
#include "common.cuh"

namespace amt {

__global__ void __launch_bounds__(256) li_shift_kernel(const double* __restrict__ in, const double lo, double* __restrict__ out,
                                                       const int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dsub(in[i], lo);
}

__global__ void __launch_bounds__(256) li_pad_copy_kernel(const double* __restrict__ in, const int64_t n, double* __restrict__ out,
                                                          const int64_t n_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) out[i] = i < n ? in[i] : __longlong_as_double(0x7ff0000000000000ll);
}

constexpr int kSortTile = 2048;  // elements one CTA sorts / merges in shared memory (1024 threads, one pair each)

__device__ __forceinline__ void bitonic_exchange(double& a, double& b, const bool ascending) {
  if ((a > b) == ascending) {
    const double t = a;
    a = b;
    b = t;
  }
}

// every (k, j) step with k <= kSortTile: sorts each tile, direction alternating so that step k = 2*kSortTile can merge
__global__ void __launch_bounds__(1024) bitonic_tile_sort_kernel(double* __restrict__ a) {
  __shared__ double s[kSortTile];
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  const int t = threadIdx.x;
  s[t] = a[base + t];
  s[t + 1024] = a[base + t + 1024];
  __syncthreads();
  for (int k = 2; k <= kSortTile; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = 2 * t - (t & (j - 1));  // lower index of this thread's pair
      const bool asc = (((base + i) & k) == 0);
      bitonic_exchange(s[i], s[i + j], asc);
      __syncthreads();
    }
  }
  a[base + t] = s[t];
  a[base + t + 1024] = s[t + 1024];
}

// one (k, j) step with j >= kSortTile
__global__ void __launch_bounds__(256) bitonic_global_step_kernel(double* __restrict__ a, const int64_t k, const int64_t j,
                                                                  const int64_t n_pairs) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  const int64_t i = 2 * t - (t & (j - 1));
  double x = a[i], y = a[i + j];
  if ((x > y) == ((i & k) == 0)) {
    a[i] = y;
    a[i + j] = x;
  }
}

// the steps j = kSortTile/2 .. 1 of stage k (k > kSortTile) inside shared memory
__global__ void __launch_bounds__(1024) bitonic_tile_merge_kernel(double* __restrict__ a, const int64_t k) {
  __shared__ double s[kSortTile];
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  const int t = threadIdx.x;
  s[t] = a[base + t];
  s[t + 1024] = a[base + t + 1024];
  __syncthreads();
  const bool asc = ((base & k) == 0);  // k > kSortTile: one direction per tile
  for (int j = kSortTile >> 1; j > 0; j >>= 1) {
    const int i = 2 * t - (t & (j - 1));
    bitonic_exchange(s[i], s[i + j], asc);
    __syncthreads();
  }
  a[base + t] = s[t];
  a[base + t + 1024] = s[t + 1024];
}

// positive doubles order like their bit patterns: the minimum gap is an atomicMin over uint64
__global__ void __launch_bounds__(256) li_min_gap_kernel(const double* __restrict__ sorted, const int64_t n,
                                                         unsigned long long* __restrict__ out_bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t best = 0x7ff0000000000000ull;
  if (i + 1 < n) {
    const double d = dsub(sorted[i + 1], sorted[i]);
    if (d > 0.0) best = (uint64_t)__double_as_longlong(d);
  }
  best = warp_min_u64(best);
  if ((threadIdx.x & 31) == 0 && best != 0x7ff0000000000000ull) atomicMin(out_bits, (unsigned long long)best);
}

constexpr int kSplitBlock = 1024;

__global__ void __launch_bounds__(kSplitBlock) li_split_count_kernel(const double* __restrict__ x, const int64_t n, const double t,
                                                                     int32_t* __restrict__ block_counts) {
  const int64_t i = (int64_t)blockIdx.x * kSplitBlock + threadIdx.x;
  const int c = __syncthreads_count(i < n && x[i] > t);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive prefix of the block counts (64-bit), one CTA; totals[0] = samples above t, totals[1] = the rest
__global__ void __launch_bounds__(1024) li_split_scan_kernel(const int32_t* __restrict__ block_counts, const int64_t n_blocks,
                                                             int64_t* __restrict__ block_offsets, const int64_t n,
                                                             int64_t* __restrict__ totals) {
  __shared__ int64_t partial[1024];
  const int t = threadIdx.x;
  const int64_t per = (n_blocks + 1023) / 1024;
  const int64_t lo = t * per, hi = lo + per < n_blocks ? lo + per : n_blocks;
  int64_t sum = 0;
  for (int64_t b = lo; b < hi; ++b) sum += block_counts[b];
  partial[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    const int64_t v = t >= o ? partial[t - o] : 0;
    __syncthreads();
    partial[t] += v;
    __syncthreads();
  }
  int64_t run = t > 0 ? partial[t - 1] : 0;
  for (int64_t b = lo; b < hi; ++b) {
    block_offsets[b] = run;
    run += block_counts[b];
  }
  if (t == 1023) {
    totals[0] = partial[1023];
    totals[1] = n - partial[1023];
  }
}

__global__ void __launch_bounds__(kSplitBlock) li_split_scatter_kernel(const double* __restrict__ x, const int64_t n, const double t,
                                                                       const int64_t* __restrict__ block_offsets,
                                                                       double* __restrict__ above, double* __restrict__ rest) {
  __shared__ int warp_counts[kSplitBlock / 32];
  const int64_t block_start = (int64_t)blockIdx.x * kSplitBlock;
  const int64_t i = block_start + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool valid = i < n;
  const double v = valid ? x[i] : 0.0;
  const bool fg = valid && v > t;
  const unsigned ballot = __ballot_sync(0xffffffffu, fg);
  if (lane == 0) warp_counts[warp] = __popc(ballot);
  __syncthreads();
  int before = 0;
  for (int wi = 0; wi < warp; ++wi) before += warp_counts[wi];
  const int rank = before + __popc(ballot & ((1u << lane) - 1u));  // samples above t before this one, in this block
  if (!valid) return;
  const int64_t fg_off = block_offsets[blockIdx.x];
  if (fg)
    above[fg_off + rank] = v;
  else
    rest[(block_start - fg_off) + (threadIdx.x - rank)] = v;
}

static int64_t next_pow2(int64_t n) {
  int64_t p = kSortTile;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace amt

extern "C" {

int amt_li_shift_f64(const double* data, int64_t n, double lo, double* out, amt_stream_t stream) {
  using namespace amt;
  if (!data || !out || n <= 0) return AMT_ERR_INVALID;
  li_shift_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(data, lo, out, n);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

size_t amt_li_min_gap_scratch_bytes(int64_t n) {
  if (n <= 0) return 0;
  return (size_t)amt::next_pow2(n) * sizeof(double) + 256;
}

int amt_li_min_gap_f64(const double* data, int64_t n, double* gap, void* scratch, size_t scratch_bytes, amt_stream_t stream) {
  using namespace amt;
  if (!data || !gap || !scratch || n <= 0 || n > (1ll << 31)) return AMT_ERR_INVALID;
  if (scratch_bytes < amt_li_min_gap_scratch_bytes(n)) return AMT_ERR_CAPACITY;
  cudaStream_t st = as_stream(stream);
  const int64_t n_pad = next_pow2(n);
  double* a = (double*)scratch;
  li_pad_copy_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, st>>>(data, n, a, n_pad);
  AMT_LAUNCH_CHECK();
  const unsigned tiles = (unsigned)(n_pad / kSortTile);
  bitonic_tile_sort_kernel<<<tiles, 1024, 0, st>>>(a);
  AMT_LAUNCH_CHECK();
  for (int64_t k = 2 * kSortTile; k <= n_pad; k <<= 1) {
    for (int64_t j = k >> 1; j >= kSortTile; j >>= 1) {
      bitonic_global_step_kernel<<<(unsigned)ceil_div(n_pad / 2, 256), 256, 0, st>>>(a, k, j, n_pad / 2);
      AMT_LAUNCH_CHECK();
    }
    bitonic_tile_merge_kernel<<<tiles, 1024, 0, st>>>(a, k);
    AMT_LAUNCH_CHECK();
  }
  // +inf = "no two distinct values"
  const unsigned long long inf_bits = 0x7ff0000000000000ull;
  AMT_CUDA_TRY(cudaMemcpyAsync(gap, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, st));
  li_min_gap_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(a, n, (unsigned long long*)gap);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

size_t amt_li_split_scratch_bytes(int64_t n) {
  if (n <= 0) return 0;
  const int64_t blocks = amt::ceil_div(n, amt::kSplitBlock);
  return (size_t)blocks * (sizeof(int32_t) + sizeof(int64_t)) + 512;
}

int amt_li_split_f64(const double* data, int64_t n, double t, double* above, double* rest, int64_t* totals, void* scratch,
                     size_t scratch_bytes, amt_stream_t stream) {
  using namespace amt;
  if (!data || !above || !rest || !totals || !scratch || n <= 0) return AMT_ERR_INVALID;
  if (scratch_bytes < amt_li_split_scratch_bytes(n)) return AMT_ERR_CAPACITY;
  cudaStream_t st = as_stream(stream);
  const int64_t blocks = ceil_div(n, kSplitBlock);
  if (blocks > 0x7fffffff) return AMT_ERR_UNSUPPORTED;
  int64_t* offsets = (int64_t*)scratch;
  int32_t* counts = (int32_t*)((char*)scratch + ((blocks * sizeof(int64_t) + 255) / 256) * 256);
  li_split_count_kernel<<<(unsigned)blocks, kSplitBlock, 0, st>>>(data, n, t, counts);
  AMT_LAUNCH_CHECK();
  li_split_scan_kernel<<<1, 1024, 0, st>>>(counts, blocks, offsets, n, totals);
  AMT_LAUNCH_CHECK();
  li_split_scatter_kernel<<<(unsigned)blocks, kSplitBlock, 0, st>>>(data, n, t, offsets, above, rest);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
