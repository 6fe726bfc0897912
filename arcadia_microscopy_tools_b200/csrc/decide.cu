// Decision-exact thresholded channel (DESIGN.md "Exact where a decision depends on it").
//
// Everything discrete the pipeline derives from the thresholded channel's difference-of-Gaussians plane D —
// percentile order statistics, the bin of every sample of the rescaled plane's 256-bin histogram, the Otsu
// threshold and the mask P > t — depends on D only through comparisons.  The tensor-core filter (tcgauss.cu)
// delivers D' with |D' - D| <= eps (a proven bound, amt_tcg_error_bound), where D is the float64 plane scipy's
// operation order produces (ref: operations.py:91).  A comparison of D' against a value that is further than the
// bound from it has the same outcome as the comparison of D; only the handful of samples INSIDE a bound of a
// deciding value are re-evaluated in scipy's exact operation order (exact_dog_warp: one warp per sample, the 129 x
// 129 taps of that one sample in the order the strip kernels of dog.cu use), and decided on their exact value:
//   * order statistics: the exact k-th smallest value lies within eps of the k-th smallest of D'; the samples
//     within 2 eps of it are the candidates, #(D' < v' - 2 eps) samples are certainly below, so the exact
//     statistic is the (k - #below)-th smallest EXACT value among the candidates (dx_rank_* kernels).  Ranks 0 and
//     n - 1 (the plane's min / max) are resolved the same way.
//   * histogram bins and the threshold comparison: samples of the rescaled plane within the (propagated) bound of
//     a bin edge or a bin centre (the Otsu threshold is a bin centre) are collected by the map kernel (map.cu),
//     re-evaluated, written back exactly and their histogram counts corrected (dx_patch_kernel in map.cu).
// A plane whose candidate lists overflow (massive ties: e.g. a synthetic constant image) raises its FOV's retry
// flag; the executor then recomputes that FOV with the float64 strip kernels (executor.cu).  Labels, counts,
// thresholds and tables are therefore bit-identical to the reference's by construction in either case.
#include "internal.cuh"

namespace amt {
namespace dx {

// D(y, x) of one plane in scipy's exact operation order, computed by one CTA of EVAL_WARPS warps; every thread
// returns the value.  The axis-0 filters of the 2 r_hi + 1 columns the axis-1 pass of this one sample needs run 32
// columns per warp, all warps at once: a warp first stages the (2 r_hi + 1) rows x 32 columns of raw samples it
// needs in shared memory (row loads of 64 bytes, all independent: one round trip to L2 instead of one per tap),
// then every lane walks its column in the strip kernels' order (dog.cu: centre tap first, then the tap pairs from
// the outside in).  The narrow filter's columns and rows are a subset of the wide filter's.
// tile: EVAL_WARPS x (2 r_hi + 1) x 32 uint16; g: (2 r_hi + 1) + (2 r_lo + 1) doubles; both shared by the CTA.
constexpr int MAX_R_HI = 64, MAX_R_LO = 4;
constexpr int G_DOUBLES = (2 * MAX_R_HI + 1) + (2 * MAX_R_LO + 1) + 1;
constexpr int TILE_U16 = (2 * MAX_R_HI + 1) * 32;
constexpr int EVAL_WARPS = 5;  // ceil((2 * 64 + 1) / 32) column groups

__device__ double exact_dog_cta(const uint16_t* __restrict__ plane, int h, int w, int y, int x, double scale,
                                const double* __restrict__ whi, int r_hi, const double* __restrict__ wlo, int r_lo,
                                uint16_t* __restrict__ tiles, double* __restrict__ g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_hi = 2 * r_hi + 1;
  uint16_t* tile = tiles + warp * TILE_U16;
  for (int c0 = warp * 32; c0 < n_hi; c0 += EVAL_WARPS * 32) {
    // column c0 + lane of the window = image column clamp(x - r_hi + c0 + lane) (mode='nearest')
    int xc = x - r_hi + c0 + lane;
    xc = xc < 0 ? 0 : (xc > w - 1 ? w - 1 : xc);
    for (int rr = 0; rr < n_hi; ++rr) {
      int yy = y - r_hi + rr;
      yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
      tile[rr * 32 + lane] = __ldg(plane + (int64_t)yy * w + xc);
    }
    __syncwarp();
    if (c0 + lane < n_hi) {
      const uint16_t* col = tile + lane;
      double acc = dmul(dmul((double)col[r_hi * 32], scale), whi[0]);
      for (int j = r_hi; j >= 1; --j) {
        const double a = dmul((double)col[(r_hi - j) * 32], scale), b = dmul((double)col[(r_hi + j) * 32], scale);
        acc = dadd(acc, dmul(dadd(a, b), whi[j]));
      }
      g[c0 + lane] = acc;
      const int dxl = c0 + lane - r_hi;  // this column's offset from x: the narrow filter needs |dxl| <= r_lo
      if (dxl >= -r_lo && dxl <= r_lo) {
        double al = dmul(dmul((double)col[r_hi * 32], scale), wlo[0]);
        for (int j = r_lo; j >= 1; --j) {
          const double a = dmul((double)col[(r_hi - j) * 32], scale), b = dmul((double)col[(r_hi + j) * 32], scale);
          al = dadd(al, dmul(dadd(a, b), wlo[j]));
        }
        g[n_hi + dxl + r_lo] = al;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // axis 1 at column x (every thread redundantly: shared-memory broadcasts)
  double hi = dmul(g[r_hi], whi[0]);
  for (int j = r_hi; j >= 1; --j) hi = dadd(hi, dmul(dadd(g[r_hi - j], g[r_hi + j]), whi[j]));
  const double* gl = g + n_hi;
  double lo = dmul(gl[r_lo], wlo[0]);
  for (int j = r_lo; j >= 1; --j) lo = dadd(lo, dmul(dadd(gl[r_lo - j], gl[r_lo + j]), wlo[j]));
  __syncthreads();
  return dsub(lo, hi);
}

// one CTA per candidate.  Candidates sit in `n_lists` lists of capacity `cap`: list l belongs to image l / lists_per_img,
// holds min(count[l], cap) pixel indices in idx[l * cap ..], and receives the exact values in val[l * cap ..].
__global__ void __launch_bounds__(EVAL_WARPS * 32)
exact_eval_kernel(const uint16_t* __restrict__ in, int64_t img_stride, int h, int w, double scale,
                  const double* __restrict__ hw_hi, int r_hi, const double* __restrict__ hw_lo, int r_lo,
                  const uint32_t* __restrict__ count, const uint32_t* __restrict__ idx, double* __restrict__ val,
                  int n_lists, int lists_per_img, int cap) {
  __shared__ double s_whi[MAX_R_HI + 1], s_wlo[MAX_R_LO + 1];
  __shared__ double s_g[G_DOUBLES];
  __shared__ uint16_t s_tile[EVAL_WARPS * TILE_U16];
  for (int i = threadIdx.x; i <= r_hi; i += blockDim.x) s_whi[i] = hw_hi[i];
  for (int i = threadIdx.x; i <= r_lo; i += blockDim.x) s_wlo[i] = hw_lo[i];
  __syncthreads();
  // the lists are short and uneven: walk them as one concatenated sequence so that the CTAs share the work evenly
  int64_t base = 0;
  const int64_t gb = blockIdx.x, n_blocks = gridDim.x;
  for (int l = 0; l < n_lists; ++l) {
    const uint32_t cl = count[l];
    const int c = (int)(cl < (uint32_t)cap ? cl : (uint32_t)cap);
    // candidate k of list l has global number base + k; this CTA takes those congruent to its index
    int64_t k0 = (gb - base) % n_blocks;
    if (k0 < 0) k0 += n_blocks;
    for (int64_t k = k0; k < c; k += n_blocks) {  // block-uniform
      const uint32_t pix = idx[(int64_t)l * cap + k];
      const int y = (int)(pix / (uint32_t)w), x = (int)(pix - (uint32_t)y * (uint32_t)w);
      const uint16_t* plane = in + (int64_t)(l / lists_per_img) * img_stride;
      const double d = exact_dog_cta(plane, h, w, y, x, scale, s_whi, r_hi, s_wlo, r_lo, s_tile, s_g);
      if (threadIdx.x == 0) val[(int64_t)l * cap + k] = d;
    }
    base += c;
  }
}

// Windows of the order statistics: list (img, k), k < N_WIN; centre[k] = the approximate statistic (k < 6: the six
// percentile ranks; 6, 7: min, max).  One pass over D': samples below the window are counted, samples inside it listed.
constexpr int N_WIN = 8;
}  // namespace dx
int g_dx_collect_threads = 1024;  // amt_tune "dx_collect_threads": threads per CTA of rank_collect_buckets_kernel (256 or 1024)
namespace dx {

__global__ void __launch_bounds__(256)
rank_collect_kernel(const double* __restrict__ dog, int64_t img_stride, int64_t n, const double* __restrict__ stats,
                    int64_t stats_stride, const uint64_t* __restrict__ mm, int64_t mm_stride, double eps2,
                    uint32_t* __restrict__ below, uint32_t* __restrict__ count, uint32_t* __restrict__ idx, int cap) {
  __shared__ uint32_t s_below[N_WIN];
  const int64_t img = blockIdx.y;
  const double* src = dog + img * img_stride;
  double lo[N_WIN], hi[N_WIN];
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) {
    const double c = k < 6 ? stats[img * stats_stride + k] : key_to_f64(mm[img * mm_stride + (k - 6)]);
    lo[k] = c - eps2;
    hi[k] = c + eps2;
  }
  if (threadIdx.x < N_WIN) s_below[threadIdx.x] = 0;
  __syncthreads();
  uint32_t nb[N_WIN];
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) nb[k] = 0;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const double d = src[i];
#pragma unroll
    for (int k = 0; k < N_WIN; ++k) {
      if (d < lo[k]) {
        nb[k] += 1;
      } else if (d <= hi[k]) {
        const uint32_t pos = atomicAdd(&count[img * N_WIN + k], 1u);
        if (pos < (uint32_t)cap) idx[(img * N_WIN + k) * cap + pos] = (uint32_t)i;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) {
    const int s = warp_sum_i32((int)nb[k]);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_below[k], (uint32_t)s);
  }
  __syncthreads();
  if (threadIdx.x < N_WIN && s_below[threadIdx.x]) atomicAdd(&below[img * N_WIN + threadIdx.x], s_below[threadIdx.x]);
}

// Variant that reads the 2-byte selection buckets the DoG pass wrote next to the plane (bucket12() is monotone, so
// a whole bucket is below / above a window whenever its code is below / above the codes of the window's ends) and
// touches the 8-byte samples of the few buckets a window end falls into only.  A per-block table maps a bucket code
// to eight bytes, one per window: 1 = certainly below, 2 = look at the value; the "below" bytes of up to 255 samples
// add up in one 64-bit register.
__global__ void __launch_bounds__(1024)
rank_collect_buckets_kernel(const double* __restrict__ dog, const uint16_t* __restrict__ buckets, int64_t img_stride, int64_t n,
                            const double* __restrict__ stats, int64_t stats_stride, const uint64_t* __restrict__ mm,
                            int64_t mm_stride, double eps2, uint32_t* __restrict__ below, uint32_t* __restrict__ count,
                            uint32_t* __restrict__ idx, int cap) {
  __shared__ uint64_t s_lut[4096];
  __shared__ uint32_t s_below[N_WIN];
  __shared__ double s_lo[N_WIN], s_hi[N_WIN];
  const int64_t img = blockIdx.y;
  const double* src = dog + img * img_stride;
  const uint16_t* bsrc = buckets + img * img_stride;
  if (threadIdx.x < N_WIN) {
    const int k = threadIdx.x;
    const double c = k < 6 ? stats[img * stats_stride + k] : key_to_f64(mm[img * mm_stride + (k - 6)]);
    s_lo[k] = c - eps2;
    s_hi[k] = c + eps2;
    s_below[k] = 0;
  }
  __syncthreads();
  uint32_t blo[N_WIN], bhi[N_WIN];
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) blo[k] = bucket12(s_lo[k]), bhi[k] = bucket12(s_hi[k]);
  for (int code = threadIdx.x; code < 4096; code += blockDim.x) {
    uint64_t e = 0;
#pragma unroll
    for (int k = 0; k < N_WIN; ++k) e |= (uint64_t)((uint32_t)code < blo[k] ? 1u : ((uint32_t)code <= bhi[k] ? 2u : 0u)) << (8 * k);
    s_lut[code] = e;
  }
  __syncthreads();
  uint32_t nb[N_WIN];
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) nb[k] = 0;
  auto slow = [&](uint64_t e, int64_t i) {  // a sample whose bucket holds a window end: decided on its value
    const double d = src[i];
#pragma unroll
    for (int k = 0; k < N_WIN; ++k) {
      if (((e >> (8 * k)) & 2u) == 0) continue;
      if (d < s_lo[k]) {
        nb[k] += 1;
      } else if (d <= s_hi[k]) {
        const uint32_t pos = atomicAdd(&count[img * N_WIN + k], 1u);
        if (pos < (uint32_t)cap) idx[(img * N_WIN + k) * cap + pos] = (uint32_t)i;
      }
    }
  };
  const int64_t n8 = n >> 3;  // n % 8 == 0 (checked by the caller)
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  int pending = 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n8; q += step) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(bsrc) + q);
    const uint32_t w4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint32_t code = (w4[t >> 1] >> (16 * (t & 1))) & 0xffffu;
      const uint64_t e = s_lut[code & 4095u];
      acc += e & 0x0101010101010101ull;
      if (e & 0x0202020202020202ull) slow(e, 8 * q + t);
    }
    pending += 8;
    if (pending > 240) {  // the byte counters hold 255
#pragma unroll
      for (int k = 0; k < N_WIN; ++k) nb[k] += (uint32_t)(acc >> (8 * k)) & 0xffu;
      acc = 0;
      pending = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < N_WIN; ++k) {
    nb[k] += (uint32_t)(acc >> (8 * k)) & 0xffu;
    const int sum = warp_sum_i32((int)nb[k]);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&s_below[k], (uint32_t)sum);
  }
  __syncthreads();
  if (threadIdx.x < N_WIN && s_below[threadIdx.x]) atomicAdd(&below[img * N_WIN + threadIdx.x], s_below[threadIdx.x]);
}

// One block per (image, window): the exact statistic is the (rank - below)-th smallest exact value of the list.
__global__ void __launch_bounds__(256)
rank_resolve_kernel(const uint32_t* __restrict__ below, const uint32_t* __restrict__ count, const double* __restrict__ val,
                    int cap, const int64_t* __restrict__ ranks /* 6 */, int64_t n, double* __restrict__ stats,
                    int64_t stats_stride, uint64_t* __restrict__ mm, int64_t mm_stride, int32_t* __restrict__ retry) {
  extern __shared__ double s_val[];
  const int64_t img = blockIdx.y;
  const int k = blockIdx.x;
  const uint32_t c = count[img * N_WIN + k];
  const int64_t rank = k < 6 ? ranks[k] : (k == 6 ? 0 : n - 1);
  const int64_t want = rank - (int64_t)below[img * N_WIN + k];
  if (c > (uint32_t)cap || want < 0 || want >= (int64_t)c) {  // list overflow / an inconsistent window: recompute this image
    if (threadIdx.x == 0) retry[img] = 1;
    return;
  }
  const double* v = val + (img * N_WIN + k) * cap;
  for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) s_val[i] = v[i];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
    const double x = s_val[i];
    uint32_t less = 0, equal_before = 0;
    for (uint32_t j = 0; j < c; ++j) {
      const double y = s_val[j];
      less += y < x ? 1u : 0u;
      equal_before += (y == x && j < i) ? 1u : 0u;
    }
    if ((int64_t)(less + equal_before) == want) {
      if (k < 6)
        stats[img * stats_stride + k] = x;
      else
        mm[img * mm_stride + (k - 6)] = f64_to_key(x);
    }
  }
}

int rank_exact(const uint16_t* in, int64_t in_img_stride, const double* dog, const uint16_t* buckets, int64_t dog_img_stride,
               int64_t n_img, int h, int w, double scale, const double* hw_hi, int r_hi, const double* hw_lo, int r_lo, double eps,
               const int64_t* ranks_dev, double* stats, int64_t stats_stride, uint64_t* mm, int64_t mm_stride,
               uint32_t* scratch_u32, double* scratch_val, int cap, int32_t* retry, cudaStream_t st) {
  if (r_hi > MAX_R_HI || r_lo > MAX_R_LO || cap < 1 || cap > 4096) return AMT_ERR_UNSUPPORTED;
  const int64_t n = (int64_t)h * w;
  uint32_t* below = scratch_u32;
  uint32_t* count = below + n_img * N_WIN;
  uint32_t* idx = count + n_img * N_WIN;
  AMT_CUDA_TRY(cudaMemsetAsync(below, 0, (size_t)n_img * N_WIN * 2 * sizeof(uint32_t), st));
  int64_t blocks = ceil_div(n, 256 * 8);
  const int64_t cap_blocks = ceil_div((int64_t)kNumSMs * 8, n_img);
  if (blocks > cap_blocks) blocks = cap_blocks;
  if (buckets != nullptr && n % 8 == 0 && ((uintptr_t)buckets % 16) == 0 && (dog_img_stride % 8) == 0) {
    // every CTA builds a 4096-entry table first: fewer, larger CTAs (g_dx_collect_threads = 1024: two per SM) spend less on it
    const int thr = g_dx_collect_threads;
    int64_t bb = ceil_div(n, (int64_t)thr * 8);
    const int64_t cb = ceil_div((int64_t)kNumSMs * (2048 / thr), n_img);
    if (bb > cb) bb = cb;
    rank_collect_buckets_kernel<<<dim3((unsigned)bb, (unsigned)n_img), thr, 0, st>>>(
        dog, buckets, dog_img_stride, n, stats, stats_stride, mm, mm_stride, 2.0 * eps, below, count, idx, cap);
  } else
    rank_collect_kernel<<<dim3((unsigned)blocks, (unsigned)n_img), 256, 0, st>>>(dog, dog_img_stride, n, stats, stats_stride, mm,
                                                                               mm_stride, 2.0 * eps, below, count, idx, cap);
  AMT_LAUNCH_CHECK();
  exact_eval_kernel<<<kNumSMs * 4, EVAL_WARPS * 32, 0, st>>>(in, in_img_stride, h, w, scale, hw_hi, r_hi, hw_lo, r_lo, count, idx,
                                                          scratch_val, (int)(n_img * N_WIN), N_WIN, cap);
  AMT_LAUNCH_CHECK();
  rank_resolve_kernel<<<dim3(N_WIN, (unsigned)n_img), 256, (size_t)cap * sizeof(double), st>>>(
      below, count, scratch_val, cap, ranks_dev, n, stats, stats_stride, mm, mm_stride, retry);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int exact_eval(const uint16_t* in, int64_t in_img_stride, int h, int w, double scale, const double* hw_hi, int r_hi,
               const double* hw_lo, int r_lo, const uint32_t* count, const uint32_t* idx, double* val, int n_lists, int cap,
               cudaStream_t st) {
  if (r_hi > MAX_R_HI || r_lo > MAX_R_LO) return AMT_ERR_UNSUPPORTED;
  exact_eval_kernel<<<kNumSMs * 4, EVAL_WARPS * 32, 0, st>>>(in, in_img_stride, h, w, scale, hw_hi, r_hi, hw_lo, r_lo, count, idx,
                                                          val, n_lists, 1, cap);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace dx
}  // namespace amt
