// Exact order statistics (np.percentile's floor/ceil ranks), plane min/max, uint16 histograms.
//
// Reference path: operations.py:47 and :94 np.percentile(..., method='linear') [3p]; the
// min()==max() guards at operations.py:43 and :201.  SURVEY.md 8a item 3.
//
// float64 selection is exact (no approximate quantiles): one streaming pass builds a
// 4096-bin histogram per plane with a MONOTONE value->bin map (any monotone map partitions the
// order statistics correctly, so floating-point rounding in the map cannot change the answer),
// a one-block "locate" kernel finds the bin of every requested rank, a second streaming pass
// compacts only the elements of those bins, and a per-rank block finishes with an 8x8-bit MSD
// radix select over order-preserving 64-bit keys.  Elements equal to the plane min / max are
// counted apart (clipped-at-zero and saturated planes put most pixels there).  If a bin holds
// more than the candidate capacity the per-rank block simply radix-selects the whole plane
// (slow, always correct).  All of it is HBM-bound streaming: 2 reads of the plane.
//
// uint16 selection / histogram: one pass with a block-private 65536-bin histogram of packed
// 16-bit counters in 128 KB of shared memory (a block never sees more than 32768 samples
// between flushes, so a counter cannot overflow).

#include "common.cuh"

namespace amt {

extern int g_pass_ctas;  // CTAs per SM of the streaming passes (amt_tune "pass_ctas"), core.cu

constexpr int SEL_NB = 4096;
constexpr int SEL_CHUNK = 2048;  // elements per block in the streaming kernels (256 threads x 8)

struct SelDesc {
  const double* src;
  int64_t n;
  int64_t k;
  double direct;  // result when mode == 0
  int32_t mode;   // 0 direct, 1 candidate list, 2 whole plane
  int32_t list;   // candidate list id (mode 1)
};

struct SelPlane {
  double mn, mx, scale;
  uint32_t eq_min, eq_max;
  int32_t n_lists;
  int32_t list_bin[AMT_MAX_RANKS];
  uint32_t list_fill[AMT_MAX_RANKS];  // running fill counters of the candidate lists
  SelDesc desc[AMT_MAX_RANKS];
};

struct SelRanks {
  int64_t r[AMT_MAX_RANKS];
  int n;
};

__device__ __forceinline__ int sel_bin(double x, double mn, double scale) {
  // monotone in x; NaN / inf products collapse to bin 0 / NB-1
  const double t = dmul(dsub(x, mn), scale);
  const int b = __double2int_rz(t);  // saturates (NaN -> 0): no separate range test on the FP64 pipe
  return b < 0 ? 0 : (b > SEL_NB - 1 ? SEL_NB - 1 : b);
}

// ------------------------------------------------------------------ min / max
template <typename T>
__device__ __forceinline__ uint64_t to_key(T v);
template <>
__device__ __forceinline__ uint64_t to_key<double>(double v) {
  return f64_to_key(v);
}
template <>
__device__ __forceinline__ uint64_t to_key<uint16_t>(uint16_t v) {
  return (uint64_t)v;
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ data, int64_t n, uint64_t* __restrict__ mm) {
  const T* p = data + (int64_t)blockIdx.y * n;
  uint64_t kmin = ~0ull, kmax = 0ull;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += step) {
    const uint64_t k = to_key<T>(p[i]);
    kmin = k < kmin ? k : kmin;
    kmax = k > kmax ? k : kmax;
  }
  __shared__ uint64_t s[16];
  kmin = warp_min_u64(kmin);
  kmax = warp_max_u64(kmax);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s[warp] = kmin;
    s[8 + warp] = kmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      kmin = s[i] < kmin ? s[i] : kmin;
      kmax = s[8 + i] > kmax ? s[8 + i] : kmax;
    }
    atomicMin((unsigned long long*)&mm[2 * blockIdx.y], (unsigned long long)kmin);
    atomicMax((unsigned long long*)&mm[2 * blockIdx.y + 1], (unsigned long long)kmax);
  }
}

__global__ void minmax_decode_kernel(const uint64_t* __restrict__ mm, int is_f64, int64_t n2, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) out[i] = is_f64 ? key_to_f64(mm[i]) : (double)mm[i];
}

int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st);  // gauss.cu

template <typename T>
static int minmax_launch(const T* data, int64_t n_img, int64_t n, uint64_t* mm, cudaStream_t st) {
  if (!data || !mm || n_img <= 0 || n <= 0 || n_img > 65535) return AMT_ERR_INVALID;
  AMT_TRY(minmax_init(mm, n_img, st));
  int64_t bx = ceil_div(n, 256 * 8);
  const int64_t cap = ceil_div((int64_t)kNumSMs * 8, n_img);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  minmax_kernel<T><<<dim3((unsigned)bx, (unsigned)n_img), 256, 0, st>>>(data, n, mm);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

// ------------------------------------------------------------------ f64 selection
__global__ void sel_prepare_kernel(const uint64_t* __restrict__ mm, SelPlane* __restrict__ planes,
                                   uint32_t* __restrict__ hist, int64_t n_img) {
  // one block per plane: decode min/max, zero the histogram and counters
  const int64_t img = blockIdx.x;
  uint32_t* h = hist + img * SEL_NB;
  for (int i = threadIdx.x; i < SEL_NB; i += blockDim.x) h[i] = 0;
  if (threadIdx.x == 0) {
    SelPlane& p = planes[img];
    const double mn = key_to_f64(mm[2 * img]), mx = key_to_f64(mm[2 * img + 1]);
    p.mn = mn;
    p.mx = mx;
    const double range = dsub(mx, mn);
    p.scale = (range > 0.0) ? ddiv((double)SEL_NB, range) : 0.0;
    p.eq_min = 0;
    p.eq_max = 0;
    p.n_lists = 0;
    for (int i = 0; i < AMT_MAX_RANKS; ++i) p.list_fill[i] = 0;
  }
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(const double* __restrict__ data, int64_t n, SelPlane* __restrict__ planes, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[SEL_NB];
  __shared__ uint32_t s_eq[2];
  const int64_t img = blockIdx.y;
  const double* p = data + img * n;
  for (int i = threadIdx.x; i < SEL_NB; i += 256) sh[i] = 0;
  if (threadIdx.x < 2) s_eq[threadIdx.x] = 0;
  __syncthreads();
  const double mn = planes[img].mn, mx = planes[img].mx, scale = planes[img].scale;
  uint32_t eqmin = 0, eqmax = 0;
  const int64_t step = (int64_t)gridDim.x * 256;
  // four independent loads in flight per thread (the pass is HBM-bound)
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += 4 * step) {
    double x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = (i + e * step < n) ? __ldg(p + i + e * step) : mn;
    const int valid = (int)((n - i + step - 1) / step);  // samples of this iteration inside the plane
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < valid) {
        if (x[e] == mn) {
          ++eqmin;
        } else if (x[e] == mx) {
          ++eqmax;
        } else {
          atomicAdd(&sh[sel_bin(x[e], mn, scale)], 1u);
        }
      }
    }
  }
  eqmin = (uint32_t)warp_sum_i32((int)eqmin);
  eqmax = (uint32_t)warp_sum_i32((int)eqmax);
  if ((threadIdx.x & 31) == 0) {
    if (eqmin) atomicAdd(&s_eq[0], eqmin);
    if (eqmax) atomicAdd(&s_eq[1], eqmax);
  }
  __syncthreads();
  uint32_t* h = hist + img * SEL_NB;
  for (int i = threadIdx.x; i < SEL_NB; i += 256) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&h[i], c);
  }
  if (threadIdx.x == 0) {
    if (s_eq[0]) atomicAdd(&planes[img].eq_min, s_eq[0]);
    if (s_eq[1]) atomicAdd(&planes[img].eq_max, s_eq[1]);
  }
}

// one block of 1024 threads per plane
__global__ void __launch_bounds__(1024)
sel_locate_kernel(const double* __restrict__ data, int64_t n, SelRanks ranks, SelPlane* __restrict__ planes,
                  const uint32_t* __restrict__ hist, double* __restrict__ cand, int64_t cap) {
  __shared__ uint32_t s_excl[SEL_NB];
  __shared__ uint32_t s_warp[32];
  __shared__ int s_bin[AMT_MAX_RANKS];
  const int64_t img = blockIdx.x;
  const uint32_t* h = hist + img * SEL_NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // exclusive scan of 4096 bins: 4 consecutive bins per thread
  uint32_t v[4], tsum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = h[tid * 4 + i];
    tsum += v[i];
  }
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
  }
  if (tid < AMT_MAX_RANKS) s_bin[tid] = -1;
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - tsum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s_excl[tid * 4 + i] = run;
    run += v[i];
  }
  __syncthreads();
  SelPlane& p = planes[img];
  const int64_t eqmin = p.eq_min, eqmax = p.eq_max;
  // which bin holds each interior rank
  for (int q = 0; q < ranks.n; ++q) {
    const int64_t k = ranks.r[q];
    if (k < eqmin || k >= n - eqmax) continue;
    const uint32_t k2 = (uint32_t)(k - eqmin);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = tid * 4 + i;
      const uint32_t lo = s_excl[b];
      if (v[i] != 0 && k2 >= lo && k2 < lo + v[i]) s_bin[q] = b;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n_lists = 0;
    for (int q = 0; q < ranks.n; ++q) {
      SelDesc d;
      const int64_t k = ranks.r[q];
      d.src = data + img * n;
      d.n = n;
      d.k = k;
      d.direct = 0.0;
      d.list = -1;
      if (k < eqmin) {
        d.mode = 0;
        d.direct = p.mn;
      } else if (k >= n - eqmax) {
        d.mode = 0;
        d.direct = p.mx;
      } else {
        const int b = s_bin[q];
        const uint32_t cnt = h[b];
        if ((int64_t)cnt <= cap) {
          int l = -1;
          for (int j = 0; j < n_lists; ++j)
            if (p.list_bin[j] == b) l = j;
          if (l < 0) {
            l = n_lists++;
            p.list_bin[l] = b;
          }
          d.mode = 1;
          d.list = l;
          d.src = cand + (img * AMT_MAX_RANKS + l) * cap;
          d.n = cnt;
          d.k = (k - eqmin) - (int64_t)s_excl[b];
        } else {
          d.mode = 2;  // whole-plane radix select with the original rank
        }
      }
      p.desc[q] = d;
    }
    p.n_lists = n_lists;
  }
}

__global__ void __launch_bounds__(256)
sel_compact_kernel(const double* __restrict__ data, int64_t n, SelPlane* __restrict__ planes,
                   double* __restrict__ cand, int64_t cap) {
  __shared__ uint32_t s_cnt[AMT_MAX_RANKS];
  __shared__ uint32_t s_base[AMT_MAX_RANKS];
  __shared__ uint8_t s_lut[SEL_NB];  // bin -> candidate list id (0xff: not wanted)
  const int64_t img = blockIdx.y;
  SelPlane& pl = planes[img];
  const int n_lists = pl.n_lists;
  if (n_lists == 0) return;
  for (int i = threadIdx.x; i < SEL_NB / 4; i += 256) reinterpret_cast<uint32_t*>(s_lut)[i] = 0xffffffffu;
  if (threadIdx.x < AMT_MAX_RANKS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < n_lists) s_lut[pl.list_bin[threadIdx.x]] = (uint8_t)threadIdx.x;
  __syncthreads();
  const double mn = pl.mn, mx = pl.mx, scale = pl.scale;
  const double* p = data + img * n;
  const int64_t base = (int64_t)blockIdx.x * SEL_CHUNK;
  double x[8];
  int slot[8];  // (list << 24) | local position, or -1
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int64_t i = base + e * 256 + threadIdx.x;
    x[e] = (i < n) ? __ldg(p + i) : mn;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    slot[e] = -1;
    if (x[e] != mn && x[e] != mx) {  // also rejects the padding of a partial chunk
      const int l = s_lut[sel_bin(x[e], mn, scale)];
      if (l != 0xff) slot[e] = (l << 24) | (int)atomicAdd(&s_cnt[l], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < n_lists && s_cnt[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(&pl.list_fill[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (slot[e] >= 0) {
      const int l = slot[e] >> 24;
      const int64_t pos = (int64_t)s_base[l] + (slot[e] & 0xffffff);
      if (pos < cap) cand[(img * AMT_MAX_RANKS + l) * cap + pos] = x[e];
    }
  }
}

__global__ void bucket12_kernel(const double* __restrict__ data, uint16_t* __restrict__ out, int64_t n) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = (uint16_t)bucket12(data[i]);
}

// ---- bucketed variant: the DoG's second pass has already written bucket12(x) (uint16) for every
// sample, so the histogram pass reads 2 instead of 8 bytes per sample and the compaction pass reads
// the buckets plus only the few float64 values whose bucket holds a wanted rank.  Minimum / maximum
// ties are not counted apart here (eq_min = eq_max = 0): every sample goes through its bucket.
__global__ void __launch_bounds__(256)
sel_hist_buckets_kernel(const uint16_t* __restrict__ buckets, int64_t n, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[SEL_NB];
  const int64_t img = blockIdx.y;
  const uint16_t* p = buckets + img * n;
  for (int i = threadIdx.x; i < SEL_NB; i += 256) sh[i] = 0;
  __syncthreads();
  const int64_t n8 = n >> 3;  // 8 buckets per 16-byte load (planes are 16-byte aligned: n % 8 == 0)
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n8; q += 2 * step) {
    const bool second = q + step < n8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p) + q);
    uint4 b = make_uint4(0, 0, 0, 0);
    if (second) b = __ldg(reinterpret_cast<const uint4*>(p) + q + step);
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&sh[wa[e] & 0xfffu], 1u);
      atomicAdd(&sh[(wa[e] >> 16) & 0xfffu], 1u);
    }
    if (second) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(&sh[wb[e] & 0xfffu], 1u);
        atomicAdd(&sh[(wb[e] >> 16) & 0xfffu], 1u);
      }
    }
  }
  __syncthreads();
  uint32_t* h = hist + img * SEL_NB;
  for (int i = threadIdx.x; i < SEL_NB; i += 256) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&h[i], c);
  }
}

// grid (chunks, planes); block 256.  No block barriers: a warp takes 32 x 8 consecutive samples per
// iteration (one 16-byte bucket load per lane, two iterations in flight), and for every candidate list
// the lanes' counts are scanned with shuffles, one lane reserves the warp's range with a single global
// atomic, and the wanted float64 values (a few per cent of the samples) are gathered and appended.
__global__ void __launch_bounds__(256)
sel_compact_buckets_kernel(const uint16_t* __restrict__ buckets, const double* __restrict__ data, int64_t n,
                           SelPlane* __restrict__ planes, double* __restrict__ cand, int64_t cap) {
  __shared__ uint8_t s_lut[SEL_NB];  // bucket -> candidate list id (0xff: not wanted)
  const int64_t img = blockIdx.y;
  SelPlane& pl = planes[img];
  const int n_lists = pl.n_lists;
  if (n_lists == 0) return;
  for (int i = threadIdx.x; i < SEL_NB / 4; i += 256) reinterpret_cast<uint32_t*>(s_lut)[i] = 0xffffffffu;
  __syncthreads();
  if (threadIdx.x < n_lists) s_lut[pl.list_bin[threadIdx.x]] = (uint8_t)threadIdx.x;
  __syncthreads();
  const uint16_t* kp = buckets + img * n;
  const double* dp = data + img * n;
  const int lane = threadIdx.x & 31;
  const int64_t n8 = n >> 3;
  const int64_t warp0 = ((int64_t)blockIdx.x * 256 + (threadIdx.x & ~31));  // first 8-sample group of this warp
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t q0 = warp0; q0 < n8; q0 += 2 * step) {  // warp-uniform trip count
    uint4 a[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t q = q0 + u * step + lane;
      a[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (q < n8) a[u] = __ldg(reinterpret_cast<const uint4*>(kp) + q);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t q = q0 + u * step + lane;
      const uint32_t w[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
      uint32_t lists = 0;  // 4 bits per sample: list id, 0xf = not wanted
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t bk = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
        uint32_t l = 0xf;
        if (bk < (uint32_t)SEL_NB) {  // the padding value 0xffff is not a bucket
          const uint32_t t = s_lut[bk];
          l = t == 0xff ? 0xf : t;
        }
        lists |= l << (4 * e);
      }
      if (!__any_sync(0xffffffffu, lists != 0xffffffffu)) continue;
      if (n_lists <= 7) {
        // every list at once: the lanes' per-list counts ride in one 64-bit word (9 bits per list: a warp holds at most
        // 256 samples), ONE shuffle scan serves them all, lane l reserves list l's range with its own atomic
        unsigned long long cnt = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t l = (lists >> (4 * e)) & 0xfu;
          cnt += l != 0xfu ? 1ull << (9 * l) : 0ull;
        }
        unsigned long long incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (lane < n_lists) {
          const uint32_t tl = (uint32_t)(total >> (9 * lane)) & 0x1ffu;
          if (tl) base = atomicAdd(&pl.list_fill[lane], tl);
        }
        unsigned long long excl = incl - cnt;  // samples of every list in the lanes before this one
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t l = (lists >> (4 * e)) & 0xfu;
          const uint32_t b = __shfl_sync(0xffffffffu, base, l & 7u);  // every lane takes part; l == 0xf reads lane 7, unused
          if (l != 0xfu) {
            const int64_t pos = (int64_t)b + (int64_t)((excl >> (9 * l)) & 0x1ffull);
            if (pos < cap) cand[(img * AMT_MAX_RANKS + l) * cap + pos] = __ldg(dp + 8 * q + e);
            excl += 1ull << (9 * l);
          }
        }
        continue;
      }
      for (int l = 0; l < n_lists; ++l) {
        uint32_t mine = 0;  // bit e: sample e goes to list l
#pragma unroll
        for (int e = 0; e < 8; ++e) mine |= (((lists >> (4 * e)) & 0xfu) == (uint32_t)l) ? (1u << e) : 0u;
        const int c = __popc(mine);
        if (!__any_sync(0xffffffffu, c != 0)) continue;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(&pl.list_fill[l], (uint32_t)incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        int64_t pos = (int64_t)base + (incl - c);
        double* dst = cand + (img * AMT_MAX_RANKS + l) * cap;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if ((mine >> e) & 1u) {
            if (pos < cap) dst[pos] = __ldg(dp + 8 * q + e);
            ++pos;
          }
        }
      }
    }
  }
}

// one block per (rank, plane): 8 passes of 8 bits, most significant first
__global__ void __launch_bounds__(1024)
sel_radix_kernel(const SelPlane* __restrict__ planes, int n_ranks, double* __restrict__ out) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ int64_t s_k;
  const int q = blockIdx.x;
  const int64_t img = blockIdx.y;
  const SelDesc d = planes[img].desc[q];
  if (d.mode == 0) {
    if (threadIdx.x == 0) out[img * n_ranks + q] = d.direct;
    return;
  }
  const double* __restrict__ src = d.src;
  const int64_t n = d.n;
  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_k = d.k;
  }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t prefix = s_prefix;
    const uint64_t himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
      const uint64_t key = f64_to_key(src[i]);
      if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // warp 0: each lane owns 8 consecutive digits
      const int lane = threadIdx.x;
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        c[i] = s_hist[lane * 8 + i];
        tot += c[i];
      }
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int64_t k = s_k;
      int64_t run = (int64_t)(incl - tot);
      __syncwarp();
      if (k >= run && k < run + (int64_t)tot) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k >= run && k < run + (int64_t)c[i]) {
            s_prefix = prefix | ((uint64_t)(lane * 8 + i) << shift);
            s_k = k - run;
          }
          run += c[i];
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[img * n_ranks + q] = key_to_f64(s_prefix);
}

static int64_t sel_cap(int64_t n) {
  int64_t cap = n / 8;
  if (cap < 4096) cap = 4096;
  return (cap + 255) / 256 * 256;
}

// ------------------------------------------------------------------ uint16 histogram / selection
constexpr int U16_CHUNK = 32768;  // samples per block: a packed 16-bit counter cannot overflow

__global__ void __launch_bounds__(1024)
hist_u16_kernel(const uint16_t* __restrict__ data, int64_t n, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh[];  // 32768 words = 65536 packed 16-bit counters
  const int64_t img = blockIdx.y;
  const uint16_t* p = data + img * n;
  for (int i = threadIdx.x; i < 32768; i += 1024) sh[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * U16_CHUNK;
  const int64_t end = base + U16_CHUNK < n ? base + U16_CHUNK : n;
  for (int64_t i = base + threadIdx.x; i < end; i += 1024) {
    const uint32_t v = p[i];
    atomicAdd(&sh[v >> 1], 1u << ((v & 1u) * 16));
  }
  __syncthreads();
  uint32_t* h = hist + img * 65536;
  for (int i = threadIdx.x; i < 32768; i += 1024) {
    const uint32_t wv = sh[i];
    if (wv) {
      if (wv & 0xffffu) atomicAdd(&h[2 * i], wv & 0xffffu);
      if (wv >> 16) atomicAdd(&h[2 * i + 1], wv >> 16);
    }
  }
}

int hist_u16(const uint16_t* data, int64_t n_img, int64_t n, uint32_t* hist, cudaStream_t st) {
  if (!data || !hist || n_img <= 0 || n <= 0 || n_img > 65535) return AMT_ERR_INVALID;
  AMT_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)n_img * 65536 * sizeof(uint32_t), st));
  const size_t smem = 32768 * sizeof(uint32_t);
  AMT_CUDA_TRY(cudaFuncSetAttribute(hist_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hist_u16_kernel<<<dim3((unsigned)ceil_div(n, U16_CHUNK), (unsigned)n_img), 1024, smem, st>>>(data, n, hist);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

// one block of 1024 threads per plane: scan 65536 bins, emit the value holding each rank
__global__ void __launch_bounds__(1024)
sel_u16_locate_kernel(const uint32_t* __restrict__ hist, SelRanks ranks, double* __restrict__ out) {
  __shared__ uint32_t s_warp[32];
  const int64_t img = blockIdx.x;
  const uint32_t* h = hist + img * 65536;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t tsum = 0;
  for (int i = 0; i < 64; ++i) tsum += h[tid * 64 + i];
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  const int64_t lo = (int64_t)s_warp[warp] + incl - tsum;
  for (int q = 0; q < ranks.n; ++q) {
    const int64_t k = ranks.r[q];
    if (k >= lo && k < lo + (int64_t)tsum) {
      int64_t run = lo;
      for (int i = 0; i < 64; ++i) {
        const uint32_t c = h[tid * 64 + i];
        if (k >= run && k < run + (int64_t)c) {
          out[img * ranks.n + q] = (double)(tid * 64 + i);
          break;
        }
        run += c;
      }
    }
  }
}

}  // namespace amt

extern "C" {

int amt_minmax_f64(const double* data, int64_t n_img, int64_t n, uint64_t* mm, amt_stream_t stream) {
  return amt::minmax_launch<double>(data, n_img, n, mm, amt::as_stream(stream));
}
int amt_minmax_u16(const uint16_t* data, int64_t n_img, int64_t n, uint64_t* mm, amt_stream_t stream) {
  return amt::minmax_launch<uint16_t>(data, n_img, n, mm, amt::as_stream(stream));
}
int amt_minmax_decode(const uint64_t* mm, int is_f64, int64_t n_img, double* out, amt_stream_t stream) {
  using namespace amt;
  if (!mm || !out || n_img <= 0) return AMT_ERR_INVALID;
  minmax_decode_kernel<<<(unsigned)ceil_div(2 * n_img, 256), 256, 0, as_stream(stream)>>>(mm, is_f64, 2 * n_img, out);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

static size_t sel_head_bytes(int64_t n_img) {
  const size_t b = (size_t)n_img * (sizeof(amt::SelPlane) + amt::SEL_NB * sizeof(uint32_t));
  return (b + 255) / 256 * 256;
}

// plane descriptors + histograms + AMT_MAX_RANKS candidate lists of n/8 samples per plane
size_t amt_select_f64_scratch_bytes(int64_t n_img, int64_t n) {
  return sel_head_bytes(n_img) + (size_t)n_img * AMT_MAX_RANKS * (size_t)amt::sel_cap(n) * sizeof(double);
}

int amt_select_f64(const double* data, int64_t n_img, int64_t n, const int64_t* ranks_host, int n_ranks,
                   const uint64_t* minmax_keys, double* out_vals, void* scratch, size_t scratch_bytes,
                   amt_stream_t stream) {
  using namespace amt;
  if (!data || !ranks_host || !minmax_keys || !out_vals || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || n <= 0 || n_ranks <= 0 || n_ranks > AMT_MAX_RANKS || n_img > 65535) return AMT_ERR_INVALID;
  if (n >= (1ll << 32)) return AMT_ERR_CAPACITY;
  if (scratch_bytes < amt_select_f64_scratch_bytes(n_img, n)) return AMT_ERR_CAPACITY;
  SelRanks ranks;
  ranks.n = n_ranks;
  for (int i = 0; i < n_ranks; ++i) {
    if (ranks_host[i] < 0 || ranks_host[i] >= n) return AMT_ERR_INVALID;
    ranks.r[i] = ranks_host[i];
  }
  cudaStream_t st = as_stream(stream);
  const int64_t cap = sel_cap(n);
  char* base = (char*)scratch;
  SelPlane* planes = (SelPlane*)base;
  uint32_t* hist = (uint32_t*)(base + (size_t)n_img * sizeof(SelPlane));
  double* cand = (double*)(base + sel_head_bytes(n_img));

  sel_prepare_kernel<<<(unsigned)n_img, 256, 0, st>>>(minmax_keys, planes, hist, n_img);
  AMT_LAUNCH_CHECK();
  int64_t bx = ceil_div(n, 256 * 16);
  const int64_t capb = ceil_div((int64_t)kNumSMs * g_pass_ctas, n_img);
  if (bx > capb) bx = capb;
  if (bx < 1) bx = 1;
  sel_hist_kernel<<<dim3((unsigned)bx, (unsigned)n_img), 256, 0, st>>>(data, n, planes, hist);
  AMT_LAUNCH_CHECK();
  sel_locate_kernel<<<(unsigned)n_img, 1024, 0, st>>>(data, n, ranks, planes, hist, cand, cap);
  AMT_LAUNCH_CHECK();
  sel_compact_kernel<<<dim3((unsigned)ceil_div(n, SEL_CHUNK), (unsigned)n_img), 256, 0, st>>>(data, n, planes, cand, cap);
  AMT_LAUNCH_CHECK();
  sel_radix_kernel<<<dim3((unsigned)n_ranks, (unsigned)n_img), 1024, 0, st>>>(planes, n_ranks, out_vals);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

size_t amt_select_u16_scratch_bytes(int64_t n_img) { return (size_t)n_img * 65536 * sizeof(uint32_t); }

int amt_hist_u16(const uint16_t* data, int64_t n_img, int64_t n, uint32_t* hist65536, amt_stream_t stream) {
  return amt::hist_u16(data, n_img, n, hist65536, amt::as_stream(stream));
}

int amt_select_u16(const uint16_t* data, int64_t n_img, int64_t n, const int64_t* ranks_host, int n_ranks,
                   double* out_vals, void* scratch, size_t scratch_bytes, amt_stream_t stream) {
  using namespace amt;
  if (!data || !ranks_host || !out_vals || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || n <= 0 || n_ranks <= 0 || n_ranks > AMT_MAX_RANKS) return AMT_ERR_INVALID;
  if (n >= (1ll << 32)) return AMT_ERR_CAPACITY;
  if (scratch_bytes < amt_select_u16_scratch_bytes(n_img)) return AMT_ERR_CAPACITY;
  SelRanks ranks;
  ranks.n = n_ranks;
  for (int i = 0; i < n_ranks; ++i) {
    if (ranks_host[i] < 0 || ranks_host[i] >= n) return AMT_ERR_INVALID;
    ranks.r[i] = ranks_host[i];
  }
  cudaStream_t st = as_stream(stream);
  uint32_t* hist = (uint32_t*)scratch;
  AMT_TRY(hist_u16(data, n_img, n, hist, st));
  sel_u16_locate_kernel<<<(unsigned)n_img, 1024, 0, st>>>(hist, ranks, out_vals);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"

namespace amt {

// Same contract as amt_select_f64 (same scratch layout and size), with the bucket12() plane the DoG's
// second pass wrote next to `data`.  n must be a multiple of 8 and both planes 16-byte aligned.
int select_f64_bucketed(const double* data, const uint16_t* buckets, int64_t n_img, int64_t n, const int64_t* ranks_host,
                        int n_ranks, const uint64_t* minmax_keys, double* out_vals, void* scratch, size_t scratch_bytes,
                        cudaStream_t st) {
  if (!data || !buckets || !ranks_host || !minmax_keys || !out_vals || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || n <= 0 || n % 8 != 0 || n_ranks <= 0 || n_ranks > AMT_MAX_RANKS || n_img > 65535) return AMT_ERR_INVALID;
  if (n >= (1ll << 32)) return AMT_ERR_CAPACITY;
  if (scratch_bytes < amt_select_f64_scratch_bytes(n_img, n)) return AMT_ERR_CAPACITY;
  SelRanks ranks;
  ranks.n = n_ranks;
  for (int i = 0; i < n_ranks; ++i) {
    if (ranks_host[i] < 0 || ranks_host[i] >= n) return AMT_ERR_INVALID;
    ranks.r[i] = ranks_host[i];
  }
  const int64_t cap = sel_cap(n);
  char* base = (char*)scratch;
  SelPlane* planes = (SelPlane*)base;
  uint32_t* hist = (uint32_t*)(base + (size_t)n_img * sizeof(SelPlane));
  double* cand = (double*)(base + sel_head_bytes(n_img));
  sel_prepare_kernel<<<(unsigned)n_img, 256, 0, st>>>(minmax_keys, planes, hist, n_img);
  AMT_LAUNCH_CHECK();
  int64_t bx = ceil_div(n, 256 * 8 * 16);
  const int64_t capb = ceil_div((int64_t)kNumSMs * g_pass_ctas, n_img);
  if (bx > capb) bx = capb;
  if (bx < 1) bx = 1;
  sel_hist_buckets_kernel<<<dim3((unsigned)bx, (unsigned)n_img), 256, 0, st>>>(buckets, n, hist);
  AMT_LAUNCH_CHECK();
  sel_locate_kernel<<<(unsigned)n_img, 1024, 0, st>>>(data, n, ranks, planes, hist, cand, cap);
  AMT_LAUNCH_CHECK();
  int64_t cx = ceil_div(n, 256 * 8 * 8);
  if (cx > capb) cx = capb;
  if (cx < 1) cx = 1;
  sel_compact_buckets_kernel<<<dim3((unsigned)cx, (unsigned)n_img), 256, 0, st>>>(buckets, data, n, planes, cand, cap);
  AMT_LAUNCH_CHECK();
  sel_radix_kernel<<<dim3((unsigned)n_ranks, (unsigned)n_img), 1024, 0, st>>>(planes, n_ranks, out_vals);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt

extern "C" {

// Test / standalone entry: bucket12() of every sample of `data` (what amt_dog2d's second pass writes on
// the executor path), so that the bucketed selection can be checked against amt_select_f64.
int amt_select_f64_bucketed(const double* data, const uint16_t* buckets, int64_t n_img, int64_t n,
                            const int64_t* ranks_host, int n_ranks, const uint64_t* minmax_keys, double* out_vals,
                            void* scratch, size_t scratch_bytes, amt_stream_t stream) {
  return amt::select_f64_bucketed(data, buckets, n_img, n, ranks_host, n_ranks, minmax_keys, out_vals, scratch,
                                  scratch_bytes, amt::as_stream(stream));
}

int amt_bucket12(const double* data, uint16_t* buckets, int64_t n, amt_stream_t stream) {
  using namespace amt;
  if (!data || !buckets || n <= 0) return AMT_ERR_INVALID;
  bucket12_kernel<<<kNumSMs * 8, 256, 0, as_stream(stream)>>>(data, buckets, n);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
