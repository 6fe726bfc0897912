// Per-cell quantification: one pass over (labels, C uint16 channels) -> exact integer
// accumulators per label -> float64 regionprops table.
//
// Reference path: masks.py:286-289 and :317-326 ski.measure.regionprops_table [3p]
// (SURVEY.md 8a item 10).  The reference re-scans the label image once per property group and
// once more per channel, in a Python loop over regions; here every statistic of every channel
// comes out of a single streaming pass.
//
// Reduce kernel: a warp scans 8 rows x 256 columns for 8-pixel strips that contain foreground and
// hands them out densely; a lane folds its strip (two 128-bit label loads, one 128-bit load per
// channel) into runs of equal label in registers and flushes each run with 64-bit integer atomics
// into a SoA table [field][label].  All accumulators are integers (counts, coordinate sums up to order
// 2, intensity sum and sum of squares, min / max), so area, bbox and intensity sums are exact
// and the float statistics are computed once, in finalize, from exact numerators.  HBM-bound:
// 4 + 2*C bytes per pixel; the atomics go to L2.

#include "internal.cuh"

namespace amt {

enum {
  F_COUNT = 0, F_SR, F_SC, F_SRR, F_SCC, F_SRC, F_RMIN, F_RMAX, F_CMIN, F_CMAX, F_BASE
};
enum { CF_SUM = 0, CF_SUMSQ, CF_MIN, CF_MAX, CF_PER };
static_assert(F_BASE == AMT_ACC_BASE && CF_PER == AMT_ACC_PER_CHANNEL, "accumulator layout");

constexpr int MAX_CH = 8;

__global__ void acc_init_kernel(uint64_t* __restrict__ acc, int n_fields, int64_t max_labels, int64_t total) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int f = (int)((i / max_labels) % n_fields);
    bool is_min = (f == F_RMIN || f == F_CMIN);
    if (f >= F_BASE) is_min = ((f - F_BASE) % CF_PER) == CF_MIN;
    acc[i] = is_min ? ~0ull : 0ull;
  }
}

struct Run {
  int label;
  uint32_t n, sx;
  uint64_t sxx;
  uint32_t vs[MAX_CH], vmin[MAX_CH], vmax[MAX_CH];
  uint64_t vss[MAX_CH];
};

template <int C>
__device__ __forceinline__ void flush_run(const Run& r, const int y, uint64_t* __restrict__ acc, const int64_t max_labels) {
  if (r.label <= 0 || r.label > max_labels) return;
  unsigned long long* a = (unsigned long long*)acc + (r.label - 1);
  const uint64_t n = r.n, yy = (uint64_t)y;
  atomicAdd(a + F_COUNT * max_labels, n);
  atomicAdd(a + F_SR * max_labels, n * yy);
  atomicAdd(a + F_SC * max_labels, (uint64_t)r.sx);
  atomicAdd(a + F_SRR * max_labels, n * yy * yy);
  atomicAdd(a + F_SCC * max_labels, r.sxx);
  atomicAdd(a + F_SRC * max_labels, yy * (uint64_t)r.sx);
  atomicMin(a + F_RMIN * max_labels, yy);
  atomicMax(a + F_RMAX * max_labels, yy);
  // c_min / c_max are folded in by the caller through sx bounds: see below
#pragma unroll
  for (int c = 0; c < C; ++c) {
    unsigned long long* ac = a + (F_BASE + c * CF_PER) * max_labels;
    atomicAdd(ac + CF_SUM * max_labels, (uint64_t)r.vs[c]);
    atomicAdd(ac + CF_SUMSQ * max_labels, r.vss[c]);
    atomicMin(ac + CF_MIN * max_labels, (uint64_t)r.vmin[c]);
    atomicMax(ac + CF_MAX * max_labels, (uint64_t)r.vmax[c]);
  }
}

// One 8-pixel strip (row y, columns x0 .. x0+7) of image img: fold runs of equal label, flush them.
template <int C, bool VEC>
__device__ __forceinline__ void reduce_strip(const int32_t* __restrict__ labels, const uint16_t* __restrict__ channels,
                                             const int64_t img_stride, const int64_t chan_stride, const int h, const int w,
                                             const int64_t max_labels, uint64_t* __restrict__ acc, const int n_fields,
                                             const int64_t img, const int y, const int x0) {
  const int64_t row = img * (int64_t)h * w + (int64_t)y * w;
  int lab[8];
  if (VEC) {
    const int4 a = ld_nc_int4(labels + row + x0);
    const int4 b = ld_nc_int4(labels + row + x0 + 4);
    lab[0] = a.x; lab[1] = a.y; lab[2] = a.z; lab[3] = a.w;
    lab[4] = b.x; lab[5] = b.y; lab[6] = b.z; lab[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) lab[i] = (x0 + i < w) ? labels[row + x0 + i] : 0;
  }
  // (only strips with foreground are queued: the channel loads below do not wait for the labels)
  uint32_t val[C > 0 ? C : 1][8];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const uint16_t* src = channels + img * img_stride + c * chan_stride + (int64_t)y * w + x0;
    if (VEC) {
      const int4 q = ld_nc_int4(src);
      val[c][0] = (uint32_t)q.x & 0xffffu; val[c][1] = (uint32_t)q.x >> 16;
      val[c][2] = (uint32_t)q.y & 0xffffu; val[c][3] = (uint32_t)q.y >> 16;
      val[c][4] = (uint32_t)q.z & 0xffffu; val[c][5] = (uint32_t)q.z >> 16;
      val[c][6] = (uint32_t)q.w & 0xffffu; val[c][7] = (uint32_t)q.w >> 16;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) val[c][i] = (x0 + i < w) ? src[i] : 0;
    }
  }

  uint64_t* acc_img = acc + img * (int64_t)n_fields * max_labels;
  Run r;
  r.label = 0;
  int run_x0 = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int l = lab[i];
    if (l != r.label) {
      if (r.label > 0 && r.label <= max_labels) {
        flush_run<C>(r, y, acc_img, max_labels);
        unsigned long long* a = (unsigned long long*)acc_img + (r.label - 1);
        atomicMin(a + F_CMIN * max_labels, (uint64_t)run_x0);
        atomicMax(a + F_CMAX * max_labels, (uint64_t)(x0 + i - 1));
      }
      r.label = l;
      r.n = 0; r.sx = 0; r.sxx = 0;
      run_x0 = x0 + i;
#pragma unroll
      for (int c = 0; c < C; ++c) { r.vs[c] = 0; r.vss[c] = 0; r.vmin[c] = 0xffffffffu; r.vmax[c] = 0; }
    }
    if (l > 0) {
      const uint32_t x = (uint32_t)(x0 + i);
      r.n += 1; r.sx += x; r.sxx += (uint64_t)x * x;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t v = val[c][i];
        r.vs[c] += v; r.vss[c] += (uint64_t)v * v;
        r.vmin[c] = v < r.vmin[c] ? v : r.vmin[c];
        r.vmax[c] = v > r.vmax[c] ? v : r.vmax[c];
      }
    }
  }
  if (r.label > 0 && r.label <= max_labels) {
    flush_run<C>(r, y, acc_img, max_labels);
    unsigned long long* a = (unsigned long long*)acc_img + (r.label - 1);
    atomicMin(a + F_CMIN * max_labels, (uint64_t)run_x0);
    atomicMax(a + F_CMAX * max_labels, (uint64_t)(x0 + 7 < w ? x0 + 7 : w - 1));
  }
}

// grid: (ceil(w / 1024), ceil(h / 8), n_img); block 128 threads = 4 warps; a warp owns 8 rows x 256
// columns = 256 strips of 8 pixels.  ~85 % of the strips of a cell image are pure background, and a
// warp that maps lanes to strips one to one runs the whole fold / flush path for the 3-5 lanes that
// have work (ncu: 99 M warp instructions per 33.5 Mpx launch).  So the warp first scans its strips
// (label loads only, one ballot per row), then hands the foreground strips out densely, 32 at a
// time: the expensive path runs with full warps (31-42 M instructions).
// A per-CTA shared-memory table of privatised accumulators was tried on top of this and was SLOWER
// (162 vs 124 us): after the dense hand-out, neighbouring lanes hold the same label, so every
// shared-memory atomic is a 32-way same-address conflict, while the L2 absorbs the same pattern.
template <int C, bool VEC>
__global__ void __launch_bounds__(128)
region_reduce_kernel(const int32_t* __restrict__ labels, const uint16_t* __restrict__ channels, const int64_t img_stride,
                     const int64_t chan_stride, const int h, const int w, const int64_t max_labels,
                     uint64_t* __restrict__ acc, const int n_fields) {
  const int64_t img = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int y_base = blockIdx.y * 8;
  const int x_base = blockIdx.x * 1024 + warp * 256;
  if (x_base >= w) return;
  const int x0 = x_base + lane * 8;
  const int64_t plane = img * (int64_t)h * w;
  unsigned fg[8];
  {
    int any[8];  // all 16 label loads of the scan in flight before the first ballot
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int y = y_base + r;
      any[r] = 0;
      if (y < h && x0 < w) {
        const int32_t* p = labels + plane + (int64_t)y * w + x0;
        if (VEC) {
          const int4 a = __ldg(reinterpret_cast<const int4*>(p)), b = __ldg(reinterpret_cast<const int4*>(p) + 1);
          any[r] = a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) any[r] |= (x0 + i < w) ? p[i] : 0;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) fg[r] = __ballot_sync(0xffffffffu, any[r] != 0);
  }
  int total = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) total += __popc(fg[r]);
  for (int base = 0; base < total; base += 32) {
    const int idx = base + lane;
    if (idx < total) {
      int r = 0, before = 0, run = 0;
      unsigned m = fg[0];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {  // the row whose ballot holds work item idx
        if (idx >= run) { r = rr; m = fg[rr]; before = run; }
        run += __popc(fg[rr]);
      }
      const int ln = __fns(m, 0, idx - before + 1);  // lane of the (idx - before)-th foreground strip of row r
      reduce_strip<C, VEC>(labels, channels, img_stride, chan_stride, h, w, max_labels, acc, n_fields, img, y_base + r,
                           x_base + ln * 8);
    }
  }
}

// exact (N*S2 - S1a*S1b) as a double; the difference is formed in 128-bit integers
__device__ __forceinline__ double exact_cov_num(uint64_t n, uint64_t s2, uint64_t s1a, uint64_t s1b) {
  const unsigned __int128 p = (unsigned __int128)n * s2;
  const unsigned __int128 q = (unsigned __int128)s1a * s1b;
  const bool neg = q > p;
  const unsigned __int128 d = neg ? q - p : p - q;
  const double v = (double)(uint64_t)(d >> 64) * 18446744073709551616.0 + (double)(uint64_t)d;
  return neg ? -v : v;
}

__global__ void region_finalize_kernel(const uint64_t* __restrict__ acc, const int32_t* __restrict__ counts,
                                       const int n_channels, const int64_t max_labels, double* __restrict__ table) {
  const int64_t img = blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t K = counts[img];
  if (K > max_labels) K = max_labels;
  if (k >= K) return;
  const int n_fields = AMT_ACC_FIELDS(n_channels);
  const int n_cols = AMT_TABLE_COLS(n_channels);
  const uint64_t* a = acc + img * (int64_t)n_fields * max_labels + k;
  double* t = table + img * (int64_t)n_cols * max_labels + k;
  auto A = [&](int f) -> uint64_t { return a[(int64_t)f * max_labels]; };
  auto T = [&](int c) -> double& { return t[(int64_t)c * max_labels]; };
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint64_t n = A(F_COUNT);
  T(0) = (double)(k + 1);
  T(1) = (double)n;
  if (n == 0) {  // label id without pixels (cannot happen after relabelling; keep the row defined)
    for (int c = 2; c < n_cols; ++c) T(c) = nan;
    return;
  }
  const double N = (double)n;
  T(2) = (double)A(F_RMIN);
  T(3) = (double)A(F_CMIN);
  T(4) = (double)(A(F_RMAX) + 1);
  T(5) = (double)(A(F_CMAX) + 1);
  T(6) = (double)A(F_SR) / N;
  T(7) = (double)A(F_SC) / N;
  // central second moments from exact integer numerators: mu20 = (N*Srr - Sr^2)/N, ...
  const double mu20 = exact_cov_num(n, A(F_SRR), A(F_SR), A(F_SR)) / N;
  const double mu02 = exact_cov_num(n, A(F_SCC), A(F_SC), A(F_SC)) / N;
  const double mu11 = exact_cov_num(n, A(F_SRC), A(F_SR), A(F_SC)) / N;
  // skimage inertia_tensor: diag = (sum(mu20, mu02) - mu) / mu00, off-diag = -mu11 / mu00
  const double S = mu20 + mu02;
  const double ta = (S - mu20) / N, tc = (S - mu02) / N, tb = -mu11 / N;
  const double half_tr = 0.5 * (ta + tc);
  const double dev = hypot(0.5 * (ta - tc), tb);
  double l1 = half_tr + dev, l2 = half_tr - dev;
  l1 = l1 > 0.0 ? l1 : 0.0;
  l2 = l2 > 0.0 ? l2 : 0.0;
  T(8) = l1;
  T(9) = l2;
  T(10) = 4.0 * sqrt(l1);
  T(11) = 4.0 * sqrt(l2);
  T(12) = (l1 == 0.0) ? 0.0 : sqrt(1.0 - l2 / l1);
  const double PI = 3.141592653589793;
  if (ta - tc == 0.0)
    T(13) = (tb < 0.0) ? PI / 4.0 : -PI / 4.0;
  else
    T(13) = 0.5 * atan2(-2.0 * tb, tc - ta);
  T(14) = nan;  // perimeter   (amt_region_shape)
  T(15) = nan;  // area_convex (amt_region_shape)
  for (int c = 0; c < n_channels; ++c) {
    const int f = F_BASE + c * CF_PER;
    const int col = AMT_TABLE_BASE + c * AMT_TABLE_PER_CHANNEL;
    const uint64_t s = A(f + CF_SUM), ss = A(f + CF_SUMSQ);
    T(col + 0) = (double)s;
    T(col + 1) = (double)s / N;
    T(col + 2) = (double)A(f + CF_MAX);
    T(col + 3) = (double)A(f + CF_MIN);
    const double varnum = exact_cov_num(n, ss, s, s);  // N*Sxx - Sx^2 >= 0
    T(col + 4) = sqrt((varnum > 0.0 ? varnum : 0.0)) / N;
  }
}

template <int C>
static int reduce_dispatch(const int32_t* labels, const uint16_t* channels, int64_t img_stride, int64_t chan_stride,
                           int64_t n_img, int h, int w, int64_t max_labels, uint64_t* acc, cudaStream_t st) {
  const int n_fields = AMT_ACC_FIELDS(C);
  dim3 grid((unsigned)ceil_div(w, 1024), (unsigned)ceil_div(h, 8), (unsigned)n_img);
  bool vec = (w % 8 == 0) && (((uintptr_t)labels) % 16 == 0);
  if (C > 0)
    vec = vec && (((uintptr_t)channels) % 16 == 0) && (img_stride % 8 == 0) && (chan_stride % 8 == 0);
  if (vec)
    region_reduce_kernel<C, true><<<grid, 128, 0, st>>>(labels, channels, img_stride, chan_stride, h, w, max_labels, acc, n_fields);
  else
    region_reduce_kernel<C, false><<<grid, 128, 0, st>>>(labels, channels, img_stride, chan_stride, h, w, max_labels, acc, n_fields);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int region_reduce(const int32_t* labels, const uint16_t* channels, int n_channels, int64_t img_stride, int64_t chan_stride,
                  int64_t n_img, int64_t h, int64_t w, int64_t max_labels, uint64_t* acc, cudaStream_t st) {
  if (!labels || !acc || n_img <= 0 || h <= 0 || w <= 0 || max_labels <= 0) return AMT_ERR_INVALID;
  if (n_channels < 0 || n_channels > MAX_CH || (n_channels > 0 && !channels)) return AMT_ERR_INVALID;
  if (ceil_div(h, 8) > 65535 || n_img > 65535 || h * w >= (1ll << 31)) return AMT_ERR_CAPACITY;
  const int n_fields = AMT_ACC_FIELDS(n_channels);
  const int64_t total = n_img * (int64_t)n_fields * max_labels;
  int64_t ib = ceil_div(total, 256);
  if (ib > kNumSMs * 8) ib = kNumSMs * 8;
  acc_init_kernel<<<(unsigned)ib, 256, 0, st>>>(acc, n_fields, max_labels, total);
  AMT_LAUNCH_CHECK();
  switch (n_channels) {
    case 0: return reduce_dispatch<0>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 1: return reduce_dispatch<1>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 2: return reduce_dispatch<2>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 3: return reduce_dispatch<3>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 4: return reduce_dispatch<4>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 5: return reduce_dispatch<5>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 6: return reduce_dispatch<6>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    case 7: return reduce_dispatch<7>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
    default: return reduce_dispatch<8>(labels, channels, img_stride, chan_stride, n_img, (int)h, (int)w, max_labels, acc, st);
  }
}

int region_finalize(const uint64_t* acc, const int32_t* counts, int n_channels, int64_t n_img, int64_t max_labels,
                    double* table, cudaStream_t st) {
  if (!acc || !counts || !table || n_img <= 0 || max_labels <= 0 || n_channels < 0 || n_channels > MAX_CH)
    return AMT_ERR_INVALID;
  dim3 grid((unsigned)ceil_div(max_labels, 128), (unsigned)n_img);
  region_finalize_kernel<<<grid, 128, 0, st>>>(acc, counts, n_channels, max_labels, table);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt

extern "C" {

int amt_region_reduce(const int32_t* labels, const uint16_t* channels, int n_channels, int64_t img_stride,
                      int64_t chan_stride, int64_t n_img, int64_t h, int64_t w, int64_t max_labels, uint64_t* acc,
                      amt_stream_t stream) {
  return amt::region_reduce(labels, channels, n_channels, img_stride, chan_stride, n_img, h, w, max_labels, acc,
                            amt::as_stream(stream));
}

int amt_region_finalize(const uint64_t* acc, const int32_t* counts, int n_channels, int64_t n_img, int64_t max_labels,
                        double* table, amt_stream_t stream) {
  return amt::region_finalize(acc, counts, n_channels, n_img, max_labels, table, amt::as_stream(stream));
}

size_t amt_region_shape_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_labels) {
  return amt::region_shape_scratch_bytes(n_img, h, w, max_labels);
}

int amt_region_shape(const int32_t* labels, const uint64_t* acc, int n_channels, const int32_t* counts, int64_t n_img,
                     int64_t h, int64_t w, int64_t max_labels, double* table, void* scratch, size_t scratch_bytes,
                     amt_stream_t stream) {
  return amt::region_shape(labels, acc, n_channels, counts, n_img, h, w, max_labels, table, scratch, scratch_bytes,
                           amt::as_stream(stream));
}

}  // extern "C"
