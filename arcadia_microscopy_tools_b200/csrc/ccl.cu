// Connected-component labelling, border clearing and sequential relabelling.
//
// Reference path: masks.py:38-65 (_process_mask): ski.segmentation.clear_border (:56) [3p] then
// ski.measure.label (:63) [3p] for bool masks or ski.segmentation.relabel_sequential (:65) [3p]
// for integer masks.  SURVEY.md 8a items 7-9.
//
// Run-based block union-find, 8-connectivity, on "same non-zero value" adjacency (binary masks
// are the one-value case; integer masks give clear_border's re-labelling by value+connectivity).
// The work is proportional to the number of RUNS, not pixels; pixels only pay for the streaming
// read of the input and the streaming write of the labels:
//  A. tile:     a CTA takes a 64x64 tile.  Warps load it row by row (lane = pixel) and turn every
//               row into 64-bit masks with ballots: foreground, run starts and, for integer
//               masks, "same value as N / NW / NE".  One thread per row then walks that row's
//               RUNS with bit scans and unions each run with the runs of the row above it
//               touches (shared-memory union-find, roots = smallest index).  Every pixel is
//               written once with its tile root as a GLOBAL pixel index; tile roots are appended
//               to a per-image list.  The input is read once (thresholded on the fly if it is a
//               float64 plane).
//  B. seams:    only the pixels on tile seams union across tiles (global atomicMin): the
//               warp-level boundary-merge pass, ~5 % of the pixels.
//  C. roots:    only the listed tile roots are flattened to their global root (the smallest
//               pixel index of the component = its first pixel in raster order, which is the
//               numbering key of scipy.ndimage.label / skimage); global roots set their bit in a
//               one-bit-per-pixel bitmap.  Border pixels clear the bit of their component.
//  D. number:   a popcount prefix over the bitmap words ranks the surviving roots in raster
//               order (integer masks: surviving roots mark their VALUE, a presence-table scan
//               ranks the values = relabel_sequential).
//  E. final:    pixel -> tile root -> global root -> id, in place, 16-byte accesses.
// Pixel traffic: A 4-12 B, E 8 B + cached gathers; everything else touches runs / roots only.

#include "common.cuh"

namespace amt {

#ifndef AMT_CCL_TH
#define AMT_CCL_TH 32
#endif
// amt_tune "ccl_touch_filter": 1 (default) = integer masks with border clearing are labelled only where it can matter
// (the pixels of values that occur on the image border), 0 = every pixel
int g_ccl_touch_filter = 1;
constexpr int TW = 64, TH = AMT_CCL_TH;  // tile of kernel A (TW is the width of the 64-bit row masks)
constexpr int TILE_WARPS = TH / 8, TILE_THREADS = TILE_WARPS * 32;

template <int KIND> struct CclRaw { typedef uint8_t type; };
template <> struct CclRaw<1> { typedef double type; };
template <> struct CclRaw<2> { typedef int32_t type; };

template <int KIND>
__device__ __forceinline__ typename CclRaw<KIND>::type ccl_raw(const void* in, const int64_t idx) {
  return __ldg((const typename CclRaw<KIND>::type*)in + idx);
}
template <int KIND>
__device__ __forceinline__ int ccl_cook(const typename CclRaw<KIND>::type raw, const double thr) {
  if (KIND == 0) return raw ? 1 : 0;
  if (KIND == 1) return raw > thr ? 1 : 0;
  return (int)raw;
}

template <int KIND>
__device__ __forceinline__ int ccl_value(const void* in, const double thr, const int64_t idx) {
  if (KIND == 0) return ((const uint8_t*)in)[idx] ? 1 : 0;
  if (KIND == 1) return ((const double*)in)[idx] > thr ? 1 : 0;
  return ((const int32_t*)in)[idx];
}

// ---------------------------------------------------------------- union-find (global / shared)
__device__ __forceinline__ int uf_find(const int32_t* L, int a) {
  while (true) {
    const int p = __ldcg(L + a);
    if (p == a) return a;
    a = p;
  }
}
__device__ __forceinline__ void uf_union(int32_t* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      const int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}
__device__ __forceinline__ int suf_find(const volatile int* L, int a) {
  while (true) {
    const int p = L[a];
    if (p == a) return a;
    a = p;
  }
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
  bool done;
  do {
    a = suf_find(L, a);
    b = suf_find(L, b);
    if (a < b) {
      const int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// ---------------------------------------------------------------- A. tile labelling
__device__ __forceinline__ uint64_t run_mask_from(const uint64_t stops, const int a) {
  // bits a .. b where b + 1 is the first set bit of `stops` above a (or 64)
  const uint64_t above = (a == 63) ? 0ull : (stops >> (a + 1)) << (a + 1);
  const int end = above ? (__ffsll((long long)above) - 1) : 64;  // exclusive
  const uint64_t upto = (end == 64) ? ~0ull : ((1ull << end) - 1ull);
  return upto & ~((1ull << a) - 1ull);
}

// grid (ceil(w/TW), ceil(h/TH), planes); block TH/8 warps; warp -> rows warp, warp + TH/8, ... (8 rows each)
template <int KIND>
__global__ void __launch_bounds__(TILE_THREADS)
ccl_tile_kernel(const void* __restrict__ in, const int64_t in_stride, const double* __restrict__ thresholds,
                int32_t* __restrict__ L, int32_t* __restrict__ rootlist, int32_t* __restrict__ rootcnt, const int h,
                const int w, const int32_t* __restrict__ touch, int32_t* __restrict__ present, const int nval) {
  __shared__ int s_lab[TH * TW];                  // union-find over run starts (tile-local pixel index)
  __shared__ uint64_t s_fg[TH], s_brk[TH];        // per row: foreground, run starts
  __shared__ uint64_t s_n[KIND == 2 ? TH : 1], s_nw[KIND == 2 ? TH : 1], s_ne[KIND == 2 ? TH : 1];
  __shared__ int s_val[KIND == 2 ? TH * TW : 1];
  const int64_t img = blockIdx.z;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double thr = (KIND == 1) ? thresholds[img] : 0.0;
  const int64_t in_base = img * in_stride;
  const int64_t plane = img * (int64_t)h * w;

  int vals[8][2];
  {
    // every global load of the tile is issued before the first use (16 in flight per thread)
    typename CclRaw<KIND>::type raw[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int y = ty0 + warp + TILE_WARPS * q;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int x = tx0 + half * 32 + lane;
        raw[q][half] = (y < h && x < w) ? ccl_raw<KIND>(in, in_base + (int64_t)y * w + x) : (typename CclRaw<KIND>::type)0;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int y = ty0 + warp + TILE_WARPS * q;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int x = tx0 + half * 32 + lane;
        vals[q][half] = (y < h && x < w) ? ccl_cook<KIND>(raw[q][half], thr) : 0;
      }
    }
  }
  if (KIND == 2 && touch != nullptr) {
    // Integer masks with border clearing: only a VALUE that occurs on the image border can lose a fragment, so only
    // the pixels of those values need connectivity.  Every other in-range value is present as it stands (marked here,
    // once per run) and its pixels are background for the labelling; a tile without a pixel of a border value is done.
    const int32_t* tp = touch + img * (int64_t)nval;
    int32_t* pp = present + img * (int64_t)nval;
    int tch[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int v = vals[q][half];
        v = (v > 0 && v < nval) ? v : 0;  // values outside the declared range label nothing (relabel_final reports them)
        vals[q][half] = v;
        tch[q][half] = v ? __ldg(tp + v) : 0;
      }
    int any = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int last = 0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int v = vals[q][half];
        int vl = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) vl = half ? last : 0;
        last = __shfl_sync(0xffffffffu, v, 31);
        if (v != 0 && !tch[q][half]) {
          if (v != vl) pp[v] = 1;
          vals[q][half] = 0;
        }
        any |= vals[q][half];
      }
    }
    if (!__syncthreads_or(any)) return;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int lr = warp + TILE_WARPS * q;
    uint32_t fgw[2], bkw[2];
    int last = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lc = half * 32 + lane;
      const int v = vals[q][half];
      int vl = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0) vl = half ? last : 0;  // the run may continue from the left half; tile edge: a start
      last = __shfl_sync(0xffffffffu, v, 31);
      const bool start = v != 0 && v != vl;
      fgw[half] = __ballot_sync(0xffffffffu, v != 0);
      bkw[half] = __ballot_sync(0xffffffffu, start);
      if (start) s_lab[lr * TW + lc] = lr * TW + lc;
      if (KIND == 2) s_val[lr * TW + lc] = v;
    }
    if (lane == 0) {
      s_fg[lr] = ((uint64_t)fgw[1] << 32) | fgw[0];
      s_brk[lr] = ((uint64_t)bkw[1] << 32) | bkw[0];
    }
  }
  __syncthreads();
  if (KIND == 2) {  // same-value masks against the row above (ballots; the values are still in registers)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int lr = warp + TILE_WARPS * q;
      uint32_t nw_[2], n_[2], ne_[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int lc = half * 32 + lane, v = vals[q][half];
        const int* up = s_val + (lr - 1) * TW + lc;
        const bool ok = v != 0 && lr > 0;
        n_[half] = __ballot_sync(0xffffffffu, ok && up[0] == v);
        nw_[half] = __ballot_sync(0xffffffffu, ok && lc > 0 && up[-1] == v);
        ne_[half] = __ballot_sync(0xffffffffu, ok && lc < TW - 1 && up[1] == v);
      }
      if (lane == 0) {
        s_n[lr] = ((uint64_t)n_[1] << 32) | n_[0];
        s_nw[lr] = ((uint64_t)nw_[1] << 32) | nw_[0];
        s_ne[lr] = ((uint64_t)ne_[1] << 32) | ne_[0];
      }
    }
    __syncthreads();
  }

  // link: one thread per row walks its runs and unions each with the runs it touches in the row above
  if (threadIdx.x < TH && threadIdx.x > 0) {
    const int lr = threadIdx.x;
    const uint64_t fg = s_fg[lr], brk = s_brk[lr];
    const uint64_t ufg = s_fg[lr - 1], ubrk = s_brk[lr - 1];
    uint64_t en, enw, ene;
    if (KIND == 2) {
      en = s_n[lr]; enw = s_nw[lr]; ene = s_ne[lr];
    } else {
      en = fg & ufg; enw = fg & (ufg << 1); ene = fg & (ufg >> 1);
    }
    if (en | enw | ene) {
      const uint64_t stops = brk | ~fg, ustops = ubrk | ~ufg;
      uint64_t todo = brk;
      while (todo) {
        const int a = __ffsll((long long)todo) - 1;
        todo &= todo - 1;
        const uint64_t R = run_mask_from(stops, a);
        // pixels of the row above that carry this run's value and touch it
        uint64_t A = (en & R) | ((enw & R) >> 1) | ((ene & R) << 1);
        while (A) {
          const int x = __ffsll((long long)A) - 1;
          const int ua = 63 - __clzll((long long)(ubrk & ((x == 63) ? ~0ull : ((2ull << x) - 1ull))));
          suf_union(s_lab, lr * TW + a, (lr - 1) * TW + ua);
          A &= ~run_mask_from(ustops, ua);
        }
      }
    }
  }
  __syncthreads();
  // flatten the runs (find everything first, write after a barrier: every entry then holds its root);
  // tile roots join the per-image list
  {
    const int lr = threadIdx.x & (TH - 1);
    int roots[4];  // runs 4*k + (threadIdx.x / TH) of row lr: four threads share a row
    int nr = 0;
    bool overflow = false;
    uint64_t todo = s_brk[lr];
    int k = 0;
    while (todo) {
      const int a = __ffsll((long long)todo) - 1;
      todo &= todo - 1;
      if ((k++ & 3) != (int)(threadIdx.x / TH)) continue;
      const int r = suf_find(s_lab, lr * TW + a);
      if (nr < 4) roots[nr++] = (a << 16) | r; else overflow = true;
    }
    __syncthreads();
    for (int i = 0; i < nr; ++i) {
      const int a = roots[i] >> 16, r = roots[i] & 0xffff, idx = lr * TW + a;
      if (r == idx) {
        const int pos = atomicAdd(rootcnt + img, 1);
        rootlist[plane + pos] = (ty0 + lr) * w + tx0 + a;
      } else {
        s_lab[idx] = r;
      }
    }
    if (overflow) {  // more than 16 runs in the row: the remaining ones (k >= 16) one at a time
      uint64_t rest = s_brk[lr];
      int kk = 0;
      while (rest) {
        const int a = __ffsll((long long)rest) - 1;
        rest &= rest - 1;
        const int mine = (kk++ & 3) == (int)(threadIdx.x / TH);
        if (!mine || kk <= 16) continue;
        const int idx = lr * TW + a;
        const int r = suf_find(s_lab, idx);
        if (r == idx) {
          const int pos = atomicAdd(rootcnt + img, 1);
          rootlist[plane + pos] = (ty0 + lr) * w + tx0 + a;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int lr = warp + TILE_WARPS * q, y = ty0 + lr;
    const uint64_t fg = s_fg[lr], brk = s_brk[lr];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lc = half * 32 + lane, x = tx0 + lc;
      if (y < h && x < w) {
        int out = -1;
        if ((uint32_t)(fg >> (half * 32)) != 0u) {  // warp-uniform: most row segments are all background
          if ((fg >> lc) & 1ull) {
            const int a = 63 - __clzll((long long)(brk & ((lc == 63) ? ~0ull : ((2ull << lc) - 1ull))));
            const int r = suf_find(s_lab, lr * TW + a);  // one hop for flattened entries
            out = (ty0 + (r >> 6)) * w + tx0 + (r & 63);
          }
        }
        L[plane + (int64_t)y * w + x] = out;
      }
    }
  }
}

// ---------------------------------------------------------------- B. seams between tiles
// one thread per seam pixel: rows y = k*TH (k >= 1), columns x = k*TW - 1 and x = k*TW (k >= 1)
template <int KIND>
__global__ void __launch_bounds__(256)
ccl_seam_kernel(const void* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ L, const int h, const int w,
                const int n_hseams, const int n_vseams, const int32_t* __restrict__ touch, const int nval) {
  const int64_t img = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t n_h = (int64_t)n_hseams * w;
  const int64_t n_v = (int64_t)n_vseams * 2 * h;
  if (t >= n_h + n_v) return;
  int x, y;
  if (t < n_h) {
    y = (int)(t / w + 1) * TH;
    x = (int)(t % w);
  } else {
    const int64_t u = t - n_h;
    const int k = (int)(u / (2 * h));
    const int rem = (int)(u % (2 * h));
    x = (k + 1) * TW - 1 + (rem & 1);
    y = rem >> 1;
  }
  if (x >= w || y >= h) return;
  int32_t* Lp = L + img * (int64_t)h * w;
  const int p = y * w + x;
  int v = 1;
  const int32_t* lab = nullptr;
  if (KIND == 2) {
    lab = (const int32_t*)in + img * in_stride;
    v = lab[p];
    if (v == 0) return;
    if (touch != nullptr && (v < 0 || v >= nval || !touch[img * (int64_t)nval + v])) return;  // not labelled (see the tile kernel)
  } else if (Lp[p] < 0) {
    return;
  }
  const int tcx = x / TW, tcy = y / TH;
  auto try_union = [&](int qx, int qy) {
    if (qx < 0 || qx >= w || qy < 0) return;
    if (qx / TW == tcx && qy / TH == tcy) return;  // same tile: done in shared memory
    const int q = qy * w + qx;
    const bool same = (KIND == 2) ? (lab[q] == v) : (Lp[q] >= 0);
    if (same) uf_union(Lp, p, q);
  };
  try_union(x - 1, y);
  try_union(x - 1, y - 1);
  try_union(x, y - 1);
  try_union(x + 1, y - 1);
}

__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  // blockDim.x threads (multiple of 32, <= 1024)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int wv = lane < nwarp ? s_warp[lane] : 0;
    int wi = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - wv;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  return s_warp[warp] + incl - v;
}

// ---------------------------------------------------------------- C. roots: flatten, mark, border
// grid (ceil(cap/256), planes): the listed tile roots point at their global root; global roots set
// their bit (integer masks decide later, from the surviving bits, which VALUES are present)
__global__ void __launch_bounds__(256)
ccl_roots_kernel(int32_t* __restrict__ L, const int32_t* __restrict__ rootlist, const int32_t* __restrict__ rootcnt,
                 uint32_t* __restrict__ rootbits, const int64_t npx, const int64_t words) {
  const int64_t img = blockIdx.y;
  const int n = rootcnt[img];
  int32_t* Lp = L + img * npx;
  const int32_t* list = rootlist + img * npx;
  uint32_t* bits = rootbits + img * words;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int t = list[i];
    const int g = uf_find(Lp, t);
    if (g == t)
      atomicOr(bits + (t >> 5), 1u << (t & 31));
    else
      Lp[t] = g;  // racing readers see the old parent or g: both are ancestors
  }
}

// integer masks: the values that occur on the image border (the only ones clear_border can take a fragment from)
__global__ void __launch_bounds__(256)
ccl_touch_kernel(const int32_t* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ touch, const int h, const int w,
                 const int nval) {
  const int64_t img = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= 2 * w + 2 * h) return;
  int x, y;
  if (t < w) { y = 0; x = t; }
  else if (t < 2 * w) { y = h - 1; x = t - w; }
  else if (t < 2 * w + h) { x = 0; y = t - 2 * w; }
  else { x = w - 1; y = t - 2 * w - h; }
  const int v = in[img * in_stride + (int64_t)y * w + x];
  if (v > 0 && v < nval) touch[img * (int64_t)nval + v] = 1;
}

// components with a pixel on the image border lose their root bit (clear_border)
__global__ void __launch_bounds__(256)
ccl_border_kernel(const int32_t* __restrict__ L, uint32_t* __restrict__ rootbits, const int h, const int w,
                  const int64_t words, const int32_t* __restrict__ vals, const int64_t vals_stride, const int nval) {
  const int64_t img = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int n_border = 2 * w + 2 * h;
  if (t >= n_border) return;
  int x, y;
  if (t < w) { y = 0; x = t; }
  else if (t < 2 * w) { y = h - 1; x = t - w; }
  else if (t < 2 * w + h) { x = 0; y = t - 2 * w; }
  else { x = w - 1; y = t - 2 * w - h; }
  const int32_t* Lp = L + img * (int64_t)h * w;
  if (vals != nullptr) {  // filtered integer masks: only pixels of in-range values were labelled
    const int v = vals[img * vals_stride + y * w + x];
    if (v <= 0 || v >= nval) return;
  }
  const int r = Lp[y * w + x];
  if (r < 0) return;
  const int g = uf_find(Lp, r);
  atomicAnd(rootbits + img * words + (g >> 5), ~(1u << (g & 31)));
}

// ---------------------------------------------------------------- D. numbering
// prefix[i] = number of root bits in words [0, i) of the plane; counts[img] = their total.  Two
// levels: 1024-word blocks scanned by one CTA each (coalesced 16-byte loads), then one CTA per plane
// scans the block totals; ccl_gid_kernel adds the two.
constexpr int RANK_BLOCK = 1024;  // words per CTA (256 threads x 4)

__global__ void __launch_bounds__(256)
ccl_rank_block_kernel(const uint32_t* __restrict__ rootbits, int32_t* __restrict__ prefix, int32_t* __restrict__ blocksum,
                      const int64_t words, const int nblocks) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.y;
  const uint32_t* bits = rootbits + img * words;
  int32_t* pre = prefix + img * words;
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  int c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k] = (i0 + k < words) ? __popc(bits[i0 + k]) : 0;
  int run = block_excl_scan(c[0] + c[1] + c[2] + c[3], s_warp, &s_total);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i0 + k < words) pre[i0 + k] = run;
    run += c[k];
  }
  if (threadIdx.x == 0) blocksum[img * nblocks + blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(1024)
ccl_rank_top_kernel(int32_t* __restrict__ blocksum, const int nblocks, int32_t* __restrict__ counts) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.x;
  int32_t* bs = blocksum + img * nblocks;
  int carry = 0;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int x = i < nblocks ? bs[i] : 0;
    const int ex = block_excl_scan(x, s_warp, &s_total);
    if (i < nblocks) bs[i] = carry + ex;
    carry += s_total;
    __syncthreads();
  }
  if (counts != nullptr && threadIdx.x == 0) counts[img] = carry;
}

// integer masks: the value of every surviving component is present
__global__ void __launch_bounds__(256)
ccl_present_kernel(const int32_t* __restrict__ in, const int64_t in_stride, const int32_t* __restrict__ rootlist,
                   const int32_t* __restrict__ rootcnt, const uint32_t* __restrict__ rootbits, const int64_t npx,
                   const int64_t words, int32_t* __restrict__ present, const int64_t nval) {
  const int64_t img = blockIdx.y;
  const int n = rootcnt[img];
  const int32_t* list = rootlist + img * npx;
  const uint32_t* bits = rootbits + img * words;
  const int32_t* lab = in + img * in_stride;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int t = list[i];
    if ((bits[t >> 5] >> (t & 31)) & 1u) {
      const int v = lab[t];
      if (v > 0 && v < nval) present[img * nval + v] = 1;
    }
  }
}

// ---------------------------------------------------------------- E. ids of the tile roots, final gather
// gid[t] for every listed tile root t: the consecutive id of its component (MODE 0; 0 = removed), or
// 1 / -1 = survived / removed (MODE 1, integer masks).  gid is a separate array, so the in-place
// pixel pass below never reads an entry another thread rewrites.
template <int MODE>
__global__ void __launch_bounds__(256)
ccl_gid_kernel(const int32_t* __restrict__ L, const int32_t* __restrict__ rootlist, const int32_t* __restrict__ rootcnt,
               const uint32_t* __restrict__ rootbits, const int32_t* __restrict__ prefix,
               const int32_t* __restrict__ blocksum, const int nblocks, int32_t* __restrict__ gid, const int64_t npx,
               const int64_t words) {
  const int64_t img = blockIdx.y;
  const int n = rootcnt[img];
  const int32_t* Lp = L + img * npx;
  const int32_t* list = rootlist + img * npx;
  const uint32_t* bits = rootbits + img * words;
  const int32_t* pre = prefix + img * words;
  int32_t* gp = gid + img * npx;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int t = list[i];
    const int g = Lp[t];  // flattened by ccl_roots_kernel (a global root points at itself)
    const uint32_t wd = bits[g >> 5], bit = 1u << (g & 31);
    const bool alive = (wd & bit) != 0;
    if (MODE == 0)
      gp[t] = alive ? blocksum[img * nblocks + (g >> 5) / RANK_BLOCK] + pre[g >> 5] + __popc(wd & (bit - 1u)) + 1 : 0;
    else
      gp[t] = alive ? 1 : -1;
  }
}

__global__ void __launch_bounds__(256)
ccl_final_kernel(int32_t* __restrict__ L, const int32_t* __restrict__ gid, const int64_t npx, const int vec) {
  const int64_t img = blockIdx.y;
  int32_t* Lp = L + img * npx;
  const int32_t* gp = gid + img * npx;
  const int64_t step = (int64_t)gridDim.x * 256 * 4;
  for (int64_t p0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; p0 < npx; p0 += step) {
    if (vec && p0 + 3 < npx) {
      int4 q = *reinterpret_cast<const int4*>(Lp + p0);
      if ((q.x & q.y & q.z & q.w) < 0) {  // four background pixels
        *reinterpret_cast<int4*>(Lp + p0) = make_int4(0, 0, 0, 0);
        continue;
      }
      q.x = q.x >= 0 ? __ldg(gp + q.x) : 0;
      q.y = q.y >= 0 ? __ldg(gp + q.y) : 0;
      q.z = q.z >= 0 ? __ldg(gp + q.z) : 0;
      q.w = q.w >= 0 ? __ldg(gp + q.w) : 0;
      *reinterpret_cast<int4*>(Lp + p0) = q;
    } else {
      for (int i = 0; i < 4 && p0 + i < npx; ++i) {
        const int r = Lp[p0 + i];
        Lp[p0 + i] = r >= 0 ? gp[r] : 0;
      }
    }
  }
}

// integer masks: out = rank of the pixel's value if its fragment survived (use_ccl) else 0
__global__ void __launch_bounds__(256)
relabel_final_kernel(const int32_t* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ L,
                     const int32_t* __restrict__ aux, const int64_t npx, const int32_t* __restrict__ rank,
                     const int64_t nval, const int use_ccl, const int vec, int32_t* __restrict__ value_overflow,
                     const int32_t* __restrict__ touch) {
  const int64_t img = blockIdx.y;
  const int32_t* lab = in + img * in_stride;
  int32_t* Lp = L + img * npx;
  const int32_t* ap = aux + img * npx;
  const int32_t* rk = rank + img * nval;
  const int32_t* tp = touch != nullptr ? touch + img * nval : nullptr;
  auto one = [&](int v, int root) -> int {
    if (v >= nval) {  // outside the declared range: background here, reported to the caller
      if (value_overflow != nullptr) value_overflow[img] = 1;
      return 0;
    }
    if (v <= 0) return 0;
    if (use_ccl && __ldg(ap + root) == -1) return 0;
    return __ldg(rk + v);
  };
  // filtered labelling (touch != null): only the pixels of a border value carry a tile root; everybody else keeps its value's rank
  auto one_filtered = [&](int v, int64_t p) -> int {
    if (v >= nval) {
      if (value_overflow != nullptr) value_overflow[img] = 1;
      return 0;
    }
    if (v <= 0) return 0;
    if (__ldg(tp + v) && __ldg(ap + Lp[p]) == -1) return 0;
    return __ldg(rk + v);
  };
  const int64_t step = (int64_t)gridDim.x * 256 * 4;
  for (int64_t p0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; p0 < npx; p0 += step) {
    if (vec && p0 + 3 < npx && use_ccl && tp != nullptr) {
      const int4 v = *reinterpret_cast<const int4*>(lab + p0);
      int4 o = make_int4(0, 0, 0, 0);
      if ((v.x | v.y | v.z | v.w) != 0) o = make_int4(one_filtered(v.x, p0), one_filtered(v.y, p0 + 1), one_filtered(v.z, p0 + 2), one_filtered(v.w, p0 + 3));
      *reinterpret_cast<int4*>(Lp + p0) = o;
    } else if (use_ccl && tp != nullptr) {
      for (int i = 0; i < 4 && p0 + i < npx; ++i) Lp[p0 + i] = one_filtered(lab[p0 + i], p0 + i);
    } else if (vec && p0 + 3 < npx) {
      const int4 v = *reinterpret_cast<const int4*>(lab + p0);
      int4 r = make_int4(0, 0, 0, 0);
      if (use_ccl) r = *reinterpret_cast<const int4*>(Lp + p0);
      *reinterpret_cast<int4*>(Lp + p0) = make_int4(one(v.x, r.x), one(v.y, r.y), one(v.z, r.z), one(v.w, r.w));
    } else {
      for (int i = 0; i < 4 && p0 + i < npx; ++i) Lp[p0 + i] = one(lab[p0 + i], use_ccl ? Lp[p0 + i] : 0);
    }
  }
}

// integer masks without border clearing: presence straight from the pixels
__global__ void __launch_bounds__(256)
present_mark_kernel(const int32_t* __restrict__ in, const int64_t in_stride, const int64_t npx,
                    int32_t* __restrict__ present, const int64_t nval) {
  const int64_t img = blockIdx.y;
  const int32_t* lab = in + img * in_stride;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npx; p += step) {
    const int v = lab[p];
    if (v > 0 && v < nval) present[img * nval + v] = 1;
  }
}

// one block per plane: in-place scan of `len` ints, total to counts[img]
__global__ void __launch_bounds__(1024)
scan_kernel(int32_t* __restrict__ vals, int len, int32_t* __restrict__ counts, int inclusive) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.x;
  int32_t* v = vals + img * len;
  int carry = 0;
  for (int base = 0; base < len; base += 1024) {
    const int i = base + threadIdx.x;
    const int x = i < len ? v[i] : 0;
    const int ex = block_excl_scan(x, s_warp, &s_total);
    if (i < len) v[i] = carry + ex + (inclusive ? x : 0);
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[img] = carry;
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct LabelScratch {
  int32_t* gid;       // npx per image: id / survival flag of the tile roots (sparse)
  int32_t* rootlist;  // npx per image: the tile roots (capacity: every pixel a root)
  int32_t* rootcnt;   // 1 per image
  uint32_t* rootbits; // one bit per pixel and image: "is a surviving global root"
  int32_t* prefix;    // one int per bitmap word (prefix inside its 1024-word block)
  int32_t* blocksum;  // one int per 1024-word block (exclusive prefix over the plane after ccl_rank_top_kernel)
  int nblocks;
  int32_t* present;
  int32_t* touch;     // integer masks: value occurs on the image border (right behind `present`: cleared together)
  int64_t words;
  size_t zero_bytes;  // rootcnt + rootbits are contiguous and cleared per call
  size_t total;
};

static LabelScratch label_scratch_layout(void* base, int64_t n_img, int64_t h, int64_t w, int64_t max_value) {
  LabelScratch s;
  const int64_t npx = h * w;
  s.words = ceil_div(npx, 32);
  size_t off = 0;
  s.gid = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * npx * sizeof(int32_t));
  s.rootlist = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * npx * sizeof(int32_t));
  s.rootcnt = (int32_t*)((char*)base + off);
  const size_t zero_start = off;
  off += align256((size_t)n_img * sizeof(int32_t));
  s.rootbits = (uint32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.words * sizeof(uint32_t));
  s.zero_bytes = off - zero_start;
  s.prefix = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.words * sizeof(int32_t));
  s.nblocks = (int)ceil_div(s.words, RANK_BLOCK);
  s.blocksum = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.nblocks * sizeof(int32_t));
  s.present = (int32_t*)((char*)base + off);
  off += (size_t)n_img * (size_t)(max_value + 1) * sizeof(int32_t);
  s.touch = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * (size_t)(max_value + 1) * sizeof(int32_t));
  s.total = off;
  return s;
}

// tile -> seams -> roots (-> border): afterwards L[p] = tile root of p (or -1), L[tile root] = global
// root, and rootbits marks the surviving global roots
template <int KIND>
static int ccl_core(const void* in, int64_t in_stride, const double* thresholds, int64_t n_img, int h, int w,
                    int clear_border, int32_t* L, const LabelScratch& s, cudaStream_t st, const int32_t* touch = nullptr,
                    int nval = 0) {
  const int64_t npx = (int64_t)h * w;
  AMT_CUDA_TRY(cudaMemsetAsync(s.rootcnt, 0, s.zero_bytes, st));
  dim3 tgrid((unsigned)ceil_div(w, TW), (unsigned)ceil_div(h, TH), (unsigned)n_img);
  ccl_tile_kernel<KIND><<<tgrid, TILE_THREADS, 0, st>>>(in, in_stride, thresholds, L, s.rootlist, s.rootcnt, h, w, touch, s.present, nval);
  AMT_LAUNCH_CHECK();
  const int n_hseams = (int)ceil_div(h, TH) - 1, n_vseams = (int)ceil_div(w, TW) - 1;
  const int64_t seam_px = (int64_t)n_hseams * w + (int64_t)n_vseams * 2 * h;
  if (seam_px > 0) {
    ccl_seam_kernel<KIND><<<dim3((unsigned)ceil_div(seam_px, 256), (unsigned)n_img), 256, 0, st>>>(in, in_stride, L, h, w,
                                                                                                 n_hseams, n_vseams, touch, nval);
    AMT_LAUNCH_CHECK();
  }
  ccl_roots_kernel<<<dim3(64, (unsigned)n_img), 256, 0, st>>>(L, s.rootlist, s.rootcnt, s.rootbits, npx, s.words);
  AMT_LAUNCH_CHECK();
  if (clear_border) {
    ccl_border_kernel<<<dim3((unsigned)ceil_div(2 * (int64_t)w + 2 * h, 256), (unsigned)n_img), 256, 0, st>>>(
        L, s.rootbits, h, w, s.words, touch != nullptr ? (const int32_t*)in : nullptr, in_stride, nval);
    AMT_LAUNCH_CHECK();
  }
  return AMT_OK;
}

int label_launch(const void* in, int in_kind, int64_t in_stride, const double* thresholds, int64_t max_value,
                 int64_t n_img, int64_t h, int64_t w, int clear_border, int32_t* labels_out, int32_t* counts,
                 void* scratch, size_t scratch_bytes, cudaStream_t st, int32_t* value_overflow) {
  if (!in || !labels_out || !counts || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || h <= 0 || w <= 0 || in_kind < 0 || in_kind > 2 || max_value < 0) return AMT_ERR_INVALID;
  if (in_kind == 1 && !thresholds) return AMT_ERR_INVALID;
  if (h * w >= (1ll << 31) - 4096 || n_img > 65535 || ceil_div(h, TH) > 65535) return AMT_ERR_CAPACITY;
  if (in_kind == 2 && in == (const void*)labels_out) return AMT_ERR_INVALID;
  const int64_t npx = h * w;
  const int64_t mv = in_kind == 2 ? max_value : 0;
  LabelScratch s = label_scratch_layout(scratch, n_img, h, w, mv);
  if (scratch_bytes < s.total) return AMT_ERR_CAPACITY;
  const int vec = (npx % 4 == 0) && (((uintptr_t)labels_out) % 16 == 0);
  int64_t sb = ceil_div(npx, 256 * 16);
  const int64_t cap = ceil_div((int64_t)kNumSMs * 8, n_img);
  if (sb > cap) sb = cap;
  if (sb < 1) sb = 1;
  const dim3 sgrid((unsigned)sb, (unsigned)n_img);
  const dim3 rgrid(64, (unsigned)n_img);

  if (in_kind == 2) {
    const int64_t nval = max_value + 1;
    const bool filtered = clear_border && g_ccl_touch_filter && nval < (1ll << 31);
    AMT_CUDA_TRY(cudaMemsetAsync(s.present, 0, (size_t)n_img * nval * sizeof(int32_t) * (filtered ? 2 : 1), st));
    if (clear_border) {
      if (filtered) {
        ccl_touch_kernel<<<dim3((unsigned)ceil_div(2 * w + 2 * h, 256), (unsigned)n_img), 256, 0, st>>>(
            (const int32_t*)in, in_stride, s.touch, (int)h, (int)w, (int)nval);
        AMT_LAUNCH_CHECK();
      }
      AMT_TRY(ccl_core<2>(in, in_stride, nullptr, n_img, (int)h, (int)w, 1, labels_out, s, st, filtered ? s.touch : nullptr, (int)nval));
      ccl_present_kernel<<<rgrid, 256, 0, st>>>((const int32_t*)in, in_stride, s.rootlist, s.rootcnt, s.rootbits, npx,
                                                s.words, s.present, nval);
      AMT_LAUNCH_CHECK();
      ccl_gid_kernel<1><<<rgrid, 256, 0, st>>>(labels_out, s.rootlist, s.rootcnt, s.rootbits, s.prefix, s.blocksum,
                                               s.nblocks, s.gid, npx, s.words);
      AMT_LAUNCH_CHECK();
    } else {
      present_mark_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, npx, s.present, nval);
      AMT_LAUNCH_CHECK();
    }
    scan_kernel<<<(unsigned)n_img, 1024, 0, st>>>(s.present, (int)nval, counts, 1);
    AMT_LAUNCH_CHECK();
    relabel_final_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, labels_out, s.gid, npx, s.present, nval,
                                                clear_border, vec && (in_stride % 4 == 0) && (((uintptr_t)in) % 16 == 0),
                                                value_overflow, filtered ? s.touch : nullptr);
    AMT_LAUNCH_CHECK();
    return AMT_OK;
  }
  if (in_kind == 0)
    AMT_TRY(ccl_core<0>(in, in_stride, nullptr, n_img, (int)h, (int)w, clear_border, labels_out, s, st));
  else
    AMT_TRY(ccl_core<1>(in, in_stride, thresholds, n_img, (int)h, (int)w, clear_border, labels_out, s, st));
  ccl_rank_block_kernel<<<dim3((unsigned)s.nblocks, (unsigned)n_img), 256, 0, st>>>(s.rootbits, s.prefix, s.blocksum,
                                                                                    s.words, s.nblocks);
  AMT_LAUNCH_CHECK();
  ccl_rank_top_kernel<<<(unsigned)n_img, 1024, 0, st>>>(s.blocksum, s.nblocks, counts);
  AMT_LAUNCH_CHECK();
  ccl_gid_kernel<0><<<rgrid, 256, 0, st>>>(labels_out, s.rootlist, s.rootcnt, s.rootbits, s.prefix, s.blocksum, s.nblocks,
                                           s.gid, npx, s.words);
  AMT_LAUNCH_CHECK();
  ccl_final_kernel<<<sgrid, 256, 0, st>>>(labels_out, s.gid, npx, vec);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt

extern "C" {

size_t amt_label_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_value) {
  return amt::label_scratch_layout(nullptr, n_img, h, w, max_value < 0 ? 0 : max_value).total;
}

int amt_label(const void* in, int in_kind, const double* thresholds, int64_t max_value, int64_t n_img, int64_t h,
              int64_t w, int clear_border, int32_t* labels_out, int32_t* counts, void* scratch, size_t scratch_bytes,
              amt_stream_t stream) {
  return amt::label_launch(in, in_kind, h * w, thresholds, max_value, n_img, h, w, clear_border, labels_out, counts,
                           scratch, scratch_bytes, amt::as_stream(stream), nullptr);
}

}  // extern "C"
