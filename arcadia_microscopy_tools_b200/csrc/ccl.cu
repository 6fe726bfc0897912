// Connected-component labelling, border clearing and sequential relabelling.
//
// Reference path: masks.py:38-65 (_process_mask): ski.segmentation.clear_border (:56) [3p] then
// ski.measure.label (:63) [3p] for bool masks or ski.segmentation.relabel_sequential (:65) [3p]
// for integer masks.  SURVEY.md 8a items 7-9.
//
// Block-based union-find CCL, 8-connectivity, on "same non-zero value" adjacency (binary masks
// are the one-value case; integer masks give clear_border's re-labelling by value+connectivity):
//  A. tile:     a CTA labels a 64x32 tile entirely in shared memory — warp-level run linking
//               (one ballot per 32-pixel row segment), then lock-free unions (shared-memory
//               atomicMin, roots only) with the row above using the N / W / NW / NE decision
//               tree — and writes each pixel's tile root as a GLOBAL pixel index (one 4-byte
//               store per pixel; the input is read once, thresholded on the fly if it is a
//               float64 plane).
//  B. seams:    only the pixels on tile seams union across tiles (global atomicMin).  This is
//               the warp-level boundary-merge pass: ~6 % of the pixels.
//  C. compress: every pixel points at its root.  Roots are the smallest pixel index of the
//               component = its first pixel in raster order, which is exactly the numbering
//               key of scipy.ndimage.label / skimage.  Border pixels flag their root; each CTA
//               leaves its roots as an ordered list (block scan, no arrival-order atomics).
//  D. number:   one CTA per plane walks the root lists in raster order and gives the
//               surviving roots consecutive ids (integer masks: marks the surviving VALUES,
//               which a presence-table scan then ranks = relabel_sequential).
//  E. final:    gather the id through the root, in place, 16-byte accesses.
// Pixel traffic: A 4-12 B, C 8 B, E 8 B + a cached gather: HBM-bound streaming.

#include "common.cuh"

namespace amt {

constexpr int TW = 64, TH = 32;      // tile of kernel A
constexpr int CBLK = 1024;           // pixels per CTA in the linear kernels (256 threads x 4)
constexpr int ROOT_CAP = CBLK;       // root-list slots per CTA (trivial upper bound: every pixel a root)

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
// grid (ceil(w/64), ceil(h/32), planes); block 256 = 8 warps; warp -> rows warp, warp+8, ...
template <int KIND>
__global__ void __launch_bounds__(256)
ccl_tile_kernel(const void* __restrict__ in, const int64_t in_stride, const double* __restrict__ thresholds,
                int32_t* __restrict__ L, int32_t* __restrict__ aux, const int h, const int w) {
  __shared__ int s_lab[TH * TW];
  __shared__ int s_val[KIND == 2 ? TH * TW : 1];
  const int64_t img = blockIdx.z;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double thr = (KIND == 1) ? thresholds[img] : 0.0;
  const int64_t in_base = img * in_stride;
  const int64_t plane = img * (int64_t)h * w;

#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int lr = warp + 8 * q, y = ty0 + lr;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lc = half * 32 + lane, x = tx0 + lc;
      const int v = (y < h && x < w) ? ccl_value<KIND>(in, thr, in_base + (int64_t)y * w + x) : 0;
      const int vl = __shfl_up_sync(0xffffffffu, v, 1);
      const unsigned brk = __ballot_sync(0xffffffffu, lane == 0 || v != vl);
      const int start = 31 - __clz((int)(brk & (0xffffffffu >> (31 - lane))));
      s_lab[lr * TW + lc] = v ? (lr * TW + half * 32 + start) : -1;
      if (KIND == 2) s_val[lr * TW + lc] = v;
    }
  }
  __syncthreads();

#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int lr = warp + 8 * q;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lc = half * 32 + lane;
      const int p = lr * TW + lc;
      bool fg, n_s, w_s, nw_s, ne_s;
      if (KIND == 2) {
        const int v = s_val[p];
        fg = v != 0;
        w_s = fg && lc > 0 && s_val[p - 1] == v;
        n_s = fg && lr > 0 && s_val[p - TW] == v;
        nw_s = fg && lr > 0 && lc > 0 && s_val[p - TW - 1] == v;
        ne_s = fg && lr > 0 && lc < TW - 1 && s_val[p - TW + 1] == v;
      } else {
        fg = s_lab[p] >= 0;  // labels only ever move between non-negative values
        w_s = fg && lc > 0 && s_lab[p - 1] >= 0;
        n_s = fg && lr > 0 && s_lab[p - TW] >= 0;
        nw_s = fg && lr > 0 && lc > 0 && s_lab[p - TW - 1] >= 0;
        ne_s = fg && lr > 0 && lc < TW - 1 && s_lab[p - TW + 1] >= 0;
      }
      if (fg) {
        if (n_s) {
          // W and NW both set: p ~ W (run link) ~ NW ~ N already, except across a segment seam
          if (!(w_s && nw_s && lane != 0)) suf_union(s_lab, p, p - TW);
        } else {
          if (w_s) {
            if (lane == 0) suf_union(s_lab, p, p - 1);  // inside a segment the run is linked
          } else if (nw_s) {
            suf_union(s_lab, p, p - TW - 1);
          }
          if (ne_s) suf_union(s_lab, p, p - TW + 1);
        }
      }
    }
  }
  __syncthreads();

#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int lr = warp + 8 * q, y = ty0 + lr;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lc = half * 32 + lane, x = tx0 + lc;
      if (y < h && x < w) {
        const int p = lr * TW + lc;
        int out = -1;
        if (s_lab[p] >= 0) {
          const int r = suf_find(s_lab, p);
          out = (ty0 + r / TW) * w + tx0 + (r % TW);
          if (r == p) aux[plane + (int64_t)y * w + x] = 0;  // flag / id slot of a (tile) root
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
                const int n_hseams, const int n_vseams) {
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

// ---------------------------------------------------------------- C. compress + roots + border flags
// grid (ceil(npx/1024), planes); block 256; thread -> 4 consecutive pixels
__global__ void __launch_bounds__(256)
ccl_compress_kernel(int32_t* __restrict__ L, int32_t* __restrict__ aux, const int h, const int w, const int clear_border,
                    int32_t* __restrict__ blockcnt, int32_t* __restrict__ rootbuf, const int nblk, const int vec) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.y;
  const int npx = h * w;
  int32_t* Lp = L + img * (int64_t)npx;
  int32_t* ap = aux + img * (int64_t)npx;
  const int p0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  int lab[4] = {-1, -1, -1, -1};
  if (vec && p0 + 3 < npx) {
    const int4 q = *reinterpret_cast<const int4*>(Lp + p0);
    lab[0] = q.x; lab[1] = q.y; lab[2] = q.z; lab[3] = q.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + i < npx) lab[i] = Lp[p0 + i];
  }
  int n_roots = 0;
  int roots[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (lab[i] >= 0) {
      const int p = p0 + i;
      const int r = (lab[i] == p) ? p : uf_find(Lp, lab[i]);
      lab[i] = r;
      if (r == p) roots[n_roots++] = p;
      if (clear_border) {
        const int y = p / w, x = p - y * w;
        if (y == 0 || x == 0 || y == h - 1 || x == w - 1) ap[r] = -1;
      }
    }
  }
  if (vec && p0 + 3 < npx) {
    *reinterpret_cast<int4*>(Lp + p0) = make_int4(lab[0], lab[1], lab[2], lab[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + i < npx) Lp[p0 + i] = lab[i];
  }
  const int off = block_excl_scan(n_roots, s_warp, &s_total);
  int32_t* dst = rootbuf + (img * nblk + blockIdx.x) * (int64_t)ROOT_CAP + off;
  for (int i = 0; i < n_roots; ++i) dst[i] = roots[i];
  if (threadIdx.x == 0) blockcnt[img * nblk + blockIdx.x] = s_total;
}

// ---------------------------------------------------------------- D. numbering (one CTA per plane)
// MODE 0: aux[root] = consecutive id of the surviving roots in raster order (0 for removed),
//         counts[img] = number of survivors.
// MODE 1: integer masks — present[value of root] = 1 for surviving roots (ranked later).
template <int MODE>
__global__ void __launch_bounds__(1024)
ccl_number_kernel(int32_t* __restrict__ aux, const int32_t* __restrict__ blockcnt, const int32_t* __restrict__ rootbuf,
                  const int nblk, const int64_t npx, int32_t* __restrict__ counts, const int32_t* __restrict__ in,
                  const int64_t in_stride, int32_t* __restrict__ present, const int64_t nval) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.x;
  int32_t* ap = aux + img * npx;
  const int32_t* bc = blockcnt + img * nblk;
  const int32_t* rb = rootbuf + img * nblk * (int64_t)ROOT_CAP;
  if (MODE == 1) {
    const int32_t* lab = in + img * in_stride;
    for (int b = threadIdx.x; b < nblk; b += 1024) {
      const int c = bc[b];
      for (int i = 0; i < c; ++i) {
        const int r = rb[(int64_t)b * ROOT_CAP + i];
        if (ap[r] != -1) {
          const int v = lab[r];
          if (v > 0 && v < nval) present[img * nval + v] = 1;
        }
      }
    }
    return;
  }
  int carry = 0;
  for (int base = 0; base < nblk; base += 1024) {
    const int b = base + threadIdx.x;
    int surv = 0;
    if (b < nblk) {
      const int c = bc[b];
      for (int i = 0; i < c; ++i) surv += (ap[rb[(int64_t)b * ROOT_CAP + i]] != -1);
    }
    int k = carry + block_excl_scan(surv, s_warp, &s_total);
    if (b < nblk) {
      const int c = bc[b];
      for (int i = 0; i < c; ++i) {
        const int r = rb[(int64_t)b * ROOT_CAP + i];
        ap[r] = (ap[r] != -1) ? ++k : 0;
      }
    }
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[img] = carry;
}

// ---------------------------------------------------------------- E. final gather (in place)
__global__ void __launch_bounds__(256)
ccl_final_kernel(int32_t* __restrict__ L, const int32_t* __restrict__ aux, const int64_t npx, const int vec) {
  const int64_t img = blockIdx.y;
  int32_t* Lp = L + img * npx;
  const int32_t* ap = aux + img * npx;
  const int64_t step = (int64_t)gridDim.x * 256 * 4;
  for (int64_t p0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; p0 < npx; p0 += step) {
    if (vec && p0 + 3 < npx) {
      int4 q = *reinterpret_cast<const int4*>(Lp + p0);
      q.x = q.x >= 0 ? __ldg(ap + q.x) : 0;
      q.y = q.y >= 0 ? __ldg(ap + q.y) : 0;
      q.z = q.z >= 0 ? __ldg(ap + q.z) : 0;
      q.w = q.w >= 0 ? __ldg(ap + q.w) : 0;
      *reinterpret_cast<int4*>(Lp + p0) = q;
    } else {
      for (int i = 0; i < 4 && p0 + i < npx; ++i) {
        const int r = Lp[p0 + i];
        Lp[p0 + i] = r >= 0 ? ap[r] : 0;
      }
    }
  }
}

// integer masks: out = rank of the pixel's value if its fragment survived (use_ccl) else 0
__global__ void __launch_bounds__(256)
relabel_final_kernel(const int32_t* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ L,
                     const int32_t* __restrict__ aux, const int64_t npx, const int32_t* __restrict__ rank,
                     const int64_t nval, const int use_ccl, const int vec) {
  const int64_t img = blockIdx.y;
  const int32_t* lab = in + img * in_stride;
  int32_t* Lp = L + img * npx;
  const int32_t* ap = aux + img * npx;
  const int32_t* rk = rank + img * nval;
  auto one = [&](int v, int root) -> int {
    if (v <= 0 || v >= nval) return 0;
    if (use_ccl && __ldg(ap + root) == -1) return 0;
    return __ldg(rk + v);
  };
  const int64_t step = (int64_t)gridDim.x * 256 * 4;
  for (int64_t p0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; p0 < npx; p0 += step) {
    if (vec && p0 + 3 < npx) {
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
  int32_t* aux;
  int32_t* blockcnt;
  int32_t* rootbuf;
  int32_t* present;
  int nblk;
  size_t total;
};

static LabelScratch label_scratch_layout(void* base, int64_t n_img, int64_t h, int64_t w, int64_t max_value) {
  LabelScratch s;
  const int64_t npx = h * w;
  s.nblk = (int)ceil_div(npx, CBLK);
  size_t off = 0;
  s.aux = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * npx * sizeof(int32_t));
  s.blockcnt = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.nblk * sizeof(int32_t));
  s.rootbuf = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.nblk * ROOT_CAP * sizeof(int32_t));
  s.present = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * (size_t)(max_value + 1) * sizeof(int32_t));
  s.total = off;
  return s;
}

template <int KIND>
static int ccl_core(const void* in, int64_t in_stride, const double* thresholds, int64_t n_img, int h, int w,
                    int clear_border, int32_t* L, const LabelScratch& s, int vec, cudaStream_t st) {
  dim3 tgrid((unsigned)ceil_div(w, TW), (unsigned)ceil_div(h, TH), (unsigned)n_img);
  ccl_tile_kernel<KIND><<<tgrid, 256, 0, st>>>(in, in_stride, thresholds, L, s.aux, h, w);
  AMT_LAUNCH_CHECK();
  const int n_hseams = (int)ceil_div(h, TH) - 1, n_vseams = (int)ceil_div(w, TW) - 1;
  const int64_t seam_px = (int64_t)n_hseams * w + (int64_t)n_vseams * 2 * h;
  if (seam_px > 0) {
    ccl_seam_kernel<KIND><<<dim3((unsigned)ceil_div(seam_px, 256), (unsigned)n_img), 256, 0, st>>>(in, in_stride, L, h, w,
                                                                                                 n_hseams, n_vseams);
    AMT_LAUNCH_CHECK();
  }
  ccl_compress_kernel<<<dim3((unsigned)s.nblk, (unsigned)n_img), 256, 0, st>>>(L, s.aux, h, w, clear_border, s.blockcnt,
                                                                               s.rootbuf, s.nblk, vec);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int label_launch(const void* in, int in_kind, int64_t in_stride, const double* thresholds, int64_t max_value,
                 int64_t n_img, int64_t h, int64_t w, int clear_border, int32_t* labels_out, int32_t* counts,
                 void* scratch, size_t scratch_bytes, cudaStream_t st) {
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

  if (in_kind == 2) {
    const int64_t nval = max_value + 1;
    AMT_CUDA_TRY(cudaMemsetAsync(s.present, 0, (size_t)n_img * nval * sizeof(int32_t), st));
    if (clear_border) {
      AMT_TRY(ccl_core<2>(in, in_stride, nullptr, n_img, (int)h, (int)w, 1, labels_out, s, vec, st));
      ccl_number_kernel<1><<<(unsigned)n_img, 1024, 0, st>>>(s.aux, s.blockcnt, s.rootbuf, s.nblk, npx, counts,
                                                             (const int32_t*)in, in_stride, s.present, nval);
      AMT_LAUNCH_CHECK();
    } else {
      present_mark_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, npx, s.present, nval);
      AMT_LAUNCH_CHECK();
    }
    scan_kernel<<<(unsigned)n_img, 1024, 0, st>>>(s.present, (int)nval, counts, 1);
    AMT_LAUNCH_CHECK();
    relabel_final_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, labels_out, s.aux, npx, s.present, nval,
                                                clear_border, vec && (in_stride % 4 == 0) && (((uintptr_t)in) % 16 == 0));
    AMT_LAUNCH_CHECK();
    return AMT_OK;
  }
  if (in_kind == 0)
    AMT_TRY(ccl_core<0>(in, in_stride, nullptr, n_img, (int)h, (int)w, clear_border, labels_out, s, vec, st));
  else
    AMT_TRY(ccl_core<1>(in, in_stride, thresholds, n_img, (int)h, (int)w, clear_border, labels_out, s, vec, st));
  ccl_number_kernel<0><<<(unsigned)n_img, 1024, 0, st>>>(s.aux, s.blockcnt, s.rootbuf, s.nblk, npx, counts, nullptr, 0,
                                                         nullptr, 0);
  AMT_LAUNCH_CHECK();
  ccl_final_kernel<<<sgrid, 256, 0, st>>>(labels_out, s.aux, npx, vec);
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
                           scratch, scratch_bytes, amt::as_stream(stream));
}

}  // extern "C"
