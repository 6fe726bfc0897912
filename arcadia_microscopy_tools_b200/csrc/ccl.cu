// Connected-component labelling, border clearing and sequential relabelling.
//
// Reference path: masks.py:38-65 (_process_mask): ski.segmentation.clear_border (:56) [3p] then
// ski.measure.label (:63) [3p] for bool masks or ski.segmentation.relabel_sequential (:65) [3p]
// for integer masks.  SURVEY.md 8a items 7-9.
//
// Union-find CCL, 8-connectivity, on "same non-zero value" adjacency (binary masks are the
// one-value case; integer masks give clear_border's re-labelling by value + connectivity).
//  1. init:     warp-level run linking — each pixel's parent is the start of its horizontal run
//               inside the warp's 32-pixel segment (one ballot, no memory traffic).
//  2. merge:    lock-free unions (atomicMin on the parent array, roots only) with the row above
//               and across segment boundaries; redundant unions are skipped with the
//               N / W / NW decision tree.
//  3. compress: every pixel points at its root.  Roots are the smallest linear index of the
//               component = its first pixel in raster order, which is exactly the numbering
//               key scipy.ndimage.label / skimage use.  Border pixels flag their root.
//  4. number:   surviving roots get consecutive ids by a raster-order prefix sum (block counts,
//               one-block scan, block-local ranks) — deterministic, no arrival-order atomics.
//  5. final:    gather the id through the root.  Integer masks instead keep their values,
//               drop flagged fragments, and are renumbered by a presence table scan (sorted
//               value order = relabel_sequential).
// All passes are HBM-bound streaming over 4-byte labels.

#include "common.cuh"

namespace amt {

constexpr int CCL_BLK = 1024;  // pixels per block in the linear (numbering) kernels

template <int KIND>
__device__ __forceinline__ int ccl_value(const void* in, const double thr, const int64_t idx) {
  if (KIND == 0) return ((const uint8_t*)in)[idx] ? 1 : 0;
  if (KIND == 1) return ((const double*)in)[idx] > thr ? 1 : 0;
  return ((const int32_t*)in)[idx];
}

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

// block (32, 8); grid (ceil(w/32), ceil(h/8), n_img).  L holds plane-local linear indices.
template <int KIND>
__global__ void __launch_bounds__(256)
ccl_init_kernel(const void* __restrict__ in, const int64_t in_stride, const double* __restrict__ thresholds,
                int32_t* __restrict__ L, int32_t* __restrict__ aux, int h, int w) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (y >= h) return;  // warp-uniform
  const int64_t plane = img * (int64_t)h * w;
  const double thr = (KIND == 1) ? thresholds[img] : 0.0;
  const int lane = threadIdx.x;
  const int v = (x < w) ? ccl_value<KIND>(in, thr, img * in_stride + (int64_t)y * w + x) : 0;
  const int vl = __shfl_up_sync(0xffffffffu, v, 1);
  const unsigned brk = __ballot_sync(0xffffffffu, lane == 0 || v != vl);
  const int start = 31 - __clz((int)(brk & (0xffffffffu >> (31 - lane))));
  if (x < w) {
    const int64_t idx = plane + (int64_t)y * w + x;
    L[idx] = v ? (y * w + (int)blockIdx.x * 32 + start) : -1;
    aux[idx] = 0;
  }
}

template <int KIND>
__global__ void __launch_bounds__(256)
ccl_merge_kernel(const void* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ L, int h, int w) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (y >= h || x >= w) return;
  const int64_t plane = img * (int64_t)h * w;
  int32_t* Lp = L + plane;
  const int p = y * w + x;
  int v;
  bool n_s, w_s, nw_s, ne_s;
  if (KIND == 2) {
    const int32_t* lab = (const int32_t*)in + img * in_stride;
    v = lab[p];
    if (v == 0) return;
    w_s = x > 0 && lab[p - 1] == v;
    n_s = y > 0 && lab[p - w] == v;
    nw_s = y > 0 && x > 0 && lab[p - w - 1] == v;
    ne_s = y > 0 && x < w - 1 && lab[p - w + 1] == v;
  } else {
    // binary: foreground-ness of a neighbour is readable from the parent array (>= 0)
    if (Lp[p] < 0) return;
    w_s = x > 0 && Lp[p - 1] >= 0;
    n_s = y > 0 && Lp[p - w] >= 0;
    nw_s = y > 0 && x > 0 && Lp[p - w - 1] >= 0;
    ne_s = y > 0 && x < w - 1 && Lp[p - w + 1] >= 0;
  }
  if (n_s) {
    // W and NW both set: p ~ W (run link) ~ NW ~ N already, except across a segment boundary
    if (!(w_s && nw_s && threadIdx.x != 0)) uf_union(Lp, p, p - w);
  } else {
    if (w_s) {
      if (threadIdx.x == 0) uf_union(Lp, p, p - 1);  // inside a segment the run is already linked
    } else if (nw_s) {
      uf_union(Lp, p, p - w - 1);
    }
    if (ne_s) uf_union(Lp, p, p - w + 1);
  }
}

__global__ void __launch_bounds__(256)
ccl_compress_kernel(int32_t* __restrict__ L, int32_t* __restrict__ aux, int h, int w, int clear_border) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (y >= h || x >= w) return;
  const int64_t plane = img * (int64_t)h * w;
  const int p = y * w + x;
  if (L[plane + p] < 0) return;
  const int root = uf_find(L + plane, p);
  L[plane + p] = root;
  if (clear_border && (y == 0 || x == 0 || y == h - 1 || x == w - 1)) aux[plane + root] = -1;
}

__device__ __forceinline__ int block_excl_scan_1024(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int wv = s_warp[lane];
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

// surviving roots per block of CCL_BLK pixels
__global__ void __launch_bounds__(CCL_BLK)
ccl_count_kernel(const int32_t* __restrict__ L, const int32_t* __restrict__ aux, int64_t npx, int32_t* __restrict__ blockcnt,
                 int nblk) {
  const int64_t img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * CCL_BLK + threadIdx.x;
  const int64_t plane = img * npx;
  const bool is_root = p < npx && L[plane + p] == (int32_t)p && aux[plane + p] != -1;
  const int c = __syncthreads_count(is_root);
  if (threadIdx.x == 0) blockcnt[img * nblk + blockIdx.x] = c;
}

// one block per plane: in-place exclusive scan of `len` ints, total to counts[img]
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
    const int ex = block_excl_scan_1024(x, s_warp, &s_total);
    if (i < len) v[i] = carry + ex + (inclusive ? x : 0);
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[img] = carry;
}

__global__ void __launch_bounds__(CCL_BLK)
ccl_assign_kernel(const int32_t* __restrict__ L, int32_t* __restrict__ aux, int64_t npx, const int32_t* __restrict__ blockoff,
                  int nblk) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int64_t img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * CCL_BLK + threadIdx.x;
  const int64_t plane = img * npx;
  const bool is_root = p < npx && L[plane + p] == (int32_t)p && aux[plane + p] != -1;
  const int rank = block_excl_scan_1024(is_root ? 1 : 0, s_warp, &s_total);
  if (is_root) aux[plane + p] = blockoff[img * nblk + blockIdx.x] + rank + 1;
}

__global__ void __launch_bounds__(256)
ccl_final_kernel(int32_t* __restrict__ L, const int32_t* __restrict__ aux, int64_t npx) {
  const int64_t img = blockIdx.y;
  const int64_t plane = img * npx;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npx; p += step) {
    const int r = L[plane + p];
    int out = 0;
    if (r >= 0) {
      const int a = aux[plane + r];
      out = a > 0 ? a : 0;
    }
    L[plane + p] = out;
  }
}

// integer masks: presence of every surviving value
__global__ void __launch_bounds__(256)
present_mark_kernel(const int32_t* __restrict__ in, const int64_t in_stride, const int32_t* __restrict__ L,
                    const int32_t* __restrict__ aux, int64_t npx, int32_t* __restrict__ present, int64_t nval, int use_ccl) {
  const int64_t img = blockIdx.y;
  const int64_t plane = img * npx;
  in += img * in_stride - plane;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npx; p += step) {
    const int v = in[plane + p];
    if (v > 0 && v < nval) {
      bool keep = true;
      if (use_ccl) keep = aux[plane + L[plane + p]] != -1;
      if (keep) present[img * nval + v] = 1;
    }
  }
}

__global__ void __launch_bounds__(256)
relabel_final_kernel(const int32_t* __restrict__ in, const int64_t in_stride, int32_t* __restrict__ L,
                     const int32_t* __restrict__ aux, int64_t npx, const int32_t* __restrict__ rank, int64_t nval, int use_ccl) {
  const int64_t img = blockIdx.y;
  const int64_t plane = img * npx;
  in += img * in_stride - plane;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npx; p += step) {
    const int v = in[plane + p];
    int out = 0;
    if (v > 0 && v < nval) {
      bool keep = true;
      if (use_ccl) keep = aux[plane + L[plane + p]] != -1;
      if (keep) out = rank[img * nval + v];
    }
    L[plane + p] = out;
  }
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct LabelScratch {
  int32_t* aux;
  int32_t* blockcnt;
  int32_t* present;
  int nblk;
  size_t total;
};

static LabelScratch label_scratch_layout(void* base, int64_t n_img, int64_t h, int64_t w, int64_t max_value) {
  LabelScratch s;
  const int64_t npx = h * w;
  s.nblk = (int)ceil_div(npx, CCL_BLK);
  size_t off = 0;
  s.aux = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * npx * sizeof(int32_t));
  s.blockcnt = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * s.nblk * sizeof(int32_t));
  s.present = (int32_t*)((char*)base + off);
  off += align256((size_t)n_img * (size_t)(max_value + 1) * sizeof(int32_t));
  s.total = off;
  return s;
}

template <int KIND>
static int ccl_core(const void* in, int64_t in_stride, const double* thresholds, int64_t n_img, int h, int w,
                    int clear_border, int32_t* L, int32_t* aux, cudaStream_t st) {
  dim3 block(32, 8), grid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8), (unsigned)n_img);
  ccl_init_kernel<KIND><<<grid, block, 0, st>>>(in, in_stride, thresholds, L, aux, h, w);
  AMT_LAUNCH_CHECK();
  ccl_merge_kernel<KIND><<<grid, block, 0, st>>>(in, in_stride, L, h, w);
  AMT_LAUNCH_CHECK();
  ccl_compress_kernel<<<grid, block, 0, st>>>(L, aux, h, w, clear_border);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int label_launch(const void* in, int in_kind, int64_t in_stride, const double* thresholds, int64_t max_value,
                 int64_t n_img, int64_t h, int64_t w, int clear_border, int32_t* labels_out, int32_t* counts,
                 void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (!in || !labels_out || !counts || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || h <= 0 || w <= 0 || in_kind < 0 || in_kind > 2 || max_value < 0) return AMT_ERR_INVALID;
  if (in_kind == 1 && !thresholds) return AMT_ERR_INVALID;
  if (h * w >= (1ll << 31) || n_img > 65535 || ceil_div(h, 8) > 65535) return AMT_ERR_CAPACITY;
  if (in_kind == 2 && in == (const void*)labels_out) return AMT_ERR_INVALID;
  const int64_t npx = h * w;
  const int64_t mv = in_kind == 2 ? max_value : 0;
  LabelScratch s = label_scratch_layout(scratch, n_img, h, w, mv);
  if (scratch_bytes < s.total) return AMT_ERR_CAPACITY;
  int64_t sb = ceil_div(npx, 256 * 8);
  const int64_t cap = ceil_div((int64_t)kNumSMs * 8, n_img);
  if (sb > cap) sb = cap;
  if (sb < 1) sb = 1;
  const dim3 sgrid((unsigned)sb, (unsigned)n_img);

  if (in_kind == 2) {
    const int64_t nval = max_value + 1;
    if (clear_border) AMT_TRY(ccl_core<2>(in, in_stride, nullptr, n_img, (int)h, (int)w, 1, labels_out, s.aux, st));
    AMT_CUDA_TRY(cudaMemsetAsync(s.present, 0, (size_t)n_img * nval * sizeof(int32_t), st));
    present_mark_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, labels_out, s.aux, npx, s.present, nval, clear_border);
    AMT_LAUNCH_CHECK();
    scan_kernel<<<(unsigned)n_img, 1024, 0, st>>>(s.present, (int)nval, counts, 1);
    AMT_LAUNCH_CHECK();
    relabel_final_kernel<<<sgrid, 256, 0, st>>>((const int32_t*)in, in_stride, labels_out, s.aux, npx, s.present, nval, clear_border);
    AMT_LAUNCH_CHECK();
    return AMT_OK;
  }
  if (in_kind == 0)
    AMT_TRY(ccl_core<0>(in, in_stride, nullptr, n_img, (int)h, (int)w, clear_border, labels_out, s.aux, st));
  else
    AMT_TRY(ccl_core<1>(in, in_stride, thresholds, n_img, (int)h, (int)w, clear_border, labels_out, s.aux, st));
  const dim3 lgrid((unsigned)s.nblk, (unsigned)n_img);
  ccl_count_kernel<<<lgrid, CCL_BLK, 0, st>>>(labels_out, s.aux, npx, s.blockcnt, s.nblk);
  AMT_LAUNCH_CHECK();
  scan_kernel<<<(unsigned)n_img, 1024, 0, st>>>(s.blockcnt, s.nblk, counts, 0);
  AMT_LAUNCH_CHECK();
  ccl_assign_kernel<<<lgrid, CCL_BLK, 0, st>>>(labels_out, s.aux, npx, s.blockcnt, s.nblk);
  AMT_LAUNCH_CHECK();
  ccl_final_kernel<<<sgrid, 256, 0, st>>>(labels_out, s.aux, npx);
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
