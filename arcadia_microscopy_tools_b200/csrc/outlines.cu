// Cell outlines of a label image: the pixel passes behind SegmentationMask.cell_outlines.
//
// Reference path: masks.py:229-245 (cell_outlines) -> masks.py:68-79 (_extract_outlines_cellpose:
// cellpose.utils.outlines_list -> cv2.findContours(label == n, RETR_EXTERNAL, CHAIN_APPROX_NONE),
// longest contour) or masks.py:82-115 (_extract_outlines_skimage: skimage.measure.find_contours on
// the padded crop label == n at level 0.5, longest contour) [3p].  SURVEY.md 8f rank 4.
//
//  * marching squares (skimage): a 2x2 pixel square contributes a contour segment to label L when
//    some but not all of its corners carry L; the case number ul | ur<<1 | ll<<2 | lr<<3 decides
//    which two edge midpoints the segment joins and in which direction.  outline_squares_kernel
//    looks at every square once, for every distinct label at its corners, and appends one 64-bit key
//    label<<34 | r0<<19 | c0<<4 | case per non-trivial (square, label) pair: warp-aggregated, one
//    global atomic per warp.  Sorted keys = per label, the squares in raster order, which is the
//    order skimage's Cython loop emits segments in; joining them into ordered contours (a few dozen
//    points per cell) is host work (masks.py of this package).  HBM-bound: 4 B/px read, keys ~ border
//    pixels.
//  * border following (cv2 / cellpose): OpenCV's outer-border tracer (Suzuki-Abe, 8-connected) is
//    sequential per contour but independent across contours: one thread per candidate start pixel
//    (label pixel with a different west neighbour and no same-label pixel among NW, N, NE) follows
//    its border; the start is genuine when no traced pixel precedes it in raster order.  Pass 1 keeps,
//    per label, the longest border (ties: the later start, which cv2 lists first) with one 64-bit
//    atomicMax; pass 2 re-traces the winners and writes their (y, x) points at host-computed offsets.
//    Known deviation: RETR_EXTERNAL also drops a fragment that lies inside a hole of another fragment
//    of the SAME label; here such a fragment competes by length like any other.

#include "internal.cuh"

namespace amt {

constexpr int OL_LABEL_SHIFT = 34, OL_ROW_SHIFT = 19, OL_COL_SHIFT = 4;

__global__ void __launch_bounds__(256)
outline_squares_kernel(const int32_t* __restrict__ labels, const int h, const int w, uint64_t* __restrict__ keys,
                       const unsigned long long capacity, unsigned long long* __restrict__ count) {
  const int c0 = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool inside = r0 < h - 1 && c0 < w - 1;
  int32_t q[4] = {0, 0, 0, 0};  // ul, ur, ll, lr
  if (inside) {
    const int32_t* p = labels + (int64_t)r0 * w + c0;
    q[0] = p[0];
    q[1] = p[1];
    q[2] = p[w];
    q[3] = p[w + 1];
  }
  uint64_t mine[4];
  int n = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int32_t L = q[i];
    bool fresh = L != 0;
#pragma unroll
    for (int j = 0; j < i; ++j) fresh = fresh && q[j] != L;
    if (fresh) {
      const int cs = (q[0] == L) | ((q[1] == L) << 1) | ((q[2] == L) << 2) | ((q[3] == L) << 3);
      if (cs != 15)
        mine[n++] = ((uint64_t)(uint32_t)L << OL_LABEL_SHIFT) | ((uint64_t)r0 << OL_ROW_SHIFT) |
                    ((uint64_t)c0 << OL_COL_SHIFT) | (uint64_t)cs;
    }
  }
  // warp-aggregated append
  const unsigned lane = threadIdx.x & 31;
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned long long base = 0;
  if (lane == 31) base = atomicAdd(count, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  const unsigned long long at = base + (unsigned long long)(incl - n);
  for (int i = 0; i < n; ++i)
    if (at + i < capacity) keys[at + i] = mine[i];
}

// ---- OpenCV outer-border following.  Direction codes 0..7 = E, NE, N, NW, W, SW, S, SE (x right, y down).
struct Tracer {
  const int32_t* lab;
  int h, w, L;
  __device__ __forceinline__ bool at(int y, int x) const {
    return (unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w && lab[(int64_t)y * w + x] == L;
  }
};

// dx, dy per code, packed two bits each (value + 1): E(1,0) NE(1,-1) N(0,-1) NW(-1,-1) W(-1,0) SW(-1,1) S(0,1) SE(1,1)
__device__ __forceinline__ int code_dx(int s) { return (int)((0x901au >> (2 * s)) & 3u) - 1; }  // 2,2,1,0,0,0,1,2
__device__ __forceinline__ int code_dy(int s) { return (int)((0xa901u >> (2 * s)) & 3u) - 1; }  // 1,0,0,0,1,2,2,2

// Follows the outer border that starts at (y0, x0) (its west neighbour is not L).  Returns the number
// of border points cv2 would list (CHAIN_APPROX_NONE); *min_idx = smallest raster index visited;
// points (optional): (y, x) int32 pairs.  abort_early: give up (return 0) as soon as a pixel that precedes the
// start in raster order turns up - the start is then only a local top of a border that begins earlier, and a
// ragged component with thousands of local tops must not have each of them walk its whole border.
__device__ int64_t trace_border(const Tracer& t, const int y0, const int x0, int64_t* min_idx, int32_t* points,
                                const int64_t max_points, const bool abort_early) {
  int s = 4;
  const int s_stop = 4;
  bool found = false;
  do {
    s = (s - 1) & 7;
    if (t.at(y0 + code_dy(s), x0 + code_dx(s))) {
      found = true;
      break;
    }
  } while (s != s_stop);
  int64_t mn = (int64_t)y0 * t.w + x0;
  if (!found) {  // isolated pixel
    if (points && max_points > 0) {
      points[0] = y0;
      points[1] = x0;
    }
    *min_idx = mn;
    return 1;
  }
  const int y1 = y0 + code_dy(s), x1 = x0 + code_dx(s);
  int y3 = y0, x3 = x0;
  int64_t n = 0;
  for (;;) {
    int y4, x4;
    for (;;) {
      s = (s + 1) & 7;
      y4 = y3 + code_dy(s);
      x4 = x3 + code_dx(s);
      if (t.at(y4, x4)) break;
    }
    if (points && n < max_points) {
      points[2 * n] = y3;
      points[2 * n + 1] = x3;
    }
    ++n;
    const int64_t idx = (int64_t)y3 * t.w + x3;
    if (abort_early && idx < mn) {
      *min_idx = idx;
      return 0;
    }
    mn = idx < mn ? idx : mn;
    if ((y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) || n > 4 * (int64_t)t.h * t.w) break;
    y3 = y4;
    x3 = x4;
    s = (s + 4) & 7;
  }
  *min_idx = mn;
  return n;
}

// best[L] = max over the genuine outer-border starts of label L of (length << 32 | start raster index)
__global__ void __launch_bounds__(256)
outline_find_kernel(const int32_t* __restrict__ labels, const int h, const int w, const int64_t max_labels,
                    unsigned long long* __restrict__ best) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int32_t L = labels[(int64_t)y * w + x];
  if (L <= 0 || L > max_labels) return;
  const Tracer t{labels, h, w, L};
  if (t.at(y, x - 1) || t.at(y - 1, x - 1) || t.at(y - 1, x) || t.at(y - 1, x + 1)) return;
  int64_t mn;
  const int64_t n = trace_border(t, y, x, &mn, nullptr, 0, true);
  const int64_t me = (int64_t)y * w + x;
  if (mn != me) return;  // a local top of a border that starts earlier in raster order
  const unsigned long long len = n > 0xffffffffll ? 0xffffffffull : (unsigned long long)n;
  atomicMax(&best[L - 1], (len << 32) | (unsigned long long)me);
}

// one thread per label: re-trace the winner and write its points at offsets[L-1] (in points, 2 int32 each)
__global__ void __launch_bounds__(128)
outline_write_kernel(const int32_t* __restrict__ labels, const int h, const int w, const int64_t n_labels,
                     const unsigned long long* __restrict__ best, const int64_t* __restrict__ offsets,
                     int32_t* __restrict__ points) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_labels) return;
  const unsigned long long b = best[k];
  if (b == 0) return;
  const int64_t start = (int64_t)(b & 0xffffffffull);
  const Tracer t{labels, h, w, (int)(k + 1)};
  int64_t mn;
  trace_border(t, (int)(start / w), (int)(start % w), &mn, points + 2 * offsets[k], offsets[k + 1] - offsets[k], false);
}

}  // namespace amt

extern "C" {

int amt_outline_squares(const int32_t* labels, int64_t h, int64_t w, uint64_t* keys, int64_t capacity,
                        uint64_t* count, amt_stream_t stream) {
  using namespace amt;
  if (!labels || !count || (!keys && capacity > 0) || h <= 0 || w <= 0 || capacity < 0) return AMT_ERR_INVALID;
  if (h >= (1 << 15) || w >= (1 << 15)) return AMT_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  AMT_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
  if (h < 2 || w < 2) return AMT_OK;
  dim3 grid((unsigned)ceil_div(w - 1, 32), (unsigned)ceil_div(h - 1, 8));
  outline_squares_kernel<<<grid, 256, 0, st>>>(labels, (int)h, (int)w, keys, (unsigned long long)capacity,
                                               (unsigned long long*)count);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_outline_trace_find(const int32_t* labels, int64_t h, int64_t w, int64_t max_labels, uint64_t* best,
                           amt_stream_t stream) {
  using namespace amt;
  if (!labels || !best || h <= 0 || w <= 0 || max_labels <= 0) return AMT_ERR_INVALID;
  if (h * w >= (1ll << 32) || h >= (1ll << 31) || w >= (1ll << 31)) return AMT_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  AMT_CUDA_TRY(cudaMemsetAsync(best, 0, (size_t)max_labels * sizeof(uint64_t), st));
  dim3 grid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8));
  outline_find_kernel<<<grid, 256, 0, st>>>(labels, (int)h, (int)w, max_labels, (unsigned long long*)best);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_outline_trace_write(const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, const uint64_t* best,
                            const int64_t* offsets, int32_t* points, amt_stream_t stream) {
  using namespace amt;
  if (!labels || !best || !offsets || !points || h <= 0 || w <= 0 || n_labels <= 0) return AMT_ERR_INVALID;
  if (h * w >= (1ll << 32)) return AMT_ERR_UNSUPPORTED;
  outline_write_kernel<<<(unsigned)ceil_div(n_labels, 128), 128, 0, as_stream(stream)>>>(
      labels, (int)h, (int)w, n_labels, (const unsigned long long*)best, offsets, points);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
