// Perimeter and convex-hull pixel count per label (the two scikit-image regionprops that need
// more than sums): table columns `perimeter` and `area_convex` (solidity / circularity follow).
//
// Reference path: masks.py:15-28 puts `perimeter`, `area_convex`, `solidity`, `circularity` in
// the DEFAULT property list of SegmentationMask.cell_properties (masks.py:286-297) [3p]
// skimage.measure.perimeter(neighborhood=4) and skimage.morphology.convex_hull_image.
// SURVEY.md 8a item 10.
//
//  * perimeter: a label pixel is a border pixel when one of its 4-neighbours (or the image
//    edge) is not the same label (= the crop minus its erosion by the cross); every border
//    pixel gets the code 1 + 2*(border 4-neighbours) + 10*(border diagonal neighbours) of
//    skimage's 3x3 convolution and falls in one of three weight classes (1, sqrt 2,
//    (1+sqrt 2)/2).  One pixel pass counts the classes per label with integer atomics (counts
//    are exact); the weighted sum is done once per label.
//  * convex area: skimage takes the hull of the four edge midpoints (r+-0.5, c), (r, c+-0.5)
//    of the pixels and counts the integer grid points inside or on it.  Only the leftmost /
//    rightmost pixel of each row can contribute hull vertices, so the same pixel pass records
//    per-(label,row) column extents (atomics at run ends only); then one warp per label walks
//    the left and right hull chains by gift wrapping (lanes search the next vertex) in exact
//    integer arithmetic on coordinates scaled by 2, converts every hull edge into the first /
//    last lattice column of each integer row it spans, and sums the row widths.
// All integer work is exact, so area_convex is bit-exact against the reference.

#include <climits>

#include "internal.cuh"

namespace amt {

enum { F_COUNT = 0, F_RMIN = 6, F_RMAX = 7 };  // accumulator fields (regions.cu)

struct ShapeScratch {
  int32_t* off;      // [n_img][max_labels + 1] row-table offset of every label
  int32_t* cmin;     // [n_img][cap] leftmost column per (label,row)
  int32_t* cmax;     // [n_img][cap]
  int32_t* lat_lo;   // [n_img][cap] first lattice column inside the hull
  int32_t* lat_hi;   // [n_img][cap]
  uint32_t* pcls;    // [n_img][3][max_labels] perimeter class counts
  int64_t cap;
  size_t total;
};

static size_t align256s(size_t b) { return (b + 255) / 256 * 256; }

static ShapeScratch shape_layout(void* base, int64_t n_img, int64_t h, int64_t w, int64_t max_labels) {
  ShapeScratch s;
  s.cap = h * w / 4 > 4 * h ? h * w / 4 : 4 * h;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    char* p = (char*)base + o;
    o += align256s(bytes);
    return p;
  };
  s.off = (int32_t*)take((size_t)n_img * (max_labels + 1) * 4);
  s.cmin = (int32_t*)take((size_t)n_img * s.cap * 4);
  s.cmax = (int32_t*)take((size_t)n_img * s.cap * 4);
  s.lat_lo = (int32_t*)take((size_t)n_img * s.cap * 4);
  s.lat_hi = (int32_t*)take((size_t)n_img * s.cap * 4);
  s.pcls = (uint32_t*)take((size_t)n_img * 3 * max_labels * 4);
  s.total = o;
  return s;
}

// one block per image: off[k] = sum of bbox heights of labels < k (exclusive scan), capped
__global__ void __launch_bounds__(1024)
shape_offsets_kernel(const uint64_t* __restrict__ acc, const int n_fields, const int32_t* __restrict__ counts,
                     const int64_t max_labels, int32_t* __restrict__ off, uint32_t* __restrict__ pcls, const int64_t cap) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int64_t img = blockIdx.x;
  const uint64_t* a = acc + img * (int64_t)n_fields * max_labels;
  int32_t* o = off + img * (max_labels + 1);
  int64_t K = counts[img];
  if (K > max_labels) K = max_labels;
  for (int64_t i = threadIdx.x; i < 3 * max_labels; i += 1024) pcls[img * 3 * max_labels + i] = 0;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < K; base += 1024) {
    const int64_t k = base + threadIdx.x;
    int hgt = 0;
    if (k < K && a[(int64_t)F_COUNT * max_labels + k] > 0)
      hgt = (int)(a[(int64_t)F_RMAX * max_labels + k] - a[(int64_t)F_RMIN * max_labels + k] + 1);
    int incl = hgt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int wv = s_warp[lane];
      int wi = wv;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      s_warp[lane] = wi - wv;
    }
    __syncthreads();
    const int carry = s_carry;
    const int64_t ex = (int64_t)carry + s_warp[warp] + incl - hgt;
    if (k < K) o[k] = (ex + hgt <= cap) ? (int)ex : -1;  // -1: no room, label reports NaN
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = (int)(ex + hgt > INT_MAX / 2 ? INT_MAX / 2 : ex + hgt);
    __syncthreads();
  }
}

__global__ void shape_fill_kernel(int32_t* __restrict__ cmin, int32_t* __restrict__ cmax, int64_t total) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    cmin[i] = INT_MAX;
    cmax[i] = -1;
  }
}

// pixel pass: row extents (at run ends) + perimeter classes (at border pixels)
__global__ void __launch_bounds__(256)
shape_pixel_kernel(const int32_t* __restrict__ labels, const uint64_t* __restrict__ acc, const int n_fields,
                   const int32_t* __restrict__ off, const int h, const int w, const int64_t max_labels,
                   int32_t* __restrict__ cmin, int32_t* __restrict__ cmax, uint32_t* __restrict__ pcls, const int64_t cap) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= w || y >= h) return;
  const int32_t* L = labels + img * (int64_t)h * w;
  const int p = y * w + x;
  const int l = L[p];
  if (l <= 0 || l > max_labels) return;
  auto lab = [&](int yy, int xx) -> int { return (yy < 0 || yy >= h || xx < 0 || xx >= w) ? 0 : L[yy * w + xx]; };
  const bool wl = lab(y, x - 1) == l, el = lab(y, x + 1) == l;
  // row extents: only run ends can be the extreme of a row
  const int o = off[img * (max_labels + 1) + (l - 1)];
  if (o >= 0 && (!wl || !el)) {
    const int rmin = (int)acc[img * (int64_t)n_fields * max_labels + (int64_t)F_RMIN * max_labels + (l - 1)];
    const int64_t idx = img * cap + o + (y - rmin);
    if (!wl) atomicMin(cmin + idx, x);
    if (!el) atomicMax(cmax + idx, x);
  }
  // perimeter: border pixel <=> some 4-neighbour is not this label
  auto is_border = [&](int yy, int xx) -> bool {  // (yy, xx) is known to hold label l
    return lab(yy - 1, xx) != l || lab(yy + 1, xx) != l || lab(yy, xx - 1) != l || lab(yy, xx + 1) != l;
  };
  const bool nl = lab(y - 1, x) == l, sl = lab(y + 1, x) == l;
  if (wl && el && nl && sl) return;  // interior pixel
  int n4 = 0, nd = 0;
  if (nl && is_border(y - 1, x)) ++n4;
  if (sl && is_border(y + 1, x)) ++n4;
  if (wl && is_border(y, x - 1)) ++n4;
  if (el && is_border(y, x + 1)) ++n4;
  if (lab(y - 1, x - 1) == l && is_border(y - 1, x - 1)) ++nd;
  if (lab(y - 1, x + 1) == l && is_border(y - 1, x + 1)) ++nd;
  if (lab(y + 1, x - 1) == l && is_border(y + 1, x - 1)) ++nd;
  if (lab(y + 1, x + 1) == l && is_border(y + 1, x + 1)) ++nd;
  const int code = 1 + 2 * n4 + 10 * nd;
  int cls = -1;  // skimage perimeter_weights: {5,7,15,17,25,27} -> 1, {21,33} -> sqrt2, {13,23} -> (1+sqrt2)/2
  if (code == 5 || code == 7 || code == 15 || code == 17 || code == 25 || code == 27) cls = 0;
  else if (code == 21 || code == 33) cls = 1;
  else if (code == 13 || code == 23) cls = 2;
  if (cls >= 0) atomicAdd(pcls + (img * 3 + cls) * max_labels + (l - 1), 1u);
}

__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
  long long q = a / b;
  if ((a % b != 0) && (a < 0)) --q;
  return q;
}

// one warp per label: hull chains by gift wrapping, lattice columns per row, table columns
__global__ void __launch_bounds__(256)
shape_hull_kernel(const uint64_t* __restrict__ acc, const int n_fields, const int n_cols, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ off, const int64_t max_labels, const int32_t* __restrict__ cmin,
                  const int32_t* __restrict__ cmax, int32_t* __restrict__ lat_lo, int32_t* __restrict__ lat_hi,
                  const uint32_t* __restrict__ pcls, const int64_t cap, double* __restrict__ table) {
  const int64_t img = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  int64_t K = counts[img];
  if (K > max_labels) K = max_labels;
  if (k >= K) return;
  const uint64_t* a = acc + img * (int64_t)n_fields * max_labels + k;
  double* t = table + img * (int64_t)n_cols * max_labels + k;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (a[(int64_t)F_COUNT * max_labels] == 0) return;
  if (lane == 0) {
    const uint32_t* pc = pcls + img * 3 * max_labels + k;
    const double sqrt2 = 1.4142135623730951;
    t[14 * max_labels] = (double)pc[0] + (double)pc[max_labels] * sqrt2 + (double)pc[2 * max_labels] * ((1.0 + sqrt2) / 2.0);
  }
  const int o = off[img * (max_labels + 1) + k];
  if (o < 0) {
    if (lane == 0) t[15 * max_labels] = nan;
    return;
  }
  const int rmin = (int)a[(int64_t)F_RMIN * max_labels], rmax = (int)a[(int64_t)F_RMAX * max_labels];
  const int nrows = rmax - rmin + 1;
  const int32_t* lo = cmin + img * cap + o;
  const int32_t* hi = cmax + img * cap + o;
  int32_t* llo = lat_lo + img * cap + o;
  int32_t* lhi = lat_hi + img * cap + o;

  // side 0: left chain (minimise X), side 1: right chain (maximise X).  Coordinates scaled by 2:
  // row r contributes (2r-1, 2c), (2r, 2c -+ 1), (2r+1, 2c) with c its extreme column.
  for (int side = 0; side < 2; ++side) {
    const int32_t* ext = side == 0 ? lo : hi;
    const int sgn = side == 0 ? 1 : -1;  // compare sgn*X: both chains become "minimise"
    int cy = 2 * rmin - 1, cx = 2 * ext[0];  // top vertex of the first row's extreme pixel
    const int y_last = 2 * rmax + 1;
    while (cy < y_last) {
      // next vertex: the point below with the smallest slope d(sgn*X)/dY; ties -> farthest
      long long bdx = 0, bdy = 0;
      int bx = 0, by = 0;
      bool have = false;
      const int r_first = (cy + 1) / 2 - ((cy + 1) % 2 != 0 ? 0 : 0);  // first row whose points may lie below cy
      for (int i = r_first - rmin + lane; i < nrows; i += 32) {
        if (i < 0) continue;
        const int c = ext[i];
        if ((side == 0 && c == INT_MAX) || (side == 1 && c < 0)) continue;  // label absent from this row
        const int r = rmin + i;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int py = 2 * r - 1 + q;
          const int px = 2 * c + (q == 1 ? (side == 0 ? -1 : 1) : 0);
          if (py <= cy) continue;
          const long long dy = py - cy, dx = (long long)sgn * (px - cx);
          // slope dx/dy < bdx/bdy  <=>  dx*bdy < bdx*dy ; equal slope: larger dy wins
          const long long lhs = dx * bdy, rhs = bdx * dy;
          if (!have || lhs < rhs || (lhs == rhs && dy > bdy)) {
            have = true;
            bdx = dx; bdy = dy; bx = px; by = py;
          }
        }
      }
      // warp argmin over (slope, -dy)
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const long long odx = __shfl_xor_sync(0xffffffffu, bdx, d), ody = __shfl_xor_sync(0xffffffffu, bdy, d);
        const int obx = __shfl_xor_sync(0xffffffffu, bx, d), oby = __shfl_xor_sync(0xffffffffu, by, d);
        const int ohave = __shfl_xor_sync(0xffffffffu, (int)have, d);
        if (ohave) {
          const long long lhs = odx * bdy, rhs = bdx * ody;
          if (!have || lhs < rhs || (lhs == rhs && ody > bdy)) {
            have = true;
            bdx = odx; bdy = ody; bx = obx; by = oby;
          }
        }
      }
      if (!have) break;  // cannot happen: the bottom vertex of the last row is always below
      // lattice columns of the integer rows (even Y) in (cy, by]
      for (int Y = cy + 1 + ((cy + 1) & 1) + 2 * lane; Y <= by; Y += 64) {
        // X(Y) = cx + (bx - cx) * (Y - cy) / (by - cy); lattice column c with 2c >= X (left) / 2c <= X (right)
        const long long D = by - cy;
        const long long num = (long long)cx * D + (long long)(bx - cx) * (Y - cy);
        const int r = Y / 2 - rmin;
        if (side == 0)
          llo[r] = (int)(-floor_div(-num, 2 * D));  // ceil(num / (2D))
        else
          lhi[r] = (int)floor_div(num, 2 * D);
      }
      cx = bx;
      cy = by;
    }
    __syncwarp();
  }
  long long area = 0;
  for (int i = lane; i < nrows; i += 32) {
    const int wdt = lhi[i] - llo[i] + 1;
    area += wdt > 0 ? wdt : 0;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) area += __shfl_xor_sync(0xffffffffu, area, d);
  if (lane == 0) t[15 * max_labels] = (double)area;
}

size_t region_shape_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_labels) {
  return shape_layout(nullptr, n_img, h, w, max_labels).total;
}

int region_shape(const int32_t* labels, const uint64_t* acc, int n_channels, const int32_t* counts, int64_t n_img,
                 int64_t h, int64_t w, int64_t max_labels, double* table, void* scratch, size_t scratch_bytes,
                 cudaStream_t st) {
  if (!labels || !acc || !counts || !table || !scratch) return AMT_ERR_INVALID;
  if (n_img <= 0 || h <= 0 || w <= 0 || max_labels <= 0 || n_channels < 0 || n_channels > 8) return AMT_ERR_INVALID;
  if (h * w >= (1ll << 31) || n_img > 65535 || ceil_div(h, 8) > 65535 || max_labels >= (1ll << 28)) return AMT_ERR_CAPACITY;
  ShapeScratch s = shape_layout(scratch, n_img, h, w, max_labels);
  if (scratch_bytes < s.total) return AMT_ERR_CAPACITY;
  const int n_fields = AMT_ACC_FIELDS(n_channels);
  const int n_cols = AMT_TABLE_COLS(n_channels);
  shape_offsets_kernel<<<(unsigned)n_img, 1024, 0, st>>>(acc, n_fields, counts, max_labels, s.off, s.pcls, s.cap);
  AMT_LAUNCH_CHECK();
  shape_fill_kernel<<<kNumSMs * 4, 256, 0, st>>>(s.cmin, s.cmax, n_img * s.cap);
  AMT_LAUNCH_CHECK();
  dim3 pgrid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8), (unsigned)n_img), pblock(32, 8);
  shape_pixel_kernel<<<pgrid, pblock, 0, st>>>(labels, acc, n_fields, s.off, (int)h, (int)w, max_labels, s.cmin, s.cmax,
                                               s.pcls, s.cap);
  AMT_LAUNCH_CHECK();
  shape_hull_kernel<<<dim3((unsigned)ceil_div(max_labels, 8), (unsigned)n_img), 256, 0, st>>>(
      acc, n_fields, n_cols, counts, s.off, max_labels, s.cmin, s.cmax, s.lat_lo, s.lat_hi, s.pcls, s.cap, table);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt
