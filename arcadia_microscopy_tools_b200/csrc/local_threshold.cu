// Local-window thresholds of apply_threshold: niblack, sauvola (box mean / standard deviation) and the
// comparison against a per-pixel threshold image (threshold_local).
//
// Reference path: operations.py:185-196, :214-216 -> skimage.filters.threshold_niblack /
// threshold_sauvola / threshold_local [3p] (skimage 0.25.2 filters/thresholding.py).  SURVEY.md 8f rank 4.
//
// skimage's _mean_std pads the image (np.pad mode='reflect': the edge sample is NOT repeated) by
// (w//2 + 1, w//2), builds float64 integral images of the padded image and of its square, and takes
// window sums as four-corner differences.  For integer images every one of those float64 sums is an
// exact integer as long as it stays below 2^53 (the host checks 4 * sum(x^2) < 2^53 before calling),
// so the window sums are simply the exact integer sums over the mirrored (2*ry+1) x (2*rx+1) window and
// can be formed in any order: here a shared-memory tile, row sums, then column sums, in uint64.
//   m = S / N;  g2 = Q / N;  s = sqrt(max(g2 - m*m, 0))            (each operation rounded once, as NumPy)
//   niblack: m - k*s          sauvola: m * (1 + k * (s/r - 1))      mask = x > threshold
// HBM-bound: 2 B/px read + 1 B/px mask (+ 8 B/px when the threshold image is requested).

#include "internal.cuh"

namespace amt {

constexpr int LT_TILE = 32;

__device__ __forceinline__ int mirror_index(int i, const int n) {  // np.pad 'reflect' / scipy 'mirror'
  if (n == 1) return 0;
  const int period = 2 * n - 2;
  i %= period;
  i = i < 0 ? i + period : i;
  return i < n ? i : period - i;
}

// grid (ceil(w/32), ceil(h/32), n_img); block 256 = 32 x 8
__global__ void __launch_bounds__(256)
window_threshold_kernel(const uint16_t* __restrict__ in, const int h, const int w, const int ry, const int rx,
                        const int kind, const double k, const double r, uint8_t* __restrict__ mask,
                        double* __restrict__ thr_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rows = LT_TILE + 2 * ry, cols = LT_TILE + 2 * rx;
  const int pitch = cols | 1;
  uint64_t* hq = reinterpret_cast<uint64_t*>(smem_raw);               // [rows][32] row sums of squares
  uint32_t* hs = reinterpret_cast<uint32_t*>(hq + (size_t)rows * LT_TILE);  // [rows][32] row sums
  uint16_t* raw = reinterpret_cast<uint16_t*>(hs + (size_t)rows * LT_TILE);  // [rows][pitch]
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * LT_TILE, y0 = blockIdx.y * LT_TILE;
  const int64_t plane = (int64_t)blockIdx.z * h * w;
  const uint16_t* src = in + plane;
  for (int i = tid; i < rows * cols; i += 256) {
    const int rr = i / cols, cc = i - rr * cols;
    const int y = mirror_index(y0 - ry + rr, h), x = mirror_index(x0 - rx + cc, w);
    raw[rr * pitch + cc] = __ldg(src + (int64_t)y * w + x);
  }
  __syncthreads();
  for (int i = tid; i < rows * LT_TILE; i += 256) {
    const int rr = i >> 5, cc = i & 31;
    const uint16_t* p = raw + rr * pitch + cc;
    uint32_t s = 0;
    uint64_t q = 0;
    for (int d = 0; d <= 2 * rx; ++d) {
      const uint32_t v = p[d];
      s += v;
      q += (uint64_t)(v * v);
    }
    hs[i] = s;
    hq[i] = q;
  }
  __syncthreads();
  const int cx = tid & 31;
  const double n_window = (double)((2 * ry + 1) * (2 * rx + 1));
  for (int oy = tid >> 5; oy < LT_TILE; oy += 8) {
    const int y = y0 + oy, x = x0 + cx;
    if (y >= h || x >= w) continue;
    uint64_t S = 0, Q = 0;
    for (int d = 0; d <= 2 * ry; ++d) {
      S += hs[(oy + d) * LT_TILE + cx];
      Q += hq[(oy + d) * LT_TILE + cx];
    }
    const double m = ddiv((double)S, n_window);
    const double g2 = ddiv((double)Q, n_window);
    double var = dsub(g2, dmul(m, m));
    var = var > 0.0 ? var : 0.0;
    const double s = __dsqrt_rn(var);
    const double t = kind == 0 ? dsub(m, dmul(k, s)) : dmul(m, dadd(1.0, dmul(k, dsub(ddiv(s, r), 1.0))));
    const int64_t at = plane + (int64_t)y * w + x;
    mask[at] = (double)raw[(oy + ry) * pitch + cx + rx] > t ? 1 : 0;
    if (thr_out) thr_out[at] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
threshold_gt_image_kernel(const T* __restrict__ data, const double* __restrict__ thr, const double offset, const int64_t n,
                          uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    mask[i] = (double)data[i] > dsub(thr[i], offset) ? 1 : 0;
}

// ---------------------------------------------------------------- float64 images: skimage's float route
// ref: operations.py:194-195 -> [3p] skimage.filters.thresholding._mean_std on a float image: np.pad(mode='reflect')
// by (k//2 + 1, k//2), integral images of the padded plane and of its square as SEQUENTIAL float64 cumulative sums
// along axis 0 then axis 1 (np.cumsum), window sums as ((I00 - I01) - I10) + I11 (_correlate_sparse's order), then
// m = sum / size, g2 = sumsq / size, s = sqrt(max(g2 - m*m, 0)).  Float addition does not reassociate, so the sums are
// formed in exactly that order: one thread per column for axis 0, one thread per row (through transposing
// shared-memory tiles) for axis 1.
__device__ __forceinline__ int reflect_index(int i, int n) {  // np.pad 'reflect' (no repeated edge); |overshoot| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256)
pad_square_kernel(const double* __restrict__ in, int h, int w, int ph, int pw, int oy, int ox, double* __restrict__ p,
                  double* __restrict__ q) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= pw) return;
  const double v = in[img * (int64_t)h * w + (int64_t)reflect_index(y - oy, h) * w + reflect_index(x - ox, w)];
  const int64_t o = img * (int64_t)ph * pw + (int64_t)y * pw + x;
  p[o] = v;
  q[o] = dmul(v, v);
}

__global__ void __launch_bounds__(128)
cumsum_axis0_kernel(double* __restrict__ a, double* __restrict__ b, int ph, int pw) {
  const int64_t img = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= pw) return;
  double* pa = a + img * (int64_t)ph * pw + x;
  double* pb = b + img * (int64_t)ph * pw + x;
  double sa = pa[0], sb = pb[0];
  for (int y = 1; y < ph; ++y) {
    sa = dadd(sa, pa[(int64_t)y * pw]);
    sb = dadd(sb, pb[(int64_t)y * pw]);
    pa[(int64_t)y * pw] = sa;
    pb[(int64_t)y * pw] = sb;
  }
}

// 32 rows per CTA; the row is walked left to right in 32-column tiles that pass through shared memory so that the
// global accesses are row-contiguous while every thread owns one row's running sum
__global__ void __launch_bounds__(32)
cumsum_axis1_kernel(double* __restrict__ a, int ph, int pw) {
  __shared__ double tile[32][33];
  const int64_t img = blockIdx.y;
  const int y0 = blockIdx.x * 32, lane = threadIdx.x;
  double* base = a + img * (int64_t)ph * pw;
  double run = 0.0;
  bool first = true;
  for (int x0 = 0; x0 < pw; x0 += 32) {
    for (int r = 0; r < 32; ++r)
      if (y0 + r < ph && x0 + lane < pw) tile[r][lane] = base[(int64_t)(y0 + r) * pw + x0 + lane];
    __syncwarp();
    if (y0 + lane < ph) {
      for (int c = 0; c < 32 && x0 + c < pw; ++c) {
        run = first ? tile[lane][c] : dadd(run, tile[lane][c]);
        first = false;
        tile[lane][c] = run;
      }
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r)
      if (y0 + r < ph && x0 + lane < pw) base[(int64_t)(y0 + r) * pw + x0 + lane] = tile[r][lane];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
window_threshold_f64_kernel(const double* __restrict__ in, const double* __restrict__ ip, const double* __restrict__ iq, int h,
                            int w, int pw, int ph, int k0, int k1, int kind, double kk, double r, uint8_t* __restrict__ mask,
                            double* __restrict__ thresholds) {
  const int64_t img = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const double* P = ip + img * (int64_t)ph * pw;
  const double* Q = iq + img * (int64_t)ph * pw;
  const int64_t i00 = (int64_t)y * pw + x, i01 = i00 + k1, i10 = i00 + (int64_t)k0 * pw, i11 = i10 + k1;
  const double total = (double)(k0 * k1);
  const double m = ddiv(dadd(dsub(dsub(P[i00], P[i01]), P[i10]), P[i11]), total);
  const double g2 = ddiv(dadd(dsub(dsub(Q[i00], Q[i01]), Q[i10]), Q[i11]), total);
  double var = dsub(g2, dmul(m, m));
  var = var < 0.0 ? 0.0 : var;  // np.clip(., 0, None); NaN propagates as in NumPy
  const double sd = __dsqrt_rn(var);
  const double t = kind == 0 ? dsub(m, dmul(kk, sd)) : dmul(m, dadd(1.0, dmul(kk, dsub(ddiv(sd, r), 1.0))));
  const int64_t o = img * (int64_t)h * w + (int64_t)y * w + x;
  mask[o] = in[o] > t ? 1 : 0;
  if (thresholds != nullptr) thresholds[o] = t;
}

}  // namespace amt

extern "C" {

size_t amt_window_threshold_f64_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int window_h, int window_w) {
  if (n_img <= 0 || h <= 0 || w <= 0 || window_h < 1 || window_w < 1) return 0;
  return (size_t)2 * n_img * (h + window_h) * (w + window_w) * sizeof(double);
}

int amt_window_threshold_f64(const double* data, int64_t n_img, int64_t h, int64_t w, int window_h, int window_w, int kind,
                             double k, double r, uint8_t* mask, double* thresholds, void* scratch, size_t scratch_bytes,
                             amt_stream_t stream) {
  using namespace amt;
  if (!data || !mask || !scratch || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (window_h < 1 || window_w < 1 || window_h % 2 == 0 || window_w % 2 == 0 || (kind != 0 && kind != 1)) return AMT_ERR_INVALID;
  // np.pad would reflect more than once beyond these; n_img and the padded height ride on grid.z / grid.y
  if (window_h / 2 + 1 >= h || window_w / 2 + 1 >= w || n_img > 65535 || h + window_h > 65535 || (h + window_h) * (w + window_w) >= (1ll << 31))
    return AMT_ERR_UNSUPPORTED;
  if (scratch_bytes < amt_window_threshold_f64_scratch_bytes(n_img, h, w, window_h, window_w)) return AMT_ERR_CAPACITY;
  const int ph = (int)h + window_h, pw = (int)w + window_w;
  double* p = (double*)scratch;
  double* q = p + (size_t)n_img * ph * pw;
  cudaStream_t st = as_stream(stream);
  pad_square_kernel<<<dim3((unsigned)ceil_div(pw, 256), (unsigned)ph, (unsigned)n_img), 256, 0, st>>>(
      data, (int)h, (int)w, ph, pw, window_h / 2 + 1, window_w / 2 + 1, p, q);
  AMT_LAUNCH_CHECK();
  cumsum_axis0_kernel<<<dim3((unsigned)ceil_div(pw, 128), (unsigned)n_img), 128, 0, st>>>(p, q, ph, pw);
  AMT_LAUNCH_CHECK();
  cumsum_axis1_kernel<<<dim3((unsigned)ceil_div(ph, 32), (unsigned)(2 * n_img)), 32, 0, st>>>(p, ph, pw);  // q follows p
  AMT_LAUNCH_CHECK();
  window_threshold_f64_kernel<<<dim3((unsigned)ceil_div(w, 256), (unsigned)h, (unsigned)n_img), 256, 0, st>>>(
      data, p, q, (int)h, (int)w, pw, ph, window_h, window_w, kind, k, r, mask, thresholds);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_window_threshold_u16(const uint16_t* data, int64_t n_img, int64_t h, int64_t w, int window_h, int window_w,
                             int kind, double k, double r, uint8_t* mask, double* thresholds, amt_stream_t stream) {
  using namespace amt;
  if (!data || !mask || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (window_h < 1 || window_w < 1 || window_h % 2 == 0 || window_w % 2 == 0 || (kind != 0 && kind != 1)) return AMT_ERR_INVALID;
  const int ry = window_h / 2, rx = window_w / 2;
  if (ry > 63 || rx > 63 || n_img > 65535 || h >= (1ll << 30) || w >= (1ll << 30)) return AMT_ERR_UNSUPPORTED;
  if ((ry > 0 && ry >= h) || (rx > 0 && rx >= w)) return AMT_ERR_UNSUPPORTED;  // np.pad would reflect more than once
  const int rows = LT_TILE + 2 * ry, cols = LT_TILE + 2 * rx;
  const size_t smem = (size_t)rows * LT_TILE * 12 + (size_t)rows * (cols | 1) * 2;
  if (smem > 48 * 1024)
    AMT_CUDA_TRY(cudaFuncSetAttribute(window_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(w, LT_TILE), (unsigned)ceil_div(h, LT_TILE), (unsigned)n_img);
  if (grid.y > 65535) return AMT_ERR_UNSUPPORTED;
  window_threshold_kernel<<<grid, 256, smem, as_stream(stream)>>>(data, (int)h, (int)w, ry, rx, kind, k, r, mask, thresholds);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_threshold_gt_image(const void* data, int in_dtype, int64_t n, const double* thresholds, double offset,
                           uint8_t* mask, amt_stream_t stream) {
  using namespace amt;
  if (!data || !thresholds || !mask || n <= 0) return AMT_ERR_INVALID;
  const int64_t want = ceil_div(n, 256);
  const unsigned blocks = (unsigned)(want < (int64_t)kNumSMs * 16 ? want : (int64_t)kNumSMs * 16);
  if (in_dtype == AMT_F64)
    threshold_gt_image_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>((const double*)data, thresholds, offset, n, mask);
  else if (in_dtype == AMT_U16)
    threshold_gt_image_kernel<uint16_t><<<blocks, 256, 0, as_stream(stream)>>>((const uint16_t*)data, thresholds, offset, n, mask);
  else
    return AMT_ERR_UNSUPPORTED;
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
