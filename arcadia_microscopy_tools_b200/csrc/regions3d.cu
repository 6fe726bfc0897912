// Per-object quantification of a 3-D label volume (confocal z-stacks, BASELINE config 4).
//
// The reference's SegmentationMask is 2-D only (masks.py:171-172), so there is no reference entry
// point; the semantics are skimage.measure.regionprops_table's on a 3-D label image (SURVEY.md 8a
// item 10 and note N3): area = voxel count, 6-component half-open bbox, 3-component centroid,
// inertia tensor T_ii = (sum_k mu_kk - mu_ii)/mu0, T_ij = -mu_ij/mu0, its eigenvalues (descending,
// clipped at 0), axis_major_length = sqrt(10 (e0 + e1 - e2)), axis_minor_length =
// sqrt(10 max(-e0 + e1 + e2, 0)), and per channel sum / mean / max / min / std of the voxels.
//
// Same design as regions.cu: one streaming pass, each thread owns 8 consecutive voxels of one
// (z, y) row, folds runs of equal label in registers and flushes them with 64-bit integer atomics
// into a SoA table [field][label]; all accumulators are exact integers and the float statistics
// are formed once, from exact 128-bit numerators, in the finalize kernel.  HBM-bound: 4 + 2*C
// bytes per voxel.

#include "internal.cuh"

namespace amt {

enum {
  G_COUNT = 0, G_SZ, G_SY, G_SX, G_SZZ, G_SYY, G_SXX, G_SZY, G_SZX, G_SYX,
  G_ZMIN, G_ZMAX, G_YMIN, G_YMAX, G_XMIN, G_XMAX, G_BASE
};
enum { GC_SUM = 0, GC_SUMSQ, GC_MIN, GC_MAX, GC_PER };
static_assert(G_BASE == AMT_ACC3D_BASE && GC_PER == AMT_ACC_PER_CHANNEL, "3-D accumulator layout");

constexpr int MAX_CH3 = 8;

__global__ void acc3d_init_kernel(uint64_t* __restrict__ acc, int n_fields, int64_t max_labels, int64_t total) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int f = (int)(i / max_labels);
    bool is_min = (f == G_ZMIN || f == G_YMIN || f == G_XMIN);
    if (f >= G_BASE) is_min = ((f - G_BASE) % GC_PER) == GC_MIN;
    acc[i] = is_min ? ~0ull : 0ull;
  }
}

// grid (ceil(w / 1024), h, d); block 128; thread -> 8 voxels of row (z = blockIdx.z, y = blockIdx.y)
template <int C>
__global__ void __launch_bounds__(128)
region_reduce3d_kernel(const int32_t* __restrict__ labels, const uint16_t* __restrict__ channels, const int64_t chan_stride,
                       const int h, const int w, const int64_t max_labels, uint64_t* __restrict__ acc) {
  const int z = blockIdx.z, y = blockIdx.y;
  const int x0 = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (x0 >= w) return;
  const int64_t row = ((int64_t)z * h + y) * w;
  int lab[8];
  int any = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    lab[i] = (x0 + i < w) ? __ldg(labels + row + x0 + i) : 0;
    any |= lab[i];
  }
  if (any == 0) return;
  uint32_t val[C > 0 ? C : 1][8];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const uint16_t* src = channels + c * chan_stride + row + x0;
#pragma unroll
    for (int i = 0; i < 8; ++i) val[c][i] = (x0 + i < w) ? __ldg(src + i) : 0;
  }
  unsigned long long* base = (unsigned long long*)acc;
  int cur = 0, run_x0 = 0;
  uint32_t n = 0, sx = 0;
  uint64_t sxx = 0;
  uint32_t vs[C > 0 ? C : 1], vmin[C > 0 ? C : 1], vmax[C > 0 ? C : 1];
  uint64_t vss[C > 0 ? C : 1];
  auto flush = [&](int x_last) {
    if (cur <= 0 || cur > max_labels) return;
    unsigned long long* a = base + (cur - 1);
    const uint64_t nn = n, zz = (uint64_t)z, yy = (uint64_t)y;
    atomicAdd(a + G_COUNT * max_labels, nn);
    atomicAdd(a + G_SZ * max_labels, nn * zz);
    atomicAdd(a + G_SY * max_labels, nn * yy);
    atomicAdd(a + G_SX * max_labels, (uint64_t)sx);
    atomicAdd(a + G_SZZ * max_labels, nn * zz * zz);
    atomicAdd(a + G_SYY * max_labels, nn * yy * yy);
    atomicAdd(a + G_SXX * max_labels, sxx);
    atomicAdd(a + G_SZY * max_labels, nn * zz * yy);
    atomicAdd(a + G_SZX * max_labels, zz * (uint64_t)sx);
    atomicAdd(a + G_SYX * max_labels, yy * (uint64_t)sx);
    atomicMin(a + G_ZMIN * max_labels, zz);
    atomicMax(a + G_ZMAX * max_labels, zz);
    atomicMin(a + G_YMIN * max_labels, yy);
    atomicMax(a + G_YMAX * max_labels, yy);
    atomicMin(a + G_XMIN * max_labels, (uint64_t)run_x0);
    atomicMax(a + G_XMAX * max_labels, (uint64_t)x_last);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      unsigned long long* ac = a + (G_BASE + c * GC_PER) * max_labels;
      atomicAdd(ac + GC_SUM * max_labels, (uint64_t)vs[c]);
      atomicAdd(ac + GC_SUMSQ * max_labels, vss[c]);
      atomicMin(ac + GC_MIN * max_labels, (uint64_t)vmin[c]);
      atomicMax(ac + GC_MAX * max_labels, (uint64_t)vmax[c]);
    }
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int l = lab[i];
    if (l != cur) {
      flush(x0 + i - 1);
      cur = l;
      n = 0; sx = 0; sxx = 0;
      run_x0 = x0 + i;
#pragma unroll
      for (int c = 0; c < C; ++c) { vs[c] = 0; vss[c] = 0; vmin[c] = 0xffffffffu; vmax[c] = 0; }
    }
    if (l > 0) {
      const uint32_t x = (uint32_t)(x0 + i);
      n += 1; sx += x; sxx += (uint64_t)x * x;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t v = val[c][i];
        vs[c] += v; vss[c] += (uint64_t)v * v;
        vmin[c] = v < vmin[c] ? v : vmin[c];
        vmax[c] = v > vmax[c] ? v : vmax[c];
      }
    }
  }
  flush(x0 + 7 < w ? x0 + 7 : w - 1);
}

__device__ __forceinline__ double exact_cov_num3(uint64_t n, uint64_t s2, uint64_t s1a, uint64_t s1b) {
  const unsigned __int128 p = (unsigned __int128)n * s2;
  const unsigned __int128 q = (unsigned __int128)s1a * s1b;
  const bool neg = q > p;
  const unsigned __int128 d = neg ? q - p : p - q;
  const double v = (double)(uint64_t)(d >> 64) * 18446744073709551616.0 + (double)(uint64_t)d;
  return neg ? -v : v;
}

// eigenvalues of a symmetric 3x3 matrix by cyclic Jacobi rotations (robust for the degenerate
// spectra of spheres and flat objects; converges to double precision in a few sweeps)
__device__ void sym3_eigvals(double a00, double a11, double a22, double a01, double a02, double a12, double (&ev)[3]) {
  double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        const int r = 3 - p - q;
        const double app = A[p][p], aqq = A[q][q], apq = A[p][q], arp = A[r][p], arq = A[r][q];
        A[p][p] = app - t * apq;
        A[q][q] = aqq + t * apq;
        A[p][q] = A[q][p] = 0.0;
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
      }
    }
  }
  double e0 = A[0][0], e1 = A[1][1], e2 = A[2][2], t;
  if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
  if (e1 < e2) { t = e1; e1 = e2; e2 = t; }
  if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
  ev[0] = e0; ev[1] = e1; ev[2] = e2;
}

__global__ void region_finalize3d_kernel(const uint64_t* __restrict__ acc, const int64_t count, const int n_channels,
                                         const int64_t max_labels, double* __restrict__ table) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t K = count > max_labels ? max_labels : count;
  if (k >= K) return;
  const int n_cols = AMT_TABLE3D_COLS(n_channels);
  const uint64_t* a = acc + k;
  double* t = table + k;
  auto A = [&](int f) -> uint64_t { return a[(int64_t)f * max_labels]; };
  auto T = [&](int c) -> double& { return t[(int64_t)c * max_labels]; };
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint64_t n = A(G_COUNT);
  T(0) = (double)(k + 1);
  T(1) = (double)n;
  if (n == 0) {
    for (int c = 2; c < n_cols; ++c) T(c) = nan;
    return;
  }
  const double N = (double)n;
  T(2) = (double)A(G_ZMIN); T(3) = (double)A(G_YMIN); T(4) = (double)A(G_XMIN);
  T(5) = (double)(A(G_ZMAX) + 1); T(6) = (double)(A(G_YMAX) + 1); T(7) = (double)(A(G_XMAX) + 1);
  T(8) = (double)A(G_SZ) / N; T(9) = (double)A(G_SY) / N; T(10) = (double)A(G_SX) / N;
  const double mzz = exact_cov_num3(n, A(G_SZZ), A(G_SZ), A(G_SZ)) / N;
  const double myy = exact_cov_num3(n, A(G_SYY), A(G_SY), A(G_SY)) / N;
  const double mxx = exact_cov_num3(n, A(G_SXX), A(G_SX), A(G_SX)) / N;
  const double mzy = exact_cov_num3(n, A(G_SZY), A(G_SZ), A(G_SY)) / N;
  const double mzx = exact_cov_num3(n, A(G_SZX), A(G_SZ), A(G_SX)) / N;
  const double myx = exact_cov_num3(n, A(G_SYX), A(G_SY), A(G_SX)) / N;
  const double S = mzz + myy + mxx;
  double ev[3];
  sym3_eigvals((S - mzz) / N, (S - myy) / N, (S - mxx) / N, -mzy / N, -mzx / N, -myx / N, ev);
  for (int i = 0; i < 3; ++i) ev[i] = ev[i] > 0.0 ? ev[i] : 0.0;
  T(11) = ev[0]; T(12) = ev[1]; T(13) = ev[2];
  T(14) = sqrt(10.0 * (ev[0] + ev[1] - ev[2]));
  const double m = -ev[0] + ev[1] + ev[2];
  T(15) = sqrt(10.0 * (m > 0.0 ? m : 0.0));
  for (int c = 0; c < n_channels; ++c) {
    const int f = G_BASE + c * GC_PER;
    const int col = AMT_TABLE3D_BASE + c * AMT_TABLE_PER_CHANNEL;
    const uint64_t s = A(f + GC_SUM), ss = A(f + GC_SUMSQ);
    T(col + 0) = (double)s;
    T(col + 1) = (double)s / N;
    T(col + 2) = (double)A(f + GC_MAX);
    T(col + 3) = (double)A(f + GC_MIN);
    const double varnum = exact_cov_num3(n, ss, s, s);
    T(col + 4) = sqrt((varnum > 0.0 ? varnum : 0.0)) / N;
  }
}

template <int C>
static void reduce3d_launch(const int32_t* labels, const uint16_t* channels, int64_t chan_stride, int d, int h, int w,
                            int64_t max_labels, uint64_t* acc, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(w, 8 * 128), (unsigned)h, (unsigned)d);
  region_reduce3d_kernel<C><<<grid, 128, 0, st>>>(labels, channels, chan_stride, h, w, max_labels, acc);
}

}  // namespace amt

extern "C" {

int amt_region_reduce3d(const int32_t* labels, const uint16_t* channels, int n_channels, int64_t chan_stride, int64_t d,
                        int64_t h, int64_t w, int64_t max_labels, uint64_t* acc, amt_stream_t stream) {
  using namespace amt;
  if (!labels || !acc || d <= 0 || h <= 0 || w <= 0 || max_labels <= 0) return AMT_ERR_INVALID;
  if (n_channels < 0 || n_channels > MAX_CH3 || (n_channels > 0 && !channels)) return AMT_ERR_INVALID;
  if (h > 65535 || d > 65535 || w >= (1ll << 30)) return AMT_ERR_CAPACITY;
  // coordinate sums stay below 2^64: n * z^2 summed over the object
  cudaStream_t st = as_stream(stream);
  const int n_fields = AMT_ACC3D_FIELDS(n_channels);
  const int64_t total = (int64_t)n_fields * max_labels;
  int64_t ib = ceil_div(total, 256);
  if (ib > kNumSMs * 8) ib = kNumSMs * 8;
  acc3d_init_kernel<<<(unsigned)ib, 256, 0, st>>>(acc, n_fields, max_labels, total);
  AMT_LAUNCH_CHECK();
  switch (n_channels) {
    case 0: reduce3d_launch<0>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 1: reduce3d_launch<1>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 2: reduce3d_launch<2>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 3: reduce3d_launch<3>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 4: reduce3d_launch<4>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 5: reduce3d_launch<5>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 6: reduce3d_launch<6>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    case 7: reduce3d_launch<7>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
    default: reduce3d_launch<8>(labels, channels, chan_stride, (int)d, (int)h, (int)w, max_labels, acc, st); break;
  }
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_region_finalize3d(const uint64_t* acc, int64_t count, int n_channels, int64_t max_labels, double* table,
                          amt_stream_t stream) {
  using namespace amt;
  if (!acc || !table || count < 0 || max_labels <= 0 || n_channels < 0 || n_channels > MAX_CH3) return AMT_ERR_INVALID;
  if (count == 0) return AMT_OK;
  region_finalize3d_kernel<<<(unsigned)ceil_div(max_labels, 128), 128, 0, as_stream(stream)>>>(acc, count, n_channels,
                                                                                              max_labels, table);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
