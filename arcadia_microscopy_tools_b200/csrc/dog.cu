// Fused 2-D difference of Gaussians: strip-walking ring-buffer kernels for the hot path.
//
// Reference path: operations.py:91 ski.filters.difference_of_gaussians [3p] (see gauss.cu for
// the arithmetic contract: float64, scipy's operation order, no FMA — bit-identical).
//
// These kernels are FP64-pipe-bound (3 DP instructions per tap pair, 2*(1+3*64)+2*(1+3*2) = 400
// per sample for sigma = 0.6 / 16).  On B200 a DP warp instruction holds the SMSP's issue port
// for two cycles, so every non-DP instruction costs DP throughput (ncu: issue-slot model
// 2*DP + other = 97 % of the cycles).  The design goal is therefore "nothing but DADD / DMUL and
// three LDS per tap step", and as little as possible outside that loop:
//  * ONE kernel shape for both axes.  A CTA owns a strip of 32 positions along the contiguous
//    axis (= lanes) and walks the WHOLE filter axis, S = WARPS*R samples per step, keeping the
//    samples it needs in a shared-memory ring of S-row blocks.  Every pass writes its result
//    TRANSPOSED (out[col][row]): pass 1 filters the image along axis 0 and writes G^T, pass 2
//    filters G^T along its axis 0 (= image axis 1) and writes lo - hi back in image layout.
//    Loads are always lane-contiguous, and a thread's R outputs are contiguous in the output;
//    they leave through a small per-warp staging tile so that every store instruction writes
//    whole 128-byte lines (store_transposed).
//  * the block of S rows the NEXT step needs is fetched while this step computes: uint16 input
//    as two/four 8-byte loads per thread held in registers and converted (x * 1/65535) on the
//    way into the ring, float64 input with cp.async straight into the ring.  One barrier per
//    step; the prologue (S + 2r rows) is paid once per strip, not once per 512 rows.
//  * the left / right tap pointers advance by R rows per unrolled iteration and wrap there; R
//    mirror rows behind the ring make every access inside an iteration pointer + constant.
//  * the narrow (sigma_lo, r <= 4) filter needs R + 2*r_lo samples per thread; they are taken
//    from the ring (pass 1) or straight from global memory (pass 2) into registers.
//  * the ring wraps (only R mirror rows), so 2-3 CTAs of 128 threads fit per SM and run out of
//    phase: while one is in a non-DP phase (fill, stores, barrier) the others keep the pipe busy.
//  * any plane whose width is a multiple of 4 and height a multiple of 2 takes this path (narrow
//    last strip, short last step); other shapes and radii use the tile kernels of gauss.cu.
//  * the second pass also writes bucket12() of every output (2 bytes) for the percentile
//    selection that follows (select.cu), which then reads those instead of the 8-byte samples.
// The inner loop is conv_ring below: conv_exact of conv.cuh (used by the tile kernels of gauss.cu)
// on a wrap-around ring.

#include <cstdlib>
#include <cstring>

#include "conv.cuh"

namespace amt {

constexpr int PV_TW = 32;  // lanes along the contiguous axis
constexpr int RLO_MAX = 4;  // largest narrow radius handled in registers

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// acc[o] = x[o]*w[0] + sum_{j = r..1} (x[o-j] + x[o+j]) * w[j], scipy's order, for r <= RLO_MAX;
// xs[i] holds sample i - RLO_MAX relative to output 0 (static indices only: stays in registers).
// FMA = true (amt_tune "dog_fma", off by default): the multiply and the accumulation contract into one
// fused multiply-add, 2 DP instructions per tap pair instead of 3.  Results then differ from scipy's in
// the last bits (relative 1e-15 on the filtered planes, far inside the reference's 1e-5 tolerance) and
// bit-exact labels are no longer guaranteed by construction, only overwhelmingly likely.
template <int R, bool FMA>
__device__ __forceinline__ void conv_small(const double (&xs)[R + 2 * RLO_MAX], const double* __restrict__ hw,
                                           const int r, double (&acc)[R]) {
  const double w0 = hw[0];
#pragma unroll
  for (int o = 0; o < R; ++o) acc[o] = dmul(xs[o + RLO_MAX], w0);
#pragma unroll
  for (int j = RLO_MAX; j >= 1; --j) {
    if (j <= r) {
      const double wj = hw[j];
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const double pair = dadd(xs[o + RLO_MAX - j], xs[o + RLO_MAX + j]);
        acc[o] = FMA ? __fma_rn(pair, wj, acc[o]) : dadd(acc[o], dmul(pair, wj));
      }
    }
  }
}

// A warp's 32 (lanes) x R (outputs per thread) tile leaves TRANSPOSED: out[col][row .. row+R).
// Written straight from the accumulators, every store instruction would touch 32 different
// 128-byte lines (16 bytes of each) and the LSU spends one cycle per line: ncu showed those
// stores holding their source registers (and the scoreboards the next loads share) for ~2000
// cycles per step.  Staged through a per-warp shared-memory tile (pitch R+2 doubles:
// conflict-free 16-byte writes), R/2 lanes cover one column's whole R-output run, so an
// instruction touches 64/R lines instead of 32 and needs no CTA barrier, only __syncwarp.
template <int R>
__device__ __forceinline__ void store_transposed(double* __restrict__ sw, const double (&v)[R], double* __restrict__ dst,
                                                 const int n, const int lane, const int rows_left, const int cols_valid) {
  constexpr int P = R + 2;     // stage pitch (doubles)
  constexpr int LPR = R / 2;   // lanes per column run
  constexpr int CPI = 32 / LPR;  // columns per store instruction
#pragma unroll
  for (int o = 0; o < R; o += 2) *reinterpret_cast<double2*>(sw + lane * P + o) = make_double2(v[o], v[o + 1]);
  __syncwarp();
  const int c = lane / LPR, q = 2 * (lane % LPR);
  if (rows_left >= R && cols_valid == PV_TW) {  // whole tile inside the plane (warp-uniform): no per-store tests
#pragma unroll
    for (int i = 0; i < LPR; ++i) {
      const double2 t = *reinterpret_cast<const double2*>(sw + (i * CPI + c) * P + q);
      *reinterpret_cast<double2*>(dst + (int64_t)(i * CPI + c) * n + q) = t;
    }
  } else if (q < rows_left) {
    // rows_left is even (n and the run start are): a pair of outputs is all inside or all outside the plane
#pragma unroll
    for (int i = 0; i < LPR; ++i) {
      const double2 t = *reinterpret_cast<const double2*>(sw + (i * CPI + c) * P + q);
      if (i * CPI + c < cols_valid) *reinterpret_cast<double2*>(dst + (int64_t)(i * CPI + c) * n + q) = t;
    }
  }
  __syncwarp();
}

// The wide filter over a wrap-around ring of N rows (row pitch 32 doubles, `ring0` already offset
// by the lane): acc[o], o < R, centred at ring row c + o; r % R == 0, c % R == 0, N % R == 0.
// Same arithmetic and register-window scheme as conv_exact (conv.cuh); the left / right sample
// pointers advance by R rows per unrolled iteration and wrap there, so inside an iteration every
// access is pointer + constant.  Rows [N, N+R) of the ring mirror rows [0, R).
template <int R, bool FMA>
__device__ __forceinline__ void conv_ring(const double* __restrict__ ring0, const int N, const int c,
                                          const double* __restrict__ hw, const int r, double (&acc)[R]) {
  double L[R], Rt[R];
  {
    const double* pc = ring0 + c * PV_TW;
    const double w0 = hw[0];
#pragma unroll
    for (int o = 0; o < R; ++o) acc[o] = dmul(pc[o * PV_TW], w0);
  }
  int rl = c - r, rr = c + r - R;
  rl = rl < 0 ? rl + N : rl;
  rr = rr >= N ? rr - N : rr;
  const double* pL = ring0 + rl * PV_TW;  // sample (-j) of output 0
  const double* pR = ring0 + rr * PV_TW;  // sample (+j - R) of output 0
  const double* const ring_end = ring0 + N * PV_TW;
#pragma unroll
  for (int o = 0; o < R; ++o) {
    L[o] = pL[o * PV_TW];
    Rt[o] = pR[(R + o) * PV_TW];
  }
  double wj = hw[r];
  for (int j = r; j >= R; j -= R) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
      double t[R];
#pragma unroll
      for (int o = 0; o < R; ++o) t[o] = dadd(L[(o + u) % R], Rt[(o - u + R) % R]);
      // the two window slots that just became dead take the next step's samples right away
      L[u] = pL[(R + u) * PV_TW];
      Rt[R - 1 - u] = pR[(R - 1 - u) * PV_TW];
      const double wn = hw[j - u - 1];
      if constexpr (FMA) {
#pragma unroll
        for (int o = 0; o < R; ++o) acc[o] = __fma_rn(t[o], wj, acc[o]);
      } else {
#pragma unroll
        for (int o = 0; o < R; ++o) t[o] = dmul(t[o], wj);
#pragma unroll
        for (int o = 0; o < R; ++o) acc[o] = dadd(acc[o], t[o]);
      }
      wj = wn;
    }
    pL += R * PV_TW;
    pL = pL == ring_end ? ring0 : pL;
    pR = pR == ring0 ? ring_end : pR;
    pR -= R * PV_TW;
  }
}

// grid (ceil(inner/32) * planes); block (32, WARPS).  n = length of the filter axis (even), inner = length
// of the contiguous axis (multiple of 4 for uint16 input, of 2 for float64), input plane layout [n][inner],
// output [inner][n].  The last strip of a plane may be narrower than 32 and the last step shorter than S.
// FIRST pass : in = image (InT), out_a = G_hi^T, out_b = G_lo^T
// SECOND pass: in = G_hi^T (double), in_lo = G_lo^T, out_a = G_lo - G_hi in image layout
// Ring: N = (nb+1)*S rows; block b (S rows) holds samples y = b*S - r_hi + [0, S), clamped
// (mode='nearest'), in slot b % (nb+1).  Rows [N, N+R) mirror [0, R) and rows [-4, 0) mirror
// [N-4, N), so that aligned 2R-row runs and the narrow filter's R+8-row runs never wrap.
// FMA: 0 = scipy's exact operation order everywhere, 1 = contracted everywhere, 2 = per plane: plane p keeps the
// exact order iff p % exact_every == exact_offset (the executor's segmentation channel, whose plane decides
// labels), the others are contracted (their only product is a float plane with a 1e-5 tolerance).
template <int R, int WARPS, int MIN_CTAS, typename InT, bool SECOND, int FMA>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
dog_strip_kernel(const InT* __restrict__ in, const double* __restrict__ in_lo, const double scale,
                 double* __restrict__ out_a, double* __restrict__ out_b, const int n, const int inner,
                 const double* __restrict__ hw_lo, const int r_lo, const double* __restrict__ hw_hi, const int r_hi,
                 const int nb, uint64_t* __restrict__ minmax, uint16_t* __restrict__ buckets, const int n_items,
                 const int exact_every, const int exact_offset, const int plane_mul, const int plane_add) {
  constexpr int S = WARPS * R;  // rows per step
  constexpr int NT = WARPS * 32;
  constexpr int FRONT = RLO_MAX, BACK = R;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t s_mm[2 * WARPS];
  const int m = nb + 1;
  const int N = m * S;
  double* ring = smem + FRONT * PV_TW;  // row 0
  double* whi = ring + (size_t)(N + BACK) * PV_TW;
  double* wlo = whi + ((r_hi + 2) & ~1);
  double* stage = wlo + ((r_lo + 2) & ~1) + (size_t)threadIdx.y * (PV_TW * (R + 2));  // this warp's store tile
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * PV_TW + tx;
  for (int i = tid; i <= r_hi; i += NT) whi[i] = hw_hi[i];
  for (int i = tid; i <= r_lo; i += NT) wlo[i] = hw_lo[i];

  // item = plane * (inner/32) + strip; one item per CTA unless the launch is persistent (round robin)
  const int strips = (inner + PV_TW - 1) / PV_TW;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int plane_log = item / strips;
  const int plane_idx = plane_log * plane_mul + plane_add;  // a launch may cover every plane_mul-th plane only
  const bool contract = FMA == 1 || (FMA == 2 && plane_idx % exact_every != exact_offset);  // CTA-uniform
  const int x0 = (item - plane_log * strips) * PV_TW;
  const int64_t plane = (int64_t)plane_idx * n * inner;
  const InT* src = in + plane + x0;
  const int cols_valid = inner - x0 < PV_TW ? inner - x0 : PV_TW;  // a multiple of the load granule
  if (item != (int)blockIdx.x) __syncthreads();  // the previous strip's last step is done with the ring

  // ---- ring fill
  constexpr bool U16 = sizeof(InT) == 2;
  constexpr int ITEMS = U16 ? R / 4 : R / 2;  // per thread and block: 4 uint16 or 2 doubles each
  constexpr int TPR = U16 ? 8 : 16;           // threads per row
  constexpr int NRAW = U16 ? ITEMS : 1;
  const int fill_row = tid / TPR, fill_col = (U16 ? 4 : 2) * (tid % TPR);
  // lanes beyond a narrow last strip re-read its last granule (their ring columns are never used)
  const int load_col = fill_col < cols_valid ? fill_col : cols_valid - (U16 ? 4 : 2);
  // b = logical block, pb = its ring slot (b % m, tracked by the caller: no runtime modulo)
  auto fetch_block = [&](int b, int pb, uint2 (&raw)[NRAW]) {  // global -> registers (uint16) or -> ring (float64, async)
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int rb = it * (NT / TPR) + fill_row;
      int y = b * S + rb - r_hi;
      y = y < 0 ? 0 : (y > n - 1 ? n - 1 : y);
      if constexpr (U16) {
        raw[it] = __ldg(reinterpret_cast<const uint2*>(src + y * inner + load_col));
      } else {
        const int prow = pb * S + rb;
        const double* g = reinterpret_cast<const double*>(src) + y * inner + load_col;
        double* d = ring + prow * PV_TW + fill_col;
        cp_async16(d, g);
        if (prow < BACK) cp_async16(d + N * PV_TW, g);
        if (prow >= N - FRONT) cp_async16(d - N * PV_TW, g);
      }
    }
  };
  auto commit_block = [&](int pb, const uint2 (&raw)[NRAW]) {  // registers -> ring (uint16 only)
    if constexpr (U16) {
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int prow = pb * S + it * (NT / TPR) + fill_row;
        const double2 v01 = make_double2(dmul((double)(raw[it].x & 0xffffu), scale), dmul((double)(raw[it].x >> 16), scale));
        const double2 v23 = make_double2(dmul((double)(raw[it].y & 0xffffu), scale), dmul((double)(raw[it].y >> 16), scale));
        double* d = ring + prow * PV_TW + fill_col;
        *reinterpret_cast<double2*>(d) = v01;
        *reinterpret_cast<double2*>(d + 2) = v23;
        if (prow < BACK) {
          *reinterpret_cast<double2*>(d + N * PV_TW) = v01;
          *reinterpret_cast<double2*>(d + N * PV_TW + 2) = v23;
        }
        if (prow >= N - FRONT) {
          *reinterpret_cast<double2*>(d - N * PV_TW) = v01;
          *reinterpret_cast<double2*>(d - N * PV_TW + 2) = v23;
        }
      }
    }
  };

  for (int b0 = 0; b0 < nb; b0 += 3) {  // prologue: the window of step 0, three blocks in flight
    uint2 pro[3][NRAW];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (b0 + i < nb) fetch_block(b0 + i, b0 + i, pro[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (b0 + i < nb) commit_block(b0 + i, pro[i]);
  }
  if constexpr (!U16) cp_async_wait_all();
  __syncthreads();

  uint64_t kmin = ~0ull, kmax = 0ull;
  const int n_steps = (n + S - 1) / S;
  const int64_t out_col = plane + (int64_t)x0 * n;  // transposed output: [col][row]
  const double* ring_lane = ring + tx;
  int pk = 0;  // k % m
  for (int k = 0; k < n_steps; ++k) {
    const bool more = k + 1 < n_steps;
    uint2 raw[NRAW];
    const int pf = pk == 0 ? m - 1 : pk - 1;  // (k + nb) % m: the slot of block k-1, free since the last barrier
    if (more) fetch_block(k + nb, pf, raw);
    const int yb = k * S + ty * R;
    const int rows_left = n - yb;  // even; <= 0: nothing of this thread's run is inside the plane
    const bool lane_ok = tx < cols_valid;

    double xs[R + 2 * RLO_MAX];
    if constexpr (SECOND) {  // narrow operand straight from global memory (L2), in flight during the hi filter
      const double* lo_col = in_lo + plane + x0 + (lane_ok ? tx : cols_valid - 1);
      if (yb >= RLO_MAX && yb + R + RLO_MAX <= n) {  // interior: one base, constant strides
        const double* p = lo_col + (yb - RLO_MAX) * inner;
#pragma unroll
        for (int i = 0; i < R + 2 * RLO_MAX; ++i) xs[i] = __ldg(p + i * inner);
      } else {
#pragma unroll
        for (int i = 0; i < R + 2 * RLO_MAX; ++i) {
          int y = yb + i - RLO_MAX;
          y = y < 0 ? 0 : (y > n - 1 ? n - 1 : y);
          xs[i] = __ldg(lo_col + y * inner);
        }
      }
    }

    int c = pk * S + r_hi + ty * R;  // ring row of this thread's output 0 (sample y sits at ring row y + r_hi)
    c = c >= N ? c - N : c;
    double acc[R];
    if (FMA != 0 && contract)
      conv_ring<R, true>(ring_lane, N, c, whi, r_hi, acc);
    else
      conv_ring<R, false>(ring_lane, N, c, whi, r_hi, acc);
    if constexpr (!SECOND) {
      store_transposed<R>(stage, acc, out_a + out_col + yb, n, tx, rows_left, cols_valid);
      const double* col = ring_lane + c * PV_TW;
#pragma unroll
      for (int i = 0; i < R + 2 * RLO_MAX; ++i) xs[i] = col[(i - RLO_MAX) * PV_TW];  // front / back mirrors: never wraps
      if (FMA != 0 && contract)
        conv_small<R, true>(xs, wlo, r_lo, acc);
      else
        conv_small<R, false>(xs, wlo, r_lo, acc);
      store_transposed<R>(stage, acc, out_b + out_col + yb, n, tx, rows_left, cols_valid);
    } else {
      double acc_lo[R];
      if (FMA != 0 && contract)
        conv_small<R, true>(xs, wlo, r_lo, acc_lo);
      else
        conv_small<R, false>(xs, wlo, r_lo, acc_lo);
#pragma unroll
      for (int o = 0; o < R; ++o) acc[o] = dsub(acc_lo[o], acc[o]);
      store_transposed<R>(stage, acc, out_a + out_col + yb, n, tx, rows_left, cols_valid);
      if (lane_ok && rows_left > 0) {
        if (buckets != nullptr) {  // bucket12 of every output, same layout as out_a: R consecutive uint16 per lane
          uint32_t packed[R / 2];
#pragma unroll
          for (int o = 0; o < R; o += 2) packed[o / 2] = bucket12(acc[o]) | (bucket12(acc[o + 1]) << 16);
          uint16_t* kd = buckets + out_col + (int64_t)tx * n + yb;
          if (rows_left >= R && (n & 7) == 0) {
#pragma unroll
            for (int o = 0; o < R / 2; o += 4)
              *reinterpret_cast<uint4*>(kd + 2 * o) = make_uint4(packed[o], packed[o + 1], packed[o + 2], packed[o + 3]);
          } else {
#pragma unroll
            for (int o = 0; o < R / 2; ++o)
              if (2 * o < rows_left) *reinterpret_cast<uint32_t*>(kd + 2 * o) = packed[o];
          }
        }
        if (minmax != nullptr) {
          if (rows_left < R) {  // short last step: outputs beyond the plane repeat a valid one
#pragma unroll
            for (int o = 1; o < R; ++o) acc[o] = o < rows_left ? acc[o] : acc[0];
          }
#pragma unroll
          for (int o = 0; o < R; ++o) {
            const uint64_t key = f64_to_key(acc[o]);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
          }
        }
      }
    }
    if (more) {
      commit_block(pf, raw);
      if constexpr (!U16) cp_async_wait_all();
      __syncthreads();
    }
    pk = (pk + 1 == m) ? 0 : pk + 1;
  }
  if (SECOND && minmax != nullptr) {
    kmin = warp_min_u64(kmin);
    kmax = warp_max_u64(kmax);
    if (tx == 0) {
      s_mm[ty] = kmin;
      s_mm[WARPS + ty] = kmax;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 1; i < WARPS; ++i) {
        kmin = s_mm[i] < kmin ? s_mm[i] : kmin;
        kmax = s_mm[WARPS + i] > kmax ? s_mm[WARPS + i] : kmax;
      }
      atomicMin((unsigned long long*)&minmax[2 * plane_idx], (unsigned long long)kmin);
      atomicMax((unsigned long long*)&minmax[2 * plane_idx + 1], (unsigned long long)kmax);
    }
  }
  }  // item loop
}

int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st);  // gauss.cu

// ---- tuning knobs (amt_tune).  dog_variant: 0 = 4 warps x 8 outputs per thread (3 CTAs / SM fit),
// 1 = 4 warps x 16 (2 CTAs), 2 = 8 warps x 8 (2 CTAs).  dog_ctas: resident CTAs per SM (0 = as
// many as fit; otherwise the shared-memory request is padded so that no more than that many fit,
// which leaves the rest of the SM to the HBM-bound kernels of the other stream).
static int g_dog_variant = 1;
static int g_dog_ctas = 0;
static int g_dog_persistent = 0;
static int g_dog_generic = 0;
static int g_dog_fma = 0;
static int g_dog_exact_every = 0;   // amt_tune: the stand-alone amt_dog2d* entry points run the per-plane kernel
static int g_dog_exact_offset = 0;  // (what the executor launches) when every > 0; for bench / profiling only
extern int g_stream_ctas;     // gauss.cu
extern int g_stream_pad_kb;   // gauss.cu
extern int g_exec_swap_prio;  // executor.cu
extern int g_pass_ctas;       // core.cu
extern int g_exec_buckets;    // executor.cu
extern int g_exec_tc;         // executor.cu
extern int g_exec_fused_lo;   // executor.cu
extern int g_exec_given_stream;
extern int g_ccl_touch_filter;  // ccl.cu
extern int g_dx_collect_threads;  // decide.cu
extern int g_exec_host_narrow, g_exec_host_threads, g_exec_host_rle, g_exec_rle_share;
extern int g_exec_copy_only;  // executor.cu
static int g_dog_only_exact = 0;  // amt_tune: the stand-alone axis0 / axis1 entry points cover the exact planes only
namespace tc { extern int g_tcg_debug; int set_suspend_ns(int ns); }  // tcgauss.cu

constexpr size_t kSmemMax = 227 * 1024;
constexpr size_t kSmemPerSM = 228 * 1024;

struct DogPlan {
  bool fast;
  int R, warps, nb, ctas;
  size_t smem;
};

static bool aligned16(const void* q) { return ((uintptr_t)q) % 16 == 0; }

static DogPlan dog_plan(int in_dtype, int64_t n_img, int64_t h, int64_t w, int r_lo, int r_hi) {
  DogPlan p{};
  int variant = g_dog_variant;
  if (variant == 1 && r_hi % 16 != 0) variant = 0;  // 16 outputs per thread need 16 | radius; 8 | radius still runs fast
  p.R = variant == 1 ? 16 : 8;
  p.warps = variant == 2 ? 8 : 4;
  const int S = p.R * p.warps;
  p.nb = 1 + (2 * r_hi + S - 1) / S;
  const size_t rows = (size_t)(p.nb + 1) * S + RLO_MAX + p.R;
  p.smem = (rows * PV_TW + ((r_hi + 2) & ~1) + ((r_lo + 2) & ~1) + (size_t)p.warps * PV_TW * (p.R + 2)) * sizeof(double);
  // resident CTAs per SM: what fits (shared memory, and the register budget __launch_bounds__ was given),
  // capped by the dog_ctas knob
  const int max_ctas = variant == 0 ? 3 : 2;
  int fit = (int)(kSmemPerSM / (p.smem + 1024));
  fit = fit > max_ctas ? max_ctas : fit;
  p.ctas = (g_dog_ctas > 0 && g_dog_ctas < fit) ? g_dog_ctas : fit;
  if (p.ctas < fit) {  // pad the request so that no more than p.ctas CTAs fit (each CTA also reserves 1 KB)
    const size_t pad = kSmemPerSM / (p.ctas + 1) + 1 - 1024;
    if (p.smem < pad && pad <= kSmemMax) p.smem = pad;
  }
  // widths: 8-byte loads of 4 uint16 (16-byte cp.async of 2 doubles) along the contiguous axis of either pass,
  // 16-byte stores of output pairs along the other
  const bool dims_ok = (in_dtype == AMT_U16) ? (w % 4 == 0 && h % 2 == 0) : (w % 2 == 0 && h % 2 == 0);
  p.fast = !g_dog_generic && p.smem <= kSmemMax && fit >= 1 && n_img * (((h > w ? h : w) + 31) / 32) < (1ll << 31) &&
           dims_ok && h >= 2 && w >= 4 && h * w < (1ll << 31) && r_lo <= RLO_MAX && r_hi >= r_lo && r_hi >= p.R &&
           r_hi % p.R == 0 && (in_dtype == AMT_U16 || in_dtype == AMT_F64);
  return p;
}

struct DogMix {
  int every, offset;  // every > 0: plane p is exact iff p % every == offset, the other planes are contracted
  bool only;          // every > 0 and only: the launch covers the exact planes alone (the others belong to tcgauss.cu)
};

template <int R, int WARPS, int MIN_CTAS, typename InT, bool SECOND, int FMA>
static int launch_strip(const DogPlan& p, const InT* in, const double* in_lo, double scale, double* out_a, double* out_b,
                        int64_t planes, int64_t n, int64_t inner, const double* hw_lo, int r_lo, const double* hw_hi,
                        int r_hi, uint64_t* minmax, uint16_t* buckets, cudaStream_t st, DogMix mix) {
  auto kernel = dog_strip_kernel<R, WARPS, MIN_CTAS, InT, SECOND, FMA>;
  AMT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  const bool subset = mix.only && mix.every > 0;
  if (subset) planes /= mix.every;
  const int64_t items = planes * ((inner + PV_TW - 1) / PV_TW);
  // one CTA per strip by default (the hardware scheduler balances the tail better than a static
  // round robin); dog_persistent = 1 launches exactly the resident CTAs and lets them loop
  const int64_t resident = g_dog_persistent ? (int64_t)kNumSMs * p.ctas : items;
  dim3 grid((unsigned)(items < resident ? items : resident)), block(PV_TW, WARPS);
  kernel<<<grid, block, p.smem, st>>>(in, in_lo, scale, out_a, out_b, (int)n, (int)inner, hw_lo, r_lo, hw_hi, r_hi, p.nb,
                                      minmax, buckets, (int)items, mix.every > 0 ? mix.every : 1, mix.offset,
                                      subset ? mix.every : 1, subset ? mix.offset : 0);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

template <typename InT, bool SECOND, int FMA>
static int launch_variant_fma(const DogPlan& p, const InT* in, const double* in_lo, double scale, double* out_a,
                              double* out_b, int64_t planes, int64_t n, int64_t inner, const double* hw_lo, int r_lo,
                              const double* hw_hi, int r_hi, uint64_t* minmax, uint16_t* buckets, cudaStream_t st,
                              DogMix mix) {
  if (p.R == 16)
    return launch_strip<16, 4, 2, InT, SECOND, FMA>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo,
                                                     hw_hi, r_hi, minmax, buckets, st, mix);
  if (p.warps == 8)
    return launch_strip<8, 8, 2, InT, SECOND, FMA>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo,
                                                    hw_hi, r_hi, minmax, buckets, st, mix);
  return launch_strip<8, 4, 3, InT, SECOND, FMA>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo, hw_hi,
                                                  r_hi, minmax, buckets, st, mix);
}

// amt_tune "dog_fma" contracts every plane; otherwise mix.every > 0 selects the per-plane kernel
template <typename InT, bool SECOND>
static int launch_variant(const DogPlan& p, const InT* in, const double* in_lo, double scale, double* out_a,
                          double* out_b, int64_t planes, int64_t n, int64_t inner, const double* hw_lo, int r_lo,
                          const double* hw_hi, int r_hi, uint64_t* minmax, uint16_t* buckets, cudaStream_t st,
                          DogMix mix) {
  if (g_dog_fma)
    return launch_variant_fma<InT, SECOND, 1>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo, hw_hi,
                                              r_hi, minmax, buckets, st, mix);
  if (mix.every > 0 && !mix.only)
    return launch_variant_fma<InT, SECOND, 2>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo, hw_hi,
                                              r_hi, minmax, buckets, st, mix);
  return launch_variant_fma<InT, SECOND, 0>(p, in, in_lo, scale, out_a, out_b, planes, n, inner, hw_lo, r_lo, hw_hi,
                                            r_hi, minmax, buckets, st, mix);
}

// pass 1: image (h x w) -> tmp_hi, tmp_lo.  Fast path: both TRANSPOSED (w x h); generic: image layout.
static int dog_axis0(const DogPlan& p, const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h,
                     int64_t w, const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo,
                     double* tmp_hi, cudaStream_t st, DogMix mix) {
  if (!p.fast) return dog_axis0_generic(in, in_dtype, in_scale, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, tmp_lo, tmp_hi, st);
  if (in_dtype == AMT_U16)
    return launch_variant<uint16_t, false>(p, (const uint16_t*)in, nullptr, in_scale, tmp_hi, tmp_lo, n_img, h, w,
                                           hw_lo, r_lo, hw_hi, r_hi, nullptr, nullptr, st, mix);
  return launch_variant<double, false>(p, (const double*)in, nullptr, 1.0, tmp_hi, tmp_lo, n_img, h, w, hw_lo, r_lo,
                                       hw_hi, r_hi, nullptr, nullptr, st, mix);
}

// pass 2.  Fast path: the transposed planes (w x h) filtered along their axis 0, lo - hi written back
// transposed (= image layout); generic: image-layout planes through the tile kernel of gauss.cu.
static int dog_axis1(const DogPlan& p, const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img,
                     int64_t h, int64_t w, const double* hw_lo, int r_lo, const double* hw_hi, int r_hi,
                     uint64_t* minmax, uint16_t* buckets, cudaStream_t st, DogMix mix) {
  if (!p.fast) return dog_axis1_generic(tmp_lo, tmp_hi, out, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, minmax, st);
  return launch_variant<double, true>(p, tmp_hi, tmp_lo, 1.0, out, nullptr, n_img, w, h, hw_lo, r_lo, hw_hi, r_hi,
                                      minmax, buckets, st, mix);
}

// true iff dog2d would take the strip kernels for this problem (pointer alignment aside)
bool dog2d_fast(int in_dtype, int64_t n_img, int64_t h, int64_t w, int r_lo, int r_hi) {
  return dog_plan(in_dtype, n_img, h, w, r_lo, r_hi).fast;
}

// buckets (optional): n_img*h*w uint16 receiving bucket12() of every output sample; *buckets_written
// tells the caller whether the strip kernels ran (the generic fallback does not produce them).
int dog2d(const void* in, int in_dtype, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
          const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
          uint64_t* minmax, cudaStream_t st, uint16_t* buckets, bool* buckets_written, int exact_every,
          int exact_offset, bool only_exact) {
  if (buckets_written) *buckets_written = false;
  if (exact_every < 0 || (exact_every > 0 && (exact_offset < 0 || exact_offset >= exact_every))) return AMT_ERR_INVALID;
  if (only_exact && (exact_every <= 0 || n_img % exact_every != 0)) return AMT_ERR_INVALID;
  const DogMix mix{exact_every, exact_offset, only_exact};
  if (!in || !out || !tmp_lo || !tmp_hi || !hw_lo || !hw_hi) return AMT_ERR_INVALID;
  if (n_img <= 0 || h <= 0 || w <= 0 || r_lo < 0 || r_hi < 0) return AMT_ERR_INVALID;
  if (in_dtype != AMT_U16 && in_dtype != AMT_F64) return AMT_ERR_UNSUPPORTED;
  if (minmax && !only_exact) AMT_TRY(minmax_init(minmax, n_img, st));  // a subset launch leaves the others' keys alone
  DogPlan p = dog_plan(in_dtype, n_img, h, w, r_lo, r_hi);
  // the strip kernels use 16-byte accesses; odd pointers take the generic tile kernels for BOTH passes
  p.fast = p.fast && aligned16(in) && aligned16(out) && aligned16(tmp_lo) && aligned16(tmp_hi);
  if (only_exact && !p.fast) return AMT_ERR_UNSUPPORTED;  // the generic tile kernels have no plane subset
  if (buckets && !(p.fast && aligned16(buckets))) buckets = nullptr;
  if (buckets_written) *buckets_written = buckets != nullptr;
  AMT_TRY(dog_axis0(p, in, in_dtype, in_scale, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, tmp_lo, tmp_hi, st, mix));
  return dog_axis1(p, tmp_lo, tmp_hi, out, n_img, h, w, hw_lo, r_lo, hw_hi, r_hi, minmax, buckets, st, mix);
}

}  // namespace amt

extern "C" {

int amt_tune(const char* key, int value) {
  using namespace amt;
  if (!key) return AMT_ERR_INVALID;
  const auto is = [&](const char* k) { return std::strcmp(key, k) == 0; };
  if (is("dog_variant")) {
    if (value < 0 || value > 2) return AMT_ERR_INVALID;
    g_dog_variant = value;
  } else if (is("dog_persistent")) {
    g_dog_persistent = value != 0;
  } else if (is("dog_ctas")) {
    if (value < 0 || value > 8) return AMT_ERR_INVALID;
    g_dog_ctas = value;
  } else if (is("stream_pad_kb")) {
    if (value < 0 || value > 200) return AMT_ERR_INVALID;
    g_stream_pad_kb = value;
  } else if (is("stream_ctas")) {
    if (value < 1 || value > 32) return AMT_ERR_INVALID;
    g_stream_ctas = value;
  } else if (is("exec_buckets")) {
    g_exec_buckets = value != 0;
  } else if (is("tcg_suspend_ns")) {
    return tc::set_suspend_ns(value);
  } else if (is("tcg_debug")) {
    tc::g_tcg_debug = value;
  } else if (is("dx_collect_threads")) {
    if (value != 256 && value != 1024) return AMT_ERR_INVALID;
    g_dx_collect_threads = value;
  } else if (is("ccl_touch_filter")) {
    g_ccl_touch_filter = value != 0;
  } else if (is("exec_copy_only")) {
    g_exec_copy_only = value != 0;
  } else if (is("dog_only_exact")) {
    g_dog_only_exact = value != 0;
  } else if (is("exec_tc")) {
    g_exec_tc = value != 0;
  } else if (is("exec_fused_lo")) {
    g_exec_fused_lo = value != 0;
  } else if (is("exec_given_stream")) {
    g_exec_given_stream = value != 0;
  } else if (is("exec_host_narrow")) {
    g_exec_host_narrow = value != 0;
  } else if (is("exec_host_rle")) {
    g_exec_host_rle = value != 0;
  } else if (is("exec_rle_share")) {
    if (value < -1 || value > 100) return AMT_ERR_INVALID;
    g_exec_rle_share = value;
  } else if (is("exec_host_threads")) {
    if (value < 1 || value > 256) return AMT_ERR_INVALID;
    g_exec_host_threads = value;
  } else if (is("pass_ctas")) {
    if (value < 1 || value > 16) return AMT_ERR_INVALID;
    g_pass_ctas = value;
  } else if (is("exec_swap_prio")) {
    g_exec_swap_prio = value != 0;
  } else if (is("dog_generic")) {
    g_dog_generic = value != 0;
  } else if (is("dog_fma")) {
    g_dog_fma = value != 0;
  } else if (is("dog_exact_every")) {
    if (value < 0 || value > 64) return AMT_ERR_INVALID;
    g_dog_exact_every = value;
  } else if (is("dog_exact_offset")) {
    if (value < 0 || value > 63) return AMT_ERR_INVALID;
    g_dog_exact_offset = value;
  } else {
    return AMT_ERR_INVALID;
  }
  return AMT_OK;
}

int amt_dog2d(const void* in, int in_dtype, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
              const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, double* tmp_lo, double* tmp_hi,
              uint64_t* minmax_keys, amt_stream_t stream) {
  return amt::dog2d(in, in_dtype, in_scale, out, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, tmp_lo, tmp_hi,
                    minmax_keys, amt::as_stream(stream), nullptr, nullptr, amt::g_dog_exact_every,
                    amt::g_dog_exact_every > 0 ? amt::g_dog_exact_offset % amt::g_dog_exact_every : 0, false);
}

// The two passes separately (bench / profiling).  The layout of tmp_lo / tmp_hi between them is
// private to the library: call both with the same arguments.
int amt_dog2d_axis0(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, double* tmp_lo, double* tmp_hi,
                    amt_stream_t stream) {
  using namespace amt;
  if (!in || !tmp_lo || !tmp_hi || !half_w_lo || !half_w_hi || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (in_dtype != AMT_U16 && in_dtype != AMT_F64) return AMT_ERR_UNSUPPORTED;
  const DogPlan p = dog_plan(in_dtype, n_img, h, w, r_lo, r_hi);
  if (p.fast && !(aligned16(in) && aligned16(tmp_lo) && aligned16(tmp_hi))) return AMT_ERR_INVALID;
  return dog_axis0(p, in, in_dtype, in_scale, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, tmp_lo, tmp_hi,
                   as_stream(stream), DogMix{g_dog_exact_every, g_dog_exact_every > 0 ? g_dog_exact_offset % g_dog_exact_every : 0,
                          g_dog_only_exact && g_dog_exact_every > 0 && n_img % g_dog_exact_every == 0});
}

int amt_dog2d_axis1(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi, uint64_t* minmax_keys,
                    amt_stream_t stream) {
  using namespace amt;
  if (!tmp_lo || !tmp_hi || !out || !half_w_lo || !half_w_hi || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  const DogPlan p = dog_plan(AMT_U16, n_img, h, w, r_lo, r_hi);  // the dtype of pass 1's input does not matter here
  if (p.fast && !(aligned16(out) && aligned16(tmp_lo) && aligned16(tmp_hi))) return AMT_ERR_INVALID;
  if (minmax_keys) AMT_TRY(minmax_init(minmax_keys, n_img, as_stream(stream)));
  return dog_axis1(p, tmp_lo, tmp_hi, out, n_img, h, w, half_w_lo, r_lo, half_w_hi, r_hi, minmax_keys, nullptr,
                   as_stream(stream), DogMix{g_dog_exact_every, g_dog_exact_every > 0 ? g_dog_exact_offset % g_dog_exact_every : 0,
                          g_dog_only_exact && g_dog_exact_every > 0 && n_img % g_dog_exact_every == 0});
}

}  // extern "C"
