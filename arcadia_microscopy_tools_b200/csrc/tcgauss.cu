// The wide Gaussian of the difference of Gaussians on the 5th-generation tensor cores.
//
// Reference path: operations.py:91 ski.filters.difference_of_gaussians -> [3p]
// scipy.ndimage.gaussian_filter(sigma = 16, mode = 'nearest', truncate = 4): radius 64, 129 taps per
// axis.  In float64 on the CUDA cores that is 400 DP instructions per sample (dog.cu) and the whole
// pipeline is FP64-issue-bound.  Here a 1-D pass is a banded Toeplitz product
//
//     out[m] = sum_k Band[m][k] * in[k],   Band[m][k] = W[k - m - 64],  |k - m - 64| <= r,
//
// evaluated EXACTLY in integers on tcgen05.mma kind::i8:
//   * the weights are integers W[t] = round(w[t] * 2^S) (S = 37 for sigma = 16: 32 significant bits,
//     sum_t W[t] == 2^S exactly) cut into four unsigned base-256 digits -> four uint8 band tiles
//     [128 outputs x 256 inputs], resident in shared memory for the life of a persistent CTA;
//   * pass 1 (image axis 0): the raw uint16 image is read BY TMA AS BYTES.  A tile of 256 rows x 64
//     bytes (32 pixels, low and high byte interleaved) is the MN-major B operand as it lies in memory;
//     the low / high byte columns come out as neighbouring accumulator columns and are recombined in
//     the epilogue: 4 digit MMAs x 8 K-steps (M = 128, N = 64, K = 32) per tile, int32 accumulators
//     in TMEM (every partial sum < 2^23: exact).  The epilogue rebuilds the 53-bit integer, adds the
//     clamped-edge ('nearest') taps, rounds to 40 bits and stores five uint8 digit planes;
//   * pass 2 (image axis 1): the five digit planes are the K-major B operand, again straight from TMA;
//     weight digit d x sample digit s accumulates into the TMEM accumulator of d + s (products of equal
//     significance share one accumulator; d + s < 2 is below 2^-45 of full scale and skipped): 17 digit
//     products x 8 K-steps (M = 128, N = 32, K = 32).  The epilogue shifts the six accumulators together
//     in 64-bit integers, converts ONCE to float64, subtracts from the narrow Gaussian (lo2d_kernel
//     below, float64, scipy's order) and writes the DoG plane, its selection buckets and min / max.
// Errors: the only approximation is the 32-bit rounding of the weights (|dw| <= 2^-37 per tap, zero
// in sum) and the 40-bit rounding between the passes: |dG| <= 129 * 2^-37 * 2 * (local contrast) in the
// worst case, ~1e-11 of the [0, 1] scale in practice (tests/test_gpu_tcgauss.py measures it); the
// reference's tolerance for filtered planes is 1e-5.  The segmentation channel, whose plane decides
// labels, never takes this path (executor.cu): it keeps dog.cu's bit-exact kernels.
//
// Kernel anatomy (both passes): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane), warps 2..9 = epilogue (TMEM lane quarter = warp % 4).  mbarrier rings: full/empty per
// operand stage, acc_full/acc_empty per TMEM buffer (two buffers: the epilogue of tile i overlaps the
// MMAs of tile i + 1).  Persistent grid: one CTA per SM, a contiguous range of tiles each.

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "internal.cuh"

namespace amt {
namespace tc {

constexpr int KBAND = 256;       // inputs per output tile along the filter axis (128 + 2 * 64)
constexpr int MT = 128;          // outputs per tile along the filter axis (UMMA M)
constexpr int HALO = 64;         // largest radius
constexpr int WD = 4;            // weight digits (base 256)
constexpr int GD = 5;            // digits of the pass-1 result (40 bits)
constexpr int JMIN = 2;          // digit products with d + s < JMIN are dropped in pass 2
constexpr int NACC2 = WD + GD - 1 - JMIN;  // accumulators of pass 2 (j = JMIN .. WD+GD-2)
constexpr int P1_NB = 64;        // pass 1: bytes (UMMA N) per tile along the contiguous axis = 32 pixels
constexpr int P1_STAGES = 4;
constexpr int P2_NR = 32;        // pass 2: rows (UMMA N) per tile
constexpr int P2_STAGES = 2;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = (2 + EPI_WARPS) * 32;
constexpr uint32_t BAND_PANEL_BYTES = MT * 128;                  // one K panel (128 bytes of K) of one digit
constexpr uint32_t BAND_BYTES = WD * 2 * BAND_PANEL_BYTES;       // 128 KB
constexpr uint32_t P1_STAGE_BYTES = KBAND * P1_NB;               // 16 KB
constexpr uint32_t P2_PANEL_BYTES = P2_NR * 128;                 // 4 KB
constexpr uint32_t P2_STAGE_BYTES = GD * 2 * P2_PANEL_BYTES;     // 40 KB

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Every wait carries a watchdog: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], uint8 x uint8 -> int32, issued by one thread for the CTA
__device__ __forceinline__ void mma_u8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format: version 1 in bits 46-47; offsets in 16-byte units)
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// Instruction descriptor, kind::i8: D = int32 (bits 4-5 = 2), A and B unsigned 8-bit (format 0), A K-major,
// B K-major or MN-major (bit 16), N >> 3 in bits 17-22, M >> 4 in bits 24-28
__host__ __device__ constexpr uint32_t idesc_u8(int m, int n, bool b_mn_major) {
  return (2u << 4) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct alignas(8) Barriers {
  uint64_t band_full;
  uint64_t full[4], empty[4];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

struct Pass1Params {
  const uint16_t* in;      // [planes][h][w]
  uint8_t* digits;         // [planes][GD][h][w]
  const uint64_t* suffix;  // suffix[j] = sum_{t >= j} W[t], j = 0 .. r + 1
  int h, w, r, shift;      // G1q = (sum + 2^(shift-1)) >> shift
  int n_sel;               // logical planes
  int tiles_y, tiles_x;    // per plane
  PlaneSel sel;
};

// ------------------------------------------------------------------ pass 1: uint16 image -> 40-bit digits, axis 0
__global__ void __launch_bounds__(NTHREADS, 1)
tcg_axis0_kernel(const __grid_constant__ CUtensorMap band_map, const __grid_constant__ CUtensorMap in_map,
                 const Pass1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* band_s = smem;
  uint8_t* stage_s = smem + BAND_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(stage_s + P1_STAGES * P1_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t tiles_total = (int64_t)p.n_sel * p.tiles_y * p.tiles_x;
  const int64_t per_cta = (tiles_total + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t t_end = t_begin + per_cta < tiles_total ? t_begin + per_cta : tiles_total;

  if (threadIdx.x == 0) {
    mbar_init(&bars->band_full, 1);
    for (int s = 0; s < P1_STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // tile t -> (logical plane, x tile, y tile), y fastest: consecutive tiles of a CTA share half their rows (L2)
  auto decode = [&](int64_t t, int& q, int& tx, int& ty) {
    ty = (int)(t % p.tiles_y);
    const int64_t u = t / p.tiles_y;
    tx = (int)(u % p.tiles_x);
    q = (int)(u / p.tiles_x);
  };

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&band_map);
      prefetch_tmap(&in_map);
      mbar_expect_tx(&bars->band_full, BAND_BYTES);
      for (int d = 0; d < WD; ++d)
        for (int pn = 0; pn < 2; ++pn)
          tma_load_2d(band_s + (d * 2 + pn) * BAND_PANEL_BYTES, &band_map, &bars->band_full, pn * 128, d * MT);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        int q, tx, ty;
        decode(t, q, tx, ty);
        mbar_wait(&bars->empty[stage], phase ^ 1);
        mbar_expect_tx(&bars->full[stage], P1_STAGE_BYTES);
        tma_load_3d(stage_s + stage * P1_STAGE_BYTES, &in_map, &bars->full[stage], tx * P1_NB, ty * MT - HALO,
                    p.sel.phys(q));
        if (++stage == P1_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_u8(MT, P1_NB, true);
      mbar_wait(&bars->band_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t t = t_begin; t < t_end; ++t, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t accphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&bars->acc_empty[buf], accphase ^ 1);
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t b_base = smem_u32(stage_s + stage * P1_STAGE_BYTES);
#pragma unroll
        for (int d = 0; d < WD; ++d) {
          const uint32_t d_tmem = tmem + buf * 256 + d * P1_NB;
#pragma unroll
          for (int ks = 0; ks < KBAND / 32; ++ks) {
            const uint64_t a_desc =
                smem_desc(smem_u32(band_s) + (d * 2 + ks / 4) * BAND_PANEL_BYTES + (ks % 4) * 32, 16, 1024, LAYOUT_SW128);
            // MN-major, 64-byte swizzle: 8 K rows of 64 bytes per atom (512 bytes), 32 K rows per MMA
            const uint64_t b_desc = smem_desc(b_base + ks * 32 * P1_NB, P1_STAGE_BYTES, 512, LAYOUT_SW64);
            mma_u8(d_tmem, a_desc, b_desc, idesc, ks > 0 ? 1u : 0u);
          }
        }
        mma_commit(&bars->empty[stage]);
        mma_commit(&bars->acc_full[buf]);
        if (++stage == P1_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lanes this warp may read
    const int hcol = ew >> 2;            // which half of the tile's 64 byte columns
    const int m = quarter * 32 + lane;   // output row inside the tile
    const int64_t hw = (int64_t)p.h * p.w;
    int64_t it = 0;
    for (int64_t t = t_begin; t < t_end; ++t, ++it) {
      int q, tx, ty;
      decode(t, q, tx, ty);
      const int buf = (int)(it & 1);
      const uint32_t accphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&bars->acc_full[buf], accphase);
      tc_fence_after();
      uint32_t v[WD][32];
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256 + hcol * 32;
#pragma unroll
      for (int d = 0; d < WD; ++d) {
        tmem_ld16(taddr + d * P1_NB, *reinterpret_cast<uint32_t(*)[16]>(&v[d][0]));
        tmem_ld16(taddr + d * P1_NB + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[d][16]));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);

      const int plane = p.sel.phys(q);
      const int y = ty * MT + m;
      const int x0 = tx * (P1_NB / 2) + hcol * 16;  // first of this thread's 16 pixels
      if (y < p.h && x0 < p.w) {
        // clamped-edge taps: rows above 0 / below h-1 all read the edge row
        uint64_t f_top = 0, f_bot = 0;
        if (y < p.r) f_top = p.suffix[y + 1];
        if (y >= p.h - p.r) f_bot = p.suffix[p.h - y];
        uint16_t e_top[16], e_bot[16];
        if (f_top | f_bot) {
          const uint16_t* row0 = p.in + (int64_t)plane * hw + x0;
          const uint16_t* row1 = row0 + (int64_t)(p.h - 1) * p.w;
          *reinterpret_cast<uint4*>(&e_top[0]) = __ldg(reinterpret_cast<const uint4*>(row0));
          *reinterpret_cast<uint4*>(&e_top[8]) = __ldg(reinterpret_cast<const uint4*>(row0) + 1);
          *reinterpret_cast<uint4*>(&e_bot[0]) = __ldg(reinterpret_cast<const uint4*>(row1));
          *reinterpret_cast<uint4*>(&e_bot[8]) = __ldg(reinterpret_cast<const uint4*>(row1) + 1);
        }
        const uint64_t half = 1ull << (p.shift - 1);
        uint32_t dig[GD][4];
#pragma unroll
        for (int g = 0; g < GD; ++g)
#pragma unroll
          for (int k = 0; k < 4; ++k) dig[g][k] = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint64_t tot = 0;
#pragma unroll
          for (int d = 0; d < WD; ++d) tot += (uint64_t)(v[d][2 * i] + (v[d][2 * i + 1] << 8)) << (8 * d);
          if (f_top | f_bot) tot += f_top * e_top[i] + f_bot * e_bot[i];
          const uint64_t g40 = (tot + half) >> p.shift;
          const uint32_t lo = (uint32_t)g40, hi = (uint32_t)(g40 >> 32);
          const int sh = 8 * (i & 3);
          dig[0][i >> 2] |= (lo & 0xffu) << sh;
          dig[1][i >> 2] |= ((lo >> 8) & 0xffu) << sh;
          dig[2][i >> 2] |= ((lo >> 16) & 0xffu) << sh;
          dig[3][i >> 2] |= (lo >> 24) << sh;
          dig[4][i >> 2] |= (hi & 0xffu) << sh;
        }
        uint8_t* dst = p.digits + ((int64_t)plane * GD) * hw + (int64_t)y * p.w + x0;
#pragma unroll
        for (int g = 0; g < GD; ++g)
          *reinterpret_cast<uint4*>(dst + g * hw) = make_uint4(dig[g][0], dig[g][1], dig[g][2], dig[g][3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

struct Pass2Params {
  const uint8_t* digits;   // [planes][GD][h][w]
  const double* lo;        // narrow Gaussian [planes][h][w] (may be null: out = G_hi)
  double* out;             // lo - G_hi
  uint16_t* buckets;       // optional
  uint64_t* minmax;        // optional [planes][2]
  const double* suffix_f;  // (double)suffix[j] * 2^-16
  double scale;            // in_scale * 2^-(S+8)
  int h, w, r;
  int n_sel, tiles_y, tiles_x;
  PlaneSel sel;
};

// ------------------------------------------------------------------ pass 2: digits -> float64 DoG, axis 1
__global__ void __launch_bounds__(NTHREADS, 1)
tcg_axis1_kernel(const __grid_constant__ CUtensorMap band_map, const __grid_constant__ CUtensorMap dig_map,
                 const Pass2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* band_s = smem;
  uint8_t* stage_s = smem + BAND_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(stage_s + P2_STAGES * P2_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t tiles_total = (int64_t)p.n_sel * p.tiles_y * p.tiles_x;
  const int64_t per_cta = (tiles_total + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t t_end = t_begin + per_cta < tiles_total ? t_begin + per_cta : tiles_total;

  if (threadIdx.x == 0) {
    mbar_init(&bars->band_full, 1);
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // x fastest: consecutive tiles of a CTA share half their columns
  auto decode = [&](int64_t t, int& q, int& tx, int& ty) {
    tx = (int)(t % p.tiles_x);
    const int64_t u = t / p.tiles_x;
    ty = (int)(u % p.tiles_y);
    q = (int)(u / p.tiles_y);
  };

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&band_map);
      prefetch_tmap(&dig_map);
      mbar_expect_tx(&bars->band_full, BAND_BYTES);
      for (int d = 0; d < WD; ++d)
        for (int pn = 0; pn < 2; ++pn)
          tma_load_2d(band_s + (d * 2 + pn) * BAND_PANEL_BYTES, &band_map, &bars->band_full, pn * 128, d * MT);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        int q, tx, ty;
        decode(t, q, tx, ty);
        mbar_wait(&bars->empty[stage], phase ^ 1);
        mbar_expect_tx(&bars->full[stage], P2_STAGE_BYTES);
        const int plane = p.sel.phys(q);
        for (int s = 0; s < GD; ++s)
          for (int pn = 0; pn < 2; ++pn)
            tma_load_3d(stage_s + stage * P2_STAGE_BYTES + (s * 2 + pn) * P2_PANEL_BYTES, &dig_map, &bars->full[stage],
                        tx * MT - HALO + pn * 128, ty * P2_NR, plane * GD + s);
        if (++stage == P2_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_u8(MT, P2_NR, false);
      mbar_wait(&bars->band_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t t = t_begin; t < t_end; ++t, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t accphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&bars->acc_empty[buf], accphase ^ 1);
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t b_base = smem_u32(stage_s + stage * P2_STAGE_BYTES);
#pragma unroll
        for (int j = JMIN; j <= WD + GD - 2; ++j) {
          const uint32_t d_tmem = tmem + buf * 256 + (j - JMIN) * P2_NR;
          bool first = true;
#pragma unroll
          for (int d = 0; d < WD; ++d) {
            const int s = j - d;
            if (s < 0 || s >= GD) continue;
#pragma unroll
            for (int ks = 0; ks < KBAND / 32; ++ks) {
              const uint64_t a_desc =
                  smem_desc(smem_u32(band_s) + (d * 2 + ks / 4) * BAND_PANEL_BYTES + (ks % 4) * 32, 16, 1024, LAYOUT_SW128);
              const uint64_t b_desc =
                  smem_desc(b_base + (s * 2 + ks / 4) * P2_PANEL_BYTES + (ks % 4) * 32, 16, 1024, LAYOUT_SW128);
              mma_u8(d_tmem, a_desc, b_desc, idesc, first ? 0u : 1u);
              first = false;
            }
          }
        }
        mma_commit(&bars->empty[stage]);
        mma_commit(&bars->acc_full[buf]);
        if (++stage == P2_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int hrow = ew >> 2;               // which 16 of the tile's 32 rows
    const int mx = quarter * 32 + lane;     // output column inside the tile
    const int64_t hw = (int64_t)p.h * p.w;
    uint64_t kmin = ~0ull, kmax = 0ull;
    int cur_plane = -1;
    auto flush = [&]() {
      if (p.minmax != nullptr && cur_plane >= 0) {
        const uint64_t a = warp_min_u64(kmin), b = warp_max_u64(kmax);
        if (lane == 0 && a <= b) {
          atomicMin((unsigned long long*)&p.minmax[2 * cur_plane], (unsigned long long)a);
          atomicMax((unsigned long long*)&p.minmax[2 * cur_plane + 1], (unsigned long long)b);
        }
      }
      kmin = ~0ull;
      kmax = 0ull;
    };
    int64_t it = 0;
    for (int64_t t = t_begin; t < t_end; ++t, ++it) {
      int q, tx, ty;
      decode(t, q, tx, ty);
      const int plane = p.sel.phys(q);
      if (plane != cur_plane) {
        flush();
        cur_plane = plane;
      }
      const int x = tx * MT + mx;
      const int y0 = ty * P2_NR + hrow * 16;
      const bool x_ok = x < p.w;
      // the narrow operand (coalesced: lanes are consecutive x), in flight while the accumulators arrive
      double lo[16];
      if (p.lo != nullptr) {
        const double* lp = p.lo + (int64_t)plane * hw + (int64_t)y0 * p.w + (x_ok ? x : p.w - 1);
#pragma unroll
        for (int n = 0; n < 16; ++n) lo[n] = (y0 + n < p.h) ? __ldg(lp + (int64_t)n * p.w) : 0.0;
      }
      // clamped-edge taps: columns left of 0 / right of w-1 all read the edge column of the same row
      double f_l = 0.0, f_r = 0.0;
      if (x_ok && x < p.r) f_l = p.suffix_f[x + 1];
      if (x_ok && x >= p.w - p.r) f_r = p.suffix_f[p.w - x];
      const bool edge_tile = tx * MT < p.r || tx * MT + MT > p.w - p.r;  // warp-uniform
      double e_l = 0.0, e_r = 0.0;  // lane n < 16: the edge samples of row y0 + n (40-bit integers, exact in float64)
      if (edge_tile && lane < 16 && y0 + lane < p.h) {
        const uint8_t* dp = p.digits + ((int64_t)plane * GD) * hw + (int64_t)(y0 + lane) * p.w;
        uint64_t a = 0, b = 0;
#pragma unroll
        for (int s = 0; s < GD; ++s) {
          a |= (uint64_t)dp[s * hw] << (8 * s);
          b |= (uint64_t)dp[s * hw + p.w - 1] << (8 * s);
        }
        e_l = (double)a;
        e_r = (double)b;
      }

      const int buf = (int)(it & 1);
      const uint32_t accphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&bars->acc_full[buf], accphase);
      tc_fence_after();
      uint32_t v[NACC2][16];
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256 + hrow * 16;
#pragma unroll
      for (int a = 0; a < NACC2; ++a) tmem_ld16(taddr + a * P2_NR, v[a]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);

      double res[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        uint64_t tot = 0;
#pragma unroll
        for (int a = 0; a < NACC2; ++a) tot += (uint64_t)v[a][n] << (8 * a);
        double g = (double)tot;
        if (edge_tile) {
          const double el = __shfl_sync(0xffffffffu, e_l, n), er = __shfl_sync(0xffffffffu, e_r, n);
          g += f_l * el + f_r * er;
        }
        g *= p.scale;
        res[n] = p.lo != nullptr ? lo[n] - g : g;
      }
      if (x_ok) {
        double* op = p.out + (int64_t)plane * hw + (int64_t)y0 * p.w + x;
        uint16_t* bp = p.buckets != nullptr ? p.buckets + (int64_t)plane * hw + (int64_t)y0 * p.w + x : nullptr;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          if (y0 + n < p.h) {
            op[(int64_t)n * p.w] = res[n];
            if (bp != nullptr) bp[(int64_t)n * p.w] = (uint16_t)bucket12(res[n]);
            const uint64_t key = f64_to_key(res[n]);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
          }
        }
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ the narrow Gaussian (radius <= 4), float64
// Exact scipy order (axis 0 first, then axis 1; acc = x0*w0; acc += (x-j + x+j) * wj for j = r..1).
// A CTA of 128 threads walks a strip of LO_TW output columns down the plane in blocks of LO_TH rows:
// thread t owns column x0 - 4 + t for the axis-0 pass (sliding window in registers, uint16 -> float64 on the
// way in), the block's axis-0 results sit in shared memory, then thread t < LO_TW produces column x0 + t.
constexpr int LO_R = 4, LO_NT = 128, LO_TW = LO_NT - 2 * LO_R, LO_TH = 32;

__global__ void __launch_bounds__(LO_NT)
lo2d_kernel(const uint16_t* __restrict__ in, double* __restrict__ out, const double scale, const int h, const int w,
            const double* __restrict__ hw_lo, const int r, const int strips, const PlaneSel sel) {
  __shared__ double vs[LO_TH][LO_NT];
  __shared__ double wsm[LO_R + 1];
  const int t = threadIdx.x;
  if (t <= LO_R) wsm[t] = t <= r ? hw_lo[t] : 0.0;
  const int q = blockIdx.x / strips, strip = blockIdx.x - q * strips;
  const int plane = sel.phys(q);
  const int x0 = strip * LO_TW;
  int xc = x0 - LO_R + t;  // the column this thread filters along axis 0 (clamped: mode='nearest')
  xc = xc < 0 ? 0 : (xc > w - 1 ? w - 1 : xc);
  const uint16_t* src = in + (int64_t)plane * h * w + xc;
  double* dst = out + (int64_t)plane * h * w;
  __syncthreads();
  const double w0 = wsm[0], w1 = wsm[1], w2 = wsm[2], w3 = wsm[3], w4 = wsm[4];
  // window[i] = sample y - 4 + i of the current row y
  double win[2 * LO_R + 1];
  auto fetch = [&](int y) {
    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    return dmul((double)__ldg(src + (int64_t)y * w), scale);
  };
#pragma unroll
  for (int i = 0; i < 2 * LO_R; ++i) win[i + 1] = fetch(i - LO_R);
  for (int yb = 0; yb < h; yb += LO_TH) {
#pragma unroll 4
    for (int yy = 0; yy < LO_TH; ++yy) {
#pragma unroll
      for (int i = 0; i < 2 * LO_R; ++i) win[i] = win[i + 1];
      win[2 * LO_R] = fetch(yb + yy + LO_R);
      double acc = dmul(win[LO_R], w0);
      if (r >= 4) acc = dadd(acc, dmul(dadd(win[LO_R - 4], win[LO_R + 4]), w4));
      if (r >= 3) acc = dadd(acc, dmul(dadd(win[LO_R - 3], win[LO_R + 3]), w3));
      if (r >= 2) acc = dadd(acc, dmul(dadd(win[LO_R - 2], win[LO_R + 2]), w2));
      if (r >= 1) acc = dadd(acc, dmul(dadd(win[LO_R - 1], win[LO_R + 1]), w1));
      vs[yy][t] = acc;
    }
    __syncthreads();
    const int x = x0 + t;
    if (t < LO_TW && x < w) {
#pragma unroll 4
      for (int yy = 0; yy < LO_TH; ++yy) {
        if (yb + yy >= h) break;
        const double* vr = &vs[yy][t + LO_R];
        double acc = dmul(vr[0], w0);
        if (r >= 4) acc = dadd(acc, dmul(dadd(vr[-4], vr[4]), w4));
        if (r >= 3) acc = dadd(acc, dmul(dadd(vr[-3], vr[3]), w3));
        if (r >= 2) acc = dadd(acc, dmul(dadd(vr[-2], vr[2]), w2));
        if (r >= 1) acc = dadd(acc, dmul(dadd(vr[-1], vr[1]), w1));
        dst[(int64_t)(yb + yy) * w + x] = acc;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// uint8 tensor (inner, rows[, planes]) with a (box_inner x box_rows) box
static int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t planes, uint32_t box_inner,
                    uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return AMT_ERR_UNSUPPORTED;
  const cuuint32_t rank = planes > 0 ? 3 : 2;
  cuuint64_t dims[3] = {inner, rows, planes};
  cuuint64_t strides[2] = {inner, inner * rows};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AMT_OK : AMT_ERR_CUDA;
}

}  // namespace tc
}  // namespace amt

struct amt_tcg {
  int device, r, S;
  std::vector<uint64_t>* W;  // W[t], t = 0..r (symmetric)
  uint8_t* band;             // device [WD][128][256]
  uint64_t* suffix;          // device [r + 2]
  double* suffix_f;          // device [r + 2]: suffix * 2^-16
  CUtensorMap band_map;
};

namespace amt {
namespace tc {

constexpr size_t P1_SMEM = 1024 + BAND_BYTES + P1_STAGES * P1_STAGE_BYTES + sizeof(Barriers);
constexpr size_t P2_SMEM = 1024 + BAND_BYTES + P2_STAGES * P2_STAGE_BYTES + sizeof(Barriers);

bool tcg_shape_ok(int64_t h, int64_t w) { return h >= 128 && w >= 128 && w % 16 == 0 && h * w < (1ll << 31); }

static int sel_count(int64_t n_img, const PlaneSel& sel, int64_t* n_sel) {
  if (sel.every <= 0) {
    *n_sel = n_img;
    return AMT_OK;
  }
  if (sel.every < 2 || sel.skip < 0 || sel.skip >= sel.every || n_img % sel.every != 0) return AMT_ERR_INVALID;
  *n_sel = n_img / sel.every * (sel.every - 1);
  return AMT_OK;
}

int tcg_axis0(const amt_tcg* g, const uint16_t* in, int64_t n_img, int64_t h, int64_t w, uint8_t* digits, PlaneSel sel,
              cudaStream_t st) {
  if (!g || !in || !digits || n_img <= 0) return AMT_ERR_INVALID;
  if (!tcg_shape_ok(h, w) || ((uintptr_t)in % 16) || ((uintptr_t)digits % 16)) return AMT_ERR_UNSUPPORTED;
  int64_t n_sel = 0;
  AMT_TRY(sel_count(n_img, sel, &n_sel));
  if (n_sel == 0) return AMT_OK;
  CUtensorMap in_map;
  AMT_TRY(make_map(&in_map, in, (uint64_t)w * 2, (uint64_t)h, (uint64_t)n_img, P1_NB, KBAND, CU_TENSOR_MAP_SWIZZLE_64B));
  Pass1Params p{};
  p.in = in;
  p.digits = digits;
  p.suffix = g->suffix;
  p.h = (int)h;
  p.w = (int)w;
  p.r = g->r;
  p.shift = g->S - 24;
  p.n_sel = (int)n_sel;
  p.tiles_y = (int)ceil_div(h, MT);
  p.tiles_x = (int)ceil_div(w * 2, P1_NB);
  p.sel = sel;
  const int64_t tiles = n_sel * p.tiles_y * p.tiles_x;
  AMT_CUDA_TRY(cudaFuncSetAttribute(tcg_axis0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P1_SMEM));
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  tcg_axis0_kernel<<<grid, NTHREADS, P1_SMEM, st>>>(g->band_map, in_map, p);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int tcg_axis1(const amt_tcg* g, const uint8_t* digits, const double* lo, double in_scale, double* out, int64_t n_img,
              int64_t h, int64_t w, uint16_t* buckets, uint64_t* minmax, PlaneSel sel, cudaStream_t st) {
  if (!g || !digits || !out || n_img <= 0) return AMT_ERR_INVALID;
  if (!tcg_shape_ok(h, w) || ((uintptr_t)digits % 16)) return AMT_ERR_UNSUPPORTED;
  int64_t n_sel = 0;
  AMT_TRY(sel_count(n_img, sel, &n_sel));
  if (n_sel == 0) return AMT_OK;
  CUtensorMap dig_map;
  AMT_TRY(make_map(&dig_map, digits, (uint64_t)w, (uint64_t)h, (uint64_t)n_img * GD, 128, P2_NR, CU_TENSOR_MAP_SWIZZLE_128B));
  Pass2Params p{};
  p.digits = digits;
  p.lo = lo;
  p.out = out;
  p.buckets = buckets;
  p.minmax = minmax;
  p.suffix_f = g->suffix_f;
  p.scale = std::ldexp(in_scale, -(g->S + 8));
  p.h = (int)h;
  p.w = (int)w;
  p.r = g->r;
  p.n_sel = (int)n_sel;
  p.tiles_y = (int)ceil_div(h, P2_NR);
  p.tiles_x = (int)ceil_div(w, MT);
  p.sel = sel;
  const int64_t tiles = n_sel * p.tiles_y * p.tiles_x;
  AMT_CUDA_TRY(cudaFuncSetAttribute(tcg_axis1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P2_SMEM));
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  tcg_axis1_kernel<<<grid, NTHREADS, P2_SMEM, st>>>(g->band_map, dig_map, p);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w, const double* hw_lo, int r_lo,
         PlaneSel sel, cudaStream_t st) {
  if (!in || !out || !hw_lo || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (r_lo < 0 || r_lo > LO_R) return AMT_ERR_UNSUPPORTED;
  int64_t n_sel = 0;
  AMT_TRY(sel_count(n_img, sel, &n_sel));
  if (n_sel == 0) return AMT_OK;
  const int strips = (int)ceil_div(w, LO_TW);
  if (n_sel * strips >= (1ll << 31)) return AMT_ERR_CAPACITY;
  lo2d_kernel<<<(unsigned)(n_sel * strips), LO_NT, 0, st>>>(in, out, in_scale, (int)h, (int)w, hw_lo, r_lo, strips, sel);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace tc
}  // namespace amt

extern "C" {

int amt_tcg_create(const double* half_w_host, int radius, int device, amt_tcg** out) {
  using namespace amt;
  using namespace amt::tc;
  if (!half_w_host || !out || radius < 1) return AMT_ERR_INVALID;
  if (radius > HALO) return AMT_ERR_UNSUPPORTED;
  for (int t = 0; t <= radius; ++t)
    if (!(half_w_host[t] > 0.0) || half_w_host[t] > half_w_host[0]) return AMT_ERR_UNSUPPORTED;
  AMT_CUDA_TRY(cudaSetDevice(device));
  // S: the largest scale that keeps the centre weight inside 32 bits (with room for the +-1 of the normalisation)
  int S = 0;
  while (S < 40 && std::ldexp(half_w_host[0], S + 1) < 4294967000.0) ++S;
  if (S < 26) return AMT_ERR_UNSUPPORTED;
  std::vector<uint64_t> W(radius + 1);
  std::vector<double> frac(radius + 1);
  int64_t sum = 0;
  for (int t = 0; t <= radius; ++t) {
    const double x = std::ldexp(half_w_host[t], S);
    const double f = std::floor(x + 0.5);
    W[t] = (uint64_t)f;
    frac[t] = x - f;  // in [-0.5, 0.5): positive = rounded down
    sum += (t == 0 ? 1 : 2) * (int64_t)W[t];
  }
  // make the integer weights sum to 2^S exactly (a constant image filters to itself): one unit on the centre tap if
  // the defect is odd, then one unit on the tap pairs whose rounding went furthest the other way
  int64_t defect = ((int64_t)1 << S) - sum;
  if (defect % 2 != 0) {
    const int64_t s1 = defect > 0 ? 1 : -1;
    W[0] = (uint64_t)((int64_t)W[0] + s1);
    defect -= s1;
  }
  while (defect != 0) {
    const int64_t s1 = defect > 0 ? 1 : -1;
    int best = -1;
    for (int t = 1; t <= radius; ++t)
      if (best < 0 || (double)s1 * frac[t] > (double)s1 * frac[best]) best = t;
    if (best < 0 || (s1 < 0 && W[best] == 0)) break;
    W[best] = (uint64_t)((int64_t)W[best] + s1);
    frac[best] -= (double)s1;
    defect -= 2 * s1;
  }
  if (defect != 0) return AMT_ERR_UNSUPPORTED;
  for (int t = 0; t <= radius; ++t)
    if (W[t] >> 32) return AMT_ERR_UNSUPPORTED;

  amt_tcg* g = new (std::nothrow) amt_tcg();
  if (!g) return AMT_ERR_CAPACITY;
  std::memset(g, 0, sizeof(*g));
  g->device = device;
  g->r = radius;
  g->S = S;
  g->W = new std::vector<uint64_t>(W);
  std::vector<uint8_t> band((size_t)WD * MT * KBAND, 0);
  for (int d = 0; d < WD; ++d)
    for (int m = 0; m < MT; ++m)
      for (int k = 0; k < KBAND; ++k) {
        const int t = std::abs(k - HALO - m);
        if (t <= radius) band[((size_t)d * MT + m) * KBAND + k] = (uint8_t)((W[t] >> (8 * d)) & 0xff);
      }
  std::vector<uint64_t> suffix(radius + 2, 0);
  std::vector<double> suffix_f(radius + 2, 0.0);
  for (int j = radius; j >= 0; --j) suffix[j] = suffix[j + 1] + W[j];
  for (int j = 0; j <= radius + 1; ++j) suffix_f[j] = std::ldexp((double)suffix[j], -16);
  auto fail = [&](int s) {
    amt_tcg_destroy(g);
    return s;
  };
  if (cudaMalloc((void**)&g->band, band.size()) != cudaSuccess || cudaMalloc((void**)&g->suffix, suffix.size() * 8) != cudaSuccess ||
      cudaMalloc((void**)&g->suffix_f, suffix_f.size() * 8) != cudaSuccess)
    return fail(AMT_ERR_CUDA);
  if (cudaMemcpy(g->band, band.data(), band.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(g->suffix, suffix.data(), suffix.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(g->suffix_f, suffix_f.data(), suffix_f.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(AMT_ERR_CUDA);
  const int ms = make_map(&g->band_map, g->band, KBAND, (uint64_t)WD * MT, 0, 128, MT, CU_TENSOR_MAP_SWIZZLE_128B);
  if (ms != AMT_OK) return fail(ms);
  *out = g;
  return AMT_OK;
}

void amt_tcg_destroy(amt_tcg* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->band) cudaFree(g->band);
  if (g->suffix) cudaFree(g->suffix);
  if (g->suffix_f) cudaFree(g->suffix_f);
  delete g->W;
  delete g;
}

int amt_tcg_weights(const amt_tcg* g, uint64_t* w_host, int* scale_bits) {
  if (!g || !w_host || !scale_bits) return AMT_ERR_INVALID;
  for (int t = 0; t <= g->r; ++t) w_host[t] = (*g->W)[t];
  *scale_bits = g->S;
  return AMT_OK;
}

int amt_tcg_supported(int64_t h, int64_t w, int radius) {
  return radius >= 1 && radius <= amt::tc::HALO && amt::tc::tcg_shape_ok(h, w) ? 1 : 0;
}

size_t amt_tcg_digit_bytes(int64_t n_img, int64_t h, int64_t w) { return (size_t)n_img * amt::tc::GD * h * w; }

int amt_tcg_axis0(const amt_tcg* g, const uint16_t* in, int64_t n_img, int64_t h, int64_t w, uint8_t* digits,
                  int skip_every, int skip_offset, amt_stream_t stream) {
  return amt::tc::tcg_axis0(g, in, n_img, h, w, digits, amt::tc::PlaneSel{skip_every, skip_offset}, amt::as_stream(stream));
}

int amt_tcg_axis1(const amt_tcg* g, const uint8_t* digits, const double* lo, double in_scale, double* out, int64_t n_img,
                  int64_t h, int64_t w, uint16_t* buckets, uint64_t* minmax_keys, int skip_every, int skip_offset,
                  amt_stream_t stream) {
  if (minmax_keys && n_img > 0) AMT_TRY(amt::minmax_init(minmax_keys, n_img, amt::as_stream(stream)));
  return amt::tc::tcg_axis1(g, digits, lo, in_scale, out, n_img, h, w, buckets, minmax_keys,
                            amt::tc::PlaneSel{skip_every, skip_offset}, amt::as_stream(stream));
}

int amt_gauss_lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
                   const double* half_w_lo, int r_lo, int skip_every, int skip_offset, amt_stream_t stream) {
  return amt::tc::lo2d(in, in_scale, out, n_img, h, w, half_w_lo, r_lo, amt::tc::PlaneSel{skip_every, skip_offset},
                       amt::as_stream(stream));
}

}  // extern "C"
