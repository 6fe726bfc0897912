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
//     [128 outputs x 256 inputs], resident in TENSOR MEMORY (the A operand) for the life of a persistent CTA;
//   * pass 1 (image axis 0): the raw uint16 image is read BY TMA AS BYTES.  A tile of 256 rows x 64
//     bytes (32 pixels, low and high byte interleaved) is the MN-major B operand as it lies in memory;
//     the low / high byte columns come out as neighbouring accumulator columns and are recombined in
//     the epilogue: 4 digit MMAs x 8 K-steps (M = 128, N = 64, K = 32) per tile, int32 accumulators
//     in TMEM (every partial sum < 2^23: exact).  The epilogue rebuilds the 53-bit integer, adds the
//     clamped-edge ('nearest') taps, rounds to 40 bits and stores five uint8 digit planes;
//   * pass 2 (image axis 1): the five digit planes are the K-major B operand, again straight from TMA;
//     weight digit d x sample digit s accumulates into the TMEM accumulator of d + s (products of equal
//     significance share one accumulator; d + s < 3 is at most 5.2e-12 of full scale and skipped): 14 digit
//     products x 8 K-steps (M = 128, N = 32, K = 32).  The epilogue shifts the five accumulators together
//     in 64-bit integers, converts ONCE to float64, subtracts from the narrow Gaussian (lo2d_kernel
//     below, float64, scipy's order) and writes the DoG plane, its selection buckets and min / max.
// Errors: the approximations are the 32-bit rounding of the weights (|dw| <= 2^-37 per tap, zero in sum), the
// 40-bit rounding between the passes and the digit products pass 2 leaves out: amt_tcg_error_bound adds them up
// (5.3e-10 of the [0, 1] scale for sigma = 16, a proof), ~1.5e-11 in practice (tests/test_gpu_tcgauss.py measures
// it); the reference's tolerance for filtered planes is 1e-5.  The thresholded channel takes this path too
// (executor.cu): every sample a decision depends on is then re-evaluated in scipy's exact order (decide.cu).
//
// Kernel anatomy (both passes): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane), warps 2..9 = epilogue (TMEM lane quarter = warp % 4).  mbarrier rings: full/empty per
// operand stage, acc_full/acc_empty for the accumulators (ONE set: the band matrix holds half of tensor memory; the
// epilogue's arithmetic overlaps the MMAs of the next tile once its tcgen05.ld have drained the accumulators).
// Persistent grid: one CTA per SM, a contiguous range of tiles each.  The warp-specialised fused pass 2 adds two halo
// warps and eight lo warps (the narrow Gaussian, float64, scipy's order): see tcg_axis1_kernel.
// Both kernels run at the board's power cap (scripts/power_probe.py): what costs energy costs time.

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "internal.cuh"

#ifndef AMT_TCG_SUSPEND_NS
#define AMT_TCG_SUSPEND_NS 20000
#endif

namespace amt {
namespace tc {

constexpr int KBAND = 256;       // inputs per output tile along the filter axis (128 + 2 * 64)
constexpr int MT = 128;          // outputs per tile along the filter axis (UMMA M)
constexpr int HALO = 64;         // largest radius
constexpr int WD = 4;            // weight digits (base 256)
constexpr int GD = 5;            // digits of the pass-1 result (40 bits)
constexpr int JMIN = 3;          // digit products with d + s < JMIN are dropped in pass 2 (amt_tcg_error_bound counts them)
constexpr int NACC2 = WD + GD - 1 - JMIN;  // accumulators of pass 2 (j = JMIN .. WD+GD-2)
constexpr int P1_NB = 64;        // pass 1: bytes (UMMA N) per tile along the contiguous axis = 32 pixels
constexpr int P1_STAGES = 8;
constexpr int P1_GROUP = 4;      // x tiles per work unit: their 4 x 32 pixels leave as 128-byte rows through one staging tile
constexpr int P2_NR = 32;        // pass 2: rows (UMMA N) per tile
constexpr int P2_STAGES = 3;        // pass 2 reading the narrow Gaussian from memory: three 72 KB stages
constexpr int P2F_STAGES = 2;       // fused pass 2: two 52 KB stages (three measured 2 % slower: 0.732 against 0.717 ms per 32 planes)
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = (2 + EPI_WARPS) * 32;
constexpr int MAX_STAGES = 8;
constexpr uint32_t P1_STAGE_BYTES = KBAND * P1_NB;               // 16 KB
constexpr uint32_t P2_PANEL_BYTES = P2_NR * 128;                 // 4 KB
constexpr uint32_t P2_DIG_BYTES = GD * 2 * P2_PANEL_BYTES;       // 40 KB of digit panels ...
constexpr uint32_t P2_LO_BYTES = P2_NR * MT * 8;                 // ... and the 32 x 128 float64 tile of the narrow Gaussian
constexpr uint32_t P2_STAGE_BYTES = P2_DIG_BYTES + P2_LO_BYTES;  // 72 KB
// pass 2 with the narrow Gaussian fused in: the stage carries a raw uint16 tile (tile + a halo of LO_HALO samples on
// every side) instead of the float64 tile, and two buffers of axis-0 results sit behind the stages
constexpr int LO_HALO = 4;                                        // largest radius of the narrow Gaussian
constexpr int P2F_RAW_HX = 8;                                     // x halo of the raw tile: TMA wants the box to start on a 16-byte boundary (8 uint16)
constexpr int P2F_RAW_W = MT + 2 * P2F_RAW_HX;                    // 144 samples = 288 bytes
constexpr int P2F_RAW_H = P2_NR + 2 * LO_HALO;                    // 40 rows
constexpr uint32_t P2F_RAW_BYTES = P2F_RAW_W * P2F_RAW_H * 2;     // 11520
constexpr uint32_t P2F_STAGE_BYTES = ((P2_DIG_BYTES + P2F_RAW_BYTES + 1023) / 1024) * 1024;  // 52 KB (panels stay 1 KB aligned)
constexpr int P2F_V_W = MT + 2 * LO_HALO;                         // axis-0 results of a tile: 136 columns (tile + LO_HALO each side)
constexpr uint32_t P2F_V_BYTES = P2_NR * P2F_V_W * 8;             // ... x 32 rows, float64
constexpr uint32_t P1_OUT_BYTES = GD * MT * 128;                 // pass 1 staging: five 128 x 128-byte digit tiles
// TMEM map (512 columns): the band matrix (A operand) lives in columns [0, 256): digit d, K step ks (32 inputs =
// 8 columns of 4 bytes) at column d * 64 + ks * 8; the accumulators in [256, 512).
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_BAND_COLS_PER_DIGIT = KBAND / 4;  // 64
constexpr uint32_t TMEM_ACC0 = 256;

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
// A waiting warp may sleep in hardware: with a suspend-time hint `mbarrier.try_wait` holds the thread for up to that
// long (it wakes when the phase completes); without it a waiting warp polls every ~30 ns, and ten of a CTA's twenty
// warps are waiting at any time -- on a part that runs these kernels AT ITS POWER CAP (scripts/power_probe.py) polling
// costs clock.  amt_tune("tcg_suspend_ns", 0) switches the hint off (A/B measurements).
__constant__ uint32_t c_suspend_ns = AMT_TCG_SUSPEND_NS;
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
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Every wait carries a watchdog: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
// By the clock, not by a spin count: with the hint one iteration lasts anything from 30 ns to 20 us.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint32_t ns = c_suspend_ns;
  const uint64_t t0 = global_timer_ns();
  while (!(ns ? mbar_try_wait_hint(bar, parity, ns) : mbar_try_wait(bar, parity))) {
    if (global_timer_ns() - t0 > 10000000000ull) __trap();  // 10 s (a pre-empted kernel keeps this clock running)
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// shared -> global tile store (bulk async group of the issuing thread); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// warpgroup-wide register re-allocation (all four warps of an aligned warpgroup execute it)
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// true in exactly one lane of a converged warp.  Code that issues tcgen05.mma must branch on THIS (not on
// lane == 0): ptxas then keeps the instruction's operands in uniform registers; under an ordinary divergent branch it
// wraps every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (measured: 60-80 clk per MMA whatever its size).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
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
// D[tmem] (+)= A[tmem] * B[smem], uint8 x uint8 -> int32, issued by one thread for the CTA.  A (the band matrix,
// the same for every tile) comes from tensor memory: with both operands in shared memory an M = 128 MMA re-reads
// its 4 KB A block every time and runs at the shared-memory operand rate (measured ~73 B/clk: 84 clk for N = 64,
// 70 clk for N = 32) instead of the math rate (N / 2 clk).
__device__ __forceinline__ void mma_u8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t addr, uint32_t* v) {
  static_assert(N == 8 || N == 16, "8 or 16 columns");
  if (N == 16) tmem_ld16(addr, v); else tmem_ld8(addr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// d = a * b + c with a 64-bit sum (IMAD.WIDE.U32).  The epilogues call it with b = a power of two that ptxas cannot
// see through -- a kernel PARAMETER (Pass1Params / Pass2Params c8, c16, c24; a constant hidden behind an
// `asm volatile mov` is folded by ptxas all the same).  With a visible power of two ptxas spends four or five
// instructions on every 64-bit term (IMAD.SHL / IMAD.HI / IADD3 / IADD3.X) and the integer combine was a third of
// both epilogues.
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}

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
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t acc_full, acc_empty;
  uint64_t lo_full[2], lo_empty[2];  // warp-specialised pass 2: narrow-Gaussian tiles handed from the lo warps to the epilogue warps
  uint64_t suffix[HALO + 2];   // pass 1: integer tail sums of the weights (clamped-edge taps)
  double suffix_f[HALO + 2];   // pass 2: the same, as float64 * 2^-(8 JMIN)
  uint32_t tmem_base;
  uint32_t pad;
};

// One-time setup shared by both passes: barriers, tensor memory, the band matrix into tensor memory.
// Band row m (all four digits, 256 bytes each) is written by the thread that owns TMEM lane m: epilogue warps with
// warp % 4 == m / 32.  K element k of a row sits in column k / 4, byte k % 4.
__device__ __forceinline__ uint32_t tcg_setup(Barriers* bars, const uint8_t* __restrict__ band, int n_stages,
                                              int empty_count, const uint64_t* __restrict__ suffix,
                                              const double* __restrict__ suffix_f, int epi_warps = EPI_WARPS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], empty_count);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, epi_warps);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->lo_full[b], epi_warps);
      mbar_init(&bars->lo_empty[b], epi_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < HALO + 2; i += blockDim.x) {
    if (suffix != nullptr) bars->suffix[i] = suffix[i];
    if (suffix_f != nullptr) bars->suffix_f[i] = suffix_f[i];
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (warp >= 2 && warp < 6) {
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int d = 0; d < WD; ++d) {
      const uint4* src = reinterpret_cast<const uint4*>(band + ((size_t)d * MT + m) * KBAND);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 x = __ldg(src + c * 4 + k);
          v[4 * k] = x.x, v[4 * k + 1] = x.y, v[4 * k + 2] = x.z, v[4 * k + 3] = x.w;
        }
        tmem_st16(lane_addr + d * TMEM_BAND_COLS_PER_DIGIT + c * 16, v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return tmem;
}

// a CTA's contiguous range of tiles, walked without divisions: (plane, slow, fast) with `fast` innermost
struct TileWalk {
  int q, slow, fast, n_slow, n_fast;
  __device__ void init(int64_t t, int n_slow_, int n_fast_) {
    n_slow = n_slow_, n_fast = n_fast_;
    fast = (int)(t % n_fast);
    const int64_t u = t / n_fast;
    slow = (int)(u % n_slow);
    q = (int)(u / n_slow);
  }
  __device__ __forceinline__ void next() {
    if (++fast == n_fast) {
      fast = 0;
      if (++slow == n_slow) slow = 0, ++q;
    }
  }
};

struct Pass1Params {
  const uint16_t* in;      // [planes][h][w]
  const uint8_t* band;     // [WD][128][256]
  const uint64_t* suffix;  // suffix[j] = sum_{t >= j} W[t], j = 0 .. HALO + 1 (zero beyond the radius)
  int h, w, r, shift;      // G1q = (sum + 2^(shift-1)) >> shift
  int n_sel;               // logical planes
  int tiles_y, groups_x;   // per plane: 128-row tiles, groups of P1_GROUP 32-pixel tiles
  PlaneSel sel;
  int dbg;                 // amt_tune "tcg_debug" (timing experiments): 1 = no MMAs, 2 = no epilogue arithmetic / stores, 4 = no TMEM loads, 8 = no stores
  uint32_t c8, c16, c24;   // 2^8, 2^16, 2^24 as kernel PARAMETERS: multipliers of the epilogue's IMAD.WIDE chain ptxas cannot fold
};

// ------------------------------------------------------------------ pass 1: uint16 image -> 40-bit digits, axis 0
// Work unit = 128 rows x 128 pixels = P1_GROUP tiles of 32 pixels (the accumulators of one tile fill the 256 TMEM
// columns next to the band).  The digits of a unit are collected in a shared-memory staging tile per digit plane
// (128 rows x 128 bytes, 128-byte swizzle: conflict-free 16-byte writes) and leave by TMA store: whole 128-byte
// lines.  (Written straight from the registers, each warp store touched 32 lines with 16 bytes each and the kernel
// spent 3/4 of its time on those stores.)
__global__ void __launch_bounds__(NTHREADS, 1)
tcg_axis0_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                 const Pass1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_s = stage_s + P1_STAGES * P1_STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(out_s + P1_OUT_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t units_total = (int64_t)p.n_sel * p.tiles_y * p.groups_x;
  const int64_t per_cta = (units_total + gridDim.x - 1) / gridDim.x;
  const int64_t u_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t u_end = u_begin + per_cta < units_total ? u_begin + per_cta : units_total;
  const int n_units = u_end > u_begin ? (int)(u_end - u_begin) : 0;

  const uint32_t tmem = tcg_setup(bars, p.band, P1_STAGES, 1, p.suffix, nullptr);
  // unit -> (logical plane, x group, y tile), y fastest: consecutive units of a CTA share half their rows (L2)
  TileWalk tw;
  tw.init(u_begin, p.groups_x, p.tiles_y);

  if (warp == 0) {
    prefetch_tmap(&in_map);
    int stage = 0;
    uint32_t phase = 0;
    for (int u = 0; u < n_units; ++u, tw.next()) {
      const int plane = p.sel.phys(tw.q);
      for (int k = 0; k < P1_GROUP; ++k) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bars->full[stage], P1_STAGE_BYTES);
          tma_load_3d(stage_s + stage * P1_STAGE_BYTES, &in_map, &bars->full[stage], (tw.slow * P1_GROUP + k) * P1_NB,
                      tw.fast * MT - HALO, plane);
        }
        __syncwarp();
        if (++stage == P1_STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop (uniform control flow and addresses); one elected lane issues
    constexpr uint32_t idesc = idesc_u8(MT, P1_NB, true);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    int stage = 0;
    uint32_t phase = 0;
    const int n_tiles = n_units * P1_GROUP;
    for (int it = 0; it < n_tiles; ++it) {
      mbar_wait(&bars->acc_empty, (uint32_t)(it & 1) ^ 1);
      mbar_wait(&bars->full[stage], phase);
      tc_fence_after();
      const uint32_t b_base = smem_u32(stage_s + stage * P1_STAGE_BYTES);
      if (elect_one()) {
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int ks = 0; ks < KBAND / 32; ++ks) {
            // B: MN-major, 64-byte swizzle: 8 K rows of 64 bytes per atom (512 bytes), 32 K rows per MMA
            const uint64_t b_desc = smem_desc(b_base + ks * 32 * P1_NB, P1_STAGE_BYTES, 512, LAYOUT_SW64);
#pragma unroll
            for (int d = 0; d < WD; ++d)
              mma_u8_ts(tm + TMEM_ACC0 + d * P1_NB, tm + d * TMEM_BAND_COLS_PER_DIGIT + ks * 8, b_desc, idesc, ks > 0 ? 1u : 0u);
          }
        }
        mma_commit(&bars->empty[stage]);
        mma_commit(&bars->acc_full);
      }
      __syncwarp();
      if (++stage == P1_STAGES) stage = 0, phase ^= 1;
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lanes this warp may read
    const int hcol = ew >> 2;            // which half of a tile's 64 byte columns
    const int m = quarter * 32 + lane;   // output row inside the tile
    const int64_t hw = (int64_t)p.h * p.w;
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + TMEM_ACC0 + hcol * 32;
    const uint32_t half = 1u << (p.shift - 1);
    uint8_t* my_row = out_s + m * 128;
    const uint32_t c8 = p.c8, c16 = p.c16, c24 = p.c24;
    int it = 0;
    for (int u = 0; u < n_units; ++u, tw.next()) {
      const int plane = p.sel.phys(tw.q);
      const int ty = tw.fast, xg = tw.slow;
      const int y = ty * MT + m;
      const bool edge_unit = ty * MT < p.r || ty * MT + MT > p.h - p.r;  // warp-uniform
      const uint64_t f_top = (edge_unit && y < p.r) ? bars->suffix[y + 1] : 0ull;
      const uint64_t f_bot = (edge_unit && y < p.h && y >= p.h - p.r) ? bars->suffix[p.h - y] : 0ull;
#pragma unroll 1
      for (int k = 0; k < P1_GROUP; ++k, ++it) {
        mbar_wait(&bars->acc_full, (uint32_t)(it & 1));
        tc_fence_after();
        uint32_t v[WD][32];
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int d = 0; d < WD; ++d) {
            tmem_ld16(taddr + d * P1_NB, &v[d][0]);
            tmem_ld16(taddr + d * P1_NB + 16, &v[d][16]);
          }
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int d = 0; d < WD; ++d)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[d][i] = (uint32_t)(it + d + i);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty);  // the MMAs of the next tile may overwrite the accumulators now
        if ((p.dbg & 2) && k == 0 && u > 0) {
          if (warp == 2 && lane == 0) tma_store_wait_read();
          epi_barrier();
        }
        if (p.dbg & 2) continue;

        const int x0 = (xg * P1_GROUP + k) * (P1_NB / 2) + hcol * 16;  // first of this thread's 16 pixels
        // 53-bit sums, rounded to 40 bits: g = (sum_d 256^d (lo_d + 256 hi_d) + half) >> shift
        //   = lo_0 + half + 2^8 (lo_1 + hi_0) + 2^16 (lo_2 + hi_1) + 2^24 (lo_3 + hi_2) + 2^32 hi_3, every part < 2^24
        uint32_t glo[16], ghi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint64_t tot = mad_wide(v[1][2 * i] + v[0][2 * i + 1], c8, 0ull);
          tot = mad_wide(v[2][2 * i] + v[1][2 * i + 1], c16, tot);
          tot = mad_wide(v[3][2 * i] + v[2][2 * i + 1], c24, tot);
          tot += ((uint64_t)v[3][2 * i + 1] << 32) | (uint64_t)(v[0][2 * i] + half);
          glo[i] = (uint32_t)tot, ghi[i] = (uint32_t)(tot >> 32);
        }
        // clamped-edge taps (mode='nearest'): rows above 0 / below h-1 all read the edge row.  Only the first and
        // the last units of a column of units get here.
        if (edge_unit && (f_top | f_bot) != 0ull && x0 < p.w) {
          const uint16_t* row0 = p.in + (int64_t)plane * hw + x0;
          const uint16_t* row1 = row0 + (int64_t)(p.h - 1) * p.w;
          uint32_t et[8], eb[8];
          *reinterpret_cast<uint4*>(&et[0]) = __ldg(reinterpret_cast<const uint4*>(row0));
          *reinterpret_cast<uint4*>(&et[4]) = __ldg(reinterpret_cast<const uint4*>(row0) + 1);
          *reinterpret_cast<uint4*>(&eb[0]) = __ldg(reinterpret_cast<const uint4*>(row1));
          *reinterpret_cast<uint4*>(&eb[4]) = __ldg(reinterpret_cast<const uint4*>(row1) + 1);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t e0 = (i & 1) ? (et[i >> 1] >> 16) : (et[i >> 1] & 0xffffu);
            const uint32_t e1 = (i & 1) ? (eb[i >> 1] >> 16) : (eb[i >> 1] & 0xffffu);
            const uint64_t tot = (((uint64_t)ghi[i] << 32) | glo[i]) + f_top * e0 + f_bot * e1;
            glo[i] = (uint32_t)tot, ghi[i] = (uint32_t)(tot >> 32);
          }
        }
        uint32_t dig[GD][4];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          uint32_t l[4], hb[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int i = 4 * g4 + kk;
            l[kk] = __funnelshift_r(glo[i], ghi[i], p.shift);  // bits [shift, shift + 32)
            hb[kk] = ghi[i] >> p.shift;                          // bits [shift + 32, shift + 40)
          }
          // 4 x 4 byte transpose: l[kk] holds digits 0..3 of pixel kk -> dig[d] holds digit d of pixels 0..3
          const uint32_t x01 = __byte_perm(l[0], l[1], 0x5140), y01 = __byte_perm(l[0], l[1], 0x7362);
          const uint32_t x23 = __byte_perm(l[2], l[3], 0x5140), y23 = __byte_perm(l[2], l[3], 0x7362);
          dig[0][g4] = __byte_perm(x01, x23, 0x5410);
          dig[1][g4] = __byte_perm(x01, x23, 0x7632);
          dig[2][g4] = __byte_perm(y01, y23, 0x5410);
          dig[3][g4] = __byte_perm(y01, y23, 0x7632);
          dig[4][g4] = __byte_perm(__byte_perm(hb[0], hb[1], 0x0040), __byte_perm(hb[2], hb[3], 0x0040), 0x5410);
        }
        if ((p.dbg & 8) && (dig[0][0] ^ dig[1][1] ^ dig[2][2] ^ dig[3][3] ^ dig[4][0]) != 0x9e3779b9u) continue;  // timing: no stores
        if (k == 0 && u > 0) {
          // the staging tile is free once the previous unit's TMA stores have READ it: waited for here, behind the
          // arithmetic of the unit's first tile, not in front of it (the epilogue warps spent 15 % of their time there)
          if (warp == 2 && lane == 0) tma_store_wait_read();
          epi_barrier();
        }
        // 16 bytes per digit plane into the staging tile: chunk c of row m sits at chunk position c ^ (m & 7)
        uint8_t* dst = my_row + (((k * 2 + hcol) ^ (m & 7)) << 4);
#pragma unroll
        for (int g = 0; g < GD; ++g)
          *reinterpret_cast<uint4*>(dst + g * (MT * 128)) = make_uint4(dig[g][0], dig[g][1], dig[g][2], dig[g][3]);
      }
      if (p.dbg & (2 | 8)) continue;
      fence_async_smem();  // this thread's staging writes become visible to the TMA engine
      epi_barrier();
      if (warp == 2 && lane == 0) {
#pragma unroll
        for (int g = 0; g < GD; ++g)
          tma_store_3d(&out_map, out_s + g * (MT * 128), xg * (P1_GROUP * P1_NB / 2), ty * MT, plane * GD + g);
        tma_store_commit();
      }
    }
    if (warp == 2 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

struct Pass2Params {
  const uint8_t* digits;   // [planes][GD][h][w]
  const double* lo;        // narrow Gaussian [planes][h][w] (may be null: out = G_hi)
  double* out;             // lo - G_hi
  uint16_t* buckets;       // optional
  uint64_t* minmax;        // optional [planes][2]
  const uint8_t* band;
  const double* suffix_f;  // (double)suffix[j] * 2^-(8 JMIN)
  double scale;            // in_scale * 2^-(S + 24 - 8 JMIN)
  int h, w, r;
  int n_sel, tiles_y, tiles_x;
  PlaneSel sel;
  int dbg;
  // fused narrow Gaussian (RT > 0): half weights hw_lo[0..r_lo] (device), the raw samples' scale
  const double* hw_lo;
  int r_lo;
  double in_scale;
  uint32_t c8, c16, c24;   // as in Pass1Params
};

// explicit shared-memory accesses (the tiles' pointers lose their address space in the alignment arithmetic, and
// generic loads cost an address translation each)
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// N consecutive axis-0 results of the narrow Gaussian down one column of a raw tile, in scipy's order (lo2d_kernel's
// arithmetic, operation by operation): acc = x0*w0; acc += (x-j + x+j) * wj for j = RT .. 1.  rt_col = shared-memory
// address of the column in the raw tile, whose row 0 is global row gy0; rows outside the plane read the edge row
// (mode='nearest') when CLAMP.  Branch-free: all loads, then all conversions, then N independent chains, so that the
// float64 latencies (the chains are 4 operations deep) overlap.
template <int RT, int N, bool CLAMP>
__device__ __forceinline__ void lo_axis0_run_t(const uint32_t rt_col, const int y_first, const int gy0, const int h,
                                               const double scale, const double (&wt)[RT + 1], const uint32_t vdst) {
  uint32_t raw[N + 2 * RT];
#pragma unroll
  for (int i = 0; i < N + 2 * RT; ++i) {
    int y = y_first - RT + i;
    if (CLAMP) y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    raw[i] = lds_u16(rt_col + (uint32_t)((y - gy0) * (P2F_RAW_W * 2)));
  }
  double sm[N + 2 * RT];
#pragma unroll
  for (int i = 0; i < N + 2 * RT; ++i) sm[i] = dmul((double)raw[i], scale);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double acc = dmul(sm[n + RT], wt[0]);
#pragma unroll
    for (int j = RT; j >= 1; --j) acc = dadd(acc, dmul(dadd(sm[n + RT - j], sm[n + RT + j]), wt[j]));
    sts_f64(vdst + (uint32_t)(n * P2F_V_W * 8), acc);
  }
}
template <int RT, int N>
__device__ __forceinline__ void lo_axis0_run(const uint32_t rt_col, const int y_first, const int gy0, const int h, const bool yedge,
                                             const double scale, const double (&wt)[RT + 1], const uint32_t vdst) {
  if (yedge)
    lo_axis0_run_t<RT, N, true>(rt_col, y_first, gy0, h, scale, wt, vdst);
  else
    lo_axis0_run_t<RT, N, false>(rt_col, y_first, gy0, h, scale, wt, vdst);
}

// Axis 1 of the narrow Gaussian on the shared-memory tile of axis-0 results (vs = its address; column c of the tile <->
// element LO_HALO + c of a row), rows 4 lw .. 4 lw + 3, results written IN PLACE.  CLAMP: columns outside the plane read
// the edge column (mode='nearest').
__device__ __forceinline__ void lds_f64x2(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// The same pass for tiles whose taps all lie inside the plane, with a third of the shared-memory traffic (the MMAs'
// operand fetch and these loads share one port, and the fused kernel is bound by it): lane l owns the column PAIRS
// 2l, 2l+1 and 64 + 2l, 64 + 2l+1; the 2 + 2 RT values a pair needs come in as 1 + RT 16-byte loads (conflict-free:
// consecutive lanes, consecutive 16 bytes) and the two results leave as one 16-byte store.
template <int RT>
__device__ __forceinline__ void lo_axis1_rows_pairs(const uint32_t vs, const int lw, const int lane, const double (&wt)[RT + 1]) {
  static_assert(RT % 2 == 0, "the first tap of a pair must be 16-byte aligned");
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const uint32_t row = vs + (uint32_t)(((4 * lw + rr) * P2F_V_W + LO_HALO) * 8);
    double acc[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 2 * lane + 64 * k;
      double v[2 + 2 * RT];
#pragma unroll
      for (int i = 0; i < 1 + RT; ++i) lds_f64x2(row + (uint32_t)((c - RT + 2 * i) * 8), v[2 * i], v[2 * i + 1]);
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        double a = dmul(v[RT + o], wt[0]);
#pragma unroll
        for (int j = RT; j >= 1; --j) a = dadd(a, dmul(dadd(v[RT + o - j], v[RT + o + j]), wt[j]));
        acc[k][o] = a;
      }
    }
    __syncwarp();  // every lane has read its neighbours' columns of this row
#pragma unroll
    for (int k = 0; k < 2; ++k) sts_f64x2(row + (uint32_t)((2 * lane + 64 * k) * 8), acc[k][0], acc[k][1]);
  }
}

template <int RT, bool CLAMP>
__device__ __forceinline__ void lo_axis1_rows(const uint32_t vs, const int lw, const int lane, const int x0, const int w,
                                              const double (&wt)[RT + 1]) {
#pragma unroll
  for (int rr = 0; rr < 4; rr += 2) {
    double acc[2][4];
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const uint32_t row = vs + (uint32_t)(((4 * lw + rr + r2) * P2F_V_W + LO_HALO) * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k;
        double a = dmul(lds_f64(row + c * 8), wt[0]);
#pragma unroll
        for (int j = RT; j >= 1; --j) {
          int cl = c - j, cr = c + j;
          if (CLAMP) {
            cl = (x0 + cl < 0 ? 0 : (x0 + cl > w - 1 ? w - 1 : x0 + cl)) - x0;
            cr = (x0 + cr < 0 ? 0 : (x0 + cr > w - 1 ? w - 1 : x0 + cr)) - x0;
          }
          a = dadd(a, dmul(dadd(lds_f64(row + cl * 8), lds_f64(row + cr * 8)), wt[j]));
        }
        acc[r2][k] = a;
      }
    }
    __syncwarp();  // every lane has read its neighbours' columns of these two rows
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const uint32_t row = vs + (uint32_t)(((4 * lw + rr + r2) * P2F_V_W + LO_HALO) * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) sts_f64(row + (lane + 32 * k) * 8, acc[r2][k]);
    }
  }
}

// ------------------------------------------------------------------ pass 2: digits -> float64 DoG, axis 1
// EW epilogue warps (8 or 16): TMEM lane quarter = warp % 4, the tile's 32 rows are shared out over the EW / 4 warps of
// a quarter (RPT rows per thread).
// RT = 0: the narrow Gaussian comes from memory (lo_map: a float64 tile per stage) or not at all.
// RT > 0: FUSED.  lo_map describes the raw uint16 planes; a stage carries the tile's raw samples with a halo, the
// epilogue warps filter them along axis 0 (each thread its own column and rows, a few threads the halo columns) into a
// double-buffered shared-memory tile while the tile's MMAs run, then along axis 1 where the result is used: the
// narrow Gaussian never exists in HBM (- 8 B/px written by a kernel of its own, - 8 B/px read here, + 2 B/px).
// WS (with RT > 0): WARP-SPECIALISED fused variant.  Twenty warps: warpgroup 0 = TMA producer, MMA issuer and two halo
// warps (axis 0 of the narrow Gaussian for the 2 RT columns beside the tile); warpgroups 1-2 = the EW = 8 epilogue
// warps; warpgroups 3-4 = eight "lo warps" that do nothing but the narrow Gaussian of the NEXT tile (both axes, result
// in place in the double-buffered shared-memory tile, handed over through lo_full / lo_empty).  Registers follow the
// roles (setmaxnreg): 48 / 152 / 64 per thread, which adds up to exactly the
// 640 x 96 the CTA is launched with (an increase can only be served from what the CTA's own warps gave back).  The epilogue warps then
// run exactly the instructions of the unfused kernel while the extra arithmetic has its own issue slots.
template <int EW, int RT, bool WS = false>
__global__ void __launch_bounds__((WS ? 4 + 2 * EW : 2 + EW) * 32, 1)
tcg_axis1_kernel(const __grid_constant__ CUtensorMap dig_map, const __grid_constant__ CUtensorMap lo_map,
                 const Pass2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool FUSED = RT > 0;
  constexpr int NSTAGE = FUSED ? P2F_STAGES : P2_STAGES;
  constexpr uint32_t STAGE_BYTES = FUSED ? P2F_STAGE_BYTES : P2_STAGE_BYTES;
  constexpr uint32_t LO_TX_BYTES = FUSED ? P2F_RAW_BYTES : P2_LO_BYTES;
  double* const vbuf = reinterpret_cast<double*>(stage_s + NSTAGE * STAGE_BYTES);  // FUSED: two axis-0 result tiles
  Barriers* bars = reinterpret_cast<Barriers*>(stage_s + NSTAGE * STAGE_BYTES + (FUSED ? 2 * P2F_V_BYTES : 0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t tiles_total = (int64_t)p.n_sel * p.tiles_y * p.tiles_x;
  const int64_t per_cta = (tiles_total + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t t_end = t_begin + per_cta < tiles_total ? t_begin + per_cta : tiles_total;
  const int n_tiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;

  // a stage is free again once the MMAs have read its digit panels AND the epilogue warps its narrow-Gaussian tile
  constexpr int RPT = P2_NR / (EW / 4);  // rows per epilogue thread
  const uint32_t tmem = tcg_setup(bars, p.band, NSTAGE, 1 + EW, nullptr, p.suffix_f, EW);
  // x fastest: consecutive tiles of a CTA share half their columns
  TileWalk tw;
  tw.init(t_begin, p.tiles_y, p.tiles_x);

  auto producer = [&]() {
    prefetch_tmap(&dig_map);
    prefetch_tmap(&lo_map);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_tiles; ++it, tw.next()) {
      mbar_wait(&bars->empty[stage], phase ^ 1);
      const int plane = p.sel.phys(tw.q);
      if (elect_one()) {
        uint8_t* dst = stage_s + stage * STAGE_BYTES;
        const bool load_lo = FUSED ? !(p.dbg & 16) : p.lo != nullptr;
        mbar_expect_tx(&bars->full[stage], load_lo ? P2_DIG_BYTES + LO_TX_BYTES : P2_DIG_BYTES);
#pragma unroll
        for (int s = 0; s < GD; ++s)
#pragma unroll
          for (int pn = 0; pn < 2; ++pn)
            tma_load_3d(dst + (s * 2 + pn) * P2_PANEL_BYTES, &dig_map, &bars->full[stage], tw.fast * MT - HALO + pn * 128,
                        tw.slow * P2_NR, plane * GD + s);
        if (FUSED && load_lo)
          tma_load_3d(dst + P2_DIG_BYTES, &lo_map, &bars->full[stage], tw.fast * MT - P2F_RAW_HX, tw.slow * P2_NR - LO_HALO, plane);
        else if (!FUSED && load_lo)
          tma_load_3d(dst + P2_DIG_BYTES, &lo_map, &bars->full[stage], tw.fast * MT, tw.slow * P2_NR, plane);
      }
      __syncwarp();
      if (++stage == NSTAGE) stage = 0, phase ^= 1;
    }
  };
  auto mma_issuer = [&]() {
    constexpr uint32_t idesc = idesc_u8(MT, P2_NR, false);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_tiles; ++it) {
      mbar_wait(&bars->acc_empty, (uint32_t)(it & 1) ^ 1);
      mbar_wait(&bars->full[stage], phase);
      tc_fence_after();
      const uint32_t b_base = smem_u32(stage_s + stage * STAGE_BYTES);
      if (elect_one()) {
        if (!(p.dbg & 1)) {
          // K step outermost, then weight digit, then sample digit: consecutive MMAs go to different accumulators
#pragma unroll
          for (int ks = 0; ks < KBAND / 32; ++ks) {
#pragma unroll
            for (int d = 0; d < WD; ++d) {
#pragma unroll
              for (int s = 0; s < GD; ++s) {
                const int j = d + s;
                if (j < JMIN) continue;
                const int d_first = j - (GD - 1) > 0 ? j - (GD - 1) : 0;  // the first product that lands in accumulator j
                const uint64_t b_desc =
                    smem_desc(b_base + (s * 2 + ks / 4) * P2_PANEL_BYTES + (ks % 4) * 32, 16, 1024, LAYOUT_SW128);
                mma_u8_ts(tm + TMEM_ACC0 + (j - JMIN) * P2_NR, tm + d * TMEM_BAND_COLS_PER_DIGIT + ks * 8, b_desc, idesc,
                          (ks == 0 && d == d_first) ? 0u : 1u);
              }
            }
          }
        }
        mma_commit(&bars->empty[stage]);
        mma_commit(&bars->acc_full);
      }
      __syncwarp();
      if (++stage == NSTAGE) stage = 0, phase ^= 1;
    }
  };
  auto lo_warps = [&]() {
    // ---- lo warps: the narrow Gaussian of every tile, one tile ahead of the epilogue warps
    const int lw = warp - (4 + EW);
    const int quarter = warp & 3, hrow = lw >> 2;
    const int mx = quarter * 32 + lane;
    double wt[RT + 1];
#pragma unroll
    for (int j = 0; j <= RT; ++j) wt[j] = j <= p.r_lo ? __ldg(p.hw_lo + j) : 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_tiles; ++it, tw.next()) {
      const int tx = tw.fast, ty = tw.slow;
      const int y0 = ty * P2_NR + hrow * RPT;
      const int b = it & 1;
      double* vb = vbuf + (size_t)b * (P2F_V_BYTES / 8);
      mbar_wait(&bars->full[stage], phase);                               // the raw tile has landed
      mbar_wait(&bars->lo_empty[b], ((uint32_t)(it >> 1) & 1u) ^ 1u);     // the epilogue warps are done with this buffer
      const uint32_t rt = smem_u32(stage_s + stage * STAGE_BYTES + P2_DIG_BYTES);  // raw tile, uint16
      const uint32_t vs = smem_u32(vb);                                            // result tile, float64
      const int gy0 = ty * P2_NR - LO_HALO;
      const bool yedge = gy0 + LO_HALO - RT < 0 || gy0 + LO_HALO + P2_NR + RT > p.h;  // warp-uniform
      const bool lo_math = !(p.dbg & 32);  // timing experiments: the hand-shakes only
      // two halves of RPT / 2 rows: 8 chains in flight fit the lo warps' 64 registers
      if (lo_math) {
      lo_axis0_run<RT, RPT / 2>(rt + (mx + P2F_RAW_HX) * 2, y0, gy0, p.h, yedge, p.in_scale, wt,
                                vs + ((hrow * RPT) * P2F_V_W + mx + LO_HALO) * 8);
      lo_axis0_run<RT, RPT / 2>(rt + (mx + P2F_RAW_HX) * 2, y0 + RPT / 2, gy0, p.h, yedge, p.in_scale, wt,
                                vs + ((hrow * RPT + RPT / 2) * P2F_V_W + mx + LO_HALO) * 8);
      }
      // every column of the axis-0 results is in place: this warp's, the other lo warps' and the halo columns, which
      // the two halo warps of warpgroup 0 compute
      asm volatile("bar.sync 3, %0;" ::"n"((EW + 2) * 32) : "memory");
      if (lane == 0) mbar_arrive(&bars->empty[stage]);            // raw tile used up (bar.sync ordered the warp's reads)
      if (++stage == NSTAGE) stage = 0, phase ^= 1;
      // axis 1, ROW-partitioned: warp lw owns rows 4 lw .. 4 lw + 3 of the tile, lane l the columns l, l + 32, l + 64,
      // l + 96.  Every read and the in-place write of a row then happen inside one warp (reads, __syncwarp, writes): no
      // second CTA-level barrier, and two rows = eight independent float64 chains are in flight per thread.
      const bool xedge = tx * MT - RT < 0 || tx * MT + MT + RT > p.w;  // warp-uniform
      if (!lo_math) {
      } else if (xedge)
        lo_axis1_rows<RT, true>(vs, lw, lane, tx * MT, p.w, wt);
      else if (RT % 2 == 0)
        lo_axis1_rows_pairs<(RT % 2 == 0 ? RT : 2)>(vs, lw, lane, reinterpret_cast<const double (&)[(RT % 2 == 0 ? RT : 2) + 1]>(wt));
      else
        lo_axis1_rows<RT, false>(vs, lw, lane, tx * MT, p.w, wt);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->lo_full[b]);
    }
  };
  auto halo_warps = [&]() {
    // ---- warps 2 and 3 of warpgroup 0: axis 0 of the narrow Gaussian for the 2 RT columns left and right of the tile
    // (2 RT x 32 single results per tile, two per thread).  On the lo warps this was a latency chain for one result
    // per thread in front of their barrier, with half of them idle.
    const int t = (warp - 2) * 32 + lane;
    double wt[RT + 1];
#pragma unroll
    for (int j = 0; j <= RT; ++j) wt[j] = j <= p.r_lo ? __ldg(p.hw_lo + j) : 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_tiles; ++it, tw.next()) {
      const int tx = tw.fast, ty = tw.slow;
      (void)tx;
      const int b = it & 1;
      mbar_wait(&bars->full[stage], phase);
      mbar_wait(&bars->lo_empty[b], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      const uint32_t rt = smem_u32(stage_s + stage * STAGE_BYTES + P2_DIG_BYTES);
      const uint32_t vs = smem_u32(vbuf + (size_t)b * (P2F_V_BYTES / 8));
      const int gy0 = ty * P2_NR - LO_HALO;
      const bool yedge = gy0 + LO_HALO - RT < 0 || gy0 + LO_HALO + P2_NR + RT > p.h;  // warp-uniform
      if (!(p.dbg & 32)) {
        for (int v = t; v < 2 * RT * P2_NR; v += 64) {
          const int hc = v >> 5, row = v & 31;
          const int c = hc < RT ? hc - RT : MT + (hc - RT);
          lo_axis0_run<RT, 1>(rt + (c + P2F_RAW_HX) * 2, ty * P2_NR + row, gy0, p.h, yedge, p.in_scale, wt,
                              vs + (row * P2F_V_W + c + LO_HALO) * 8);
        }
      }
      // a blocking barrier, not bar.arrive: a halo warp that ran a tile ahead would add its arrivals to the barrier
      // generation the lo warps have not completed yet
      asm volatile("bar.sync 3, %0;" ::"n"((EW + 2) * 32) : "memory");
      if (++stage == NSTAGE) stage = 0, phase ^= 1;
    }
  };
  auto epilogue = [&]() {
    const int ew = warp - (WS ? 4 : 2);
    const int quarter = warp & 3;
    const int hrow = ew >> 2;               // which RPT of the tile's 32 rows
    const int mx = quarter * 32 + lane;     // output column inside the tile
    const int64_t hw = (int64_t)p.h * p.w;
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + TMEM_ACC0 + hrow * RPT;
    // No -0.0 can occur (lo >= +0, g >= +0, and x - x = +0), so float64 comparisons order the values as their keys do
    double vmin = __longlong_as_double(0x7ff0000000000000ll), vmax = __longlong_as_double(0xfff0000000000000ll);
    int cur_q = -1, plane = -1;
    int stage = 0;
    uint32_t phase = 0;
    const bool has_lo = FUSED || p.lo != nullptr;
    double wt[RT + 1];
#pragma unroll
    for (int j = 0; j <= RT; ++j) wt[j] = (FUSED && j <= p.r_lo) ? __ldg(p.hw_lo + j) : 0.0;
    const uint32_t c8 = p.c8, c16 = p.c16, c24 = p.c24;
    auto flush = [&]() {
      if (p.minmax != nullptr && plane >= 0) {
        const uint64_t a = warp_min_u64(f64_to_key(vmin)), b = warp_max_u64(f64_to_key(vmax));
        if (lane == 0 && a <= b) {
          atomicMin((unsigned long long*)&p.minmax[2 * plane], (unsigned long long)a);
          atomicMax((unsigned long long*)&p.minmax[2 * plane + 1], (unsigned long long)b);
        }
      }
      vmin = __longlong_as_double(0x7ff0000000000000ll), vmax = __longlong_as_double(0xfff0000000000000ll);
    };
    for (int it = 0; it < n_tiles; ++it, tw.next()) {
      if (tw.q != cur_q) {
        flush();
        cur_q = tw.q, plane = p.sel.phys(tw.q);
      }
      const int tx = tw.fast, ty = tw.slow;
      const int x = tx * MT + mx;
      const int y0 = ty * P2_NR + hrow * RPT;
      const bool x_ok = x < p.w;
      const int rows = p.h - y0 < RPT ? p.h - y0 : RPT;  // valid rows of this thread's RPT (may be <= 0)
      // the narrow operand: this thread's 16 samples of the stage's float64 tile (lanes are consecutive x: no bank
      // conflicts) are read where they are used (holding them in registers across the accumulator loads spilled);
      // the stage goes back to the producer after that
      if (!WS) mbar_wait(&bars->full[stage], phase);
      constexpr int LT_STRIDE = WS ? P2F_V_W : MT;
      const double* lt = WS ? vbuf + (size_t)(it & 1) * (P2F_V_BYTES / 8) + (hrow * RPT) * P2F_V_W + mx + LO_HALO
                            : reinterpret_cast<const double*>(stage_s + stage * STAGE_BYTES + P2_DIG_BYTES) + (hrow * RPT) * MT + mx;
      uint64_t* const stage_empty = &bars->empty[stage];
      const double* vrow = nullptr;  // FUSED: this thread's first axis-0 result (row hrow * RPT, its own column)
      if (FUSED && !WS) {
        // axis 0 of the narrow Gaussian, while this tile's MMAs run.  Raw-tile / result-tile column c <-> global column
        // tile x0 - 8 + c (result tile: x0 - 4 + c), raw-tile row k <-> global row gy0 + k.
        const uint32_t rt = smem_u32(stage_s + stage * STAGE_BYTES + P2_DIG_BYTES);
        double* vb = vbuf + (size_t)(it & 1) * (P2F_V_BYTES / 8);
        const uint32_t vs = smem_u32(vb);
        const int gy0 = ty * P2_NR - LO_HALO;
        const bool yedge = gy0 + LO_HALO - RT < 0 || gy0 + LO_HALO + P2_NR + RT > p.h;  // warp-uniform
        lo_axis0_run<RT, RPT>(rt + (mx + P2F_RAW_HX) * 2, y0, gy0, p.h, yedge, p.in_scale, wt,
                              vs + ((hrow * RPT) * P2F_V_W + mx + LO_HALO) * 8);
        // the halo columns left and right of the tile: 2 * RT columns x 32 rows, one value per thread and round
        for (int v = ew * 32 + lane; v < 2 * RT * P2_NR; v += EW * 32) {
          const int hc = v >> 5, row = v & 31;
          const int c = hc < RT ? hc - RT : MT + (hc - RT);  // column relative to the tile's first
          lo_axis0_run<RT, 1>(rt + (c + P2F_RAW_HX) * 2, ty * P2_NR + row, gy0, p.h, yedge, p.in_scale, wt,
                              vs + (row * P2F_V_W + c + LO_HALO) * 8);
        }
        asm volatile("bar.sync 2, %0;" ::"n"(EW * 32) : "memory");  // every column of this tile's axis-0 results is in place
        vrow = vb + (hrow * RPT) * P2F_V_W + mx + LO_HALO;
        __syncwarp();
        if (lane == 0) mbar_arrive(stage_empty);  // the raw tile is used up; the digit panels go back with the MMAs' commit
      }
      if (++stage == NSTAGE) stage = 0, phase ^= 1;
      // clamped-edge taps: columns left of 0 / right of w-1 all read the edge column of the same row
      const bool edge_tile = tx * MT < p.r || tx * MT + MT > p.w - p.r;  // warp-uniform
      double f_l = 0.0, f_r = 0.0, e_l = 0.0, e_r = 0.0;
      if (edge_tile) {
        if (x_ok && x < p.r) f_l = bars->suffix_f[x + 1];
        if (x_ok && x >= p.w - p.r) f_r = bars->suffix_f[p.w - x];
        if (lane < rows) {  // lane n < 16: the edge samples of row y0 + n (40-bit integers, exact in float64)
          const uint8_t* dp = p.digits + ((int64_t)plane * GD) * hw + (int64_t)(y0 + lane) * p.w;
          uint64_t a = 0, b = 0;
#pragma unroll
          for (int s = 0; s < GD; ++s) {
            a |= (uint64_t)dp[s * hw] << (8 * s);
            b |= (uint64_t)dp[s * hw + p.w - 1] << (8 * s);
          }
          e_l = (double)a, e_r = (double)b;
        }
      }

      mbar_wait(&bars->acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t v[NACC2][RPT];
      if (!(p.dbg & 4)) {
#pragma unroll
        for (int a = 0; a < NACC2; ++a) tmem_ldn<RPT>(taddr + a * P2_NR, v[a]);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int a = 0; a < NACC2; ++a)
#pragma unroll
          for (int i = 0; i < RPT; ++i) v[a][i] = (uint32_t)(it + a + i);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);
      if (p.dbg & 2) {  // timing experiments: no epilogue arithmetic / stores, the hand-shakes stay
        if (WS) {
          mbar_wait(&bars->lo_full[it & 1], (uint32_t)(it >> 1) & 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->lo_empty[it & 1]);
        } else {
          __syncwarp();
          if (!FUSED && lane == 0) mbar_arrive(stage_empty);
        }
        continue;
      }

      // Branch-free passes over the thread's 16 samples (a per-sample `if` would put every sample in its own basic
      // block and serialise sixteen independent latency chains): integer combine + ONE conversion each, then the
      // rare edge term for all of them, then scale and subtract.
      double res[RPT];
#pragma unroll
      for (int n = 0; n < RPT; ++n) {
        // sum_a acc_a * 256^a, a = 0 .. 4 (every accumulator < 2^27, the total < 2^60): three IMAD.WIDE and one
        // 64-bit addition of (acc_4 : acc_0)
        static_assert(NACC2 == 5, "the combine below is written for five accumulators");
        uint64_t tot = mad_wide(v[1][n], c8, 0ull);
        tot = mad_wide(v[2][n], c16, tot);
        tot = mad_wide(v[3][n], c24, tot);
        tot += ((uint64_t)v[4][n] << 32) | (uint64_t)v[0][n];
        res[n] = (double)tot;
      }
      if (edge_tile) {  // warp-uniform
#pragma unroll
        for (int n = 0; n < RPT; ++n) {
          const double el = __shfl_sync(0xffffffffu, e_l, n), er = __shfl_sync(0xffffffffu, e_r, n);
          res[n] += f_l * el + f_r * er;
        }
      }
      const double scale = p.scale;
      if (WS) {
        mbar_wait(&bars->lo_full[it & 1], (uint32_t)(it >> 1) & 1u);  // the lo warps have finished this tile
#pragma unroll
        for (int n = 0; n < RPT; ++n) res[n] = lt[n * LT_STRIDE] - res[n] * scale;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->lo_empty[it & 1]);
      } else if (FUSED) {
        // axis 1 of the narrow Gaussian from the shared-memory tile (columns beyond the plane read the edge column)
        const bool xedge = tx * MT - RT < 0 || tx * MT + MT + RT > p.w;  // warp-uniform
        int ol[RT + 1], orr[RT + 1];  // column offsets of the tap pairs relative to this thread's column
#pragma unroll
        for (int j = 1; j <= RT; ++j) {
          ol[j] = -j, orr[j] = j;
          if (xedge) {
            const int xl = x - j < 0 ? 0 : (x - j > p.w - 1 ? p.w - 1 : x - j);
            const int xr = x + j > p.w - 1 ? p.w - 1 : x + j;
            ol[j] = xl - x, orr[j] = xr - x;
          }
        }
        if (xedge && x > p.w - 1) {  // a column right of the plane (not stored): keep the reads inside the tile
#pragma unroll
          for (int j = 1; j <= RT; ++j) ol[j] = 0, orr[j] = 0;
        }
#pragma unroll
        for (int n = 0; n < RPT; ++n) {
          const double* vr = vrow + n * P2F_V_W;
          double acc = dmul(vr[0], wt[0]);
#pragma unroll
          for (int j = RT; j >= 1; --j) acc = dadd(acc, dmul(dadd(vr[ol[j]], vr[orr[j]]), wt[j]));
          res[n] = acc - res[n] * scale;
        }
      } else if (has_lo) {
#pragma unroll
        for (int n = 0; n < RPT; ++n) res[n] = lt[n * LT_STRIDE] - res[n] * scale;
      } else {
#pragma unroll
        for (int n = 0; n < RPT; ++n) res[n] = res[n] * scale;
      }
      if (!FUSED) {
        __syncwarp();
        if (lane == 0) mbar_arrive(stage_empty);
      }
      if (x_ok && rows > 0 && !((p.dbg & 8) && res[0] + res[RPT - 1] + res[RPT / 2] != 0.123456)) {  // dbg 8: timing without the stores
        double* op = p.out + (int64_t)plane * hw + (int64_t)y0 * p.w + x;
        uint16_t* bp = p.buckets != nullptr ? p.buckets + (int64_t)plane * hw + (int64_t)y0 * p.w + x : nullptr;
        if (rows == RPT) {  // the whole run lies inside the plane: no per-row tests
#pragma unroll
          for (int n = 0; n < RPT; ++n) op[(int64_t)n * p.w] = res[n];
          if (bp != nullptr) {
#pragma unroll
            for (int n = 0; n < RPT; ++n) bp[(int64_t)n * p.w] = (uint16_t)bucket12(res[n]);
          }
#pragma unroll
          for (int n = 0; n < RPT; ++n) {
            vmin = res[n] < vmin ? res[n] : vmin;
            vmax = res[n] > vmax ? res[n] : vmax;
          }
        } else {
#pragma unroll
          for (int n = 0; n < RPT; ++n) {
            if (n < rows) {
              op[(int64_t)n * p.w] = res[n];
              if (bp != nullptr) bp[(int64_t)n * p.w] = (uint16_t)bucket12(res[n]);
              vmin = res[n] < vmin ? res[n] : vmin;
              vmax = res[n] > vmax ? res[n] : vmax;
            }
          }
        }
      }
    }
    flush();
  };
  if constexpr (WS) {  // registers follow the roles: every warp of a warpgroup re-allocates at the top of its branch
    if (warp < 4) {
      reg_dec<48>();
      if (warp == 0) producer();
      else if (warp == 1) mma_issuer();
      else if (RT > 0) halo_warps();
    } else if (warp < 4 + EW) {
      reg_inc<152>();
      epilogue();
    } else {
      reg_dec<64>();
      lo_warps();
    }
  } else {
    if (warp == 0) producer();
    else if (warp == 1) mma_issuer();
    else epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ the narrow Gaussian (radius <= 4), float64
// Exact scipy order (axis 0 first, then axis 1; acc = x0*w0; acc += (x-j + x+j) * wj for j = r..1).
// A CTA of 128 threads walks a strip of 128 - 2R output columns down a SEGMENT of the plane (grid.y segments, so that
// a chunk fills the machine several CTAs deep) in blocks of LO_TH rows: thread t owns column x0 - R + t for the
// axis-0 pass (sliding window in registers, uint16 -> float64 on the way in; the NEXT block's LO_TH samples are
// fetched as one batch of independent loads before this block is computed, so the walk is not a chain of exposed L2
// latencies), the block's axis-0 results sit in shared memory, then thread t < 128 - 2R produces column x0 + t.
// The radius is a template parameter: the window, the halo and the tap loops have their true size.
constexpr int LO_R = 4, LO_TH = 16;

template <int R, int LO_NT>
__global__ void __launch_bounds__(LO_NT)
lo2d_kernel(const uint16_t* __restrict__ in, double* __restrict__ out, const double scale, const int h, const int w,
            const double* __restrict__ hw_lo, const int strips, const int seg_rows, const PlaneSel sel) {
  constexpr int TWO = LO_NT - 2 * R;  // output columns of a strip
  __shared__ double vs[LO_TH][LO_NT];
  __shared__ double wsm[R + 1];
  const int t = threadIdx.x;
  if (t <= R) wsm[t] = hw_lo[t];
  const int q = blockIdx.x / strips, strip = blockIdx.x - q * strips;
  const int plane = sel.phys(q);
  const int x0 = strip * TWO;
  const int y_begin = blockIdx.y * seg_rows;
  const int y_end = y_begin + seg_rows < h ? y_begin + seg_rows : h;
  int xc = x0 - R + t;  // the column this thread filters along axis 0 (clamped: mode='nearest')
  xc = xc < 0 ? 0 : (xc > w - 1 ? w - 1 : xc);
  const uint16_t* src = in + (int64_t)plane * h * w + xc;
  double* dst = out + (int64_t)plane * h * w;
  __syncthreads();
  double wt[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) wt[j] = wsm[j];
  auto raw_at = [&](int y) -> uint16_t {
    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    return __ldg(src + (int64_t)y * w);
  };
  // win[i] = sample y - R + i of the current row y; cur[yy] = raw sample yb + yy + R (the one row yy shifts in)
  double win[2 * R + 1];
#pragma unroll
  for (int i = 0; i < 2 * R; ++i) win[i + 1] = dmul((double)raw_at(y_begin + i - R), scale);
  uint16_t cur[LO_TH];
#pragma unroll
  for (int yy = 0; yy < LO_TH; ++yy) cur[yy] = raw_at(y_begin + yy + R);
  for (int yb = y_begin; yb < y_end; yb += LO_TH) {
    uint16_t nxt[LO_TH];
    if (yb + LO_TH < y_end) {
#pragma unroll
      for (int yy = 0; yy < LO_TH; ++yy) nxt[yy] = raw_at(yb + LO_TH + yy + R);
    }
#pragma unroll
    for (int yy = 0; yy < LO_TH; ++yy) {
#pragma unroll
      for (int i = 0; i < 2 * R; ++i) win[i] = win[i + 1];
      win[2 * R] = dmul((double)cur[yy], scale);
      double acc = dmul(win[R], wt[0]);
#pragma unroll
      for (int j = R; j >= 1; --j) acc = dadd(acc, dmul(dadd(win[R - j], win[R + j]), wt[j]));
      vs[yy][t] = acc;
    }
    __syncthreads();
    const int x = x0 + t;
    if (t < TWO && x < w) {
      const int rows = y_end - yb < LO_TH ? y_end - yb : LO_TH;
      if (rows == LO_TH) {
#pragma unroll
        for (int yy = 0; yy < LO_TH; ++yy) {
          const double* vr = &vs[yy][t + R];
          double acc = dmul(vr[0], wt[0]);
#pragma unroll
          for (int j = R; j >= 1; --j) acc = dadd(acc, dmul(dadd(vr[-j], vr[j]), wt[j]));
          dst[(int64_t)(yb + yy) * w + x] = acc;
        }
      } else {
        for (int yy = 0; yy < rows; ++yy) {
          const double* vr = &vs[yy][t + R];
          double acc = dmul(vr[0], wt[0]);
#pragma unroll
          for (int j = R; j >= 1; --j) acc = dadd(acc, dmul(dadd(vr[-j], vr[j]), wt[j]));
          dst[(int64_t)(yb + yy) * w + x] = acc;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < LO_TH; ++yy) cur[yy] = nxt[yy];
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

// uint16 tensor (inner, rows, planes) with a (box_inner x box_rows) box, no swizzle, out-of-bounds samples read as zero
static int make_map_u16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t planes, uint32_t box_inner,
                        uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return AMT_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {inner, rows, planes};
  cuuint64_t strides[2] = {inner * 2, inner * rows * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AMT_OK : AMT_ERR_CUDA;
}

// float64 tensor (inner, rows, planes) with a (box_inner x box_rows) box, no swizzle
static int make_map_f64(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t planes, uint32_t box_inner,
                        uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return AMT_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {inner, rows, planes};
  cuuint64_t strides[2] = {inner * 8, inner * rows * 8};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AMT_OK : AMT_ERR_CUDA;
}

}  // namespace tc
}  // namespace amt

struct amt_tcg {
  int device, r, S;
  std::vector<uint64_t>* W;  // W[t], t = 0..r (symmetric)
  uint8_t* band;             // device [WD][128][256]
  uint64_t* suffix;          // device [HALO + 2]
  double* suffix_f;          // device [HALO + 2]: suffix * 2^-(8 JMIN)
  double weight_l1_error;    // sum over the taps of |W[t] * 2^-S - w[t]|
  double dropped_bound;      // largest sum of the digit products pass 2 drops, on the [0, 1] scale of a uint16 image
};

namespace amt {
namespace tc {

int g_tcg_debug = 0;  // amt_tune "tcg_debug"
int set_suspend_ns(int ns) {  // amt_tune "tcg_suspend_ns"
  const uint32_t v = ns < 0 ? 0u : (uint32_t)ns;
  return cudaMemcpyToSymbol(c_suspend_ns, &v, sizeof(v)) == cudaSuccess ? AMT_OK : AMT_ERR_CUDA;
}
constexpr size_t P1_SMEM = 1024 + P1_STAGES * P1_STAGE_BYTES + P1_OUT_BYTES + sizeof(Barriers);
constexpr size_t P2_SMEM = 1024 + P2_STAGES * P2_STAGE_BYTES + sizeof(Barriers);
constexpr size_t P2F_SMEM = 1024 + P2F_STAGES * P2F_STAGE_BYTES + 2 * P2F_V_BYTES + sizeof(Barriers);
static_assert(P2F_SMEM <= 227 * 1024, "fused pass 2 exceeds the shared memory of an SM");

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
  p.band = g->band;
  p.suffix = g->suffix;
  p.h = (int)h;
  p.w = (int)w;
  p.r = g->r;
  p.shift = g->S - 24;
  p.n_sel = (int)n_sel;
  CUtensorMap out_map;
  AMT_TRY(make_map(&out_map, digits, (uint64_t)w, (uint64_t)h, (uint64_t)n_img * GD, 128, MT, CU_TENSOR_MAP_SWIZZLE_128B));
  p.tiles_y = (int)ceil_div(h, MT);
  p.groups_x = (int)ceil_div(w, P1_GROUP * P1_NB / 2);
  p.sel = sel;
  p.dbg = g_tcg_debug;
  p.c8 = 1u << 8, p.c16 = 1u << 16, p.c24 = 1u << 24;
  const int64_t units = n_sel * p.tiles_y * p.groups_x;
  AMT_CUDA_TRY(cudaFuncSetAttribute(tcg_axis0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P1_SMEM));
  const int grid = (int)(units < kNumSMs ? units : kNumSMs);
  tcg_axis0_kernel<<<grid, NTHREADS, P1_SMEM, st>>>(in_map, out_map, p);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

template <int EW, int RT, bool WS = false>
static int launch_axis1(int grid, size_t smem, const CUtensorMap& dig_map, const CUtensorMap& lo_map, const Pass2Params& p,
                        cudaStream_t st) {
  AMT_CUDA_TRY(cudaFuncSetAttribute(tcg_axis1_kernel<EW, RT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tcg_axis1_kernel<EW, RT, WS><<<grid, (WS ? 4 + 2 * EW : 2 + EW) * 32, smem, st>>>(dig_map, lo_map, p);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

// raw != nullptr: the narrow Gaussian (half weights hw_lo[0..r_lo] on the device, r_lo <= 4) is computed inside the kernel
// from the raw uint16 planes and `lo` is ignored; otherwise `lo` (may be null) is the narrow Gaussian in memory.
int tcg_axis1(const amt_tcg* g, const uint8_t* digits, const double* lo, double in_scale, double* out, int64_t n_img,
              int64_t h, int64_t w, uint16_t* buckets, uint64_t* minmax, PlaneSel sel, cudaStream_t st, const uint16_t* raw,
              const double* hw_lo, int r_lo) {
  if (!g || !digits || !out || n_img <= 0) return AMT_ERR_INVALID;
  if (!tcg_shape_ok(h, w) || ((uintptr_t)digits % 16)) return AMT_ERR_UNSUPPORTED;
  if (raw != nullptr && (!hw_lo || r_lo < 0 || r_lo > LO_HALO || ((uintptr_t)raw % 16))) return AMT_ERR_UNSUPPORTED;
  int64_t n_sel = 0;
  AMT_TRY(sel_count(n_img, sel, &n_sel));
  if (n_sel == 0) return AMT_OK;
  CUtensorMap dig_map;
  AMT_TRY(make_map(&dig_map, digits, (uint64_t)w, (uint64_t)h, (uint64_t)n_img * GD, 128, P2_NR, CU_TENSOR_MAP_SWIZZLE_128B));
  Pass2Params p{};
  p.digits = digits;
  p.lo = raw != nullptr ? nullptr : lo;
  p.out = out;
  p.buckets = buckets;
  p.minmax = minmax;
  p.band = g->band;
  p.suffix_f = g->suffix_f;
  p.scale = std::ldexp(in_scale, -(g->S + 24 - 8 * JMIN));
  p.h = (int)h;
  p.w = (int)w;
  p.r = g->r;
  p.n_sel = (int)n_sel;
  p.tiles_y = (int)ceil_div(h, P2_NR);
  p.tiles_x = (int)ceil_div(w, MT);
  p.sel = sel;
  p.dbg = g_tcg_debug;
  p.c8 = 1u << 8, p.c16 = 1u << 16, p.c24 = 1u << 24;
  p.hw_lo = hw_lo;
  p.r_lo = r_lo;
  p.in_scale = in_scale;
  const int64_t tiles = n_sel * p.tiles_y * p.tiles_x;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  const bool wide = (g_tcg_debug & 0x400) != 0;  // experiment: 16 epilogue warps
  CUtensorMap lo_map = dig_map;  // unused when there is no narrow operand
  if (raw != nullptr) {
    AMT_TRY(make_map_u16(&lo_map, raw, (uint64_t)w, (uint64_t)h, (uint64_t)n_img, P2F_RAW_W, P2F_RAW_H));
    if (g_tcg_debug & 0x800) {  // experiment: the epilogue warps do the narrow Gaussian themselves (measured slower)
      if (r_lo <= 2) return wide ? launch_axis1<16, 2>(grid, P2F_SMEM, dig_map, lo_map, p, st) : launch_axis1<8, 2>(grid, P2F_SMEM, dig_map, lo_map, p, st);
      return launch_axis1<8, 4>(grid, P2F_SMEM, dig_map, lo_map, p, st);
    }
    if (r_lo <= 2) return launch_axis1<8, 2, true>(grid, P2F_SMEM, dig_map, lo_map, p, st);
    return launch_axis1<8, 4, true>(grid, P2F_SMEM, dig_map, lo_map, p, st);
  }
  if (lo != nullptr) {
    if ((uintptr_t)lo % 16) return AMT_ERR_UNSUPPORTED;
    AMT_TRY(make_map_f64(&lo_map, lo, (uint64_t)w, (uint64_t)h, (uint64_t)n_img, MT, P2_NR));
  }
  return wide ? launch_axis1<16, 0>(grid, P2_SMEM, dig_map, lo_map, p, st) : launch_axis1<8, 0>(grid, P2_SMEM, dig_map, lo_map, p, st);
}

int lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w, const double* hw_lo, int r_lo,
         PlaneSel sel, cudaStream_t st) {
  if (!in || !out || !hw_lo || n_img <= 0 || h <= 0 || w <= 0) return AMT_ERR_INVALID;
  if (r_lo < 0 || r_lo > LO_R) return AMT_ERR_UNSUPPORTED;
  int64_t n_sel = 0;
  AMT_TRY(sel_count(n_img, sel, &n_sel));
  if (n_sel == 0) return AMT_OK;
  constexpr int nt = 128;  // 256 threads per CTA measured slower (0.30 vs 0.28 ms per 32 planes)
  const int strips = (int)ceil_div(w, nt - 2 * r_lo);
  if (n_sel * strips >= (1ll << 31)) return AMT_ERR_CAPACITY;
  // vertical segments: enough CTAs to fill the machine several waves deep, whole blocks of LO_TH rows each
  int segs = (int)ceil_div((int64_t)kNumSMs * 32, n_sel * strips);
  const int max_segs = (int)ceil_div(h, 4 * LO_TH);
  segs = segs < 1 ? 1 : (segs > max_segs ? max_segs : segs);
  const int seg_rows = (int)(ceil_div(ceil_div(h, segs), LO_TH) * LO_TH);
  const dim3 grid((unsigned)(n_sel * strips), (unsigned)ceil_div(h, seg_rows));
#define AMT_LO2D(R) lo2d_kernel<R, nt><<<grid, nt, 0, st>>>(in, out, in_scale, (int)h, (int)w, hw_lo, strips, seg_rows, sel)
  switch (r_lo) {
    case 0: AMT_LO2D(0); break;
    case 1: AMT_LO2D(1); break;
    case 2: AMT_LO2D(2); break;
    case 3: AMT_LO2D(3); break;
    default: AMT_LO2D(4); break;
  }
#undef AMT_LO2D
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
  for (int t = 0; t <= radius; ++t)
    g->weight_l1_error += (t == 0 ? 1.0 : 2.0) * std::fabs(std::ldexp((double)W[t], -S) - half_w_host[t]);
  // pass 2 leaves out the products (weight digit d) x (sample digit s) with d + s < JMIN: at most
  // sum_{d + s < JMIN} 256^(d + s) * 255 * sum_t digit_d(W[t]) units of 2^-(S + 24) / 65535
  for (int d = 0; d < WD; ++d) {
    double l1 = 0.0;
    for (int t = 0; t <= radius; ++t) l1 += (t == 0 ? 1.0 : 2.0) * (double)((W[t] >> (8 * d)) & 0xff);
    for (int sd = 0; sd < GD; ++sd)
      if (d + sd < JMIN) g->dropped_bound += std::ldexp(255.0 * l1, 8 * (d + sd) - (S + 24)) / 65535.0;
  }
  std::vector<uint8_t> band((size_t)WD * MT * KBAND, 0);
  for (int d = 0; d < WD; ++d)
    for (int m = 0; m < MT; ++m)
      for (int k = 0; k < KBAND; ++k) {
        const int t = std::abs(k - HALO - m);
        if (t <= radius) band[((size_t)d * MT + m) * KBAND + k] = (uint8_t)((W[t] >> (8 * d)) & 0xff);
      }
  std::vector<uint64_t> suffix(HALO + 2, 0);  // zero beyond the radius
  std::vector<double> suffix_f(HALO + 2, 0.0);
  for (int j = radius; j >= 0; --j) suffix[j] = suffix[j + 1] + W[j];
  for (int j = 0; j <= HALO + 1; ++j) suffix_f[j] = std::ldexp((double)suffix[j], -8 * JMIN);
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
  if (encode_fn() == nullptr) return fail(AMT_ERR_UNSUPPORTED);  // no cuTensorMapEncodeTiled in this driver
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

/* |G' - G| <= this, on the [0, 1] scale of img_as_float, for every sample of every image: G' = the two tensor-core
 * passes, G = the float64 Gaussian in scipy's operation order.  Terms (X = 1 bounds the scaled samples):
 *   weights rounded to integers, both passes        2 * sum_t |W[t] 2^-S - w[t]|
 *   pass-1 result rounded to 40 bits                 2^-41 (half a unit of 2^-24 / 65535 per sample, times weights summing to 1)
 *   digit products dropped in pass 2 (d + s < 3)     sum of 256^(d+s) * 255 * |digit d of W|_1, scaled: 9e-12 for sigma = 16
 *   float64 roundings of scipy's 2 x (2 r + 2) operations and of the final conversion / scaling: < 1e-13
 * plus a 1 % margin. */
double amt_tcg_error_bound(const amt_tcg* g) {
  if (!g) return -1.0;
  return 1.01 * (2.0 * g->weight_l1_error + std::ldexp(1.0, -41) + g->dropped_bound + 1e-13);
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

int amt_tcg_axis1_dog(const amt_tcg* g, const uint8_t* digits, const uint16_t* raw, const double* half_w_lo, int r_lo,
                      double in_scale, double* out, int64_t n_img, int64_t h, int64_t w, uint16_t* buckets,
                      uint64_t* minmax_keys, int skip_every, int skip_offset, amt_stream_t stream) {
  if (!raw || !half_w_lo) return AMT_ERR_INVALID;
  if (minmax_keys && n_img > 0) AMT_TRY(amt::minmax_init(minmax_keys, n_img, amt::as_stream(stream)));
  return amt::tc::tcg_axis1(g, digits, nullptr, in_scale, out, n_img, h, w, buckets, minmax_keys,
                            amt::tc::PlaneSel{skip_every, skip_offset}, amt::as_stream(stream), raw, half_w_lo, r_lo);
}

int amt_gauss_lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
                   const double* half_w_lo, int r_lo, int skip_every, int skip_offset, amt_stream_t stream) {
  return amt::tc::lo2d(in, in_scale, out, n_img, h, w, half_w_lo, r_lo, amt::tc::PlaneSel{skip_every, skip_offset},
                       amt::as_stream(stream));
}

}  // extern "C"
