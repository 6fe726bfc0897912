// Host half of the run-length staging of label masks (amt_executor_run_host, amt_rle_encode_host).
//
// A label mask handed to SegmentationMask (ref: masks.py:138; Cellpose hands int64, model.py:215) is long runs of equal
// value by nature.  Host threads turn every row into runs {value, end column (exclusive)} in pinned staging; only the
// runs cross PCIe and rle_decode_kernel (executor.cu) writes the int32 label image.  Plain C++ (no CUDA): built with
// g++ so that the row scan can carry an AVX2 clone next to the portable one (picked once at run time).
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define AMT_HAVE_AVX2_CLONE 1
#endif

#include <vector_types.h>  // uint2

#include "../../include/amt_b200.h"

namespace amt {

namespace {

// first index in [x, W) whose value differs from cur (W when none does)
template <typename T>
inline int find_change(const T* p, int x, int W, T cur) {
  while (x + 8 <= W) {
    const T d = (T)((p[x] ^ cur) | (p[x + 1] ^ cur) | (p[x + 2] ^ cur) | (p[x + 3] ^ cur) | (p[x + 4] ^ cur) | (p[x + 5] ^ cur) |
                    (p[x + 6] ^ cur) | (p[x + 7] ^ cur));
    if (d != 0) break;
    x += 8;
  }
  while (x < W && p[x] == cur) ++x;
  return x;
}

#ifdef AMT_HAVE_AVX2_CLONE
__attribute__((target("avx2"))) inline int find_change_avx2(const int64_t* p, int x, int W, int64_t cur) {
  const __m256i c = _mm256_set1_epi64x(cur);
  while (x + 8 <= W) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + x));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + x + 4));
    const unsigned m = (unsigned)_mm256_movemask_pd(_mm256_castsi256_pd(_mm256_cmpeq_epi64(a, c))) |
                       ((unsigned)_mm256_movemask_pd(_mm256_castsi256_pd(_mm256_cmpeq_epi64(b, c))) << 4);
    if (m != 0xffu) return x + __builtin_ctz(~m);
    x += 8;
  }
  while (x < W && p[x] == cur) ++x;
  return x;
}
__attribute__((target("avx2"))) inline int find_change_avx2(const int32_t* p, int x, int W, int32_t cur) {
  const __m256i c = _mm256_set1_epi32(cur);
  while (x + 8 <= W) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + x));
    const unsigned m = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(a, c)));
    if (m != 0xffu) return x + __builtin_ctz(~m);
    x += 8;
  }
  while (x < W && p[x] == cur) ++x;
  return x;
}
__attribute__((target("avx2"))) inline int find_change_avx2(const uint16_t* p, int x, int W, uint16_t cur) {
  const __m256i c = _mm256_set1_epi16((short)cur);
  while (x + 16 <= W) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + x));
    const unsigned m = (unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi16(a, c));  // two bits per pixel
    if (m != 0xffffffffu) return x + (__builtin_ctz(~m) >> 1);
    x += 16;
  }
  while (x < W && p[x] == cur) ++x;
  return x;
}
#endif

template <typename T>
inline uint32_t run_value(T v, bool& neg) {
  if (sizeof(T) == 2) return (uint32_t)(uint16_t)v;
  const int64_t w = (int64_t)v;
  if (w < 0) {
    neg = true;
    return 0u;  // negative labels are background and raise the FOV's flag, as on the device routes
  }
  return w > 2147483647ll ? 2147483647u : (uint32_t)w;  // beyond int32: out of range for the labelling pass
}

// rows [r_lo, r_hi) of the chunk -> runs packed from slot `base` on; returns the number of runs, or -1 when they do not fit
#define AMT_RLE_ROWS_BODY(FIND)                                                                              \
  uint2* o = runs + base;                                                                                    \
  int64_t n = 0;                                                                                             \
  for (int64_t row = r_lo; row < r_hi; ++row) {                                                              \
    const T* p = in + row * W;                                                                               \
    const int64_t n0 = n;                                                                                    \
    bool neg = false;                                                                                        \
    if (cap - n >= W) { /* room for the worst case of this row: no test per run */                           \
      T cur = p[0];                                                                                          \
      int x = 1;                                                                                             \
      for (;;) {                                                                                             \
        x = FIND(p, x, W, cur);                                                                              \
        o[n].x = run_value(cur, neg), o[n].y = (uint32_t)x, ++n;                                             \
        if (x >= W) break;                                                                                   \
        cur = p[x++];                                                                                        \
      }                                                                                                      \
    } else {                                                                                                 \
      T cur = p[0];                                                                                          \
      int x = 1;                                                                                             \
      for (;;) {                                                                                             \
        x = FIND(p, x, W, cur);                                                                              \
        if (n >= cap) return -1;                                                                             \
        o[n].x = run_value(cur, neg), o[n].y = (uint32_t)x, ++n;                                             \
        if (x >= W) break;                                                                                   \
        cur = p[x++];                                                                                        \
      }                                                                                                      \
    }                                                                                                        \
    rows[row].x = (uint32_t)(base + n0), rows[row].y = (uint32_t)(n - n0);                                   \
    if (neg && negative != nullptr) __atomic_store_n(&negative[row / H], 1, __ATOMIC_RELAXED);               \
  }                                                                                                          \
  return n;

template <typename T>
int64_t encode_rows_portable(const T* in, int H, int W, int64_t r_lo, int64_t r_hi, uint2* runs, int64_t base, int64_t cap, uint2* rows,
                             int32_t* negative) {
  AMT_RLE_ROWS_BODY(find_change)
}

#ifdef AMT_HAVE_AVX2_CLONE
template <typename T>
__attribute__((target("avx2"))) int64_t encode_rows_avx2(const T* in, int H, int W, int64_t r_lo, int64_t r_hi, uint2* runs, int64_t base,
                                                         int64_t cap, uint2* rows, int32_t* negative) {
  AMT_RLE_ROWS_BODY(find_change_avx2)
}
#endif

template <typename T>
bool encode(const T* in, int H, int W, int g, uint2* runs, uint2* rows, int32_t* negative, int n_thr, std::vector<int64_t>& first,
            std::vector<int64_t>& used) {
  const int64_t R = (int64_t)g * H;
  const int64_t per_row = W / 4;
  if (n_thr > R) n_thr = (int)R;
  if (n_thr < 1 || R * W < (1 << 20)) n_thr = 1;
  first.assign(n_thr, 0), used.assign(n_thr, 0);
#ifdef AMT_HAVE_AVX2_CLONE
  static const bool avx2 = __builtin_cpu_supports("avx2");
#endif
  auto work = [&](int t) {
    const int64_t r_lo = R * t / n_thr, r_hi = R * (t + 1) / n_thr;
    first[t] = r_lo * per_row;
#ifdef AMT_HAVE_AVX2_CLONE
    if (avx2) {
      used[t] = encode_rows_avx2(in, H, W, r_lo, r_hi, runs, first[t], (r_hi - r_lo) * per_row, rows, negative);
      return;
    }
#endif
    used[t] = encode_rows_portable(in, H, W, r_lo, r_hi, runs, first[t], (r_hi - r_lo) * per_row, rows, negative);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (int t = 0; t < n_thr; ++t)
    if (used[t] < 0) return false;
  return true;
}

}  // namespace

// Thread t takes the rows [R t / T, R (t + 1) / T) of the chunk (R = g * H) and packs their runs from run slot
// first[t] = row_lo * (W / 4) on; used[t] = the runs it wrote.  rows[r] = {first run slot, number of runs} of row r.
// Returns false when a thread's runs did not fit its rows' slots (the caller then sends the chunk as the plain mask).
bool rle_encode_host(const void* in, int dtype, int H, int W, int g, uint2* runs, uint2* rows, int32_t* negative, int n_thr,
                     std::vector<int64_t>& first, std::vector<int64_t>& used) {
  if (dtype == AMT_I64) return encode((const int64_t*)in, H, W, g, runs, rows, negative, n_thr, first, used);
  if (dtype == AMT_U16) return encode((const uint16_t*)in, H, W, g, runs, rows, negative, n_thr, first, used);
  return encode((const int32_t*)in, H, W, g, runs, rows, negative, n_thr, first, used);
}

}  // namespace amt
