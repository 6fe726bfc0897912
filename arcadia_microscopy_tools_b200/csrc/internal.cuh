// Cross-file launch helpers used by the fused executor (strided variants of the C-ABI calls).
#pragma once

#include "common.cuh"

namespace amt {

int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st);

int dog2d(const void* in, int in_dtype, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
          const double* hw_lo, int r_lo, const double* hw_hi, int r_hi, double* tmp_lo, double* tmp_hi,
          uint64_t* minmax, cudaStream_t st, uint16_t* buckets = nullptr, bool* buckets_written = nullptr,
          int exact_every = 0, int exact_offset = 0,  // exact_every > 0: only planes p % every == offset keep scipy's order
          bool only_exact = false);                   // ... and only those planes are computed at all (tcgauss.cu has the rest)
bool dog2d_fast(int in_dtype, int64_t n_img, int64_t h, int64_t w, int r_lo, int r_hi);

// order statistics of float64 planes whose bucket12() values were written next to them by dog2d
int select_f64_bucketed(const double* data, const uint16_t* buckets, int64_t n_img, int64_t n, const int64_t* ranks_host,
                        int n_ranks, const uint64_t* minmax_keys, double* out_vals, void* scratch, size_t scratch_bytes,
                        cudaStream_t st);

// hist256 (optional): plane i gets a histogram iff i % hist_every == hist_offset, stored at
// hist256[(i / hist_every) * 256].
int map_launch(const void* in, int in_dtype, double* out, int64_t n_img, int64_t n, const amt_map_params* params,
               uint32_t* hist256, int hist_every, int hist_offset, cudaStream_t st,
               // decision-exact mode (decide.cu): samples of the histogram planes within the propagated cand_eps of a bin
               // edge / centre are listed (cand_count[hist plane], cand_idx[hist plane * cand_cap + k])
               double cand_eps = 0.0, uint32_t* cand_count = nullptr, uint32_t* cand_idx = nullptr, int cand_cap = 0);
int dx_patch(const uint32_t* cand_count, const uint32_t* cand_idx, const double* exact_in, int cap, const amt_map_params* params,
             int hist_every, int hist_offset, int64_t n_hist, double* out, int64_t n, uint32_t* hist256, int32_t* retry,
             cudaStream_t st);
namespace dx {
// exact order statistics (6 percentile ranks + min + max) of n_img planes whose approximate plane `dog` is within eps
// of the exact difference of Gaussians of `in`; stats / mm are read (approximate) and overwritten (exact)
int rank_exact(const uint16_t* in, int64_t in_img_stride, const double* dog, const uint16_t* buckets /* or null */,
               int64_t dog_img_stride, int64_t n_img, int h, int w, double scale, const double* hw_hi, int r_hi,
               const double* hw_lo, int r_lo, double eps,
               const int64_t* ranks_dev, double* stats, int64_t stats_stride, uint64_t* mm, int64_t mm_stride,
               uint32_t* scratch_u32, double* scratch_val, int cap, int32_t* retry, cudaStream_t st);
// exact difference of Gaussians of the listed pixels: list l = image l, count[l] entries (capped at cap)
int exact_eval(const uint16_t* in, int64_t in_img_stride, int h, int w, double scale, const double* hw_hi, int r_hi,
               const double* hw_lo, int r_lo, const uint32_t* count, const uint32_t* idx, double* val, int n_lists, int cap,
               cudaStream_t st);
}  // namespace dx

int plan_dog_rescale(const double* stats, const uint64_t* mm, int64_t n_img, double g_bg, double g_lo, double g_hi,
                     double o1, double o2, amt_map_params* params, cudaStream_t st);

// mode 0 reads params[img * pstride + poffset]
int otsu_launch(const uint32_t* hist, int mode, const amt_map_params* params, int64_t pstride, int64_t poffset,
                const uint64_t* mm, int64_t n_img, double* thresholds, void* scratch, size_t scratch_bytes,
                cudaStream_t st);

// in_stride: elements between consecutive input images
int label_launch(const void* in, int in_kind, int64_t in_stride, const double* thresholds, int64_t max_value,
                 int64_t n_img, int64_t h, int64_t w, int clear_border, int32_t* labels_out, int32_t* counts,
                 void* scratch, size_t scratch_bytes, cudaStream_t st,
                 int32_t* value_overflow = nullptr);  // integer masks: [img] = 1 if a value > max_value was seen (zeroed by the caller)

int region_reduce(const int32_t* labels, const uint16_t* channels, int n_channels, int64_t img_stride, int64_t chan_stride,
                  int64_t n_img, int64_t h, int64_t w, int64_t max_labels, uint64_t* acc, cudaStream_t st);
int region_finalize(const uint64_t* acc, const int32_t* counts, int n_channels, int64_t n_img, int64_t max_labels,
                    double* table, cudaStream_t st);

}  // namespace amt

struct amt_tcg;
namespace amt {
namespace tc {
struct PlaneSel {
  int every, skip;  // every > 0: planes p with p % every == skip are left out
  __host__ __device__ int phys(int q) const {
    if (every <= 0) return q;
    const int grp = q / (every - 1), c = q - grp * (every - 1);
    return grp * every + c + (c >= skip ? 1 : 0);
  }
};
bool tcg_shape_ok(int64_t h, int64_t w);
int tcg_axis0(const amt_tcg* g, const uint16_t* in, int64_t n_img, int64_t h, int64_t w, uint8_t* digits, PlaneSel sel,
              cudaStream_t st);
int tcg_axis1(const amt_tcg* g, const uint8_t* digits, const double* lo, double in_scale, double* out, int64_t n_img,
              int64_t h, int64_t w, uint16_t* buckets, uint64_t* minmax, PlaneSel sel, cudaStream_t st,
              const uint16_t* raw = nullptr, const double* hw_lo = nullptr, int r_lo = 0);
int lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w, const double* hw_lo, int r_lo,
         PlaneSel sel, cudaStream_t st);
}  // namespace tc
size_t region_shape_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_labels);
int region_shape(const int32_t* labels, const uint64_t* acc, int n_channels, const int32_t* counts, int64_t n_img,
                 int64_t h, int64_t w, int64_t max_labels, double* table, void* scratch, size_t scratch_bytes,
                 cudaStream_t st);
}  // namespace amt
