/*
 * amt_b200.h — C ABI of libamt_b200.so: the B200 (sm_100a) implementation of the
 * arcadia-microscopy-tools per-image hot path (preprocess -> threshold/label -> quantify).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless its name ends in `_host`.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); nothing
 *    synchronises except the `*_host` executor entry points, which say so.
 *  - Return value: AMT_OK (0) or a negative amt_status.  No exceptions, no exit().
 *  - The library is stateless apart from amt_executor handles: the caller owns all
 *    buffers including scratch (sizes come from the *_scratch_bytes helpers), so every
 *    entry point is re-entrant and may be called concurrently from many host threads
 *    (the reference fans one operation out over a ThreadPoolExecutor,
 *    ref: src/arcadia_microscopy_tools/pipeline.py:145-146).
 *  - Images are C-contiguous planes, batched along a leading `n_img` axis.
 *
 * Each entry point names the reference interface it stands in for (paths relative to
 * /root/reference/src/arcadia_microscopy_tools/; [3p] = the un-vendored scikit-image /
 * scipy / numpy routine that reference line dispatches to).
 */
#ifndef AMT_B200_H
#define AMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* amt_stream_t; /* cudaStream_t */

typedef enum amt_status {
  AMT_OK = 0,
  AMT_ERR_INVALID = -1,     /* bad argument (maps to ValueError / TypeError in the shim) */
  AMT_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
  AMT_ERR_CAPACITY = -3,    /* a caller-provided capacity (labels, scratch, radius) is too small */
  AMT_ERR_UNSUPPORTED = -4  /* dtype / rank combination outside the hot path */
} amt_status;

typedef enum amt_dtype {
  AMT_U8 = 0,  /* also bool masks (0/1) */
  AMT_U16 = 1,
  AMT_I32 = 2,
  AMT_F64 = 3,
  AMT_I64 = 4  /* host label masks of amt_executor_run_host only (the reference's mask dtype, masks.py:138) */
} amt_dtype;

int amt_version(void);
const char* amt_strerror(int status);
/* Last CUDA error string seen by this host thread (diagnostics only). */
const char* amt_last_cuda_error(void);
/* Number of kernels this library has launched in the calling process (bench gpu_launches). */
uint64_t amt_launch_count(void);
/* FP64 issue-rate probe (non-FMA DMUL + DADD chains, the Gaussian's instruction mix without
 * memory traffic): bench.py times it to get the measured DP-pipe roofline.  scratch: at least
 * 148*8*256 doubles.  *dp_instructions = thread-level DP instructions of one launch. */
int amt_fp64_probe(int iters, double* scratch, uint64_t* dp_instructions, amt_stream_t stream);

/* ------------------------------------------------------------------ Gaussian / DoG
 * ref: operations.py:91  ski.filters.difference_of_gaussians -> [3p] scipy.ndimage.
 * gaussian_filter(mode='nearest', truncate=4.0) -> correlate1d symmetric loop.  float64,
 * scipy's summation order, no FMA contraction: bit-identical to scipy.
 *
 * `half_w` holds weights[c-j] for j = 0..radius (the centre and one side of the symmetric
 * kernel, computed on the host with NumPy exactly as scipy's _gaussian_kernel1d does).
 */

/* One 1-D pass along the middle axis of a C-contiguous (outer, n, inner) array.
 * inner == 1 filters along the contiguous axis.  in_dtype AMT_U16 (values multiplied by
 * in_scale, i.e. img_as_float's 1/65535) or AMT_F64 (in_scale ignored). */
int amt_gaussian_axis(const void* in, int in_dtype, double in_scale, double* out,
                      int64_t outer, int64_t n, int64_t inner,
                      const double* half_w, int radius, amt_stream_t stream);
/* The same pass with the boundary extension named: AMT_EXTEND_NEAREST (scipy mode='nearest', what
 * amt_gaussian_axis does) or AMT_EXTEND_REFLECT (scipy mode='reflect', the default of
 * skimage.filters.threshold_local: operations.py:193). */
#define AMT_EXTEND_NEAREST 0
#define AMT_EXTEND_REFLECT 1
int amt_gaussian_axis_mode(const void* in, int in_dtype, double in_scale, double* out,
                           int64_t outer, int64_t n, int64_t inner,
                           const double* half_w, int radius, int mode, amt_stream_t stream);

/* Fused 2-D difference of Gaussians over a batch of planes: out = G_lo(x) - G_hi(x).
 * tmp_lo / tmp_hi: caller scratch, n_img*h*w doubles each (axis-0 pass results).
 * minmax_keys (optional, may be NULL): n_img*2 uint64, receives the order-preserving keys
 * of min and max of each output plane (see amt_minmax_f64). */
int amt_dog2d(const void* in, int in_dtype, double in_scale, double* out,
              int64_t n_img, int64_t h, int64_t w,
              const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi,
              double* tmp_lo, double* tmp_hi, uint64_t* minmax_keys, amt_stream_t stream);

/* The two kernels of amt_dog2d, callable on their own (per-kernel timing, custom staging):
 * axis 0 pass of both filters (input -> tmp_lo, tmp_hi), then axis 1 pass + subtraction
 * (+ min/max keys).  The layout of tmp_lo / tmp_hi between the two calls is private to the
 * library (transposed planes on the fast path): call both with the same shapes and radii. */
int amt_dog2d_axis0(const void* in, int in_dtype, double in_scale, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi,
                    double* tmp_lo, double* tmp_hi, amt_stream_t stream);
int amt_dog2d_axis1(const double* tmp_lo, const double* tmp_hi, double* out, int64_t n_img, int64_t h, int64_t w,
                    const double* half_w_lo, int r_lo, const double* half_w_hi, int r_hi,
                    uint64_t* minmax_keys, amt_stream_t stream);

/* ------------------------------------------------------------------ tensor-core Gaussian (tcgen05 + TMA)
 * ref: operations.py:91 — the sigma_high Gaussian of difference_of_gaussians for planes nothing discrete is
 * derived from (every channel except the thresholded one): a 1-D pass is a banded Toeplitz product on
 * tcgen05.mma kind::i8 with the weights rounded to 32 significant bits (base-256 digits) and exact integer
 * accumulation, operands staged by TMA (csrc/tcgauss.cu).  Results equal scipy's to ~1e-10 of the [0, 1]
 * scale (the reference tolerance for filtered planes is 1e-5); they are NOT bit-identical, so the
 * thresholded channel keeps amt_dog2d.
 *
 * amt_tcg_create: half_w_host = weights[c-j], j = 0..radius (radius <= 64, i.e. sigma <= 16 at truncate 4).
 * amt_tcg_weights: the integer weights W[j] (host array of radius+1) and S with w[j] ~ W[j] * 2^-S,
 *   sum_j W == 2^S exactly (test hook: lets a CPU checker restate the integer pipeline bit for bit).
 * amt_tcg_supported: 1 if (h, w, radius) can take this path (h, w >= 128, w % 16 == 0).
 * amt_tcg_axis0: uint16 planes -> `digits`, five uint8 planes per image ([n_img][5][h][w]) holding
 *   round(sum_t W[t] * x[y+t] / 2^(S-24)) (40 bits), rows clamped at the edges (mode='nearest').
 * amt_tcg_axis1: digits -> out = lo - G_hi (lo = the narrow Gaussian, may be NULL: out = G_hi), G_hi in
 *   units of in_scale * input; optional bucket12 codes (uint16 per sample) and min / max keys.
 * amt_gauss_lo2d: the narrow Gaussian (radius <= 4) of uint16 planes in float64, scipy's order (bit-identical).
 * skip_every > 0: planes p with p % skip_every == skip_offset are left untouched (the executor's
 *   segmentation channel); n_img must then be a multiple of skip_every. */
typedef struct amt_tcg amt_tcg;
int amt_tcg_create(const double* half_w_host, int radius, int device, amt_tcg** out);
void amt_tcg_destroy(amt_tcg* g);
int amt_tcg_weights(const amt_tcg* g, uint64_t* w_host, int* scale_bits);
int amt_tcg_supported(int64_t h, int64_t w, int radius);
/* Proven bound on |tensor-core Gaussian - float64 Gaussian in scipy's order| per sample, [0, 1] input scale
 * (see csrc/tcgauss.cu); what the decision-exact mode of the executor widens its comparisons by. */
double amt_tcg_error_bound(const amt_tcg* g);
size_t amt_tcg_digit_bytes(int64_t n_img, int64_t h, int64_t w);
int amt_tcg_axis0(const amt_tcg* g, const uint16_t* in, int64_t n_img, int64_t h, int64_t w, uint8_t* digits,
                  int skip_every, int skip_offset, amt_stream_t stream);
int amt_tcg_axis1(const amt_tcg* g, const uint8_t* digits, const double* lo, double in_scale, double* out,
                  int64_t n_img, int64_t h, int64_t w, uint16_t* buckets, uint64_t* minmax_keys,
                  int skip_every, int skip_offset, amt_stream_t stream);
/* amt_tcg_axis1 with the narrow Gaussian fused in (the executor's path): out = G_lo(raw) - G_hi, where G_lo (half weights
 * half_w_lo[0..r_lo] on the device, r_lo <= 4) is computed inside the kernel from the raw uint16 planes in scipy's order
 * (bit-identical to amt_gauss_lo2d followed by amt_tcg_axis1; the narrow Gaussian never exists in HBM). */
int amt_tcg_axis1_dog(const amt_tcg* g, const uint8_t* digits, const uint16_t* raw, const double* half_w_lo, int r_lo,
                      double in_scale, double* out, int64_t n_img, int64_t h, int64_t w, uint16_t* buckets,
                      uint64_t* minmax_keys, int skip_every, int skip_offset, amt_stream_t stream);
int amt_gauss_lo2d(const uint16_t* in, double in_scale, double* out, int64_t n_img, int64_t h, int64_t w,
                   const double* half_w_lo, int r_lo, int skip_every, int skip_offset, amt_stream_t stream);

/* Tuning knobs (process-wide; set before launching work; bench / profiling only — defaults are
 * the shipped configuration).  Keys: "dog_variant" 0 = 8 warps x 8 outputs per thread,
 * 1 = 4 warps x 16, 2 = 8 warps x 16; "dog_solo" 1 = one DoG CTA per SM (leaves half of the SM
 * to the HBM-bound kernels of the other stream), 0 = as many as fit; "dog_generic" 1 = force
 * the generic tile kernels; "dog_fma" 1 = contract the DoG's multiply-adds (2 instead of 3 DP
 * instructions per tap pair; the filtered planes then differ from scipy's in the last bits: opt-in,
 * off by default); "tcg_debug" switches parts of the tensor-core Gaussian off for timing experiments
 * (csrc/tcgauss.cu); "tcg_suspend_ns" = suspend-time hint of the tensor-core kernels' mbarrier waits (default 20000;
 * 0 = plain polling). */
int amt_tune(const char* key, int value);

/* Pinned host staging without torch (ref: nikon.py:25-43 / leica.py:52-80 decode into host arrays; the
 * north star feeds the device from pinned, double-buffered staging).  write_combined = 1: write-combined
 * pages for buffers the host only fills.  Free with amt_host_free. */
int amt_host_alloc(size_t bytes, int write_combined, void** out);
int amt_host_free(void* ptr);

/* Test hook: *mismatches = number of i for which the plane-constant division sequence of the map
 * kernel (reciprocal + two FMA corrections) differs bitwise from __ddiv_rn(a[i], b[i]). */
int amt_selftest_div(const double* a, const double* b, int64_t n, uint64_t* mismatches, amt_stream_t stream);

/* Flat grey-scale erosion (is_max = 0) / dilation (is_max = 1) along one axis of an (outer, n, inner)
 * array: out[i] = min / max of in over the window [i - left, i - left + size), mode='reflect'
 * (= scipy.ndimage.minimum_filter1d / maximum_filter1d with origin = left - size/2).  dtype AMT_U16 or
 * AMT_F64, out has the same dtype; minuend (optional, same dtype): out = minuend - result, the final
 * step of scipy.ndimage.white_tophat.  Extension: the reference has no morphology op. */
int amt_minmax_filter_axis(const void* in, int dtype, void* out, const void* minuend, int64_t outer, int64_t n,
                           int64_t inner, int size, int left, int is_max, amt_stream_t stream);

/* Pixel-interleaved frames (n_frames, n_pix, C) uint16 -> channel planes (n_frames, C, n_pix): the
 * layout step of nd2.ND2File.asarray (nikon.py:25-43) for raw ND2 "ImageDataSeq" payloads that were
 * memcpy'd to the device unchanged. */
int amt_deinterleave_u16(const uint16_t* in_yxc, uint16_t* out_cyx, int64_t n_frames, int64_t n_pix, int n_channels,
                         amt_stream_t stream);

/* out = a - b elementwise (N-D DoG fallback: two full Gaussians then subtract). */
int amt_sub_f64(const double* a, const double* b, double* out, int64_t n, amt_stream_t stream);

/* ------------------------------------------------------------------ min / max
 * ref: operations.py:43, :201 (intensities.min() == intensities.max() guards).
 * minmax_keys: n_img*2 uint64 = {key(min), key(max)}; key(x) is the order-preserving map
 * bits ^ (sign ? ~0 : 1<<63) for float64, the value itself for uint16.
 * amt_minmax_decode writes n_img*2 doubles {min, max}. */
int amt_minmax_f64(const double* data, int64_t n_img, int64_t n, uint64_t* minmax_keys, amt_stream_t stream);
int amt_minmax_u16(const uint16_t* data, int64_t n_img, int64_t n, uint64_t* minmax_keys, amt_stream_t stream);
int amt_minmax_decode(const uint64_t* minmax_keys, int is_f64, int64_t n_img, double* out_minmax, amt_stream_t stream);

/* ------------------------------------------------------------------ exact order statistics
 * ref: operations.py:47, :94  np.percentile (method 'linear') needs the floor/ceil order
 * statistics of the flattened plane.  `ranks_host`: n_ranks (<= AMT_MAX_RANKS) zero-based
 * ranks, the same for every plane.  out_vals: n_img*n_ranks doubles.
 * f64: needs minmax_keys of the data (amt_minmax_f64 / amt_dog2d) and scratch of
 * amt_select_f64_scratch_bytes(n_img, n).  u16: scratch of amt_select_u16_scratch_bytes. */
#define AMT_MAX_RANKS 8
size_t amt_select_f64_scratch_bytes(int64_t n_img, int64_t n);
int amt_select_f64(const double* data, int64_t n_img, int64_t n, const int64_t* ranks_host, int n_ranks,
                   const uint64_t* minmax_keys, double* out_vals, void* scratch, size_t scratch_bytes,
                   amt_stream_t stream);
size_t amt_select_u16_scratch_bytes(int64_t n_img);
/* Bucketed variant of amt_select_f64 (same scratch, same results): `buckets` holds amt_bucket12() of every
 * sample (12-bit monotone bucket: sign, float32 exponent and top 6 mantissa bits).  The executor's DoG writes
 * the buckets in its second pass, so the two selection passes read 2 B instead of 8 B per sample.
 * n must be a multiple of 8, both planes 16-byte aligned. */
int amt_bucket12(const double* data, uint16_t* buckets, int64_t n, amt_stream_t stream);
int amt_select_f64_bucketed(const double* data, const uint16_t* buckets, int64_t n_img, int64_t n,
                            const int64_t* ranks_host, int n_ranks, const uint64_t* minmax_keys, double* out_vals,
                            void* scratch, size_t scratch_bytes, amt_stream_t stream);
int amt_select_u16(const uint16_t* data, int64_t n_img, int64_t n, const int64_t* ranks_host, int n_ranks,
                   double* out_vals, void* scratch, size_t scratch_bytes, amt_stream_t stream);

/* ------------------------------------------------------------------ elementwise maps
 * Per-plane parameters live in device memory so that they can be produced by device code
 * (amt_plan_dog_rescale) without a host round trip.
 *   flags bit0 SUBCLIP : y = max(x - lvl, 0)                   ref: operations.py:97
 *   flags bit1 RESCALE : y = ((clip(y,p1,p2) - p1)/(p2 - p1))*(o2-o1) + o1, or
 *                        clip(y,o1,o2) when p1 == p2           ref: operations.py:50-54 [3p]
 *   flags bit2 FILL    : y = o1 (constant plane)               ref: operations.py:43-44
 */
typedef struct amt_map_params {
  double lvl, p1, p2, o1, o2;
  double hist_first, hist_last; /* range of the fused 256-bin histogram (written by plan) */
  int32_t flags;
  int32_t pad;
} amt_map_params;
#define AMT_MAP_SUBCLIP 1
#define AMT_MAP_RESCALE 2
#define AMT_MAP_FILL 4

/* in_dtype AMT_U16 or AMT_F64; out float64.  hist256 (optional): n_img*256 uint32 counts of
 * the OUTPUT plane with numpy's uniform-bin semantics over [hist_first, hist_last]
 * (edges = np.linspace restated on device); must be zeroed by the caller. */
int amt_map(const void* in, int in_dtype, double* out, int64_t n_img, int64_t n,
            const amt_map_params* params, uint32_t* hist256, amt_stream_t stream);

/* Device-side planning for subtract_background_dog -> rescale_by_percentile on a DoG plane:
 * from 6 order statistics per plane (ranks lo/hi of `percentile`, of p_lo and of p_hi, as
 * produced by amt_select_f64) and the host-computed lerp fractions, fill amt_map_params
 * (SUBCLIP|RESCALE, or FILL for a constant plane) exactly as NumPy would.
 * ref: operations.py:94-97 then operations.py:41-54. */
int amt_plan_dog_rescale(const double* order_stats /* n_img*6 */, const uint64_t* minmax_keys,
                         int64_t n_img, double g_bg, double g_lo, double g_hi, double o1, double o2,
                         amt_map_params* params, amt_stream_t stream);

/* ------------------------------------------------------------------ histogram + Otsu
 * ref: operations.py:186/:214 ski.filters.threshold_otsu [3p]: float images -> 256 uniform
 * bins over [min,max]; integer images -> one bin per value over [min,max]; float32 counts
 * and class weights, float64 class means, first maximum.  thresholds: n_img doubles. */
int amt_hist256_f64(const double* data, int64_t n_img, int64_t n, const uint64_t* minmax_keys,
                    uint32_t* hist256, amt_stream_t stream);
int amt_hist_u16(const uint16_t* data, int64_t n_img, int64_t n, uint32_t* hist65536, amt_stream_t stream);
/* np.histogram(plane, nbins, range=(min, max)) of float64 planes for ANY nbins (threshold_*'s `nbins` argument,
 * operations.py:214 forwards it): edges = n_img * (nbins + 1) doubles, np.linspace(min, max, nbins + 1) computed
 * by the host; uniform-formula candidate corrected against the edges exactly as NumPy does. */
int amt_hist_f64(const double* data, int64_t n_img, int64_t n, const double* edges, int nbins, uint32_t* hist,
                 amt_stream_t stream);
/* np.sum of every contiguous float64 plane in NumPy's PAIRWISE order (bit-identical; np.mean = sum / n):
 * threshold_mean on float images (operations.py:191).  scratch: amt_pairwise_sum_scratch_bytes. */
size_t amt_pairwise_sum_scratch_bytes(int64_t n_img, int64_t n);
int amt_pairwise_sum_f64(const double* data, int64_t n_img, int64_t n, double* sums, void* scratch, size_t scratch_bytes,
                         amt_stream_t stream);
/* mode 0: float histogram, 256 bins, range from params[i].hist_first/last;
 * mode 1: float histogram, 256 bins, range from minmax_keys;
 * mode 2: uint16 exact histogram (65536 bins), range from minmax_keys. */
size_t amt_otsu_scratch_bytes(int mode, int64_t n_img); /* 0 for the 256-bin modes */
int amt_otsu(const uint32_t* hist, int mode, const amt_map_params* params, const uint64_t* minmax_keys,
             int64_t n_img, double* thresholds, void* scratch, size_t scratch_bytes, amt_stream_t stream);
/* mask = data > threshold[img]  (uint8 0/1).  ref: operations.py:216 */
int amt_threshold_gt(const void* data, int in_dtype, int64_t n_img, int64_t n, const double* thresholds,
                     uint8_t* mask, amt_stream_t stream);
/* Local-window thresholds (ref: operations.py:193-195 -> skimage threshold_local / _niblack / _sauvola [3p]).
 * amt_window_threshold_u16: per pixel, mean m and standard deviation s of the window_h x window_w box
 * (odd sizes <= 127, np.pad 'reflect' borders, exact integer window sums); kind 0 = niblack
 * t = m - k*s, kind 1 = sauvola t = m*(1 + k*(s/r - 1)); mask = data > t; thresholds (optional,
 * may be NULL) receives t.  Exact as long as 4*sum(data^2) < 2^53 per image (the caller checks).
 * amt_threshold_gt_image: mask[i] = data[i] > thresholds[i] - offset (threshold_local's comparison). */
int amt_window_threshold_u16(const uint16_t* data, int64_t n_img, int64_t h, int64_t w, int window_h, int window_w,
                             int kind, double k, double r, uint8_t* mask, double* thresholds, amt_stream_t stream);
/* The same two methods on float64 images: scikit-image's float route (np.pad 'reflect', float64 integral images built
 * with np.cumsum's sequential additions along axis 0 then 1, window sums in _correlate_sparse's order), reproduced
 * operation by operation.  scratch: amt_window_threshold_f64_scratch_bytes (two padded planes per image). */
size_t amt_window_threshold_f64_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int window_h, int window_w);
int amt_window_threshold_f64(const double* data, int64_t n_img, int64_t h, int64_t w, int window_h, int window_w,
                             int kind, double k, double r, uint8_t* mask, double* thresholds, void* scratch,
                             size_t scratch_bytes, amt_stream_t stream);
/* threshold_li on float64 images (ref: operations.py:186 -> [3p] ski.filters.threshold_li, the non-integer branch):
 * the per-pixel pieces of scikit-image's iteration on image - min; the scalar recurrence stays with the caller.
 *  amt_li_shift_f64:    out[i] = data[i] - lo (rounded per element);
 *  amt_li_min_gap_f64:  *gap (device) = min(np.diff(np.unique(data))), +inf when all values are equal; the plane must hold
 *                       finite values; scratch: amt_li_min_gap_scratch_bytes (a padded copy that is sorted);
 *  amt_li_split_f64:    above = data[data > t], rest = data[~(data > t)], both in raster order (NumPy's boolean-mask
 *                       indexing); totals (device) = {len(above), len(rest)}; above / rest hold n doubles each. */
int amt_li_shift_f64(const double* data, int64_t n, double lo, double* out, amt_stream_t stream);
size_t amt_li_min_gap_scratch_bytes(int64_t n);
int amt_li_min_gap_f64(const double* data, int64_t n, double* gap, void* scratch, size_t scratch_bytes, amt_stream_t stream);
size_t amt_li_split_scratch_bytes(int64_t n);
int amt_li_split_f64(const double* data, int64_t n, double t, double* above, double* rest, int64_t* totals, void* scratch,
                     size_t scratch_bytes, amt_stream_t stream);
int amt_threshold_gt_image(const void* data, int in_dtype, int64_t n, const double* thresholds, double offset,
                           uint8_t* mask, amt_stream_t stream);

/* ------------------------------------------------------------------ labelling
 * ref: masks.py:38-65 (_process_mask): clear_border [3p] then measure.label [3p] (bool) or
 * relabel_sequential [3p] (integer masks).  Output labels int32, 1..K consecutive; K per
 * plane in `counts`.  Bool masks: 8-connectivity, components numbered in raster order of
 * their first pixel.  Integer masks: clear_border removes border-touching connected
 * FRAGMENTS of equal value, then values are renumbered in ascending order.
 * in_kind: 0 = uint8 mask, 1 = float64 plane compared `> thresholds[img]` on the fly,
 *          2 = int32 label plane (values in [0, max_value]). */
size_t amt_label_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_value);
int amt_label(const void* in, int in_kind, const double* thresholds, int64_t max_value,
              int64_t n_img, int64_t h, int64_t w, int clear_border,
              int32_t* labels_out, int32_t* counts, void* scratch, size_t scratch_bytes,
              amt_stream_t stream);

/* ------------------------------------------------------------------ per-cell quantification
 * ref: masks.py:286-289, :317-326 ski.measure.regionprops_table [3p].
 * One pass over labels + C uint16 channel planes accumulates exact integer statistics per
 * label; amt_region_finalize turns them into the float64 table.
 * channels: plane c of image i starts at channels + i*img_stride + c*chan_stride (elements).
 * acc: n_img * AMT_ACC_FIELDS(C) * max_labels uint64 (zero/identity-initialised by the call).
 * table: n_img * AMT_TABLE_COLS(C) * max_labels float64, column-major per image (SoA). */
#define AMT_ACC_BASE 10
#define AMT_ACC_PER_CHANNEL 4
#define AMT_ACC_FIELDS(C) (AMT_ACC_BASE + AMT_ACC_PER_CHANNEL * (C))
/* acc field order: count, sum_r, sum_c, sum_rr, sum_cc, sum_rc, r_min, r_max, c_min, c_max,
 * then per channel: sum, sum_sq, min, max. */
#define AMT_TABLE_BASE 16
#define AMT_TABLE_PER_CHANNEL 5
#define AMT_TABLE_COLS(C) (AMT_TABLE_BASE + AMT_TABLE_PER_CHANNEL * (C))
/* table column order: label, area, bbox-0..3, centroid-0, centroid-1, inertia eigval 1, 2,
 * axis_major_length, axis_minor_length, eccentricity, orientation, perimeter, area_convex,
 * then per channel: intensity_sum, intensity_mean, intensity_max, intensity_min,
 * intensity_std.  perimeter / area_convex are filled by amt_region_shape (NaN otherwise). */
int amt_region_reduce(const int32_t* labels, const uint16_t* channels, int n_channels,
                      int64_t img_stride, int64_t chan_stride, int64_t n_img, int64_t h, int64_t w,
                      int64_t max_labels, uint64_t* acc, amt_stream_t stream);
int amt_region_finalize(const uint64_t* acc, const int32_t* counts, int n_channels, int64_t n_img,
                        int64_t max_labels, double* table, amt_stream_t stream);
/* 3-D label volumes (z-stacks; extension, SegmentationMask itself is 2-D only: masks.py:171-172).
 * labels: d*h*w int32 (1..K consecutive, 0 = background); channels: C volumes of d*h*w uint16,
 * chan_stride elements apart.  acc: AMT_ACC3D_FIELDS(C) * max_labels uint64 (initialised by the call).
 * table (column-major, AMT_TABLE3D_COLS(C) x max_labels float64): label, area, bbox-0..5 (half-open,
 * z y x), centroid-0..2, inertia_tensor_eigvals-0..2, axis_major_length, axis_minor_length, then per
 * channel sum, mean, max, min, std -- skimage.measure.regionprops_table's 3-D definitions. */
#define AMT_ACC3D_BASE 16
#define AMT_ACC3D_FIELDS(C) (AMT_ACC3D_BASE + AMT_ACC_PER_CHANNEL * (C))
#define AMT_TABLE3D_BASE 16
#define AMT_TABLE3D_COLS(C) (AMT_TABLE3D_BASE + AMT_TABLE_PER_CHANNEL * (C))
int amt_region_reduce3d(const int32_t* labels, const uint16_t* channels, int n_channels, int64_t chan_stride,
                        int64_t d, int64_t h, int64_t w, int64_t max_labels, uint64_t* acc, amt_stream_t stream);
int amt_region_finalize3d(const uint64_t* acc, int64_t count, int n_channels, int64_t max_labels, double* table,
                          amt_stream_t stream);

/* perimeter (4-neighbourhood, skimage weights) and convex-hull pixel count per label.
 * ref: masks.py:15-28 defaults 'perimeter', 'area_convex', 'solidity'. scratch from
 * amt_region_shape_scratch_bytes. */
size_t amt_region_shape_scratch_bytes(int64_t n_img, int64_t h, int64_t w, int64_t max_labels);
int amt_region_shape(const int32_t* labels, const uint64_t* acc, int n_channels, const int32_t* counts,
                     int64_t n_img, int64_t h, int64_t w, int64_t max_labels, double* table,
                     void* scratch, size_t scratch_bytes, amt_stream_t stream);

/* ------------------------------------------------------------------ cell outlines
 * ref: masks.py:229-245 SegmentationMask.cell_outlines -> masks.py:82-115 (_extract_outlines_skimage,
 * skimage.measure.find_contours at level 0.5 on the padded crop label == n) and masks.py:68-79
 * (_extract_outlines_cellpose -> cv2.findContours(label == n, RETR_EXTERNAL, CHAIN_APPROX_NONE)).
 * labels: one h x w int32 label image (1..K, 0 = background), device memory.
 *
 * amt_outline_squares: marching-squares case of every 2x2 pixel square for every label at its
 * corners: keys[i] = label<<34 | r0<<19 | c0<<4 | case (case = ul | ur<<1 | ll<<2 | lr<<3, never 0 or
 * 15), in no particular order (sort them: per label, raster order of the squares = the order
 * find_contours emits its segments in).  *count (device) = number of keys the image holds; only the
 * first `capacity` are stored, call again with a larger buffer when *count > capacity.
 *
 * amt_outline_trace_find: best[L-1] = (points << 32 | raster index of the start pixel) of the longest
 * outer border among the 8-connected fragments of label L, 0 when L is absent (ties: later start).
 * amt_outline_trace_write: writes the border of label k+1 as (y, x) int32 pairs, in OpenCV's point
 * order, at points[2*offsets[k] .. 2*offsets[k+1]) (offsets: n_labels+1 device int64; an empty range
 * skips the label). */
int amt_outline_squares(const int32_t* labels, int64_t h, int64_t w, uint64_t* keys, int64_t capacity,
                        uint64_t* count, amt_stream_t stream);
int amt_outline_trace_find(const int32_t* labels, int64_t h, int64_t w, int64_t max_labels, uint64_t* best,
                           amt_stream_t stream);
int amt_outline_trace_write(const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, const uint64_t* best,
                            const int64_t* offsets, int32_t* points, amt_stream_t stream);

/* ------------------------------------------------------------------ fused FOV executor
 * The native runtime for the batch path (bench + MicroscopyImage batch pipeline): owns its
 * streams, device scratch and pinned double-buffered staging, and runs the whole workload W
 * of SURVEY.md 8(d) for a batch of fields of view with no host round trip inside a chunk:
 * per channel DoG -> percentile background -> clip -> percentile rescale; Otsu on the
 * segmentation channel; threshold + CCL + clear_border; per-cell table over all raw channels;
 * plus the same table for a caller-given integer label mask (clear_border + relabel).
 */
typedef struct amt_executor amt_executor;

typedef struct amt_fov_config {
  int32_t device;          /* CUDA device ordinal */
  int32_t n_channels;      /* C */
  int32_t height, width;   /* Y, X */
  int32_t seg_channel;     /* channel whose preprocessed plane is thresholded */
  int32_t chunk_fovs;      /* FOVs processed per launch wave (scratch is sized for this) */
  int32_t max_labels;      /* table capacity per FOV and per mask */
  int32_t max_label_value; /* largest value allowed in a given label mask */
  int32_t quantify_given_mask; /* 1: also clear_border + relabel + quantify the given mask */
  int32_t with_shape;          /* 1: also fill perimeter / area_convex (amt_region_shape) */
  int32_t given_label_dtype;   /* amt_executor_run_host only: dtype of the host label masks.  AMT_I32 (0 = default),
                                  AMT_I64 (what the reference hands SegmentationMask: model.py:215, masks.py:138;
                                  copied as int64 and narrowed on the device) or AMT_U16 (Cellpose's own mask dtype
                                  below 65536 cells; a quarter of the int64 PCIe bytes) */
  int32_t exact_all_channels;  /* 0 (default): scipy's exact operation order for the segmentation channel, whose
                                  plane decides the labels; the other channels, which only yield float planes,
                                  take the faster filter named by plane_filter.  1: exact order for every channel
                                  (every preprocessed plane bit-identical to the reference's) */
  double low_sigma, high_sigma;  /* subtract_background_dog */
  double bg_percentile;
  double pct_lo, pct_hi;         /* rescale_by_percentile percentile_range */
  double out_lo, out_hi;         /* rescale_by_percentile out_range */
  int32_t plane_filter;          /* how the channels that are NOT thresholded get their sigma_high Gaussian when
                                    exact_all_channels == 0: AMT_FILTER_TENSOR_CORE (0, default) = tcgen05 integer
                                    Toeplitz passes (amt_tcg_*; planes equal to scipy's to ~1e-10 of the [0, 1] scale)
                                    whenever amt_tcg_supported(height, width, radius) and n_channels >= 2, else the
                                    next mode; AMT_FILTER_FMA (1) = float64 with fused multiply-adds (~1e-15) */
  int32_t seg_plane_filter;      /* the thresholded channel when the tensor-core path is on: AMT_SEG_DECISION_EXACT
                                    (0, default) = tensor-core filter too; every decision derived from its plane
                                    (order statistics, histogram bins, the mask) is taken on exactly re-evaluated
                                    samples wherever the filter's proven error bound could change it (csrc/decide.cu),
                                    so thresholds, labels, counts and tables stay bit-identical to the reference's;
                                    its float plane is then within the bound like the other channels'.
                                    AMT_SEG_FLOAT64 (1) = the float64 strip kernels in scipy's order for that channel
                                    (its plane bit-identical as well) */
} amt_fov_config;
#define AMT_FILTER_TENSOR_CORE 0
#define AMT_FILTER_FMA 1
#define AMT_SEG_DECISION_EXACT 0
#define AMT_SEG_FLOAT64 1

/* half_w_*_host: NumPy-computed half kernels (radius+1 doubles each). */
int amt_executor_create(const amt_fov_config* cfg, const double* half_w_lo_host, int r_lo,
                        const double* half_w_hi_host, int r_hi, amt_executor** out);
void amt_executor_destroy(amt_executor* ex);
size_t amt_executor_device_bytes(const amt_executor* ex);
/* 1 if this executor filters the non-thresholded channels on the tensor cores (plane_filter resolved at creation). */
int amt_executor_uses_tensor_cores(const amt_executor* ex);
/* 1 if the thresholded channel runs in decision-exact mode (seg_plane_filter resolved at creation). */
int amt_executor_decision_exact(const amt_executor* ex);
/* Fields of view recomputed with the float64 kernels since creation because a candidate list of the decision-exact
 * mode overflowed (massive ties, e.g. constant images).  Their results are exact like everybody else's. */
int64_t amt_executor_retry_count(const amt_executor* ex);
/* Host -> device bytes the last amt_executor_run_host batch copied (images + label masks as they crossed PCIe).  Host
 * label masks (any of the three dtypes) cross as per-row runs of equal value, encoded by host threads into pinned
 * staging inside the call and decoded on the device (amt_tune "exec_host_rle", default on; "exec_host_threads"): a
 * segmentation mask is long runs by nature, ~1 MB instead of 8.4 MB (uint16) / 33.5 MB (int64) per 2048 x 2048 mask of
 * ~2000 cells.  A chunk whose runs do not fit the staging (under 4 pixels per run on average) is sent as the plain
 * mask instead; amt_executor_last_plain_mask_chunks counts those chunks of the last batch. */
int64_t amt_executor_last_h2d_bytes(const amt_executor* ex);
/* The host half of that route on its own (no GPU involved; tests and host-side timing): n_fov label masks of
 * height x width (dtype AMT_I64 / AMT_I32 / AMT_U16) -> runs[2 k] = value, runs[2 k + 1] = end column (exclusive) of run k;
 * rows[2 r] = first run slot of row r, rows[2 r + 1] = its number of runs.  `runs` holds n_fov * height * (width / 4)
 * slots; thread t packs the runs of its rows from slot first_row(t) * (width / 4) on.  negative[f] = 1 when FOV f holds a
 * negative label (stored as background); values beyond int32 saturate.  AMT_ERR_CAPACITY when the runs do not fit. */
int amt_rle_encode_host(const void* labels_host, int dtype, int32_t n_fov, int32_t height, int32_t width, int32_t n_threads,
                        uint32_t* runs, uint32_t* rows, int32_t* negative, int64_t* n_runs);
int64_t amt_executor_last_plain_mask_chunks(const amt_executor* ex);
/* Encoding costs host time, plain masks cost PCIe time: per chunk the executor splits the masks between the two routes
 * so that both finish together (the plain ones cross right behind the images while the host threads encode the others),
 * from the encode time per mask and the image-copy rate it measured on the chunks before.  With a GPU to itself and
 * enough host threads every mask is encoded; eight ranks sharing one host's cores and PCIe root send some masks plain.
 * amt_tune("exec_rle_share", p) fixes the encoded share at p per cent instead (-1 = balanced, the default).
 * amt_executor_last_rle_masks: masks of the last batch that crossed as runs. */
int64_t amt_executor_last_rle_masks(const amt_executor* ex);

/* Per-stage device time (CUDA events after every stage of both executor streams; adds a few microseconds per
 * chunk, off by default).  amt_executor_set_profiling(ex, 1) zeroes the counters; every amt_executor_run_device call
 * that follows adds its chunks; amt_executor_stage_ms copies the sums (milliseconds, AMT_N_STAGES entries) and
 * the number of chunks they cover.  The DoG stages run on their own stream one chunk ahead of the others, so the
 * sum over stages exceeds the wall time of a run. */
#define AMT_STAGE_DOG_EXACT 0     /* float64 DoG in scipy's order (the thresholded channel, or every plane) */
#define AMT_STAGE_DOG_LO 1        /* narrow Gaussian of the tensor-core planes */
#define AMT_STAGE_DOG_TC0 2       /* tensor-core wide Gaussian, axis 0 */
#define AMT_STAGE_DOG_TC1 3       /* tensor-core wide Gaussian, axis 1, + subtraction, buckets, min / max */
#define AMT_STAGE_SELECT 4        /* order statistics (percentiles) */
#define AMT_STAGE_MAP 5           /* plan + subtract / clip / rescale map + 256-bin histogram */
#define AMT_STAGE_LABEL_THR 6     /* Otsu scan + threshold + CCL + clear_border + numbering */
#define AMT_STAGE_REGIONS_THR 7   /* per-cell tables of the threshold mask */
#define AMT_STAGE_LABEL_GIVEN 8   /* clear_border + relabel_sequential of the given mask */
#define AMT_STAGE_REGIONS_GIVEN 9 /* per-cell tables of the given mask (+ status) */
#define AMT_N_STAGES 10
int amt_executor_set_profiling(amt_executor* ex, int enable);
int amt_executor_stage_ms(const amt_executor* ex, double* stage_ms, int64_t* n_chunks);

/* Per-FOV status bits (status[i] == 0: FOV i is complete and exact).  A field of view that trips one of these
 * does not disturb the others of the batch (the reference maps its per-image loop the same way:
 * model.py:276-288); counts_* always hold the TRUE number of cells, tables hold min(count, max_labels) columns. */
#define AMT_FOV_THR_CAPACITY 1       /* threshold mask: more cells than max_labels, table truncated to the first max_labels */
#define AMT_FOV_GIVEN_CAPACITY 2     /* given mask: idem */
#define AMT_FOV_GIVEN_VALUE_RANGE 4  /* given mask holds a value > max_label_value (such pixels count as background) */
#define AMT_FOV_THR_EMPTY 8          /* no cell left in the threshold mask after removing edge cells: the reference's
                                        _process_mask raises ValueError here (masks.py:57-60) */
#define AMT_FOV_GIVEN_EMPTY 16       /* idem for the given mask */
#define AMT_FOV_CONSTANT_PLANE 32    /* the thresholded plane is constant: apply_threshold returns all-False
                                        (operations.py:199-202) */
#define AMT_FOV_GIVEN_NEGATIVE 64    /* int64 host mask with a negative value (SegmentationMask raises, masks.py:173-176);
                                        such pixels count as background */

/* Device-resident batch.  fovs: n_fov*C*H*W uint16; given_labels: n_fov*H*W int32 or NULL.
 * Outputs (device): tables_thr / tables_given: n_fov*AMT_TABLE_COLS(C)*max_labels float64;
 * counts_thr / counts_given: n_fov int32; thresholds: n_fov float64; labels_thr /
 * labels_given (optional, may be NULL): n_fov*H*W int32; preprocessed (optional):
 * n_fov*C*H*W float64; status (optional): n_fov int32 of AMT_FOV_* bits.
 * Asynchronous on the executor's streams; amt_executor_sync waits. */
int amt_executor_run_device(amt_executor* ex, const uint16_t* fovs, const int32_t* given_labels,
                            int64_t n_fov, double* tables_thr, int32_t* counts_thr,
                            double* tables_given, int32_t* counts_given, double* thresholds,
                            int32_t* labels_thr, int32_t* labels_given, double* preprocessed, int32_t* status);
/* Host-fed batch: inputs and outputs are HOST pointers (pinned memory recommended; pageable
 * works but serialises).  Copies are double-buffered against compute on separate streams.
 * Synchronous: returns when every output byte is on the host. */
int amt_executor_run_host(amt_executor* ex, const uint16_t* fovs_host, const void* given_labels_host,
                          int64_t n_fov, double* tables_thr_host, int32_t* counts_thr_host,
                          double* tables_given_host, int32_t* counts_given_host, double* thresholds_host,
                          int32_t* status_host);
int amt_executor_sync(amt_executor* ex);
/* CUDA-event time (ms) of the last run_device / run_host call's device work. */
float amt_executor_last_ms(amt_executor* ex);

#ifdef __cplusplus
}
#endif
#endif /* AMT_B200_H */
