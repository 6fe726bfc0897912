// Fused field-of-view executor: the native runtime of the batch path.
//
// Runs workload W for a batch of FOVs with no host round trip inside a chunk:
//   for every channel   P_c = rescale_by_percentile(subtract_background_dog(x_c, lo, hi, pct_bg),
//                                                   (pct_lo, pct_hi), (out_lo, out_hi))
//                       ref: operations.py:57-97 then operations.py:10-54
//   segmentation chan.  m = apply_threshold(P_s, "otsu")            ref: operations.py:135-216
//   SegmentationMask(m, {ch: x_c}, remove_edge_cells=True)          ref: masks.py:38-65, 247-328
//   SegmentationMask(given_labels, {ch: x_c}, remove_edge_cells=True)
// FOVs are independent, so a batch is cut into chunks of `chunk_fovs` and every kernel is
// launched over all planes of a chunk (grid.y / grid.z = plane).  The host-fed entry point
// double-buffers H2D copies, compute and D2H copies on three streams.

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "internal.cuh"

struct amt_executor {
  amt_fov_config cfg;
  int r_lo, r_hi;
  int64_t ranks[6];
  double g_bg, g_lo, g_hi;
  // s_dog runs the FP64-bound DoG of chunk i+1 while s_compute (higher priority) finishes chunk i.
  // Measured (AMT_TRACE, scripts/overlap_probe.py): both sides are instruction-issue-bound, so the
  // overlap only recovers the other kernels' memory stalls (~0.5 ms of a 5.9 ms chunk)
  cudaStream_t s_compute, s_dog, s_in, s_out;
  cudaStream_t s_given;        // the given-mask chain (label + per-cell tables): depends on the inputs only
  cudaEvent_t ev_fork, ev_join;
  cudaEvent_t ev_start, ev_stop;
  cudaEvent_t ev_in[2], ev_done[2], ev_out[2];
  cudaEvent_t ev_dog_done[2], ev_dog_free[2];
  int64_t chunks_issued;
  // device buffers
  double *hw_lo, *hw_hi;
  double *tmp_lo, *tmp_hi, *dog[2], *pre;
  uint64_t* mm[2];
  amt_tcg* tcg;              // tensor-core Gaussian of the planes that are not thresholded (null: float64 everywhere)
  // decision-exact thresholded channel (decide.cu): every plane takes the tensor-core filter; the samples a decision
  // hinges on are re-evaluated in scipy's exact order
  bool dx;                   // mode resolved at creation
  bool force_exact;          // set while a field of view whose candidate lists overflowed is recomputed in float64
  double dx_eps;             // |D' - D| <= dx_eps (amt_tcg_error_bound)
  int64_t* dx_ranks;         // device copy of the six percentile ranks
  uint32_t* dx_rank_u32;     // below / count / idx of the order-statistic windows
  double* dx_rank_val;
  uint32_t* dx_bin_count;    // per FOV of a chunk
  uint32_t* dx_bin_idx;
  double* dx_bin_val;
  int32_t* retry_dev;        // per FOV of the current run_device batch (grown on demand)
  int64_t retry_cap;
  int32_t* retry_slot[2];    // host-fed path: per FOV of a chunk
  int32_t* retry_host;       // pinned, per FOV of a run_host batch (grown on demand)
  int64_t retry_host_cap;
  int64_t retries;           // fields of view recomputed in float64 since creation
  struct {
    bool pending;
    const uint16_t* fovs;
    const int32_t* given;
    int64_t n_fov;
    double *tables_thr, *tables_given, *thresholds, *preprocessed;
    int32_t *counts_thr, *counts_given, *labels_thr, *labels_given, *status;
  } last;
  uint8_t* digits;           // its 40-bit intermediate, five uint8 planes per image
  uint16_t* buckets[2];      // bucket12() of the DoG planes (written by the DoG's second pass)
  bool buckets_valid[2];
  double* stats;
  amt_map_params* params;
  void* sel_scratch;
  size_t sel_bytes;
  uint32_t* hist256;
  double* thr;
  int32_t *lab_thr, *lab_given;
  void* label_scratch;
  size_t label_bytes;
  uint64_t* acc;
  void* shape_scratch;
  size_t shape_bytes;
  void* label_scratch_g;     // the same three for the given-mask chain, which runs on its own stream
  uint64_t* acc_g;
  void* shape_scratch_g;
  // host-path staging (two slots)
  uint16_t* in_slot[2];
  int32_t* given_slot[2];
  uint16_t* given16_slot[2];  // raw uint16 label masks (given_label_dtype == AMT_U16), widened on the device
  int64_t* given64_slot[2];   // raw int64 label masks (given_label_dtype == AMT_I64), narrowed on the device
  uint16_t* given16_host[2];  // pinned: int64 masks narrowed to uint16 by host threads inside amt_executor_run_host (42 instead of 67 MB per FOV over PCIe)
  bool host_narrow;           // int64 masks, max_label_value < 65535 and amt_tune("exec_host_narrow") != 0
  int32_t* neg_host;          // per FOV of a run_host batch: a negative label was seen by the host narrowing / encoding
  // run-length staging of host label masks (amt_tune "exec_host_rle", any of the three dtypes): host threads turn every
  // row into (value, end column) runs in pinned staging; only the runs cross PCIe (~1 instead of 8.4 / 16.8 / 33.5 MB per
  // 2048 x 2048 mask of ~2000 cells) and rle_decode_kernel writes the int32 label image
  bool host_rle;
  uint2* rle_host[2];         // pinned: run slots, W/4 per row (a thread packs the runs of its rows from its first row's slot)
  uint2* rle_slot[2];         // device mirror (same sparse layout)
  uint2* rle_rows_host[2];    // pinned: per row {first run slot, number of runs}
  uint2* rle_rows_slot[2];
  int64_t last_h2d_bytes;     // bytes the last run_host batch copied host -> device
  // how many masks of a chunk take the run-length route is balanced per chunk between the host threads (seconds per
  // encoded mask, rle_a) and PCIe (bytes per second of the image copy, rle_B, timed with ev_img0/1): the others cross
  // as plain masks right behind the images while the host threads encode
  cudaEvent_t ev_img0[2], ev_img1[2];
  int64_t slot_img_bytes[2];
  double rle_a, rle_B, rle_c;   // running means; 0 = not measured yet.  rle_c: bytes of runs + row table per encoded mask
  int64_t last_rle_masks, last_plain_masks;
  int64_t rle_fallback_chunks;  // chunks of the last run_host batch whose runs did not fit (sent as plain masks)
  int64_t neg_host_cap;
  bool slot_uploaded[2];      // the staging slot has an H2D in flight or behind it (its ev_in is valid)
  int32_t* flag_slot[2];      // per FOV of a chunk: [0, chunk) value > max_label_value seen, [chunk, 2 chunk) negative value seen
  int32_t* status_slot[2];
  int32_t* flags_dev;         // the same two flag arrays for the device-resident entry point
  double *tab_thr_slot[2], *tab_given_slot[2], *thr_slot[2];
  int32_t *cnt_thr_slot[2], *cnt_given_slot[2];
  bool host_slots;
  size_t device_bytes;
  float last_ms;
  // AMT_TRACE=1: an event after every stage of both streams, dumped (ms since the batch start) by
  // amt_executor_run_device once the batch has finished.  Debugging aid; off by default.
  bool trace;
  std::vector<cudaEvent_t>* trace_events;
  std::vector<const char*>* trace_names;
  // amt_executor_set_profiling: an event after every stage; per-stage device milliseconds of the last run
  bool profile;
  std::vector<int>* trace_stage;   // stage id each event closes (-1: a "begin" mark), same order as trace_events
  std::vector<int>* trace_stream;  // 0 = s_dog, 1 = s_compute, 2 = s_given
  double stage_ms[AMT_N_STAGES];
  int64_t profiled_chunks;
};

namespace amt {

// amt_tune "exec_swap_prio": 1 (default) = the stream of the short HBM-bound kernels has the high
// priority and the long-running DoG CTAs the low one, 0 = the opposite
int g_exec_swap_prio = 1;
// amt_tune "exec_buckets": 1 (default) = the DoG's second pass also writes bucket12() of its output and the
// percentile selection reads those 2-byte buckets instead of the 8-byte planes; 0 = plain amt_select_f64
int g_exec_buckets = 1;
// amt_tune "exec_tc": 0 switches the tensor-core path off in executors that have one (A/B timing in one process)
int g_exec_tc = 1;
// amt_tune "exec_fused_lo": 1 = the narrow Gaussian inside pass 2 (amt_tcg_axis1_dog: 16 B/px less traffic, bit-identical).
// Off by default: measured SLOWER on B200 (1.11 ms against 0.65 + 0.28 ms per 32 planes): the ~450 extra instructions per
// epilogue warp and tile run at the same ~7 clk per instruction as the rest of that latency-bound epilogue (2 warps
// per scheduler at 168 registers), see profiles/r02_tcgauss_experiments.md.
int g_exec_fused_lo = 1;
// decision-exact mode: capacities of the candidate lists per plane and window / per plane (overflow = float64 retry)
constexpr int kDxRankCap = 1024;
constexpr int kDxBinCap = 16384;
// amt_tune "exec_copy_only": 1 = amt_executor_run_host performs every H2D / D2H copy of a batch with the same staging,
// streams and events but launches no kernel: the copy-only ceiling the host-fed path is measured against
int g_exec_copy_only = 0;
// amt_tune "exec_given_stream": 1 (default) = label + tables of the given masks run on a stream of their own next to
// select / map / label / tables of the thresholded channel (they depend on the inputs only); 0 = one after the other
int g_exec_given_stream = 1;
// amt_tune "exec_host_narrow": 1 (default) = int64 label masks of a host-fed batch are narrowed to uint16 by host threads
// into pinned staging before the H2D copy (values above 65534 saturate and are reported as out of range, as on the
// device route); 0 = the int64 masks cross PCIe and are narrowed on the device.  "exec_host_threads": threads used.
int g_exec_host_narrow = 1;
int g_exec_host_threads = 8;
// amt_tune "exec_host_rle": 1 (default) = host label masks of a host-fed batch (int64, int32 or uint16) cross PCIe as
// per-row runs of equal value, encoded by host threads into pinned staging inside the call and decoded on the device
// (a label mask of ~2000 cells: ~1 MB instead of 8.4 MB as uint16).  A chunk whose runs do not fit the staging (fewer
// than 4 pixels per run on average over a thread's rows) is sent the plain way.  0 = always the plain way.
int g_exec_host_rle = 1;
// amt_tune "exec_rle_share": -1 (default) = the number of masks of a chunk that take the run-length route is balanced
// against PCIe from the measured encode and copy rates; 0..100 = that fixed percentage of a chunk's masks
int g_exec_rle_share = -1;
int minmax_init(uint64_t* mm, int64_t n_img, cudaStream_t st);  // gauss.cu

static int dmalloc(amt_executor* ex, void** p, size_t bytes) {
  AMT_CUDA_TRY(cudaMalloc(p, bytes));
  ex->device_bytes += bytes;
  return AMT_OK;
}

// stage: the AMT_STAGE_* id whose work this mark closes, or -1 for a mark that only opens a stream's timeline
static void trace_mark(amt_executor* ex, cudaStream_t st, const char* name, int stage = -1) {
  if (!ex->trace && !ex->profile) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  ex->trace_events->push_back(e);
  ex->trace_names->push_back(name);
  ex->trace_stage->push_back(stage);
  ex->trace_stream->push_back(st == ex->s_dog ? 0 : (st == ex->s_given ? 2 : 1));
}

static void trace_dump(amt_executor* ex) {
  if (!ex->trace && !ex->profile) return;
  cudaDeviceSynchronize();
  if (ex->profile) {
    // a stage lasts from the previous mark on ITS stream to its own mark
    int last[3] = {-1, -1, -1};
    for (size_t i = 0; i < ex->trace_events->size(); ++i) {
      const int sid = (*ex->trace_stream)[i], stage = (*ex->trace_stage)[i];
      if (stage >= 0 && last[sid] >= 0) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, (*ex->trace_events)[last[sid]], (*ex->trace_events)[i]) == cudaSuccess)
          ex->stage_ms[stage] += ms;
      }
      last[sid] = (int)i;
    }
  }
  for (size_t i = 0; i < ex->trace_events->size(); ++i) {
    if (ex->trace) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ex->ev_start, (*ex->trace_events)[i]);
      std::fprintf(stderr, "amt-trace %9.3f ms  %s\n", ms, (*ex->trace_names)[i]);
    }
    cudaEventDestroy((*ex->trace_events)[i]);
  }
  ex->trace_events->clear();
  ex->trace_names->clear();
  ex->trace_stage->clear();
  ex->trace_stream->clear();
}

static void rank_pair(int64_t n, double q, int64_t* lo, int64_t* hi, double* gamma) {
  // numpy: virtual index (n-1) * (q/100); floor / ceil neighbours; gamma = v - floor(v)
  const double quant = q / 100.0;
  const double v = (double)(n - 1) * quant;
  int64_t l = (int64_t)std::floor(v);
  if (l < 0) l = 0;
  if (l > n - 1) l = n - 1;
  *lo = l;
  *hi = (l + 1 < n - 1) ? l + 1 : n - 1;
  *gamma = v - (double)l;
}

// stage A1 on s_dog: DoG of one chunk into dog[slot] / mm[slot]
static int enqueue_dog(amt_executor* ex, const uint16_t* in, int g, cudaEvent_t wait_input) {
  const amt_fov_config& c = ex->cfg;
  const int slot = (int)(ex->chunks_issued & 1);
  const int64_t planes = (int64_t)g * c.n_channels;
  if (wait_input) AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_dog, wait_input, 0));
  if (ex->chunks_issued >= 2) AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_dog, ex->ev_dog_free[slot], 0));
  trace_mark(ex, ex->s_dog, "dog: begin");
  if (ex->tcg != nullptr && g_exec_tc && ex->dx && !ex->force_exact) {
    // decision-exact mode: every plane on the tensor cores; process_chunk re-evaluates the deciding samples exactly
    const tc::PlaneSel all{0, 0};
    uint16_t* bk = g_exec_buckets ? ex->buckets[slot] : nullptr;
    ex->buckets_valid[slot] = bk != nullptr;
    AMT_TRY(minmax_init(ex->mm[slot], planes, ex->s_dog));
    const bool fused = g_exec_fused_lo && ex->r_lo <= 4;  // the narrow Gaussian inside pass 2 (no kernel, no plane of its own)
    if (!fused) {
      AMT_TRY(tc::lo2d(in, 1.0 / 65535.0, ex->tmp_lo, planes, c.height, c.width, ex->hw_lo, ex->r_lo, all, ex->s_dog));
      trace_mark(ex, ex->s_dog, "dog: narrow Gaussian done", AMT_STAGE_DOG_LO);
    }
    AMT_TRY(tc::tcg_axis0(ex->tcg, in, planes, c.height, c.width, ex->digits, all, ex->s_dog));
    trace_mark(ex, ex->s_dog, "dog: tensor-core axis 0 done", AMT_STAGE_DOG_TC0);
    AMT_TRY(tc::tcg_axis1(ex->tcg, ex->digits, ex->tmp_lo, 1.0 / 65535.0, ex->dog[slot], planes, c.height, c.width, bk,
                          ex->mm[slot], all, ex->s_dog, fused ? in : nullptr, ex->hw_lo, ex->r_lo));
    AMT_CUDA_TRY(cudaEventRecord(ex->ev_dog_done[slot], ex->s_dog));
    trace_mark(ex, ex->s_dog, "dog: end", AMT_STAGE_DOG_TC1);
    return AMT_OK;
  }
  if (ex->tcg != nullptr && g_exec_tc) {
    // the thresholded channel: float64 in scipy's order (strip kernels over every C-th plane); the others: narrow
    // Gaussian in float64 (into tmp_lo, whose planes of these channels the strip kernels do not touch), wide Gaussian
    // on the tensor cores.  All three write disjoint planes of dog / buckets / mm.
    const tc::PlaneSel sel{c.n_channels, c.seg_channel};
    uint16_t* bk = g_exec_buckets ? ex->buckets[slot] : nullptr;
    AMT_TRY(minmax_init(ex->mm[slot], planes, ex->s_dog));
    AMT_TRY(dog2d(in, AMT_U16, 1.0 / 65535.0, ex->dog[slot], planes, c.height, c.width, ex->hw_lo, ex->r_lo, ex->hw_hi,
                  ex->r_hi, ex->tmp_lo, ex->tmp_hi, ex->mm[slot], ex->s_dog, bk, &ex->buckets_valid[slot], c.n_channels,
                  c.seg_channel, true));
    trace_mark(ex, ex->s_dog, "dog: exact planes done", AMT_STAGE_DOG_EXACT);
    const bool fused = g_exec_fused_lo && ex->r_lo <= 4;
    if (!fused) {
      AMT_TRY(tc::lo2d(in, 1.0 / 65535.0, ex->tmp_lo, planes, c.height, c.width, ex->hw_lo, ex->r_lo, sel, ex->s_dog));
      trace_mark(ex, ex->s_dog, "dog: narrow Gaussian done", AMT_STAGE_DOG_LO);
    }
    AMT_TRY(tc::tcg_axis0(ex->tcg, in, planes, c.height, c.width, ex->digits, sel, ex->s_dog));
    trace_mark(ex, ex->s_dog, "dog: tensor-core axis 0 done", AMT_STAGE_DOG_TC0);
    AMT_TRY(tc::tcg_axis1(ex->tcg, ex->digits, ex->tmp_lo, 1.0 / 65535.0, ex->dog[slot], planes, c.height, c.width, bk,
                          ex->mm[slot], sel, ex->s_dog, fused ? in : nullptr, ex->hw_lo, ex->r_lo));
    AMT_CUDA_TRY(cudaEventRecord(ex->ev_dog_done[slot], ex->s_dog));
    trace_mark(ex, ex->s_dog, "dog: end", AMT_STAGE_DOG_TC1);
    return AMT_OK;
  }
  AMT_TRY(dog2d(in, AMT_U16, 1.0 / 65535.0, ex->dog[slot], planes, c.height, c.width, ex->hw_lo, ex->r_lo, ex->hw_hi,
                ex->r_hi, ex->tmp_lo, ex->tmp_hi, ex->mm[slot], ex->s_dog, g_exec_buckets ? ex->buckets[slot] : nullptr,
                &ex->buckets_valid[slot],
                // only the segmentation channel's plane decides anything discrete (threshold -> labels); the other
                // channels yield float planes only and take the contracted filter unless the caller asks otherwise
                (c.exact_all_channels || c.n_channels < 2) ? 0 : c.n_channels, c.seg_channel));
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_dog_done[slot], ex->s_dog));
  trace_mark(ex, ex->s_dog, "dog: end", AMT_STAGE_DOG_EXACT);
  return AMT_OK;
}

__global__ void fov_status_kernel(const int32_t* __restrict__ cnt_thr, const int32_t* __restrict__ cnt_given,
                                  const int32_t* __restrict__ value_overflow, const int32_t* __restrict__ negative,
                                  const amt_map_params* __restrict__ params, int n_channels, int seg_channel, int max_labels,
                                  int n_fov, int32_t* __restrict__ status, const int32_t* __restrict__ retry_flags,
                                  int32_t* __restrict__ retry_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_fov) return;
  if (retry_out != nullptr) retry_out[i] = retry_flags[i];
  if (status == nullptr) return;
  int s = 0;
  if (cnt_thr[i] > max_labels) s |= AMT_FOV_THR_CAPACITY;
  if (cnt_thr[i] == 0) s |= AMT_FOV_THR_EMPTY;
  const amt_map_params& p = params[(int64_t)i * n_channels + seg_channel];
  if (p.hist_first == p.hist_last) s |= AMT_FOV_CONSTANT_PLANE;
  if (cnt_given != nullptr) {
    if (cnt_given[i] > max_labels) s |= AMT_FOV_GIVEN_CAPACITY;
    if (cnt_given[i] == 0) s |= AMT_FOV_GIVEN_EMPTY;
    if (value_overflow[i]) s |= AMT_FOV_GIVEN_VALUE_RANGE;
    if (negative[i]) s |= AMT_FOV_GIVEN_NEGATIVE;
  }
  status[i] = s;
}

// label + per-cell tables of the given masks of one chunk (own scratch: may run next to the thresholded channel's chain)
static int given_chain(amt_executor* ex, const uint16_t* in, const int32_t* given, int g, double* tab_given, int32_t* cnt_given,
                       int32_t* lab_given, int32_t* flags, cudaStream_t sg) {
  const amt_fov_config& c = ex->cfg;
  const int C = c.n_channels;
  const int64_t H = c.height, W = c.width, HW = H * W;
  AMT_TRY(label_launch(given, 2, HW, nullptr, c.max_label_value, g, H, W, 1, lab_given, cnt_given, ex->label_scratch_g,
                       ex->label_bytes, sg, flags));
  trace_mark(ex, sg, "    given: label done", AMT_STAGE_LABEL_GIVEN);
  AMT_TRY(region_reduce(lab_given, in, C, (int64_t)C * HW, HW, g, H, W, c.max_labels, ex->acc_g, sg));
  AMT_TRY(region_finalize(ex->acc_g, cnt_given, C, g, c.max_labels, tab_given, sg));
  if (c.with_shape)
    AMT_TRY(region_shape(lab_given, ex->acc_g, C, cnt_given, g, H, W, c.max_labels, tab_given, ex->shape_scratch_g,
                         ex->shape_bytes, sg));
  trace_mark(ex, sg, "    given: tables done", AMT_STAGE_REGIONS_GIVEN);
  return AMT_OK;
}

// everything after the DoG, on s_compute
static int process_chunk(amt_executor* ex, const uint16_t* in, const int32_t* given, int g, double* tab_thr,
                         int32_t* cnt_thr, double* tab_given, int32_t* cnt_given, double* thr_out, int32_t* lab_thr_out,
                         int32_t* lab_given_out, double* pre_out, int32_t* flags, int32_t* status, int32_t* retry_out) {
  const amt_fov_config& c = ex->cfg;
  const int C = c.n_channels;
  const int64_t H = c.height, W = c.width, HW = H * W;
  const int64_t planes = (int64_t)g * C;
  const int slot = (int)(ex->chunks_issued & 1);
  cudaStream_t st = ex->s_compute;
  double* pre = pre_out ? pre_out : ex->pre;
  int32_t* lab_thr = lab_thr_out ? lab_thr_out : ex->lab_thr;
  int32_t* lab_given = lab_given_out ? lab_given_out : ex->lab_given;
  double* thr = thr_out ? thr_out : ex->thr;
  double* dog = ex->dog[slot];
  uint64_t* mm = ex->mm[slot];

  // stage A2: order statistics -> plan -> map (+ histogram of the segmentation planes)
  AMT_CUDA_TRY(cudaStreamWaitEvent(st, ex->ev_dog_done[slot], 0));
  trace_mark(ex, st, "  rest: begin (dog of this chunk done)");
  // the given-mask chain needs the inputs only: it runs next to everything below on its own stream
  const bool with_given = c.quantify_given_mask && given;
  cudaStream_t sg = g_exec_given_stream ? ex->s_given : st;
  if (with_given && sg != st) {
    AMT_CUDA_TRY(cudaEventRecord(ex->ev_fork, st));
    AMT_CUDA_TRY(cudaStreamWaitEvent(sg, ex->ev_fork, 0));
    trace_mark(ex, sg, "    given: begin");
    AMT_TRY(given_chain(ex, in, given, g, tab_given, cnt_given, lab_given, flags, sg));
  }
  if (ex->buckets_valid[slot] && HW % 8 == 0)
    AMT_TRY(select_f64_bucketed(dog, ex->buckets[slot], planes, HW, ex->ranks, 6, mm, ex->stats, ex->sel_scratch,
                                ex->sel_bytes, st));
  else
    AMT_TRY(amt_select_f64(dog, planes, HW, ex->ranks, 6, mm, ex->stats, ex->sel_scratch, ex->sel_bytes, st));
  const bool dx = ex->tcg != nullptr && g_exec_tc && ex->dx && !ex->force_exact;
  int32_t* retry_flags = flags + 2 * c.chunk_fovs;
  if (dx) {
    // the thresholded planes' order statistics and min / max become exact (decide.cu)
    const int64_t seg = c.seg_channel;
    AMT_TRY(dx::rank_exact(in + seg * HW, (int64_t)C * HW, dog + seg * HW,
                           ex->buckets_valid[slot] ? ex->buckets[slot] + seg * HW : nullptr, (int64_t)C * HW, g, (int)H, (int)W,
                           1.0 / 65535.0,
                           ex->hw_hi, ex->r_hi, ex->hw_lo, ex->r_lo, ex->dx_eps, ex->dx_ranks, ex->stats + seg * 6,
                           (int64_t)C * 6, mm + seg * 2, (int64_t)C * 2, ex->dx_rank_u32, ex->dx_rank_val, kDxRankCap,
                           retry_flags, st));
  }
  trace_mark(ex, st, "  rest: select done", AMT_STAGE_SELECT);
  AMT_TRY(plan_dog_rescale(ex->stats, mm, planes, ex->g_bg, ex->g_lo, ex->g_hi, c.out_lo, c.out_hi, ex->params, st));
  AMT_CUDA_TRY(cudaMemsetAsync(ex->hist256, 0, (size_t)g * 256 * sizeof(uint32_t), st));
  if (dx) {
    AMT_CUDA_TRY(cudaMemsetAsync(ex->dx_bin_count, 0, (size_t)g * sizeof(uint32_t), st));
    AMT_TRY(map_launch(dog, AMT_F64, pre, planes, HW, ex->params, ex->hist256, C, c.seg_channel, st, ex->dx_eps,
                       ex->dx_bin_count, ex->dx_bin_idx, kDxBinCap));
    // the samples next to a histogram edge or a bin centre: exact value, exact bin
    AMT_TRY(dx::exact_eval(in + (int64_t)c.seg_channel * HW, (int64_t)C * HW, (int)H, (int)W, 1.0 / 65535.0, ex->hw_hi, ex->r_hi,
                           ex->hw_lo, ex->r_lo, ex->dx_bin_count, ex->dx_bin_idx, ex->dx_bin_val, g, kDxBinCap, st));
    AMT_TRY(dx_patch(ex->dx_bin_count, ex->dx_bin_idx, ex->dx_bin_val, kDxBinCap, ex->params, C, c.seg_channel, g, pre, HW,
                     ex->hist256, retry_flags, st));
  } else {
    AMT_TRY(map_launch(dog, AMT_F64, pre, planes, HW, ex->params, ex->hist256, C, c.seg_channel, st));
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_dog_free[slot], st));
  trace_mark(ex, st, "  rest: map done", AMT_STAGE_MAP);
  // stage B: Otsu -> threshold + CCL + clear_border
  AMT_TRY(otsu_launch(ex->hist256, 0, ex->params, C, c.seg_channel, nullptr, g, thr, nullptr, 0, st));
  AMT_TRY(label_launch(pre + (int64_t)c.seg_channel * HW, 1, (int64_t)C * HW, thr, 0, g, H, W, 1, lab_thr, cnt_thr,
                       ex->label_scratch, ex->label_bytes, st));
  trace_mark(ex, st, "  rest: otsu+label(thr) done", AMT_STAGE_LABEL_THR);
  // stage C: per-cell tables over the raw channels
  AMT_TRY(region_reduce(lab_thr, in, C, (int64_t)C * HW, HW, g, H, W, c.max_labels, ex->acc, st));
  AMT_TRY(region_finalize(ex->acc, cnt_thr, C, g, c.max_labels, tab_thr, st));
  if (c.with_shape)
    AMT_TRY(region_shape(lab_thr, ex->acc, C, cnt_thr, g, H, W, c.max_labels, tab_thr, ex->shape_scratch, ex->shape_bytes, st));
  trace_mark(ex, st, "  rest: regions(thr) done", AMT_STAGE_REGIONS_THR);
  if (with_given && sg == st) AMT_TRY(given_chain(ex, in, given, g, tab_given, cnt_given, lab_given, flags, st));
  if (with_given) {
    if (sg != st) {  // join
      AMT_CUDA_TRY(cudaEventRecord(ex->ev_join, sg));
      AMT_CUDA_TRY(cudaStreamWaitEvent(st, ex->ev_join, 0));
    }
  }
  if (status != nullptr || (dx && retry_out != nullptr)) {
    fov_status_kernel<<<(unsigned)ceil_div(g, 128), 128, 0, st>>>(cnt_thr, with_given ? cnt_given : nullptr, flags,
                                                                   flags + c.chunk_fovs, ex->params, C, c.seg_channel,
                                                                   c.max_labels, g, status, retry_flags,
                                                                   dx ? retry_out : nullptr);
    AMT_LAUNCH_CHECK();
  }
  trace_mark(ex, st, "  rest: end");
  if (ex->profile) ex->profiled_chunks += 1;
  ex->chunks_issued += 1;
  return AMT_OK;
}

// int64 host masks (the reference's dtype): values beyond int32 saturate (and are then reported as out of range by
// the labelling pass), negative values become background and raise the FOV's flag
__global__ void narrow_i64_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n_per_fov, int64_t n_fov,
                                  int32_t* __restrict__ negative) {
  const int64_t n2 = n_per_fov / 2;  // pairs per FOV (n_per_fov is even)
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2 * n_fov; i += step) {
    const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(in) + i);
    int2 r;
    r.x = v.x > 2147483647ll ? 2147483647 : (v.x < 0 ? 0 : (int)v.x);
    r.y = v.y > 2147483647ll ? 2147483647 : (v.y < 0 ? 0 : (int)v.y);
    if ((v.x | v.y) < 0) negative[i / n2] = 1;
    reinterpret_cast<int2*>(out)[i] = r;
  }
}

__global__ void widen_u16_kernel(const uint16_t* __restrict__ in, int32_t* __restrict__ out, int64_t n8) {
  // 8 labels per thread: one 16-byte load, two 16-byte stores
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n8; i += step) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    int4 a, b;
    a.x = v.x & 0xffff; a.y = v.x >> 16; a.z = v.y & 0xffff; a.w = v.y >> 16;
    b.x = v.z & 0xffff; b.y = v.z >> 16; b.z = v.w & 0xffff; b.w = v.w >> 16;
    reinterpret_cast<int4*>(out)[2 * i] = a;
    reinterpret_cast<int4*>(out)[2 * i + 1] = b;
  }
}

// host_rle.cpp: host label masks -> per-row runs {value, end column} in pinned staging, on host threads
bool rle_encode_host(const void* in, int dtype, int H, int W, int g, uint2* runs, uint2* rows, int32_t* negative, int n_thr,
                     std::vector<int64_t>& first, std::vector<int64_t>& used);

// run-length staging -> int32 label image: one CTA per row; a thread finds the run of its first pixel by bisection over
// the row's run ends (in shared memory when they fit) and walks on from there, four pixels per 16-byte store
constexpr int kRleSmemRuns = 1024;
__global__ void __launch_bounds__(256) rle_decode_kernel(const uint2* __restrict__ runs, const uint2* __restrict__ rows,
                                                         int32_t* __restrict__ out, int W) {
  __shared__ uint2 s_runs[kRleSmemRuns];
  const uint2 ri = rows[blockIdx.x];
  const uint2* r = runs + ri.x;
  const int n = (int)ri.y;
  if (n <= kRleSmemRuns) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_runs[i] = __ldg(r + i);
    __syncthreads();
    r = s_runs;
  }
  int32_t* o = out + (int64_t)blockIdx.x * W;
  const bool vec = (W & 3) == 0;
  for (int x0 = threadIdx.x * 4; x0 < W; x0 += blockDim.x * 4) {
    int lo = 0, hi = n - 1;  // first run whose end lies beyond x0 (the last run ends at W)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)r[mid].y > x0) hi = mid; else lo = mid + 1;
    }
    uint2 cur = r[lo];
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      while (x < W && (int)cur.y <= x && lo + 1 < n) cur = r[++lo];
      v[k] = (int)cur.x;
    }
    if (vec) {
      *reinterpret_cast<int4*>(o + x0) = make_int4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x0 + k < W) o[x0 + k] = v[k];
    }
  }
}

static int alloc_host_slots(amt_executor* ex) {
  if (ex->host_slots) return AMT_OK;
  const amt_fov_config& c = ex->cfg;
  ex->host_rle = g_exec_host_rle && c.quantify_given_mask && c.width >= 16;
  ex->host_narrow = !ex->host_rle && g_exec_host_narrow && c.quantify_given_mask && c.given_label_dtype == AMT_I64 &&
                    c.max_label_value < 65535 && ((int64_t)c.height * c.width) % 8 == 0;
  const int64_t HW = (int64_t)c.height * c.width;
  const size_t tab = (size_t)c.chunk_fovs * AMT_TABLE_COLS(c.n_channels) * c.max_labels * sizeof(double);
  for (int s = 0; s < 2; ++s) {
    AMT_TRY(dmalloc(ex, (void**)&ex->in_slot[s], (size_t)c.chunk_fovs * c.n_channels * HW * sizeof(uint16_t)));
    AMT_TRY(dmalloc(ex, (void**)&ex->given_slot[s], (size_t)c.chunk_fovs * HW * sizeof(int32_t)));
    if (c.given_label_dtype == AMT_U16 || ex->host_narrow)
      AMT_TRY(dmalloc(ex, (void**)&ex->given16_slot[s], (size_t)c.chunk_fovs * HW * sizeof(uint16_t)));
    if (c.given_label_dtype == AMT_I64 && !ex->host_narrow)
      AMT_TRY(dmalloc(ex, (void**)&ex->given64_slot[s], (size_t)c.chunk_fovs * HW * sizeof(int64_t)));
    if (ex->host_narrow)
      AMT_CUDA_TRY(cudaMallocHost((void**)&ex->given16_host[s], (size_t)c.chunk_fovs * HW * sizeof(uint16_t)));
    if (ex->host_rle) {
      const size_t slots = (size_t)c.chunk_fovs * c.height * (c.width / 4);
      const size_t rows = (size_t)c.chunk_fovs * c.height;
      AMT_CUDA_TRY(cudaMallocHost((void**)&ex->rle_host[s], slots * sizeof(uint2)));
      AMT_CUDA_TRY(cudaMallocHost((void**)&ex->rle_rows_host[s], rows * sizeof(uint2)));
      AMT_TRY(dmalloc(ex, (void**)&ex->rle_slot[s], slots * sizeof(uint2)));
      AMT_TRY(dmalloc(ex, (void**)&ex->rle_rows_slot[s], rows * sizeof(uint2)));
      AMT_CUDA_TRY(cudaEventCreate(&ex->ev_img0[s]));
      AMT_CUDA_TRY(cudaEventCreate(&ex->ev_img1[s]));
    }
    AMT_TRY(dmalloc(ex, (void**)&ex->flag_slot[s], (size_t)3 * c.chunk_fovs * sizeof(int32_t)));
    AMT_TRY(dmalloc(ex, (void**)&ex->retry_slot[s], (size_t)c.chunk_fovs * sizeof(int32_t)));
    AMT_TRY(dmalloc(ex, (void**)&ex->status_slot[s], (size_t)c.chunk_fovs * sizeof(int32_t)));
    AMT_TRY(dmalloc(ex, (void**)&ex->tab_thr_slot[s], tab));
    AMT_TRY(dmalloc(ex, (void**)&ex->tab_given_slot[s], tab));
    AMT_TRY(dmalloc(ex, (void**)&ex->thr_slot[s], (size_t)c.chunk_fovs * sizeof(double)));
    AMT_TRY(dmalloc(ex, (void**)&ex->cnt_thr_slot[s], (size_t)c.chunk_fovs * sizeof(int32_t)));
    AMT_TRY(dmalloc(ex, (void**)&ex->cnt_given_slot[s], (size_t)c.chunk_fovs * sizeof(int32_t)));
  }
  ex->host_slots = true;
  return AMT_OK;
}

}  // namespace amt

extern "C" {

int amt_executor_create(const amt_fov_config* cfg, const double* half_w_lo_host, int r_lo, const double* half_w_hi_host,
                        int r_hi, amt_executor** out) {
  using namespace amt;
  if (!cfg || !half_w_lo_host || !half_w_hi_host || !out || r_lo < 0 || r_hi < 0) return AMT_ERR_INVALID;
  if (cfg->n_channels < 1 || cfg->n_channels > 8 || cfg->height < 1 || cfg->width < 1 || cfg->chunk_fovs < 1 ||
      cfg->max_labels < 1 || cfg->seg_channel < 0 || cfg->seg_channel >= cfg->n_channels || cfg->max_label_value < 0)
    return AMT_ERR_INVALID;
  if (cfg->given_label_dtype != 0 && cfg->given_label_dtype != AMT_I32 && cfg->given_label_dtype != AMT_U16 &&
      cfg->given_label_dtype != AMT_I64)
    return AMT_ERR_INVALID;
  if (cfg->given_label_dtype == AMT_U16 && cfg->max_label_value > 65535) return AMT_ERR_INVALID;
  if (!(cfg->bg_percentile >= 0 && cfg->bg_percentile <= 100) ||
      !(0 <= cfg->pct_lo && cfg->pct_lo < cfg->pct_hi && cfg->pct_hi <= 100))
    return AMT_ERR_INVALID;
  AMT_CUDA_TRY(cudaSetDevice(cfg->device));
  amt_executor* ex = new (std::nothrow) amt_executor();
  if (!ex) return AMT_ERR_CAPACITY;
  std::memset(ex, 0, sizeof(*ex));
  ex->cfg = *cfg;
  ex->r_lo = r_lo;
  ex->r_hi = r_hi;
  ex->trace = std::getenv("AMT_TRACE") != nullptr;
  ex->trace_events = new std::vector<cudaEvent_t>();
  ex->trace_names = new std::vector<const char*>();
  ex->trace_stage = new std::vector<int>();
  ex->trace_stream = new std::vector<int>();
  const int C = cfg->n_channels;
  const int64_t HW = (int64_t)cfg->height * cfg->width;
  const int64_t planes = (int64_t)cfg->chunk_fovs * C;
  rank_pair(HW, cfg->bg_percentile, &ex->ranks[0], &ex->ranks[1], &ex->g_bg);
  rank_pair(HW, cfg->pct_lo, &ex->ranks[2], &ex->ranks[3], &ex->g_lo);
  rank_pair(HW, cfg->pct_hi, &ex->ranks[4], &ex->ranks[5], &ex->g_hi);

  int st = AMT_OK;
  auto fail = [&](int s) {
    amt_executor_destroy(ex);
    return s;
  };
#define EX_TRY(e)                 \
  do {                            \
    st = (e);                     \
    if (st != AMT_OK) return fail(st); \
  } while (0)
#define EX_CUDA(e)                                  \
  do {                                              \
    cudaError_t _e = (e);                           \
    if (_e != cudaSuccess) {                        \
      set_last_cuda_error(_e);                      \
      return fail(AMT_ERR_CUDA);                    \
    }                                               \
  } while (0)
  int prio_lo = 0, prio_hi = 0;
  EX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  EX_CUDA(cudaStreamCreateWithPriority(&ex->s_compute, cudaStreamNonBlocking, g_exec_swap_prio ? prio_hi : prio_lo));
  EX_CUDA(cudaStreamCreateWithPriority(&ex->s_dog, cudaStreamNonBlocking, g_exec_swap_prio ? prio_lo : prio_hi));
  EX_CUDA(cudaStreamCreateWithFlags(&ex->s_in, cudaStreamNonBlocking));
  EX_CUDA(cudaStreamCreateWithFlags(&ex->s_out, cudaStreamNonBlocking));
  EX_CUDA(cudaStreamCreateWithPriority(&ex->s_given, cudaStreamNonBlocking, g_exec_swap_prio ? prio_hi : prio_lo));
  EX_CUDA(cudaEventCreateWithFlags(&ex->ev_fork, cudaEventDisableTiming));
  EX_CUDA(cudaEventCreateWithFlags(&ex->ev_join, cudaEventDisableTiming));
  EX_CUDA(cudaEventCreate(&ex->ev_start));
  EX_CUDA(cudaEventCreate(&ex->ev_stop));
  for (int s = 0; s < 2; ++s) {
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_in[s], cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_done[s], cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_out[s], cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_dog_done[s], cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_dog_free[s], cudaEventDisableTiming));
  }
  EX_TRY(dmalloc(ex, (void**)&ex->hw_lo, (size_t)(r_lo + 1) * sizeof(double)));
  EX_TRY(dmalloc(ex, (void**)&ex->hw_hi, (size_t)(r_hi + 1) * sizeof(double)));
  EX_CUDA(cudaMemcpy(ex->hw_lo, half_w_lo_host, (size_t)(r_lo + 1) * sizeof(double), cudaMemcpyHostToDevice));
  EX_CUDA(cudaMemcpy(ex->hw_hi, half_w_hi_host, (size_t)(r_hi + 1) * sizeof(double), cudaMemcpyHostToDevice));
  const size_t plane_f64 = (size_t)planes * HW * sizeof(double);
  EX_TRY(dmalloc(ex, (void**)&ex->tmp_lo, plane_f64));
  EX_TRY(dmalloc(ex, (void**)&ex->tmp_hi, plane_f64));
  for (int s = 0; s < 2; ++s) {
    EX_TRY(dmalloc(ex, (void**)&ex->dog[s], plane_f64));
    EX_TRY(dmalloc(ex, (void**)&ex->mm[s], (size_t)planes * 2 * sizeof(uint64_t)));
    EX_TRY(dmalloc(ex, (void**)&ex->buckets[s], (size_t)planes * HW * sizeof(uint16_t)));
  }
  if (!cfg->exact_all_channels && cfg->plane_filter == AMT_FILTER_TENSOR_CORE && C >= 2 &&
      amt_tcg_supported(cfg->height, cfg->width, r_hi) && cfg->height % 2 == 0 &&
      dog2d_fast(AMT_U16, planes, cfg->height, cfg->width, r_lo, r_hi) && r_lo <= 4) {
    const int ts = amt_tcg_create(half_w_hi_host, r_hi, cfg->device, &ex->tcg);
    if (ts == AMT_OK) {
      EX_TRY(dmalloc(ex, (void**)&ex->digits, amt_tcg_digit_bytes(planes, cfg->height, cfg->width)));
      if (cfg->seg_plane_filter == AMT_SEG_DECISION_EXACT) {
        ex->dx = true;
        ex->dx_eps = amt_tcg_error_bound(ex->tcg);
        EX_TRY(dmalloc(ex, (void**)&ex->dx_ranks, 6 * sizeof(int64_t)));
        EX_CUDA(cudaMemcpy(ex->dx_ranks, ex->ranks, 6 * sizeof(int64_t), cudaMemcpyHostToDevice));
        const size_t lists = (size_t)cfg->chunk_fovs * 8;
        EX_TRY(dmalloc(ex, (void**)&ex->dx_rank_u32, (lists * 2 + lists * kDxRankCap) * sizeof(uint32_t)));
        EX_TRY(dmalloc(ex, (void**)&ex->dx_rank_val, lists * kDxRankCap * sizeof(double)));
        EX_TRY(dmalloc(ex, (void**)&ex->dx_bin_count, (size_t)cfg->chunk_fovs * sizeof(uint32_t)));
        EX_TRY(dmalloc(ex, (void**)&ex->dx_bin_idx, (size_t)cfg->chunk_fovs * kDxBinCap * sizeof(uint32_t)));
        EX_TRY(dmalloc(ex, (void**)&ex->dx_bin_val, (size_t)cfg->chunk_fovs * kDxBinCap * sizeof(double)));
      }
    } else if (ts != AMT_ERR_UNSUPPORTED) {
      return fail(ts);
    }
  }
  EX_TRY(dmalloc(ex, (void**)&ex->pre, plane_f64));
  EX_TRY(dmalloc(ex, (void**)&ex->stats, (size_t)planes * 6 * sizeof(double)));
  EX_TRY(dmalloc(ex, (void**)&ex->params, (size_t)planes * sizeof(amt_map_params)));
  ex->sel_bytes = amt_select_f64_scratch_bytes(planes, HW);
  EX_TRY(dmalloc(ex, &ex->sel_scratch, ex->sel_bytes));
  EX_TRY(dmalloc(ex, (void**)&ex->hist256, (size_t)cfg->chunk_fovs * 256 * sizeof(uint32_t)));
  EX_TRY(dmalloc(ex, (void**)&ex->thr, (size_t)cfg->chunk_fovs * sizeof(double)));
  EX_TRY(dmalloc(ex, (void**)&ex->flags_dev, (size_t)3 * cfg->chunk_fovs * sizeof(int32_t)));
  EX_TRY(dmalloc(ex, (void**)&ex->lab_thr, (size_t)cfg->chunk_fovs * HW * sizeof(int32_t)));
  EX_TRY(dmalloc(ex, (void**)&ex->lab_given, (size_t)cfg->chunk_fovs * HW * sizeof(int32_t)));
  ex->label_bytes = amt_label_scratch_bytes(cfg->chunk_fovs, cfg->height, cfg->width, cfg->max_label_value);
  EX_TRY(dmalloc(ex, &ex->label_scratch, ex->label_bytes));
  EX_TRY(dmalloc(ex, (void**)&ex->acc,
                 (size_t)cfg->chunk_fovs * AMT_ACC_FIELDS(C) * cfg->max_labels * sizeof(uint64_t)));
  if (cfg->with_shape) {
    ex->shape_bytes = amt_region_shape_scratch_bytes(cfg->chunk_fovs, cfg->height, cfg->width, cfg->max_labels);
    EX_TRY(dmalloc(ex, &ex->shape_scratch, ex->shape_bytes));
  }
  if (cfg->quantify_given_mask) {
    EX_TRY(dmalloc(ex, &ex->label_scratch_g, ex->label_bytes));
    EX_TRY(dmalloc(ex, (void**)&ex->acc_g,
                   (size_t)cfg->chunk_fovs * AMT_ACC_FIELDS(C) * cfg->max_labels * sizeof(uint64_t)));
    if (cfg->with_shape) EX_TRY(dmalloc(ex, &ex->shape_scratch_g, ex->shape_bytes));
  }
#undef EX_TRY
#undef EX_CUDA
  *out = ex;
  return AMT_OK;
}

void amt_executor_destroy(amt_executor* ex) {
  if (!ex) return;
  cudaSetDevice(ex->cfg.device);
  cudaDeviceSynchronize();
  if (ex->tcg) amt_tcg_destroy(ex->tcg);
  if (ex->retry_host) cudaFreeHost(ex->retry_host);
  for (int s2 = 0; s2 < 2; ++s2) {
    if (ex->given16_host[s2]) cudaFreeHost(ex->given16_host[s2]);
    if (ex->rle_host[s2]) cudaFreeHost(ex->rle_host[s2]);
    if (ex->rle_rows_host[s2]) cudaFreeHost(ex->rle_rows_host[s2]);
    if (ex->rle_slot[s2]) cudaFree(ex->rle_slot[s2]);
    if (ex->rle_rows_slot[s2]) cudaFree(ex->rle_rows_slot[s2]);
    if (ex->ev_img0[s2]) cudaEventDestroy(ex->ev_img0[s2]);
    if (ex->ev_img1[s2]) cudaEventDestroy(ex->ev_img1[s2]);
  }
  std::free(ex->neg_host);
  void* bufs[] = {ex->dx_ranks, ex->dx_rank_u32, ex->dx_rank_val, ex->dx_bin_count, ex->dx_bin_idx, ex->dx_bin_val, ex->retry_dev,
                  ex->flags_dev, ex->digits, ex->hw_lo, ex->hw_hi, ex->tmp_lo, ex->tmp_hi, ex->dog[0], ex->dog[1], ex->pre, ex->mm[0], ex->mm[1],
                  ex->buckets[0], ex->buckets[1],
                  ex->stats, ex->params,
                  ex->sel_scratch, ex->hist256, ex->thr, ex->lab_thr, ex->lab_given, ex->label_scratch, ex->acc,
                  ex->shape_scratch, ex->label_scratch_g, ex->acc_g, ex->shape_scratch_g};
  for (void* b : bufs)
    if (b) cudaFree(b);
  for (int s = 0; s < 2; ++s) {
    void* sb[] = {ex->retry_slot[s], ex->given64_slot[s], ex->flag_slot[s], ex->status_slot[s], ex->in_slot[s], ex->given_slot[s], ex->given16_slot[s], ex->tab_thr_slot[s], ex->tab_given_slot[s], ex->thr_slot[s],
                  ex->cnt_thr_slot[s], ex->cnt_given_slot[s]};
    for (void* b : sb)
      if (b) cudaFree(b);
    if (ex->ev_in[s]) cudaEventDestroy(ex->ev_in[s]);
    if (ex->ev_done[s]) cudaEventDestroy(ex->ev_done[s]);
    if (ex->ev_out[s]) cudaEventDestroy(ex->ev_out[s]);
    if (ex->ev_dog_done[s]) cudaEventDestroy(ex->ev_dog_done[s]);
    if (ex->ev_dog_free[s]) cudaEventDestroy(ex->ev_dog_free[s]);
  }
  if (ex->ev_start) cudaEventDestroy(ex->ev_start);
  if (ex->ev_stop) cudaEventDestroy(ex->ev_stop);
  if (ex->s_compute) cudaStreamDestroy(ex->s_compute);
  if (ex->s_dog) cudaStreamDestroy(ex->s_dog);
  if (ex->s_in) cudaStreamDestroy(ex->s_in);
  if (ex->s_out) cudaStreamDestroy(ex->s_out);
  if (ex->s_given) cudaStreamDestroy(ex->s_given);
  if (ex->ev_fork) cudaEventDestroy(ex->ev_fork);
  if (ex->ev_join) cudaEventDestroy(ex->ev_join);
  delete ex->trace_events;
  delete ex->trace_names;
  delete ex->trace_stage;
  delete ex->trace_stream;
  delete ex;
}

size_t amt_executor_device_bytes(const amt_executor* ex) { return ex ? ex->device_bytes : 0; }
int amt_executor_uses_tensor_cores(const amt_executor* ex) { return ex && ex->tcg != nullptr ? 1 : 0; }

int amt_executor_set_profiling(amt_executor* ex, int enable) {
  if (!ex) return AMT_ERR_INVALID;
  ex->profile = enable != 0;
  for (int i = 0; i < AMT_N_STAGES; ++i) ex->stage_ms[i] = 0.0;
  ex->profiled_chunks = 0;
  return AMT_OK;
}

int amt_executor_stage_ms(const amt_executor* ex, double* stage_ms, int64_t* n_chunks) {
  if (!ex || !stage_ms) return AMT_ERR_INVALID;
  for (int i = 0; i < AMT_N_STAGES; ++i) stage_ms[i] = ex->stage_ms[i];
  if (n_chunks) *n_chunks = ex->profiled_chunks;
  return AMT_OK;
}

// Decision-exact mode, device-resident batch: fields of view whose candidate lists overflowed (retry flag) are
// recomputed one by one with the float64 strip kernels, into the same output slots.  Runs inside amt_executor_sync.
static int retry_device(amt_executor* ex) {
  using namespace amt;
  if (!ex->last.pending) return AMT_OK;
  ex->last.pending = false;
  const amt_fov_config& c = ex->cfg;
  const int64_t n_fov = ex->last.n_fov;
  std::vector<int32_t> flags((size_t)n_fov);
  AMT_CUDA_TRY(cudaMemcpy(flags.data(), ex->retry_dev, (size_t)n_fov * sizeof(int32_t), cudaMemcpyDeviceToHost));
  const int C = c.n_channels;
  const int64_t HW = (int64_t)c.height * c.width;
  const int64_t tab = (int64_t)AMT_TABLE_COLS(C) * c.max_labels;
  bool any = false;
  ex->force_exact = true;
  int rc = AMT_OK;
  for (int64_t i = 0; i < n_fov && rc == AMT_OK; ++i) {
    if (!flags[(size_t)i]) continue;
    any = true;
    ex->retries += 1;
    rc = enqueue_dog(ex, ex->last.fovs + i * C * HW, 1, nullptr);
    if (rc != AMT_OK) break;
    if (cudaMemsetAsync(ex->flags_dev, 0, (size_t)3 * c.chunk_fovs * sizeof(int32_t), ex->s_compute) != cudaSuccess) {
      rc = AMT_ERR_CUDA;
      break;
    }
    rc = process_chunk(ex, ex->last.fovs + i * C * HW, ex->last.given ? ex->last.given + i * HW : nullptr, 1,
                       ex->last.tables_thr + i * tab, ex->last.counts_thr + i,
                       ex->last.tables_given ? ex->last.tables_given + i * tab : nullptr,
                       ex->last.counts_given ? ex->last.counts_given + i : nullptr,
                       ex->last.thresholds ? ex->last.thresholds + i : nullptr,
                       ex->last.labels_thr ? ex->last.labels_thr + i * HW : nullptr,
                       ex->last.labels_given ? ex->last.labels_given + i * HW : nullptr,
                       ex->last.preprocessed ? ex->last.preprocessed + i * C * HW : nullptr, ex->flags_dev,
                       ex->last.status ? ex->last.status + i : nullptr, nullptr);
  }
  ex->force_exact = false;
  if (rc != AMT_OK) return rc;
  if (any) {
    AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_dog));
    AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_compute));
  }
  return AMT_OK;
}

int amt_executor_run_device(amt_executor* ex, const uint16_t* fovs, const int32_t* given_labels, int64_t n_fov,
                            double* tables_thr, int32_t* counts_thr, double* tables_given, int32_t* counts_given,
                            double* thresholds, int32_t* labels_thr, int32_t* labels_given, double* preprocessed,
                            int32_t* status) {
  using namespace amt;
  if (!ex || !fovs || !tables_thr || !counts_thr || n_fov <= 0) return AMT_ERR_INVALID;
  const amt_fov_config& c = ex->cfg;
  if (c.quantify_given_mask && given_labels && (!tables_given || !counts_given)) return AMT_ERR_INVALID;
  AMT_CUDA_TRY(cudaSetDevice(c.device));
  if (ex->last.pending) AMT_TRY(amt_executor_sync(ex));  // a batch whose retry check has not run yet
  const int C = c.n_channels;
  const int64_t HW = (int64_t)c.height * c.width;
  const int64_t tab = (int64_t)AMT_TABLE_COLS(C) * c.max_labels;
  const bool dx = ex->tcg != nullptr && g_exec_tc && ex->dx;
  if (dx && n_fov > ex->retry_cap) {
    if (ex->retry_dev) cudaFree(ex->retry_dev);
    ex->retry_dev = nullptr;
    AMT_CUDA_TRY(cudaMalloc((void**)&ex->retry_dev, (size_t)n_fov * sizeof(int32_t)));
    ex->retry_cap = n_fov;
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_start, ex->s_compute));
  AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_dog, ex->ev_start, 0));
  for (int64_t f0 = 0; f0 < n_fov; f0 += c.chunk_fovs) {
    const int g = (int)((n_fov - f0 < c.chunk_fovs) ? n_fov - f0 : c.chunk_fovs);
    AMT_TRY(enqueue_dog(ex, fovs + f0 * C * HW, g, nullptr));
    AMT_CUDA_TRY(cudaMemsetAsync(ex->flags_dev, 0, (size_t)3 * c.chunk_fovs * sizeof(int32_t), ex->s_compute));
    AMT_TRY(process_chunk(ex, fovs + f0 * C * HW, given_labels ? given_labels + f0 * HW : nullptr, g,
                          tables_thr + f0 * tab, counts_thr + f0, tables_given ? tables_given + f0 * tab : nullptr,
                          counts_given ? counts_given + f0 : nullptr, thresholds ? thresholds + f0 : nullptr,
                          labels_thr ? labels_thr + f0 * HW : nullptr, labels_given ? labels_given + f0 * HW : nullptr,
                          preprocessed ? preprocessed + f0 * C * HW : nullptr, ex->flags_dev,
                          status ? status + f0 : nullptr, dx ? ex->retry_dev + f0 : nullptr));
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_stop, ex->s_compute));
  if (dx) {
    ex->last.pending = true;
    ex->last.fovs = fovs, ex->last.given = given_labels, ex->last.n_fov = n_fov;
    ex->last.tables_thr = tables_thr, ex->last.tables_given = tables_given, ex->last.thresholds = thresholds;
    ex->last.preprocessed = preprocessed, ex->last.counts_thr = counts_thr, ex->last.counts_given = counts_given;
    ex->last.labels_thr = labels_thr, ex->last.labels_given = labels_given, ex->last.status = status;
  }
  trace_dump(ex);
  return AMT_OK;
}

// int64 label masks -> uint16 in pinned staging, on host threads: negative values become background and raise the FOV's
// flag, values above 65534 saturate to 65535 (beyond max_label_value by construction: reported as out of range by the
// labelling pass, like any value above max_label_value)
static void narrow_i64_host(const int64_t* in, uint16_t* out, int64_t n_per_fov, int g, int32_t* negative) {
  const int64_t total = n_per_fov * g;
  int n_thr = amt::g_exec_host_threads;
  const int hw_thr = (int)std::thread::hardware_concurrency();
  if (hw_thr > 0 && n_thr > hw_thr) n_thr = hw_thr;
  if (n_thr < 1 || total < (1 << 20)) n_thr = 1;
  std::vector<int64_t> neg_or((size_t)n_thr * g, 0);
  auto work = [&](int t) {
    const int64_t lo = total * t / n_thr, hi = total * (t + 1) / n_thr;
    int64_t i = lo;
    while (i < hi) {
      const int64_t fov = i / n_per_fov;
      const int64_t end = (fov + 1) * n_per_fov < hi ? (fov + 1) * n_per_fov : hi;
      int64_t acc = 0;
      // eight labels at a time: when none of them has a bit above the 16th (the usual case) they are simply packed into
      // two 64-bit words (4 loads, 3 shifts / ORs and one 8-byte store per four labels instead of compare / select / 2-byte
      // store per label: the narrowing has to keep up with PCIe on a few host cores)
      for (; i + 8 <= end && ((uintptr_t)(out + i) & 7) == 0; i += 8) {
        const uint64_t a0 = (uint64_t)in[i], a1 = (uint64_t)in[i + 1], a2 = (uint64_t)in[i + 2], a3 = (uint64_t)in[i + 3];
        const uint64_t a4 = (uint64_t)in[i + 4], a5 = (uint64_t)in[i + 5], a6 = (uint64_t)in[i + 6], a7 = (uint64_t)in[i + 7];
        const uint64_t any = a0 | a1 | a2 | a3 | a4 | a5 | a6 | a7;
        if ((any >> 16) == 0) {
          uint64_t* o = reinterpret_cast<uint64_t*>(out + i);
          o[0] = a0 | (a1 << 16) | (a2 << 32) | (a3 << 48);
          o[1] = a4 | (a5 << 16) | (a6 << 32) | (a7 << 48);
        } else {
          for (int k = 0; k < 8; ++k) {
            const int64_t v = in[i + k];
            acc |= v;
            out[i + k] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
          }
        }
      }
      for (; i < end; ++i) {
        const int64_t v = in[i];
        acc |= v;
        out[i] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
      }
      neg_or[(size_t)t * g + fov] |= acc;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  if (negative != nullptr)
    for (int f = 0; f < g; ++f) {
      int64_t acc = 0;
      for (int t = 0; t < n_thr; ++t) acc |= neg_or[(size_t)t * g + f];
      if (acc < 0) negative[f] = 1;
    }
}

// plain label masks of the chunk's FOVs [fa, fb) -> given_slot[s] (H2D on s_in, converted on the device)
static int upload_plain_masks(amt_executor* ex, int s, const void* given_labels_host, int64_t f0, int fa, int fb) {
  using namespace amt;
  const amt_fov_config& c = ex->cfg;
  const int64_t HW = (int64_t)c.height * c.width;
  const int n = fb - fa;
  if (n <= 0) return AMT_OK;
  int32_t* out = ex->given_slot[s] + (int64_t)fa * HW;
  if (ex->host_narrow) {
    // the staging slot's previous H2D must have left the pinned buffer before host threads overwrite it
    if (ex->slot_uploaded[s]) AMT_CUDA_TRY(cudaEventSynchronize(ex->ev_in[s]));
    ex->last_h2d_bytes += (int64_t)n * HW * sizeof(uint16_t);
    narrow_i64_host((const int64_t*)given_labels_host + (f0 + fa) * HW, ex->given16_host[s] + (int64_t)fa * HW, HW, n,
                    ex->neg_host ? ex->neg_host + f0 + fa : nullptr);
    AMT_CUDA_TRY(cudaMemcpyAsync(ex->given16_slot[s] + (int64_t)fa * HW, ex->given16_host[s] + (int64_t)fa * HW,
                                 (size_t)n * HW * sizeof(uint16_t), cudaMemcpyHostToDevice, ex->s_in));
    widen_u16_kernel<<<kNumSMs * 8, 256, 0, ex->s_in>>>(ex->given16_slot[s] + (int64_t)fa * HW, out, (int64_t)n * HW / 8);
    AMT_LAUNCH_CHECK();
  } else if (c.given_label_dtype == AMT_I64) {
    ex->last_h2d_bytes += (int64_t)n * HW * sizeof(int64_t);
    AMT_CUDA_TRY(cudaMemcpyAsync(ex->given64_slot[s] + (int64_t)fa * HW, (const int64_t*)given_labels_host + (f0 + fa) * HW,
                                 (size_t)n * HW * sizeof(int64_t), cudaMemcpyHostToDevice, ex->s_in));
    narrow_i64_kernel<<<kNumSMs * 8, 256, 0, ex->s_in>>>(ex->given64_slot[s] + (int64_t)fa * HW, out, HW, n,
                                                         ex->flag_slot[s] + c.chunk_fovs + fa);
    AMT_LAUNCH_CHECK();
  } else if (c.given_label_dtype == AMT_U16) {
    ex->last_h2d_bytes += (int64_t)n * HW * sizeof(uint16_t);
    AMT_CUDA_TRY(cudaMemcpyAsync(ex->given16_slot[s] + (int64_t)fa * HW, (const uint16_t*)given_labels_host + (f0 + fa) * HW,
                                 (size_t)n * HW * sizeof(uint16_t), cudaMemcpyHostToDevice, ex->s_in));
    widen_u16_kernel<<<kNumSMs * 8, 256, 0, ex->s_in>>>(ex->given16_slot[s] + (int64_t)fa * HW, out, (int64_t)n * HW / 8);
    AMT_LAUNCH_CHECK();
  } else {
    ex->last_h2d_bytes += (int64_t)n * HW * sizeof(int32_t);
    AMT_CUDA_TRY(cudaMemcpyAsync(out, (const int32_t*)given_labels_host + (f0 + fa) * HW, (size_t)n * HW * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, ex->s_in));
  }
  ex->last_plain_masks += n;
  return AMT_OK;
}

// how many of a chunk's g masks take the run-length route: the host threads need rle_a seconds per mask, PCIe moves rle_B
// bytes per second; n masks encoded and g - n sent plain right behind the images finish together when
//   rle_a n = (images + n rle_c + (g - n) mask_bytes) / rle_B
static int rle_masks_of_chunk(const amt_executor* ex, int g, int64_t img_bytes, int64_t mask_bytes, int n_thr) {
  if (amt::g_exec_rle_share >= 0) return (int)(((int64_t)g * amt::g_exec_rle_share + 50) / 100);
  if (ex->rle_a <= 0 || ex->rle_B <= 0) return n_thr >= 8 ? g : (g + 1) / 2;  // nothing measured yet
  const double n = ((double)img_bytes + (double)g * mask_bytes) / (ex->rle_a * ex->rle_B + (double)mask_bytes - ex->rle_c);
  const int r = (int)(n + 0.5);
  return r < 0 ? 0 : (r > g ? g : r);
}

// H2D of one chunk of a host-fed batch into staging slot s (on s_in), label masks converted on the device
static int upload_chunk(amt_executor* ex, int s, const uint16_t* fovs_host, const void* given_labels_host, int64_t f0, int g) {
  using namespace amt;
  const amt_fov_config& c = ex->cfg;
  const int C = c.n_channels;
  const int64_t HW = (int64_t)c.height * c.width;
  const bool given = c.quantify_given_mask && given_labels_host;
  const int64_t img_bytes = (int64_t)g * C * HW * sizeof(uint16_t);
  const bool rle = given && ex->host_rle;
  if (rle && ex->slot_uploaded[s]) {
    // the staging slot's previous H2D must have left the pinned buffer before host threads overwrite it; its image copy
    // gives the PCIe rate this process sees right now (every rank of a box copies at the same time)
    AMT_CUDA_TRY(cudaEventSynchronize(ex->ev_in[s]));
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ex->ev_img0[s], ex->ev_img1[s]) == cudaSuccess && ms > 0.0f) {
      const double B = (double)ex->slot_img_bytes[s] / (1e-3 * ms);
      ex->rle_B = ex->rle_B > 0 ? 0.5 * ex->rle_B + 0.5 * B : B;
    }
  }
  if (rle) AMT_CUDA_TRY(cudaEventRecord(ex->ev_img0[s], ex->s_in));
  AMT_CUDA_TRY(cudaMemcpyAsync(ex->in_slot[s], fovs_host + f0 * C * HW, (size_t)img_bytes, cudaMemcpyHostToDevice, ex->s_in));
  if (rle) AMT_CUDA_TRY(cudaEventRecord(ex->ev_img1[s], ex->s_in));
  ex->slot_img_bytes[s] = img_bytes;
  AMT_CUDA_TRY(cudaMemsetAsync(ex->flag_slot[s], 0, (size_t)3 * c.chunk_fovs * sizeof(int32_t), ex->s_in));
  ex->last_h2d_bytes += img_bytes;
  if (rle) {
    const int hw_thr = (int)std::thread::hardware_concurrency();
    const int n_thr = hw_thr > 0 && g_exec_host_threads > hw_thr ? hw_thr : g_exec_host_threads;
    const int dt = c.given_label_dtype == AMT_I64 || c.given_label_dtype == AMT_U16 ? c.given_label_dtype : AMT_I32;
    const int64_t mask_bytes = HW * (dt == AMT_I64 ? 8 : (dt == AMT_U16 ? 2 : 4));
    const int n_enc = rle_masks_of_chunk(ex, g, img_bytes, mask_bytes, n_thr);
    // the masks that cross plain go first: PCIe carries them while the host threads encode the others
    AMT_TRY(upload_plain_masks(ex, s, given_labels_host, f0, n_enc, g));
    if (n_enc > 0) {
      std::vector<int64_t> first, used;
      int32_t* neg = ex->neg_host ? ex->neg_host + f0 : nullptr;
      const auto t0 = std::chrono::steady_clock::now();
      const bool ok = rle_encode_host((const char*)given_labels_host + (size_t)f0 * mask_bytes, dt, c.height, c.width, n_enc,
                                      ex->rle_host[s], ex->rle_rows_host[s], neg, n_thr, first, used);
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (ok) {
        int64_t bytes = 0;
        for (size_t t = 0; t < first.size(); ++t) {
          if (used[t] == 0) continue;
          AMT_CUDA_TRY(cudaMemcpyAsync(ex->rle_slot[s] + first[t], ex->rle_host[s] + first[t], (size_t)used[t] * sizeof(uint2),
                                       cudaMemcpyHostToDevice, ex->s_in));
          bytes += used[t] * (int64_t)sizeof(uint2);
        }
        const size_t rows = (size_t)n_enc * c.height;
        AMT_CUDA_TRY(cudaMemcpyAsync(ex->rle_rows_slot[s], ex->rle_rows_host[s], rows * sizeof(uint2), cudaMemcpyHostToDevice, ex->s_in));
        bytes += (int64_t)(rows * sizeof(uint2));
        rle_decode_kernel<<<(unsigned)rows, 256, 0, ex->s_in>>>(ex->rle_slot[s], ex->rle_rows_slot[s], ex->given_slot[s], c.width);
        AMT_LAUNCH_CHECK();
        ex->last_h2d_bytes += bytes;
        ex->last_rle_masks += n_enc;
        const double a = sec / n_enc, cb = (double)bytes / n_enc;
        ex->rle_a = ex->rle_a > 0 ? 0.5 * ex->rle_a + 0.5 * a : a;
        ex->rle_c = ex->rle_c > 0 ? 0.5 * ex->rle_c + 0.5 * cb : cb;
      } else {
        // too ragged for the run slots: these masks cross plain as well
        ex->rle_fallback_chunks += 1;
        if (neg) std::memset(neg, 0, (size_t)n_enc * sizeof(int32_t));  // the plain route reports negatives itself
        AMT_TRY(upload_plain_masks(ex, s, given_labels_host, f0, 0, n_enc));
      }
    }
  } else if (given) {
    AMT_TRY(upload_plain_masks(ex, s, given_labels_host, f0, 0, g));
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_in[s], ex->s_in));
  ex->slot_uploaded[s] = true;
  return AMT_OK;
}

// D2H of one chunk's results from staging slot s (on s_out, after ev_done[s])
static int download_chunk(amt_executor* ex, int s, int64_t f0, int g, bool given, double* tables_thr_host,
                          int32_t* counts_thr_host, double* tables_given_host, int32_t* counts_given_host,
                          double* thresholds_host, int32_t* status_host, bool with_retry) {
  using namespace amt;
  const amt_fov_config& c = ex->cfg;
  const int64_t tab = (int64_t)AMT_TABLE_COLS(c.n_channels) * c.max_labels;
  AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_out, ex->ev_done[s], 0));
  AMT_CUDA_TRY(cudaMemcpyAsync(tables_thr_host + f0 * tab, ex->tab_thr_slot[s], (size_t)g * tab * sizeof(double),
                               cudaMemcpyDeviceToHost, ex->s_out));
  AMT_CUDA_TRY(cudaMemcpyAsync(counts_thr_host + f0, ex->cnt_thr_slot[s], (size_t)g * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, ex->s_out));
  if (given) {
    AMT_CUDA_TRY(cudaMemcpyAsync(tables_given_host + f0 * tab, ex->tab_given_slot[s], (size_t)g * tab * sizeof(double),
                                 cudaMemcpyDeviceToHost, ex->s_out));
    AMT_CUDA_TRY(cudaMemcpyAsync(counts_given_host + f0, ex->cnt_given_slot[s], (size_t)g * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, ex->s_out));
  }
  if (thresholds_host)
    AMT_CUDA_TRY(cudaMemcpyAsync(thresholds_host + f0, ex->thr_slot[s], (size_t)g * sizeof(double), cudaMemcpyDeviceToHost,
                                 ex->s_out));
  if (status_host)
    AMT_CUDA_TRY(cudaMemcpyAsync(status_host + f0, ex->status_slot[s], (size_t)g * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 ex->s_out));
  if (with_retry)
    AMT_CUDA_TRY(cudaMemcpyAsync(ex->retry_host + f0, ex->retry_slot[s], (size_t)g * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 ex->s_out));
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_out[s], ex->s_out));
  return AMT_OK;
}

static int sync_all(amt_executor* ex) {
  AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_in));
  AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_dog));
  AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_compute));
  AMT_CUDA_TRY(cudaStreamSynchronize(ex->s_out));
  return AMT_OK;
}

int amt_executor_run_host(amt_executor* ex, const uint16_t* fovs_host, const void* given_labels_host, int64_t n_fov,
                          double* tables_thr_host, int32_t* counts_thr_host, double* tables_given_host,
                          int32_t* counts_given_host, double* thresholds_host, int32_t* status_host) {
  using namespace amt;
  if (!ex || !fovs_host || !tables_thr_host || !counts_thr_host || n_fov <= 0) return AMT_ERR_INVALID;
  const amt_fov_config& c = ex->cfg;
  const bool given = c.quantify_given_mask && given_labels_host;
  if (given && (!tables_given_host || !counts_given_host)) return AMT_ERR_INVALID;
  if (c.given_label_dtype == AMT_U16 && ((int64_t)c.height * c.width) % 8 != 0) return AMT_ERR_UNSUPPORTED;
  if (c.given_label_dtype == AMT_I64 && ((int64_t)c.height * c.width) % 2 != 0) return AMT_ERR_UNSUPPORTED;
  AMT_CUDA_TRY(cudaSetDevice(c.device));
  if (ex->last.pending) AMT_TRY(amt_executor_sync(ex));
  AMT_TRY(alloc_host_slots(ex));
  ex->slot_uploaded[0] = ex->slot_uploaded[1] = false;
  ex->last_h2d_bytes = 0, ex->rle_fallback_chunks = 0, ex->last_rle_masks = 0, ex->last_plain_masks = 0;
  if ((ex->host_narrow || ex->host_rle) && given) {
    if (n_fov > ex->neg_host_cap) {
      std::free(ex->neg_host);
      ex->neg_host = (int32_t*)std::malloc((size_t)n_fov * sizeof(int32_t));
      ex->neg_host_cap = ex->neg_host ? n_fov : 0;
      if (!ex->neg_host) return AMT_ERR_CAPACITY;
    }
    std::memset(ex->neg_host, 0, (size_t)n_fov * sizeof(int32_t));
  }
  const bool dx = ex->tcg != nullptr && g_exec_tc && ex->dx && !g_exec_copy_only;
  if (dx && n_fov > ex->retry_host_cap) {
    if (ex->retry_host) cudaFreeHost(ex->retry_host);
    ex->retry_host = nullptr;
    AMT_CUDA_TRY(cudaMallocHost((void**)&ex->retry_host, (size_t)n_fov * sizeof(int32_t)));
    ex->retry_host_cap = n_fov;
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_start, ex->s_compute));
  AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_dog, ex->ev_start, 0));
  auto issue = [&](int64_t chunk, int64_t f0, int g) -> int {
    const int s = (int)(chunk & 1);
    // input slot s is free once the compute that last read it has finished
    if (chunk >= 2) AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_in, ex->ev_done[s], 0));
    AMT_TRY(upload_chunk(ex, s, fovs_host, given_labels_host, f0, g));
    // output slot s is free once its previous D2H has finished
    AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_compute, ex->ev_in[s], 0));
    if (chunk >= 2) AMT_CUDA_TRY(cudaStreamWaitEvent(ex->s_compute, ex->ev_out[s], 0));
    if (!g_exec_copy_only) {
      AMT_TRY(enqueue_dog(ex, ex->in_slot[s], g, ex->ev_in[s]));
      AMT_TRY(process_chunk(ex, ex->in_slot[s], given ? ex->given_slot[s] : nullptr, g, ex->tab_thr_slot[s],
                            ex->cnt_thr_slot[s], ex->tab_given_slot[s], ex->cnt_given_slot[s], ex->thr_slot[s], nullptr,
                            nullptr, nullptr, ex->flag_slot[s], ex->status_slot[s], ex->retry_slot[s]));
    }
    AMT_CUDA_TRY(cudaEventRecord(ex->ev_done[s], ex->s_compute));
    return download_chunk(ex, s, f0, g, given, tables_thr_host, counts_thr_host, tables_given_host, counts_given_host,
                          thresholds_host, status_host, dx && !ex->force_exact);
  };
  int64_t chunk = 0;
  for (int64_t f0 = 0; f0 < n_fov; f0 += c.chunk_fovs, ++chunk) {
    const int g = (int)((n_fov - f0 < c.chunk_fovs) ? n_fov - f0 : c.chunk_fovs);
    AMT_TRY(issue(chunk, f0, g));
  }
  AMT_CUDA_TRY(cudaEventRecord(ex->ev_stop, ex->s_compute));
  AMT_TRY(sync_all(ex));
  if (dx) {
    // fields of view whose candidate lists overflowed: once more, one by one, with the float64 strip kernels
    ex->force_exact = true;
    int rc = AMT_OK;
    for (int64_t i = 0; i < n_fov && rc == AMT_OK; ++i) {
      if (!ex->retry_host[i]) continue;
      ex->retries += 1;
      rc = issue(chunk, i, 1);  // every stream is idle: the slot-reuse waits are satisfied at once
      if (rc == AMT_OK) rc = sync_all(ex);
      ++chunk;
    }
    ex->force_exact = false;
    AMT_TRY(rc);
  }
  if ((ex->host_narrow || ex->host_rle) && given && status_host != nullptr && !g_exec_copy_only)
    for (int64_t i = 0; i < n_fov; ++i)
      if (ex->neg_host[i]) status_host[i] |= AMT_FOV_GIVEN_NEGATIVE;
  return AMT_OK;
}

int amt_executor_sync(amt_executor* ex) {
  using namespace amt;
  if (!ex) return AMT_ERR_INVALID;
  AMT_CUDA_TRY(cudaSetDevice(ex->cfg.device));
  AMT_TRY(sync_all(ex));
  return retry_device(ex);
}

int64_t amt_executor_retry_count(const amt_executor* ex) { return ex ? ex->retries : -1; }
int amt_rle_encode_host(const void* labels_host, int dtype, int32_t n_fov, int32_t height, int32_t width, int32_t n_threads,
                        uint32_t* runs, uint32_t* rows, int32_t* negative, int64_t* n_runs) {
  using namespace amt;
  if (!labels_host || !runs || !rows || n_fov < 1 || height < 1 || width < 16 || n_threads < 1) return AMT_ERR_INVALID;
  std::vector<int64_t> first, used;
  if (dtype != AMT_I64 && dtype != AMT_U16 && dtype != AMT_I32) return AMT_ERR_INVALID;
  if (negative) std::memset(negative, 0, (size_t)n_fov * sizeof(int32_t));
  if (!rle_encode_host(labels_host, dtype, height, width, n_fov, (uint2*)runs, (uint2*)rows, negative, n_threads, first, used))
    return AMT_ERR_CAPACITY;
  if (n_runs) {
    *n_runs = 0;
    for (int64_t u : used) *n_runs += u;
  }
  return AMT_OK;
}

int64_t amt_executor_last_h2d_bytes(const amt_executor* ex) { return ex ? ex->last_h2d_bytes : -1; }
int64_t amt_executor_last_plain_mask_chunks(const amt_executor* ex) { return ex ? ex->rle_fallback_chunks : -1; }
int64_t amt_executor_last_rle_masks(const amt_executor* ex) { return ex ? ex->last_rle_masks : -1; }
int amt_executor_decision_exact(const amt_executor* ex) { return ex && ex->dx ? 1 : 0; }

float amt_executor_last_ms(amt_executor* ex) {
  if (!ex) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(ex->ev_stop) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, ex->ev_start, ex->ev_stop) != cudaSuccess) return -1.0f;
  return ms;
}

}  // extern "C"
