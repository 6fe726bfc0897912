// Library bookkeeping: version, status strings, last CUDA error, launch counter.
#include <atomic>

#include "common.cuh"

namespace amt {

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
static std::atomic<uint64_t> g_launches{0};

// grid cap, in CTAs per SM, of the grid-stride streaming passes (sel_hist, map): 8 fills the machine;
// 1-2 keeps them fully resident in what the DoG CTAs of the other stream leave free
int g_pass_ctas = 8;

void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace amt

namespace amt {

// FP64 issue-rate probe: 8 independent non-FMA chains per thread (DMUL + DADD per step), the
// instruction mix of the Gaussian inner loop without its memory traffic.  Used by bench.py to
// measure the DP-pipe roofline the sigma=16 filter is bound by.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double a, double b, double* __restrict__ out) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the chains alive
}

}  // namespace amt

extern "C" {

// Launches the probe on `stream`; *dp_instructions receives the number of DP instructions
// (DMUL + DADD, thread-level) one launch executes.
int amt_fp64_probe(int iters, double* scratch, uint64_t* dp_instructions, amt_stream_t stream) {
  using namespace amt;
  if (iters <= 0 || !scratch) return AMT_ERR_INVALID;
  const int blocks = kNumSMs * 8;
  fp64_probe_kernel<<<blocks, 256, 0, as_stream(stream)>>>(iters, 0.999999, 1e-9, scratch);
  AMT_LAUNCH_CHECK();
  if (dp_instructions) *dp_instructions = (uint64_t)blocks * 256ull * 8ull * 2ull * (uint64_t)iters;
  return AMT_OK;
}

// Pinned (page-locked) host staging for callers that do not go through torch: decoded frames are written
// here once and read by the copy engine.  write_combined = 1 asks for write-combined pages (no CPU cache
// snooping on the device's reads; slow for the CPU to read back, so only for buffers the host just fills).
int amt_host_alloc(size_t bytes, int write_combined, void** out) {
  if (!out || bytes == 0) return AMT_ERR_INVALID;
  *out = nullptr;
  AMT_CUDA_TRY(cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return AMT_OK;
}

int amt_host_free(void* ptr) {
  if (!ptr) return AMT_OK;
  AMT_CUDA_TRY(cudaFreeHost(ptr));
  return AMT_OK;
}

int amt_version(void) { return 100; }  // 0.1.0

const char* amt_strerror(int status) {
  switch (status) {
    case AMT_OK: return "ok";
    case AMT_ERR_INVALID: return "invalid argument";
    case AMT_ERR_CUDA: return "CUDA error";
    case AMT_ERR_CAPACITY: return "capacity exceeded";
    case AMT_ERR_UNSUPPORTED: return "unsupported dtype or rank";
    default: return "unknown status";
  }
}

const char* amt_last_cuda_error(void) { return cudaGetErrorString(amt::g_last_cuda_error); }

uint64_t amt_launch_count(void) { return amt::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
