// Library bookkeeping: version, status strings, last CUDA error, launch counter.
#include <atomic>

#include "common.cuh"

namespace amt {

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
static std::atomic<uint64_t> g_launches{0};

void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace amt

extern "C" {

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
