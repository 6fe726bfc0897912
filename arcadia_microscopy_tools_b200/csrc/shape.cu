// Perimeter and convex-hull area per label (placeholder until the kernels land).
#include "internal.cuh"

namespace amt {

size_t region_shape_scratch_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }

int region_shape(const int32_t*, const uint64_t*, int, const int32_t*, int64_t, int64_t, int64_t, int64_t, double*, void*,
                 size_t, cudaStream_t) {
  return AMT_ERR_UNSUPPORTED;
}

}  // namespace amt
