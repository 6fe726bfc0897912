// Float-image pieces of the histogram thresholds (ref: operations.py:185-196 -> [3p] ski.filters.threshold_*):
//  * np.mean of a float64 plane, bit for bit: NumPy sums a contiguous array PAIRWISE — blocks of at most 128
//    elements are summed with eight strided accumulators, blocks are combined along a binary tree whose split
//    points depend only on the length (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum).  The host
//    lays out that tree once per length (leaf ranges + a post-order list of additions); the leaves are summed by
//    one thread each in NumPy's operation order, the tree by one thread per plane.
//  * np.histogram(plane, nbins, range=(min, max)) for any nbins: candidate bin from the uniform formula, corrected
//    against the linspace edges exactly as NumPy does (_histograms_impl.py:851-863).
#include <vector>

#include "common.cuh"

namespace amt {

__global__ void __launch_bounds__(128)
pairwise_leaf_kernel(const double* __restrict__ data, const int64_t n, const int32_t* __restrict__ leaf_start,
                     const int32_t* __restrict__ leaf_len, const int n_leaves, double* __restrict__ nodes, const int n_nodes) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n_leaves) return;
  const double* a = data + (int64_t)blockIdx.y * n + leaf_start[leaf];
  const int len = leaf_len[leaf];
  double res;
  if (len < 8) {
    res = -0.0;
    for (int i = 0; i < len; ++i) res = dadd(res, a[i]);
  } else {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < len - (len % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = dadd(r[j], a[i + j]);
    }
    res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
    for (; i < len; ++i) res = dadd(res, a[i]);
  }
  nodes[(int64_t)blockIdx.y * n_nodes + leaf] = res;
}

__global__ void pairwise_combine_kernel(double* __restrict__ nodes, const int32_t* __restrict__ sched, const int n_leaves,
                                        const int n_nodes, double* __restrict__ out, const int64_t n_img) {
  const int64_t img = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  double* s = nodes + img * n_nodes;
  for (int k = n_leaves; k < n_nodes; ++k) s[k] = dadd(s[sched[2 * (k - n_leaves)]], s[sched[2 * (k - n_leaves) + 1]]);
  out[img] = s[n_nodes - 1];
}

// host: NumPy's recursion over (offset, length) -> leaves in order, internal nodes in post-order
static int build_tree(int64_t off, int64_t n, std::vector<int32_t>& start, std::vector<int32_t>& len, std::vector<int32_t>& sched,
                      std::vector<int32_t>& internal_ids) {
  if (n <= 128) {
    start.push_back((int32_t)off);
    len.push_back((int32_t)n);
    return (int)start.size() - 1;  // leaf id
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  const int a = build_tree(off, n2, start, len, sched, internal_ids);
  const int b = build_tree(off + n2, n - n2, start, len, sched, internal_ids);
  sched.push_back(a);
  sched.push_back(b);
  internal_ids.push_back(0);
  return -(int)internal_ids.size();  // internal node k (1-based), patched to n_leaves + k - 1 below
}

__global__ void __launch_bounds__(256)
hist_f64_generic_kernel(const double* __restrict__ data, const int64_t n, const double* __restrict__ edges, const int nbins,
                        uint32_t* __restrict__ hist) {
  const int64_t img = blockIdx.y;
  const double* x = data + img * n;
  const double* e = edges + img * (nbins + 1);
  uint32_t* h = hist + img * nbins;
  const double first = e[0], last = e[nbins];
  const double denom = dsub(last, first);
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const double v = x[i];
    if (!(v >= first && v <= last)) continue;
    // numpy: f_indices = ((v - first) / (last - first)) * nbins; indices = f_indices.astype(intp); indices[indices == bins] -= 1
    int b = (int)dmul(ddiv(dsub(v, first), denom), (double)nbins);
    if (b == nbins) b -= 1;
    // then corrected against the edges: decrement where v < edges[b], increment where v >= edges[b+1] and b != nbins-1
    if (v < e[b])
      b -= 1;
    else if (v >= e[b + 1] && b != nbins - 1)
      b += 1;
    atomicAdd(&h[b], 1u);
  }
}

}  // namespace amt

extern "C" {

size_t amt_pairwise_sum_scratch_bytes(int64_t n_img, int64_t n) {
  if (n_img <= 0 || n <= 0) return 0;
  // at most 2 * ceil(n / 57) nodes: every leaf of a length > 128 range holds at least 57 elements
  const int64_t leaves = 2 * (n / 57 + 2);
  return (size_t)(n_img * 2 * leaves) * sizeof(double) + (size_t)(4 * leaves) * sizeof(int32_t) + 1024;
}

int amt_pairwise_sum_f64(const double* data, int64_t n_img, int64_t n, double* sums, void* scratch, size_t scratch_bytes,
                         amt_stream_t stream) {
  using namespace amt;
  if (!data || !sums || !scratch || n_img <= 0 || n <= 0 || n >= (1ll << 31)) return AMT_ERR_INVALID;
  std::vector<int32_t> start, len, sched, internal;
  const int root = build_tree(0, n, start, len, sched, internal);
  const int n_leaves = (int)start.size(), n_internal = (int)internal.size(), n_nodes = n_leaves + n_internal;
  for (auto& v : sched)
    if (v < 0) v = n_leaves + (-v) - 1;
  (void)root;
  const size_t node_bytes = (size_t)n_img * n_nodes * sizeof(double);
  const size_t int_count = (size_t)2 * n_leaves + 2 * (size_t)n_internal;
  if (scratch_bytes < node_bytes + int_count * sizeof(int32_t) + 256) return AMT_ERR_CAPACITY;
  cudaStream_t st = as_stream(stream);
  double* nodes = (double*)scratch;
  int32_t* d_int = (int32_t*)((char*)scratch + ((node_bytes + 255) / 256) * 256);
  std::vector<int32_t> host(int_count);
  std::copy(start.begin(), start.end(), host.begin());
  std::copy(len.begin(), len.end(), host.begin() + n_leaves);
  std::copy(sched.begin(), sched.end(), host.begin() + 2 * n_leaves);
  // pageable source: the copy is staged before the call returns, so `host` may go out of scope
  AMT_CUDA_TRY(cudaMemcpyAsync(d_int, host.data(), int_count * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  pairwise_leaf_kernel<<<dim3((unsigned)ceil_div(n_leaves, 128), (unsigned)n_img), 128, 0, st>>>(data, n, d_int, d_int + n_leaves,
                                                                                              n_leaves, nodes, n_nodes);
  AMT_LAUNCH_CHECK();
  pairwise_combine_kernel<<<(unsigned)ceil_div(n_img, 32), 32, 0, st>>>(nodes, d_int + 2 * n_leaves, n_leaves, n_nodes, sums, n_img);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_hist_f64(const double* data, int64_t n_img, int64_t n, const double* edges, int nbins, uint32_t* hist,
                 amt_stream_t stream) {
  using namespace amt;
  if (!data || !edges || !hist || n_img <= 0 || n <= 0 || nbins < 1 || nbins > (1 << 24) || n_img > 65535) return AMT_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  AMT_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)n_img * nbins * sizeof(uint32_t), st));
  int64_t blocks = ceil_div(n, 256 * 8);
  const int64_t cap = ceil_div((int64_t)kNumSMs * 8, n_img);
  if (blocks > cap) blocks = cap;
  hist_f64_generic_kernel<<<dim3((unsigned)blocks, (unsigned)n_img), 256, 0, st>>>(data, n, edges, nbins, hist);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
