// How fast can ONE warp issue independent float64 instructions on sm_100a, and how does that scale with the warps per
// scheduler?  (Round 2, step 27: the lo warps of the fused pass 2 issue one DADD / DMUL per ~17 clk each.)
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void probe(double* out, int iters, long long* clk) {
  double a[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) a[c] = 1.0 + threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, d = 1e-7;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = __dadd_rn(__dmul_rn(a[c], m), d);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int CHAINS>
void run(int warps, double* out, long long* clk) {
  const int iters = 2000;
  probe<CHAINS><<<148, warps * 32>>>(out, iters, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double dp_per_warp = 2.0 * CHAINS * iters;
  printf("chains %d warps/SM %2d (per scheduler %4.2f): %8lld clk, %.2f clk per DP instr per warp, %.3f DP warp-instr/clk/SM\n", CHAINS, warps,
         warps / 4.0, h, h / dp_per_warp, dp_per_warp * warps / h);
}
int main() {
  double* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&clk, 8);
  for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) run<1>(w, out, clk);
  for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) run<4>(w, out, clk);
  for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) run<8>(w, out, clk);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
