// Layout conversion on the way in: pixel-interleaved camera frames -> channel planes.
//
// Reference path: nikon.py:25-43 (nd2.ND2File.asarray) / leica.py:52-80 decode on the host and hand
// NumPy a (C, Y, X) array.  Raw ND2 frames ("ImageDataSeq|i!" chunks) are uncompressed uint16 in
// (Y, X, C) order, so the host side of the B200 path only memcpy's the chunk payload into pinned
// staging (nd2_raw.py) and this kernel does the transposition to the (C, Y, X) planes every other
// kernel expects.  HBM-bound: 2 B read + 2 B written per sample.

#include "common.cuh"

namespace amt {

constexpr int DI_PIX = 1024;  // pixels per CTA

// in: n_frames x n_pix x C (interleaved), out: n_frames x C x n_pix.  grid (ceil(n_pix/1024), n_frames), block 256.
__global__ void __launch_bounds__(256)
deinterleave_u16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, const int64_t n_pix, const int C) {
  extern __shared__ __align__(16) uint16_t di_tile[];
  const int64_t frame = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * DI_PIX;
  const int npx = (int)((n_pix - p0 < DI_PIX) ? n_pix - p0 : DI_PIX);
  const uint16_t* src = in + (frame * n_pix + p0) * C;
  const int n_el = npx * C;
  // coalesced load: 16-byte vectors when the chunk start is aligned, 2-byte tail otherwise
  if ((((uintptr_t)src) & 15) == 0) {
    const int nv = n_el / 8;
    for (int i = threadIdx.x; i < nv; i += 256) reinterpret_cast<uint4*>(di_tile)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (int i = nv * 8 + threadIdx.x; i < n_el; i += 256) di_tile[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < n_el; i += 256) di_tile[i] = src[i];
  }
  __syncthreads();
  for (int c = 0; c < C; ++c) {
    uint16_t* dst = out + (frame * C + c) * n_pix + p0;
    for (int p = threadIdx.x; p < npx; p += 256) dst[p] = di_tile[p * C + c];
  }
}

}  // namespace amt

extern "C" {

int amt_deinterleave_u16(const uint16_t* in_yxc, uint16_t* out_cyx, int64_t n_frames, int64_t n_pix, int n_channels,
                         amt_stream_t stream) {
  using namespace amt;
  if (!in_yxc || !out_cyx || n_frames <= 0 || n_pix <= 0 || n_channels < 1 || n_channels > 16) return AMT_ERR_INVALID;
  if (n_frames > 65535) return AMT_ERR_CAPACITY;
  const size_t smem = (size_t)DI_PIX * n_channels * sizeof(uint16_t);
  deinterleave_u16_kernel<<<dim3((unsigned)ceil_div(n_pix, DI_PIX), (unsigned)n_frames), 256, smem, as_stream(stream)>>>(
      in_yxc, out_cyx, n_pix, n_channels);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
