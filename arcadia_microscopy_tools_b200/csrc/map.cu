// Elementwise stage: background subtraction / clip / percentile rescale, the fused 256-bin
// histogram, Otsu's scan and thresholding.
//
// Reference path: operations.py:97 (np.clip(dog - level, 0, None)), operations.py:50-54
// ski.exposure.rescale_intensity [3p], operations.py:186/:214 ski.filters.threshold_otsu [3p],
// operations.py:216 (intensities > threshold).  SURVEY.md 8a items 4-6.
//
// Every arithmetic step is the separately rounded float64 operation NumPy performs (sub, max,
// min, sub, div, mul, add), so the output plane is bit-identical given identical percentiles.
// The 256-bin histogram reproduces np.histogram's uniform-bin fast path: candidate bin from
// ((x-first)/(last-first))*256, then the correction against np.linspace edges, which are
// regenerated on device with linspace's own arithmetic (i*step + start, last edge = stop).
// Histogram updates are warp-aggregated (__match_any_sync) into per-warp private shared-memory
// histograms, then flushed with one global atomic per non-empty bin.  HBM-bound: 8 B read +
// 8 B written per sample.

#include <type_traits>

#include "common.cuh"

namespace amt {

extern int g_pass_ctas;  // CTAs per SM of the streaming passes (amt_tune "pass_ctas"), core.cu

__device__ __forceinline__ double map_value(double x, const amt_map_params& p) {
  if (p.flags & AMT_MAP_FILL) return p.o1;
  double y = x;
  if (p.flags & AMT_MAP_SUBCLIP) y = fmax(dsub(y, p.lvl), 0.0);
  if (p.flags & AMT_MAP_RESCALE) {
    y = fmin(fmax(y, p.p1), p.p2);
    if (p.p1 != p.p2) {
      y = ddiv(dsub(y, p.p1), dsub(p.p2, p.p1));
      y = dadd(dmul(y, dsub(p.o2, p.o1)), p.o1);
    } else {
      y = fmin(fmax(y, p.o1), p.o2);
    }
  }
  return y;
}

// The per-plane flags are block-uniform: resolve them once, outside the sample loop.
//   mode 0: FILL   1: SUBCLIP   2: RESCALE (p1 != p2)   3: SUBCLIP + RESCALE (p1 != p2)
//   mode 4: anything else (generic path, e.g. the degenerate p1 == p2 branch)
__device__ __forceinline__ int map_mode(const amt_map_params& p) {
  if (p.flags & AMT_MAP_FILL) return 0;
  if (p.flags == AMT_MAP_SUBCLIP) return 1;
  if ((p.flags & AMT_MAP_RESCALE) && p.p1 != p.p2) return (p.flags & AMT_MAP_SUBCLIP) ? 3 : 2;
  return 4;
}

// Correctly rounded a / b for a plane-constant divisor: y = RN(1/b) once per CTA, then per sample
//   q = RN(a*y);  r = a - b*q (exact, FMA);  q' = RN(q + r*y)
// Markstein's theorem: y is the correctly rounded reciprocal and q = RN(a*y) is within one ulp of
// a/b (relative error of y <= 2^-53, plus half an ulp of rounding), so q' = RN(a/b).
// 3 DP instructions instead of __ddiv_rn's ~11 DP + ~10 integer/branch instructions per sample.
// Exact only while nothing under/overflows: `fast` requires 2^-300 <= |b| <= 2^300, and a sample
// must be 0 or >= 2^-400 in magnitude; anything else takes __ddiv_rn.  tests/test_gpu_ops.py
// checks the sequence against __ddiv_rn bit for bit (amt_selftest_div).
struct DivConst {
  double b, y;
  bool fast;
};

__device__ __forceinline__ DivConst make_div_const(double b) {
  DivConst d;
  d.b = b;
  d.y = __drcp_rn(b);
  const int e = (__double2hiint(b) >> 20) & 0x7ff;
  d.fast = e >= 1023 - 300 && e <= 1023 + 300;
  return d;
}

__device__ __forceinline__ double div_const(double a, const DivConst& d) {
  // 0 or |a| >= 2^-400: one unsigned compare on the magnitude bits (0 - 1 wraps to the maximum)
  const unsigned long long mag = (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffull;
  if (!d.fast || mag - 1ull < (((unsigned long long)(1023 - 400)) << 52) - 1ull || mag >= (((unsigned long long)(1023 + 400)) << 52))
    return ddiv(a, d.b);
  const double q0 = dmul(a, d.y);
  const double r = __fma_rn(-d.b, q0, a);
  const double q1 = __fma_rn(r, d.y, q0);
  // a = -0.0: the correction turns the quotient into +0.0; the sign is always that of a*y
  return __hiloint2double((__double2hiint(q1) & 0x7fffffff) | (__double2hiint(q0) & 0x80000000), __double2loint(q1));
}

template <int MODE>
__device__ __forceinline__ double map_value_mode(double x, const amt_map_params& p, const DivConst& den, const double gain) {
  if (MODE == 0) return p.o1;
  if (MODE == 4) return map_value(x, p);
  double y = x;
  if (MODE == 1 || MODE == 3) y = dsub(y, p.lvl);
  if (MODE == 1) y = fmax(y, 0.0);
  if (MODE == 2 || MODE == 3) {
    // MODE 3: p1 is a percentile of the clipped (non-negative) plane, so max(max(t, 0), p1) = max(t, p1)
    if (MODE == 3 && p.p1 < 0.0) y = fmax(y, 0.0);
    y = fmin(fmax(y, p.p1), p.p2);
    y = div_const(dsub(y, p.p1), den);
    // q * 1.0 + 0.0 is q exactly (q >= +0 here): the default out_range (0, 1) needs no arithmetic
    if (!(gain == 1.0 && p.o1 == 0.0)) y = dadd(dmul(y, gain), p.o1);
  }
  return y;
}

// Modes 2 / 3 on a thread's 8 samples.  The numerator a = clamp(y, p1, p2) - p1 is +0 or positive and at most
// p2 - p1, so the plane-constant division needs neither the upper-range test nor the sign fix of div_const, and
// the rare sample the three-instruction sequence cannot serve (a tiny non-zero a below 2^-400, -0, NaN) is
// flagged instead of branched on: one test per batch, __ddiv_rn for the whole batch when it fires.
template <int MODE>
__device__ __forceinline__ void map_rescale8(double (&v)[8], const amt_map_params& p, const DivConst& den, const double gain) {
  constexpr int T_LO = (1023 - 400) << 20, T_HI = (1023 + 400) << 20;
  double a[8];
  bool slow = !den.fast || !(den.b > 0.0);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    double y = v[e];
    if (MODE == 3) y = dsub(y, p.lvl);
    if (MODE == 3 && p.p1 < 0.0) y = fmax(y, 0.0);  // p1 >= 0 otherwise: max(max(t, 0), p1) = max(t, p1)
    y = fmin(fmax(y, p.p1), p.p2);
    a[e] = dsub(y, p.p1);
    const int hi = __double2hiint(a[e]);
    const bool in_range = (unsigned)(hi - T_LO) < (unsigned)(T_HI - T_LO);  // false for negative, -0, inf, NaN
    slow |= !(in_range || (hi | __double2loint(a[e])) == 0);
  }
  const bool plain_out = gain == 1.0 && p.o1 == 0.0;  // q * 1.0 + 0.0 is q exactly (q >= +0)
  if (!slow) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double q0 = dmul(a[e], den.y);
      const double r = __fma_rn(-den.b, q0, a[e]);
      const double q1 = __fma_rn(r, den.y, q0);
      v[e] = plain_out ? q1 : dadd(dmul(q1, gain), p.o1);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double q = ddiv(a[e], den.b);
      v[e] = plain_out ? q : dadd(dmul(q, gain), p.o1);
    }
  }
}

struct HistRange {
  double first, last, denom, step;
  bool step_zero;
};

__device__ __forceinline__ HistRange make_hist_range(double first, double last) {
  HistRange r;
  if (first == last) {  // numpy _get_outer_edges widens a degenerate range
    first = dsub(first, 0.5);
    last = dadd(last, 0.5);
  }
  r.first = first;
  r.last = last;
  r.denom = dsub(last, first);
  r.step = ddiv(r.denom, 256.0);
  r.step_zero = (r.step == 0.0);
  return r;
}

// np.linspace(first, last, 257)[i]
__device__ __forceinline__ double hist_edge(const HistRange& r, int i) {
  if (i >= 256) return r.last;
  if (r.step_zero) return dadd(dmul(ddiv((double)i, 256.0), r.denom), r.first);
  return dadd(dmul((double)i, r.step), r.first);
}

__device__ __forceinline__ int hist_bin(const HistRange& r, double x) {
  double f = dmul(ddiv(dsub(x, r.first), r.denom), 256.0);
  int idx = (int)f;
  idx = idx < 0 ? 0 : idx;
  if (idx >= 256) idx = 255;
  if (x < hist_edge(r, idx)) {
    idx -= 1;
  }
  idx = idx < 0 ? 0 : idx;
  if (idx != 255 && x >= hist_edge(r, idx + 1)) idx += 1;
  return idx;
}

// same binning with the 257 edges tabulated in shared memory and the plane-constant division
__device__ __forceinline__ int hist_bin_table(const HistRange& r, const DivConst& dn, const double* __restrict__ edges, double x) {
  const double f = dmul(div_const(dsub(x, r.first), dn), 256.0);
  int idx = (int)f;
  idx = idx < 0 ? 0 : idx;
  idx = idx > 255 ? 255 : idx;
  idx -= (x < edges[idx]) ? 1 : 0;
  idx = idx < 0 ? 0 : idx;
  idx += (idx != 255 && x >= edges[idx + 1]) ? 1 : 0;
  return idx;
}

__device__ __forceinline__ void hist_add_warp(uint32_t* wh, int bin, bool valid) {
  const int key = valid ? bin : -1;
  const unsigned m = __match_any_sync(0xffffffffu, key);
  if (valid && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&wh[bin], (uint32_t)__popc(m));
}

template <typename InT>
__device__ __forceinline__ double map_load(const InT* p);
template <>
__device__ __forceinline__ double map_load<double>(const double* p) {
  return *p;
}
template <>
__device__ __forceinline__ double map_load<uint16_t>(const uint16_t* p) {
  return (double)*p;
}

template <typename InT>
__device__ __forceinline__ void load4(const InT* p, double* v);
template <>
__device__ __forceinline__ void load4<double>(const double* p, double* v) {
  const int4 a = ld_nc_int4(p), b = ld_nc_int4(p + 2);
  v[0] = __hiloint2double(a.y, a.x);
  v[1] = __hiloint2double(a.w, a.z);
  v[2] = __hiloint2double(b.y, b.x);
  v[3] = __hiloint2double(b.w, b.z);
}
template <>
__device__ __forceinline__ void load4<uint16_t>(const uint16_t* p, double* v) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = (double)(a.x & 0xffffu);
  v[1] = (double)(a.x >> 16);
  v[2] = (double)(a.y & 0xffffu);
  v[3] = (double)(a.y >> 16);
}

// 16-byte streaming store (evict-first): the output planes are not re-read before the caches turn over
__device__ __forceinline__ void st_stream(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

template <typename InT, bool HIST, bool VEC>
__global__ void __launch_bounds__(256)
map_kernel(const InT* __restrict__ in, double* __restrict__ out, int64_t n, const amt_map_params* __restrict__ params,
           uint32_t* __restrict__ hist256, int hist_every, int hist_offset, const double cand_eps,
           uint32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_idx, const int cand_cap) {
  __shared__ uint32_t s_hist[HIST ? 8 * 256 : 1];
  __shared__ double s_edges[HIST ? 257 : 1];
  // HIST launch: grid.y runs over the histogram planes only (img = y * hist_every + hist_offset).
  // plain launch: grid.y runs over all planes; when hist_every > 0 the histogram planes belong to the
  // other launch and are skipped (hist_every == 0: no histogram requested, every plane is mapped here).
  const int64_t img = HIST ? (int64_t)blockIdx.y * hist_every + hist_offset : (int64_t)blockIdx.y;
  if (!HIST && hist_every > 0 && (img % hist_every) == hist_offset) return;
  const amt_map_params p = params[img];
  const InT* src = in + img * n;
  double* dst = out + img * n;
  const bool do_hist = HIST && hist256 != nullptr;
  HistRange hr;
  uint32_t* wh = s_hist + (threadIdx.x >> 5) * 256;
  if (HIST) {
    for (int i = threadIdx.x; i < 8 * 256; i += 256) s_hist[i] = 0;
    hr = make_hist_range(p.hist_first, p.hist_last);
    for (int i = threadIdx.x; i < 257; i += 256) s_edges[i] = hist_edge(hr, i);
    __syncthreads();
  }
  const DivConst hden = make_div_const(HIST ? hr.denom : 1.0);
  const DivConst den = make_div_const(dsub(p.p2, p.p1));
  const double gain = dsub(p.o2, p.o1);
  // Decision-exact mode (decide.cu): the input plane is within cand_eps of the exact one, hence the output within
  // tol of the exact output; a sample closer than that to a bin edge or to its bin's centre (the Otsu threshold is
  // a centre) is listed for exact re-evaluation.  The outer edges decide nothing (every exact value lies inside
  // [first, last] too); a constant output plane (FILL, p1 == p2) decides nothing at all.
  const bool collect = HIST && cand_count != nullptr && map_mode(p) == 3;
  const double tol = collect ? cand_eps * fabs(gain) / dsub(p.p2, p.p1) * 1.000001 + 1e-15 : -1.0;
  const int64_t img_h = HIST ? (int64_t)blockIdx.y : 0;
  const double tol256 = tol * 256.0;
  // every output sample is a quotient in [0, 1] (modes 2 / 3, out_range (0, 1)) and the histogram spans exactly [0, 1]
  const bool unit_hist = HIST && (map_mode(p) == 2 || map_mode(p) == 3) && gain == 1.0 && p.o1 == 0.0 && p.hist_first == 0.0 &&
                         p.hist_last == 1.0 && den.fast && den.b > 0.0;
  auto consider = [&](double x, int bin, int64_t index) {
    const double lo_e = s_edges[bin], hi_e = s_edges[bin + 1];
    const double centre = dmul(dadd(lo_e, hi_e), 0.5);
    if ((bin > 0 && x - lo_e <= tol) || (bin < 255 && hi_e - x <= tol) || fabs(x - centre) <= tol) {
      const uint32_t pos = atomicAdd(&cand_count[img_h], 1u);
      if (pos < (uint32_t)cand_cap) cand_idx[img_h * cand_cap + pos] = (uint32_t)index;
    }
  };
  auto run = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
    if (VEC) {
      // 4 consecutive samples per thread and iteration: two 16-byte loads (one 8-byte load for
      // uint16), two 16-byte stores; two iterations in flight per thread
      const int64_t n4 = n >> 2;
      const int64_t step = (int64_t)gridDim.x * 256;
      for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n4; q += 2 * step) {
        double v[8];
        const int64_t q2 = q + step;
        const bool second = q2 < n4;
        load4<InT>(src + 4 * q, v);
        if (second) {
          load4<InT>(src + 4 * q2, v + 4);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[4 + e] = v[e];  // a last odd group: defined values, results dropped
        }
        if constexpr (MODE == 2 || MODE == 3) {
          map_rescale8<MODE>(v, p, den, gain);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = map_value_mode<MODE>(v[e], p, den, gain);
        }
        st_stream(dst + 4 * q, v[0], v[1]);
        st_stream(dst + 4 * q + 2, v[2], v[3]);
        if (second) {
          st_stream(dst + 4 * q2, v[4], v[5]);
          st_stream(dst + 4 * q2 + 2, v[6], v[7]);
        }
        if (HIST && do_hist) {
          if (unit_hist) {
            // histogram range exactly [0, 1] (a rescaled plane whose percentiles clipped something, out_range (0, 1)):
            // np.histogram's own arithmetic is exact there, (x - 0) / 1 * 256 and edges i / 256, so the bin is
            // floor(256 x) (256 -> 255), never corrected; distances to edges and centre come from the fraction
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e < 4 || second) {
                const double t = dmul(v[e], 256.0);
                int bin = (int)t;
                bin = bin > 255 ? 255 : bin;
                atomicAdd(&wh[bin], 1u);
                if (collect) {
                  const double fr = dsub(t, (double)bin);  // exact; 1.0 for x == 1 (bin 255)
                  if ((bin > 0 && fr <= tol256) || (bin < 255 && dsub(1.0, fr) <= tol256) || fabs(dsub(fr, 0.5)) <= tol256) {
                    const uint32_t pos = atomicAdd(&cand_count[img_h], 1u);
                    if (pos < (uint32_t)cand_cap) cand_idx[img_h * cand_cap + pos] = (uint32_t)(4 * (e < 4 ? q : q2) + (e & 3));
                  }
                }
              }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e < 4 || second) {
                const int bin = hist_bin_table(hr, hden, s_edges, v[e]);
                atomicAdd(&wh[bin], 1u);
                if (collect) consider(v[e], bin, 4 * (e < 4 ? q : q2) + (e & 3));
              }
          }
        }
      }
    } else {
      const int64_t step = (int64_t)gridDim.x * 256;
      for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += step) {
        const double y = map_value_mode<MODE>(map_load<InT>(src + i), p, den, gain);
        dst[i] = y;
        if (HIST && do_hist) {
          const int bin = hist_bin_table(hr, hden, s_edges, y);
          atomicAdd(&wh[bin], 1u);
          if (collect) consider(y, bin, i);
        }
      }
    }
  };
  switch (map_mode(p)) {
    case 0: run(std::integral_constant<int, 0>{}); break;
    case 1: run(std::integral_constant<int, 1>{}); break;
    case 2: run(std::integral_constant<int, 2>{}); break;
    case 3: run(std::integral_constant<int, 3>{}); break;
    default: run(std::integral_constant<int, 4>{}); break;
  }
  if (HIST && do_hist) {
    __syncthreads();
    uint32_t c = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) c += s_hist[wv * 256 + threadIdx.x];
    if (c) atomicAdd(&hist256[(img / hist_every) * 256 + threadIdx.x], c);
  }
}

// standalone 256-bin float histogram (range from min/max keys)
__global__ void __launch_bounds__(256)
hist256_kernel(const double* __restrict__ data, int64_t n, const uint64_t* __restrict__ mm, uint32_t* __restrict__ hist256) {
  __shared__ uint32_t s_hist[8 * 256];
  const int64_t img = blockIdx.y;
  const double* src = data + img * n;
  for (int i = threadIdx.x; i < 8 * 256; i += 256) s_hist[i] = 0;
  const HistRange hr = make_hist_range(key_to_f64(mm[2 * img]), key_to_f64(mm[2 * img + 1]));
  uint32_t* wh = s_hist + (threadIdx.x >> 5) * 256;
  __syncthreads();
  const int64_t step = (int64_t)gridDim.x * 256;
  const int64_t start = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t i = start, wi = start - (threadIdx.x & 31); wi < n; i += step, wi += step) {
    const bool valid = i < n;
    hist_add_warp(wh, valid ? hist_bin(hr, src[i]) : 0, valid);
  }
  __syncthreads();
  uint32_t c = 0;
#pragma unroll
  for (int wv = 0; wv < 8; ++wv) c += s_hist[wv * 256 + threadIdx.x];
  if (c) atomicAdd(&hist256[img * 256 + threadIdx.x], c);
}

// Decision-exact mode: the listed samples of the histogram planes get their exact value (the exact difference of
// Gaussians through the plane's map) and their histogram count moves to the exact value's bin.
__global__ void __launch_bounds__(256)
dx_patch_kernel(const uint32_t* __restrict__ cand_count, const uint32_t* __restrict__ cand_idx, const double* __restrict__ exact_in,
                int cap, const amt_map_params* __restrict__ params, int hist_every, int hist_offset, double* __restrict__ out,
                int64_t n, uint32_t* __restrict__ hist256, int32_t* __restrict__ retry) {
  const int64_t img_h = blockIdx.y;
  const int64_t img = img_h * hist_every + hist_offset;
  const uint32_t c = cand_count[img_h];
  if (c > (uint32_t)cap) {  // list overflow: this image is recomputed in float64
    if (blockIdx.x == 0 && threadIdx.x == 0) retry[img_h] = 1;
    return;
  }
  const amt_map_params p = params[img];
  const HistRange hr = make_hist_range(p.hist_first, p.hist_last);
  double* dst = out + img * n;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < c; k += gridDim.x * blockDim.x) {
    const uint32_t pix = cand_idx[img_h * cap + k];
    const double approx = dst[pix];
    const double exact = map_value(exact_in[img_h * cap + k], p);
    dst[pix] = exact;
    const int b0 = hist_bin(hr, approx), b1 = hist_bin(hr, exact);
    if (b0 != b1) {
      atomicSub(&hist256[img_h * 256 + b0], 1u);
      atomicAdd(&hist256[img_h * 256 + b1], 1u);
    }
  }
}

__global__ void plan_dog_rescale_kernel(const double* __restrict__ stats, const uint64_t* __restrict__ mm, int64_t n_img,
                                        double g_bg, double g_lo, double g_hi, double o1, double o2,
                                        amt_map_params* __restrict__ params) {
  const int64_t img = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  const double* s = stats + img * 6;
  amt_map_params p;
  p.lvl = np_lerp(s[0], s[1], g_bg);  // np.percentile(dog, percentile)
  // order statistics commute with the monotone map x -> max(x - lvl, 0)
  const double a1 = fmax(dsub(s[2], p.lvl), 0.0), b1 = fmax(dsub(s[3], p.lvl), 0.0);
  const double a2 = fmax(dsub(s[4], p.lvl), 0.0), b2 = fmax(dsub(s[5], p.lvl), 0.0);
  p.p1 = np_lerp(a1, b1, g_lo);
  p.p2 = np_lerp(a2, b2, g_hi);
  p.o1 = o1;
  p.o2 = o2;
  const double cmin = fmax(dsub(key_to_f64(mm[2 * img]), p.lvl), 0.0);
  const double cmax = fmax(dsub(key_to_f64(mm[2 * img + 1]), p.lvl), 0.0);
  p.flags = (cmin == cmax) ? AMT_MAP_FILL : (AMT_MAP_SUBCLIP | AMT_MAP_RESCALE);
  p.pad = 0;
  // range of the output plane: the map is monotone, so it is attained at the clipped min / max
  amt_map_params q = p;
  q.flags &= ~AMT_MAP_SUBCLIP;
  const double fa = map_value(cmin, q), fb = map_value(cmax, q);
  p.hist_first = fmin(fa, fb);
  p.hist_last = fmax(fa, fb);
  params[img] = p;
}

// ------------------------------------------------------------------ Otsu scan
// One block (64 threads) per plane.  Warp 1 lane 0 runs the reversed cumulative sums
// (weight2 float32, mean2 numerator float64) into scratch, then warp 0 lane 0 runs the forward
// ones and keeps the first maximum of w1[i]*w2[i+1]*(m1[i]-m2[i+1])^2.  The sums are
// sequential on purpose: NumPy's cumsum is, and float32/float64 addition does not reassociate.
__global__ void otsu_kernel(const uint32_t* __restrict__ hist, int mode, const amt_map_params* __restrict__ params,
                            int64_t pstride, int64_t poffset, const uint64_t* __restrict__ mm,
                            float* __restrict__ w2_scratch, double* __restrict__ m2_scratch,
                            double* __restrict__ thresholds) {
  __shared__ float s_w2[256];
  __shared__ double s_m2[256];
  const int64_t img = blockIdx.x;
  int nb, lo_bin = 0;
  HistRange hr;
  const uint32_t* h;
  float* w2;
  double* m2;
  if (mode == 2) {
    const int vmin = (int)mm[2 * img], vmax = (int)mm[2 * img + 1];
    lo_bin = vmin;
    nb = vmax - vmin + 1;
    h = hist + img * 65536 + vmin;
    w2 = w2_scratch + img * 65536;
    m2 = m2_scratch + img * 65536;
  } else {
    nb = 256;
    h = hist + img * 256;
    w2 = s_w2;
    m2 = s_m2;
    double first, last;
    if (mode == 0) {
      first = params[img * pstride + poffset].hist_first;
      last = params[img * pstride + poffset].hist_last;
    } else {
      first = key_to_f64(mm[2 * img]);
      last = key_to_f64(mm[2 * img + 1]);
    }
    if (first == last) {  // constant plane: skimage returns that value (nothing is > it)
      if (threadIdx.x == 0) thresholds[img] = first;
      return;
    }
    hr = make_hist_range(first, last);
  }
  auto center = [&](int i) -> double {
    if (mode == 2) return (double)(lo_bin + i);
    return ddiv(dadd(hist_edge(hr, i), hist_edge(hr, i + 1)), 2.0);
  };
  if (nb < 2) {  // constant plane: skimage returns the single value
    if (threadIdx.x == 0) thresholds[img] = center(0);
    return;
  }
  if (mode != 2) {
    // 256 bins: only the cumulative sums are sequential (NumPy's order); products, class
    // means, variances and the arg-max run one bin per thread
    __shared__ float s_c[256], s_w1[256];
    __shared__ double s_prod[256], s_m1[256];
    __shared__ double s_bestv[8];
    __shared__ int s_besti[8];
    const int i = threadIdx.x;
    s_c[i] = (float)h[i];
    s_prod[i] = dmul((double)s_c[i], center(i));
    __syncthreads();
    if (i == 0) {
      float wsum = 0.0f;
      double msum = 0.0;
      for (int k = 0; k < 256; ++k) {
        wsum = __fadd_rn(wsum, s_c[k]);
        msum = dadd(msum, s_prod[k]);
        s_w1[k] = wsum;
        s_m1[k] = msum;
      }
    } else if (i == 32) {
      float wsum = 0.0f;
      double msum = 0.0;
      for (int k = 255; k >= 1; --k) {
        wsum = __fadd_rn(wsum, s_c[k]);
        msum = dadd(msum, s_prod[k]);
        w2[k] = wsum;
        m2[k] = msum;
      }
    }
    __syncthreads();
    double var = -1.0;
    int idx = i;
    if (i < 255) {
      const double mean1 = ddiv(s_m1[i], (double)s_w1[i]);
      const double mean2 = ddiv(m2[i + 1], (double)w2[i + 1]);
      const float ww = __fmul_rn(s_w1[i], w2[i + 1]);
      const double d = dsub(mean1, mean2);
      var = dmul((double)ww, dmul(d, d));
    }
    // first maximum: larger variance wins, ties go to the smaller index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, var, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > var || (ov == var && oi < idx)) {
        var = ov;
        idx = oi;
      }
    }
    if ((i & 31) == 0) {
      s_bestv[i >> 5] = var;
      s_besti[i >> 5] = idx;
    }
    __syncthreads();
    if (i == 0) {
      for (int k = 1; k < 8; ++k)
        if (s_bestv[k] > var || (s_bestv[k] == var && s_besti[k] < idx)) {
          var = s_bestv[k];
          idx = s_besti[k];
        }
      thresholds[img] = center(idx);
    }
    return;
  }
  if (threadIdx.x == 32) {
    float wsum = 0.0f;
    double msum = 0.0;
    for (int i = nb - 1; i >= 1; --i) {
      const float c = (float)h[i];
      wsum = __fadd_rn(wsum, c);
      msum = dadd(msum, dmul((double)c, center(i)));
      w2[i] = wsum;
      m2[i] = ddiv(msum, (double)wsum);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float wsum = 0.0f;
    double msum = 0.0, best = -1.0;
    int best_i = 0;
    for (int i = 0; i < nb - 1; ++i) {
      const float c = (float)h[i];
      wsum = __fadd_rn(wsum, c);
      msum = dadd(msum, dmul((double)c, center(i)));
      const double mean1 = ddiv(msum, (double)wsum);
      const float ww = __fmul_rn(wsum, w2[i + 1]);
      const double d = dsub(mean1, m2[i + 1]);
      const double var = dmul((double)ww, dmul(d, d));
      if (i == 0 || var > best) {
        best = var;
        best_i = i;
      }
    }
    thresholds[img] = center(best_i);
  }
}

// bitwise comparison of div_const against __ddiv_rn (test hook)
__global__ void selftest_div_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                    unsigned long long* __restrict__ mismatches) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (; i < n; i += step) {
    const DivConst d = make_div_const(b[i]);
    const double q = div_const(a[i], d), want = ddiv(a[i], b[i]);
    bad += __double_as_longlong(q) != __double_as_longlong(want);
  }
  if (bad) atomicAdd(mismatches, bad);
}

template <typename InT>
__global__ void __launch_bounds__(256)
threshold_gt_kernel(const InT* __restrict__ data, int64_t n, const double* __restrict__ thresholds, uint8_t* __restrict__ mask) {
  const int64_t img = blockIdx.y;
  const double t = thresholds[img];
  const InT* src = data + img * n;
  uint8_t* dst = mask + img * n;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += step)
    dst[i] = map_load<InT>(src + i) > t ? 1 : 0;
}

static unsigned stream_blocks(int64_t n, int64_t n_img, int per_thread) {
  int64_t bx = ceil_div(n, 256 * (int64_t)per_thread);
  const int64_t cap = ceil_div((int64_t)kNumSMs * g_pass_ctas, n_img);
  if (bx > cap) bx = cap;
  return (unsigned)(bx < 1 ? 1 : bx);
}

int map_launch(const void* in, int in_dtype, double* out, int64_t n_img, int64_t n, const amt_map_params* params,
               uint32_t* hist256, int hist_every, int hist_offset, cudaStream_t st, double cand_eps, uint32_t* cand_count,
               uint32_t* cand_idx, int cand_cap) {
  if (!in || !out || !params || n_img <= 0 || n <= 0 || n_img > 65535 || hist_every < 1) return AMT_ERR_INVALID;
  const bool vec = (n % 4 == 0) && (((uintptr_t)in) % 16 == 0) && (((uintptr_t)out) % 16 == 0);
  // planes with a histogram go through the HIST instantiation (66 registers, 10 KB of shared memory), all
  // others through the plain one (38 registers: twice the resident warps to cover the memory latency)
  if (hist256 && hist_offset >= n_img) return AMT_ERR_INVALID;
  const int64_t n_hist = hist256 ? (n_img - hist_offset + hist_every - 1) / hist_every : 0;
  const bool plain_needed = n_hist < n_img;
  dim3 grid(stream_blocks(n, n_img, vec ? 16 : 8), (unsigned)n_img);
  dim3 hgrid(stream_blocks(n, n_hist > 0 ? n_hist : 1, vec ? 16 : 8), (unsigned)(n_hist > 0 ? n_hist : 1));
#define AMT_MAP_LAUNCH(T, V)                                                                                      \
  do {                                                                                                            \
    if (plain_needed) {                                                                                           \
      map_kernel<T, false, V><<<grid, 256, 0, st>>>((const T*)in, out, n, params, nullptr, hist256 ? hist_every : 0, \
                                                    hist_offset, 0.0, nullptr, nullptr, 0);                       \
      count_launch();                                                                                             \
    }                                                                                                             \
    if (n_hist > 0)                                                                                               \
      map_kernel<T, true, V><<<hgrid, 256, 0, st>>>((const T*)in, out, n, params, hist256, hist_every, hist_offset, \
                                                    cand_eps, cand_count, cand_idx, cand_cap);                    \
  } while (0)
  if (in_dtype == AMT_F64) {
    if (vec) AMT_MAP_LAUNCH(double, true); else AMT_MAP_LAUNCH(double, false);
  } else if (in_dtype == AMT_U16) {
    if (vec) AMT_MAP_LAUNCH(uint16_t, true); else AMT_MAP_LAUNCH(uint16_t, false);
  } else {
    return AMT_ERR_UNSUPPORTED;
  }
#undef AMT_MAP_LAUNCH
  if (n_hist > 0) {
    AMT_LAUNCH_CHECK();
  } else {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
      set_last_cuda_error(e);
      (void)cudaGetLastError();
      return AMT_ERR_CUDA;
    }
  }
  return AMT_OK;
}

int dx_patch(const uint32_t* cand_count, const uint32_t* cand_idx, const double* exact_in, int cap, const amt_map_params* params,
             int hist_every, int hist_offset, int64_t n_hist, double* out, int64_t n, uint32_t* hist256, int32_t* retry,
             cudaStream_t st) {
  dx_patch_kernel<<<dim3(4, (unsigned)n_hist), 256, 0, st>>>(cand_count, cand_idx, exact_in, cap, params, hist_every, hist_offset,
                                                           out, n, hist256, retry);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int plan_dog_rescale(const double* stats, const uint64_t* mm, int64_t n_img, double g_bg, double g_lo, double g_hi,
                     double o1, double o2, amt_map_params* params, cudaStream_t st) {
  if (!stats || !mm || !params || n_img <= 0) return AMT_ERR_INVALID;
  plan_dog_rescale_kernel<<<(unsigned)ceil_div(n_img, 128), 128, 0, st>>>(stats, mm, n_img, g_bg, g_lo, g_hi, o1, o2, params);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int otsu_launch(const uint32_t* hist, int mode, const amt_map_params* params, int64_t pstride, int64_t poffset,
                const uint64_t* mm, int64_t n_img, double* thresholds, void* scratch, size_t scratch_bytes,
                cudaStream_t st) {
  if (!hist || !thresholds || n_img <= 0 || mode < 0 || mode > 2) return AMT_ERR_INVALID;
  if (mode == 0 && !params) return AMT_ERR_INVALID;
  if (mode != 0 && !mm) return AMT_ERR_INVALID;
  float* w2 = nullptr;
  double* m2 = nullptr;
  if (mode == 2) {
    if (!scratch || scratch_bytes < (size_t)n_img * 65536 * 12) return AMT_ERR_CAPACITY;
    m2 = (double*)scratch;
    w2 = (float*)((char*)scratch + (size_t)n_img * 65536 * 8);
  }
  otsu_kernel<<<(unsigned)n_img, 256, 0, st>>>(hist, mode, params, pstride, poffset, mm, w2, m2, thresholds);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // namespace amt

extern "C" {

int amt_map(const void* in, int in_dtype, double* out, int64_t n_img, int64_t n, const amt_map_params* params,
            uint32_t* hist256, amt_stream_t stream) {
  return amt::map_launch(in, in_dtype, out, n_img, n, params, hist256, 1, 0, amt::as_stream(stream), 0.0, nullptr, nullptr, 0);
}

int amt_plan_dog_rescale(const double* order_stats, const uint64_t* minmax_keys, int64_t n_img, double g_bg,
                         double g_lo, double g_hi, double o1, double o2, amt_map_params* params, amt_stream_t stream) {
  return amt::plan_dog_rescale(order_stats, minmax_keys, n_img, g_bg, g_lo, g_hi, o1, o2, params, amt::as_stream(stream));
}

int amt_hist256_f64(const double* data, int64_t n_img, int64_t n, const uint64_t* minmax_keys, uint32_t* hist256,
                    amt_stream_t stream) {
  using namespace amt;
  if (!data || !minmax_keys || !hist256 || n_img <= 0 || n <= 0 || n_img > 65535) return AMT_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  AMT_CUDA_TRY(cudaMemsetAsync(hist256, 0, (size_t)n_img * 256 * sizeof(uint32_t), st));
  hist256_kernel<<<dim3(stream_blocks(n, n_img, 8), (unsigned)n_img), 256, 0, st>>>(data, n, minmax_keys, hist256);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

size_t amt_otsu_scratch_bytes(int mode, int64_t n_img) { return mode == 2 ? (size_t)n_img * 65536 * 12 : 0; }

int amt_otsu(const uint32_t* hist, int mode, const amt_map_params* params, const uint64_t* minmax_keys, int64_t n_img,
             double* thresholds, void* scratch, size_t scratch_bytes, amt_stream_t stream) {
  return amt::otsu_launch(hist, mode, params, 1, 0, minmax_keys, n_img, thresholds, scratch, scratch_bytes,
                          amt::as_stream(stream));
}

int amt_threshold_gt(const void* data, int in_dtype, int64_t n_img, int64_t n, const double* thresholds, uint8_t* mask,
                     amt_stream_t stream) {
  using namespace amt;
  if (!data || !thresholds || !mask || n_img <= 0 || n <= 0 || n_img > 65535) return AMT_ERR_INVALID;
  dim3 grid(stream_blocks(n, n_img, 8), (unsigned)n_img);
  if (in_dtype == AMT_F64)
    threshold_gt_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)data, n, thresholds, mask);
  else if (in_dtype == AMT_U16)
    threshold_gt_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>((const uint16_t*)data, n, thresholds, mask);
  else
    return AMT_ERR_UNSUPPORTED;
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

int amt_selftest_div(const double* a, const double* b, int64_t n, uint64_t* mismatches, amt_stream_t stream) {
  using namespace amt;
  if (!a || !b || !mismatches || n <= 0) return AMT_ERR_INVALID;
  AMT_CUDA_TRY(cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), as_stream(stream)));
  selftest_div_kernel<<<kNumSMs * 8, 256, 0, as_stream(stream)>>>(a, b, n, (unsigned long long*)mismatches);
  AMT_LAUNCH_CHECK();
  return AMT_OK;
}

}  // extern "C"
