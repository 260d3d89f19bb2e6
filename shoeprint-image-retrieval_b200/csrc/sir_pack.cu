// HBM-bound preparation kernels: gallery pack (K5), window inverse norm (K6), probe variants
// (K4: Pillow-exact rotate / bicubic resize) and template pack (K5).
//
// Reference semantics restated (reference repo paths):
//   similarity.py:92-93   crop [:, 2:-2, 2:-2]
//   similarity.py:48-49   zero mean per channel over the cropped map
//   similarity.py:57-65   D = box(g^2) - box(g)^2 / (Hm*Wm), clamped at 0, float64
//   similarity.py:67      E = sum(t^2)
//   similarity.py:262-276 Image.rotate (nearest) / Image.resize (bicubic) per channel
#include <cuda_fp8.h>

#include "sir_common.cuh"

namespace sir {

// One warp copies a plane of n floats into its shared-memory slab with eight independent loads in flight per lane: with a
// plain one-load-per-iteration loop the pack kernels kept ~20 KB in flight per SM and reached 45 % of the HBM rate.
__device__ __forceinline__ void warp_load_plane(float* __restrict__ dst, const float* __restrict__ src, int n, int lane) {
  int i = lane;
  for (; i + 7 * 32 < n; i += 8 * 32) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(src + i + 32 * j);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[i + 32 * j] = v[j];
  }
  for (; i < n; i += 32) dst[i] = __ldg(src + i);
}

__device__ __forceinline__ uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

// ------------------------------------------------------------------------------------------
// K5 gallery: one CTA per (gallery, channel).  Reads 4 B/cell, writes 4 B/cell (hi+lo) [+4 gz].
__global__ void __launch_bounds__(256) gallery_pack_kernel(const float* __restrict__ gal, int C, int hg, int wg,
                                                           __half* __restrict__ ghi, __half* __restrict__ glo,
                                                           int32_t* __restrict__ gexp, float* __restrict__ gz, float* __restrict__ g32) {
  __shared__ double sred[32];
  __shared__ float fred[32];
  const int Hp = hg - 2 * kEdge, Wp = wg - 2 * kEdge, M = Hp * Wp;
  const size_t gc = blockIdx.x;  // g*C + c
  const float* src = gal + gc * (size_t)hg * wg;

  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const int y = i / Wp, x = i - y * Wp;
    acc += (double)src[(y + kEdge) * wg + x + kEdge];
  }
  const float mean = (float)(block_sum(acc, sred) / (double)M);

  float amax = 0.0f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const int y = i / Wp, x = i - y * Wp;
    amax = fmaxf(amax, fabsf(src[(y + kEdge) * wg + x + kEdge] - mean));
  }
  amax = block_max(amax, fred);
  int e = 0;
  if (amax > 0.0f && isfinite(amax)) {
    int ex;
    (void)frexpf(amax, &ex);  // amax = f * 2^ex, f in [0.5, 1)
    e = kGalleryPeakLog2 - ex;
  }
  if (threadIdx.x == 0) gexp[gc] = e;

  const int WP = gal_pitch(Wp);  // padded row pitch of the fp16 operands (pad cells are zero)
  for (int i = threadIdx.x; i < Hp * WP; i += blockDim.x) {
    const int y = i / WP, x = i - y * WP;
    __half h = __ushort_as_half(0), l = __ushort_as_half(0);
    float s = 0.0f;
    if (x < Wp) {
      const float z = src[(y + kEdge) * wg + x + kEdge] - mean;
      s = ldexpf(z, e);
      h = __float2half_rn(s);
      l = __float2half_rn(s - __half2float(h));
      if (gz) gz[gc * M + y * Wp + x] = z;
    }
    ghi[gc * Hp * WP + i] = h;
    glo[gc * Hp * WP + i] = l;
    if (g32) g32[gc * Hp * WP + i] = s;
  }
}

// Warp-per-channel version for maps that fit shared memory (all but the high-resolution configs): the
// channel is read once with coalesced loads into the warp's smem slab, reductions are warp shuffles
// (no block barriers); every loop walks the PADDED cropped rows in chunks of 8 cells (one division per
// 8 cells) and the operands leave as 16-byte stores.
__global__ void __launch_bounds__(256) gallery_pack_warp_kernel(const float* __restrict__ gal, long long planes, int hg, int wg,
                                                                __half* __restrict__ ghi, __half* __restrict__ glo,
                                                                int32_t* __restrict__ gexp, float* __restrict__ gz,
                                                                float* __restrict__ g32) {
  extern __shared__ float slab[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int Hp = hg - 2 * kEdge, Wp = wg - 2 * kEdge, M = Hp * Wp, HW = hg * wg, WP = gal_pitch(Wp);
  float* ch = slab + (size_t)wid * HW;
  const int c8 = WP / 8, n8 = Hp * c8;
  for (long long gc = (long long)blockIdx.x * nw + wid; gc < planes; gc += (long long)gridDim.x * nw) {
    const float* src = gal + gc * HW;
    warp_load_plane(ch, src, HW, lane);
    __syncwarp();
    double acc = 0.0;
    for (int o = lane; o < n8; o += 32) {
      const int y = o / c8, x0 = (o - y * c8) * 8;
      const float* row = ch + (y + kEdge) * wg + kEdge + x0;
      float part = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) part += (x0 + j < Wp) ? row[j] : 0.0f;
      acc += (double)part;
    }
    const float mean = (float)(warp_sum(acc) / (double)M);
    float amax = 0.0f;
    for (int o = lane; o < n8; o += 32) {
      const int y = o / c8, x0 = (o - y * c8) * 8;
      const float* row = ch + (y + kEdge) * wg + kEdge + x0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (x0 + j < Wp) amax = fmaxf(amax, fabsf(row[j] - mean));
    }
    amax = warp_max(amax);
    int e = 0;
    if (amax > 0.0f && isfinite(amax)) {
      int ex;
      (void)frexpf(amax, &ex);
      e = kGalleryPeakLog2 - ex;
    }
    if (lane == 0) gexp[gc] = e;
    for (int o = lane; o < n8; o += 32) {
      const int y = o / c8, x0 = (o - y * c8) * 8;
      const float* row = ch + (y + kEdge) * wg + kEdge + x0;
      __align__(16) __half h8[8];
      __align__(16) __half l8[8];
      __align__(16) float s8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sc = 0.0f;
        if (x0 + j < Wp) {
          const float z = row[j] - mean;
          sc = ldexpf(z, e);
          if (gz) gz[gc * M + y * Wp + x0 + j] = z;
        }
        s8[j] = sc;
        h8[j] = __float2half_rn(sc);
        l8[j] = __float2half_rn(sc - __half2float(h8[j]));
      }
      *reinterpret_cast<uint4*>(ghi + gc * Hp * WP + (size_t)o * 8) = *reinterpret_cast<const uint4*>(h8);
      *reinterpret_cast<uint4*>(glo + gc * Hp * WP + (size_t)o * 8) = *reinterpret_cast<const uint4*>(l8);
      if (g32) {
        float4* d = reinterpret_cast<float4*>(g32 + gc * Hp * WP + (size_t)o * 8);
        d[0] = *reinterpret_cast<const float4*>(s8);
        d[1] = *reinterpret_cast<const float4*>(s8 + 4);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// K6: window inverse norm through float64 summed-area tables held in shared memory.
// One CTA per (gallery, channel); dynamic smem = 2 * (Hp+1)*(Wp+1) doubles.
// The kernels are bound by instruction issue (ncu: 6.6 k warp instructions per 55x17 plane, FP64 pipe 7 % busy), so the
// per-cell loops carry their (row, column) along instead of dividing, and the table value is one MUFU.RSQ of the float64
// window energy (2 ulp of float32).
struct CellWalk {  // cell i = tid, tid + nthreads, ... of a rows x cols grid as (y, x)
  int y, x, dy, dx, cols;
  __device__ __forceinline__ CellWalk(int tid, int nthreads, int cols_) : cols(cols_) {
    y = tid / cols;
    x = tid - y * cols;
    dy = nthreads / cols;
    dx = nthreads - dy * cols;
  }
  __device__ __forceinline__ void next() {
    y += dy;
    x += dx;
    if (x >= cols) {
      x -= cols;
      ++y;
    }
  }
};

__device__ __forceinline__ void build_sat(const __half* __restrict__ ghi, const __half* __restrict__ glo, const float* __restrict__ gz,
                                          size_t gc, int Hp, int Wp, double* s1, double* s2) {
  const int W1 = Wp + 1, M = Hp * Wp, WP = gal_pitch(Wp);
  CellWalk cw(threadIdx.x, blockDim.x, W1);
  for (int i = threadIdx.x; i < (Hp + 1) * W1; i += blockDim.x, cw.next()) {
    double v = 0.0;
    if (cw.y > 0 && cw.x > 0) {
      const size_t j = gc * M + (size_t)(cw.y - 1) * Wp + (cw.x - 1);
      const size_t jp = (gc * Hp + (size_t)(cw.y - 1)) * WP + (cw.x - 1);
      v = gz ? (double)gz[j] : (double)__half2float(ghi[jp]) + (double)__half2float(glo[jp]);
    }
    s1[i] = v;
    s2[i] = v * v;
  }
  __syncthreads();
  for (int y = 1 + threadIdx.x; y <= Hp; y += blockDim.x) {  // prefix along x
    double a = 0.0, b = 0.0;
    for (int x = 1; x <= Wp; ++x) {
      a += s1[y * W1 + x];
      b += s2[y * W1 + x];
      s1[y * W1 + x] = a;
      s2[y * W1 + x] = b;
    }
  }
  __syncthreads();
  for (int x = 1 + threadIdx.x; x <= Wp; x += blockDim.x) {  // prefix along y
    double a = 0.0, b = 0.0;
    for (int y = 1; y <= Hp; ++y) {
      a += s1[y * W1 + x];
      b += s2[y * W1 + x];
      s1[y * W1 + x] = a;
      s2[y * W1 + x] = b;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void rnorm_from_sat(const double* s1, const double* s2, size_t gc, int Hp, int Wp, int Hm, int Wm,
                                               float* __restrict__ rnorm) {
  const int W1 = Wp + 1, M = Hp * Wp;
  const int a = Hm / 2, b = Wm / 2;
  const double inv_n = 1.0 / ((double)Hm * (double)Wm);
  float* out = rnorm + gc * M;
  CellWalk cw(threadIdx.x, blockDim.x, Wp);
  for (int i = threadIdx.x; i < M; i += blockDim.x, cw.next()) {
    const int r0 = max(cw.y - a, 0) * W1, r1 = min(cw.y - a + Hm, Hp) * W1;
    const int c0 = max(cw.x - b, 0), c1 = min(cw.x - b + Wm, Wp);
    float r = 0.0f;
    if (r1 > r0 && c1 > c0) {
      const double t1 = s1[r1 + c1] - s1[r0 + c1] - s1[r1 + c0] + s1[r0 + c0];
      const double t2 = s2[r1 + c1] - s2[r0 + c1] - s2[r1 + c0] + s2[r0 + c0];
      const double d = t2 - t1 * t1 * inv_n;
      // A window that is flat up to SAT round-off is the reference's "division by ~0" case
      // (similarity.py:69-70 zeroes the non-finite results; FFT noise decides the rest): call it 0.
      if (d > 1e-10 * t2) r = rsqrtf((float)d);
    }
    out[i] = r;
  }
}

__global__ void __launch_bounds__(256) window_rnorm_kernel(const __half* __restrict__ ghi, const __half* __restrict__ glo,
                                                           const float* __restrict__ gz, int Hp, int Wp, int Hm, int Wm,
                                                           float* __restrict__ rnorm) {
  extern __shared__ double sat[];
  double* s2 = sat + (size_t)(Hp + 1) * (Wp + 1);
  build_sat(ghi, glo, gz, blockIdx.x, Hp, Wp, sat, s2);
  rnorm_from_sat(sat, s2, blockIdx.x, Hp, Wp, Hm, Wm, rnorm);
}

// Several template shapes at once: the float64 summed-area tables of a channel are built once and every
// shape's window norm is read off them (ragged probe sets need one table per distinct template shape).
constexpr int kMaxRnormShapes = 24;
struct RnormShapes {
  int n;
  int hm[kMaxRnormShapes], wm[kMaxRnormShapes];
  float* out[kMaxRnormShapes];
};
__global__ void __launch_bounds__(256) window_rnorm_multi_kernel(const __half* __restrict__ ghi, const __half* __restrict__ glo,
                                                                 const float* __restrict__ gz, int Hp, int Wp, RnormShapes sh) {
  extern __shared__ double sat[];
  double* s2 = sat + (size_t)(Hp + 1) * (Wp + 1);
  build_sat(ghi, glo, gz, blockIdx.x, Hp, Wp, sat, s2);
  for (int si = 0; si < sh.n; ++si) rnorm_from_sat(sat, s2, blockIdx.x, Hp, Wp, sh.hm[si], sh.wm[si], sh.out[si]);
}

// ------------------------------------------------------------------------------------------
// K4 rotate: Pillow affine_fixed (Geometry.c) nearest neighbour.  mode 0 copy, 1 flip (180),
// 2 transpose-90, 3 transpose-270 (square maps only), 4 general 16.16 fixed point.
struct RotateCoeffs {
  int mode;
  long long a0, a1, a2, a3, a4, a5;
};

// source cell (row major index into the h x w map) of output cell (y, x), -1 = outside (filled with 0)
__device__ __forceinline__ int rotate_source(const RotateCoeffs& rc, int h, int w, int y, int x) {
  int sy, sx;
  switch (rc.mode) {
    case 0: sy = y; sx = x; break;
    case 1: sy = h - 1 - y; sx = w - 1 - x; break;
    case 2: sy = x; sx = w - 1 - y; break;
    case 3: sy = h - 1 - x; sx = y; break;
    default: {
      const long long xs = (rc.a2 + rc.a1 * y + rc.a0 * x) >> 16;
      const long long ys = (rc.a5 + rc.a4 * y + rc.a3 * x) >> 16;
      if (!(xs >= 0 && xs < w && ys >= 0 && ys < h)) return -1;
      sx = (int)xs;
      sy = (int)ys;
    }
  }
  return sy * w + sx;
}

// The rotation (and, with `transpose`, the change of orientation the host asks for) as an index map: cell p of the
// OUTPUT map ([h][w], or [w][h] when transposed) reads source cell map[p] of the unrotated [h][w] map, -1 = zero fill.
// The template pack applies it while it loads a plane, so the rotated / transposed variant maps are never written.
__global__ void __launch_bounds__(256) rotate_index_map_kernel(int* __restrict__ map, int h, int w, RotateCoeffs rc, int transpose) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= h * w) return;
  const int y = transpose ? p % h : p / w, x = transpose ? p / h : p % w;  // (y, x) of the rotated h x w map this output cell shows
  map[p] = rotate_source(rc, h, w, y, x);
}

__global__ void __launch_bounds__(256) rotate_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                     long long planes, RotateCoeffs rc) {
  // The source index depends only on (y, x): compute it once per thread, then walk the planes
  // (channels x maps) with it -- the gather is shared by all channels and all probes of that shape.
  const int hw = h * w;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const int y = p / w, x = p - y * w;
  const int srci = rotate_source(rc, h, w, y, x);
  const bool ok = srci >= 0;
  const int src = ok ? srci : 0;
  for (long long plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const float v = ok ? __ldg(in + plane * hw + src) : 0.0f;
    out[plane * hw + p] = v;
  }
}

// Plane transpose [P][h][w] -> [P][w][h].  Correlation scores are invariant under transposing both
// maps, and the correlation kernel tiles positions 16 x 8 and template rows by 8/16 taps, so the host
// picks the orientation with less padding (engine.py); planes are a few KB, the strided reads stay in L1.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                        long long planes) {
  const int hw = h * w;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;  // output cell: (x, y) of the h x w input
  if (p >= hw) return;
  const int x = p / h, y = p - x * h;
  const int src = y * w + x;
  for (long long plane = blockIdx.y; plane < planes; plane += gridDim.y) out[plane * hw + p] = __ldg(in + plane * hw + src);
}

// K4 resize: one separable bicubic pass (Pillow Resample.c, 32bpc float path): sequential double
// accumulation of float32 pixel * double weight, no FMA contraction, cast to float32.
// axis 1: along w (in [P][h][n_in] -> out [P][h][n_out]); axis 0: along h.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                       int h_in, int w_in, int h_out, int w_out, int axis,
                                                       const int* __restrict__ xmin, const int* __restrict__ cnt,
                                                       const double* __restrict__ kk, int ksize) {
  // One thread owns one output cell (y, x) and walks the planes (blockIdx.y strided): the tap range and the weights of
  // its output index are fetched once and kept in registers (bicubic enlargement has at most 5 taps; longer filters
  // fall back to reading them per plane).
  const int p = blockIdx.x * blockDim.x + threadIdx.x, hw_out = h_out * w_out, hw_in = h_in * w_in;
  if (p >= hw_out) return;
  const int y = p / w_out, x = p - y * w_out;
  const int o = axis ? x : y;
  const int lo = xmin[o], n = cnt[o];
  const double* k = kk + (size_t)o * ksize;
  const int src0 = axis ? y * w_in + lo : lo * w_in + x, step = axis ? 1 : w_in;
  if (n <= 6) {
    double kr[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) kr[j] = j < n ? k[j] : 0.0;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
      const float* src = in + (size_t)plane * hw_in + src0;
      double ss = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < n) ss = __dadd_rn(ss, __dmul_rn((double)__ldg(src + j * step), kr[j]));
      out[(size_t)plane * hw_out + p] = (float)ss;
    }
  } else {
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
      const float* src = in + (size_t)plane * hw_in + src0;
      double ss = 0.0;
      for (int j = 0; j < n; ++j) ss = __dadd_rn(ss, __dmul_rn((double)__ldg(src + j * step), k[j]));
      out[(size_t)plane * hw_out + p] = (float)ss;
    }
  }
}

// Loader resize (dataloader.py:231-237: Image.resize(size, LANCZOS) on 8-bit images): one pass of Pillow's 8 bits-per-channel
// resampler -- integer coefficients of 22 fractional bits, int32 accumulation from 1 << 21, arithmetic shift, clip to 0..255
// (Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc, normalize_coeffs_8bpc).  axis 1: along w; axis 0: along h.
// in [planes][h_in][w_in][ch] uint8 -> out [planes][h_out][w_out][ch]; bit exact with Pillow.
__global__ void __launch_bounds__(256) resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int planes, int h_in,
                                                          int w_in, int h_out, int w_out, int ch, int axis, const int* __restrict__ xmin,
                                                          const int* __restrict__ cnt, const int* __restrict__ kk, int ksize) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, hwc_out = h_out * w_out * ch;
  if (p >= hwc_out) return;
  const int c = p % ch, x = (p / ch) % w_out, y = p / (ch * w_out);
  const int o = axis ? x : y;
  const int lo = xmin[o], n = cnt[o];
  const int* k = kk + (size_t)o * ksize;
  const int src0 = axis ? (y * w_in + lo) * ch + c : (lo * w_in + x) * ch + c, step = axis ? ch : w_in * ch;
  for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const uint8_t* src = in + (size_t)plane * h_in * w_in * ch + src0;
    int ss = 1 << 21;
    for (int j = 0; j < n; ++j) ss += (int)__ldg(src + (size_t)j * step) * k[j];
    ss >>= 22;
    out[(size_t)plane * hwc_out + p] = (uint8_t)min(255, max(0, ss));
  }
}

// ------------------------------------------------------------------------------------------
// K5 templates: one CTA per (map n, channel c).
__global__ void __launch_bounds__(128) template_pack_kernel(const float* __restrict__ maps, int C, int h, int w, int Hb, int Wb, int col0,
                                                            int ncols_alloc, int row_align, __half* __restrict__ thi,
                                                            __half* __restrict__ tlo, float* __restrict__ t32,
                                                            uint8_t* __restrict__ t8b, uint8_t* __restrict__ t8l, float* __restrict__ t32p,
                                                            const int* __restrict__ gather) {
  __shared__ double sred[32];
  const int Hm = h - 2 * kEdge, Wm = w - 2 * kEdge, K = Hm * Wm;
  const int rowk = tpl_row_taps(Wb, row_align), Kpad = tpl_kpad_aligned(Hb, Wb, row_align);
  const int oy = Hb / 2 - Hm / 2, ox = Wb / 2 - Wm / 2;
  const int n = blockIdx.x / C, c = blockIdx.x - n * C;
  const float* src = maps + ((size_t)n * C + c) * (size_t)h * w;
  auto cell = [&](int i) -> float {  // cell i of the (rotated / transposed) map this column shows
    if (!gather) return src[i];
    const int gi = __ldg(gather + i);
    return gi >= 0 ? src[gi] : 0.0f;
  };

  double acc = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const int u = i / Wm, v = i - u * Wm;
    acc += (double)cell((u + kEdge) * w + v + kEdge);
  }
  const float mean = (float)(block_sum(acc, sred) / (double)K);
  double e = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const int u = i / Wm, v = i - u * Wm;
    const double z = (double)(cell((u + kEdge) * w + v + kEdge) - mean);
    e += z * z;
  }
  e = block_sum(e, sred);
  const double inv = e > 0.0 ? 1.0 / sqrt(e) : 0.0;

  const size_t col = (size_t)c * ncols_alloc + col0 + n;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    const int ub = k / rowk, u = ub - oy, v = k - ub * rowk - ox;
    float tn = 0.0f;
    if (u >= 0 && u < Hm && v >= 0 && v < Wm) {
      tn = (float)((double)(cell((u + kEdge) * w + v + kEdge) - mean) * inv);
      if (t32) t32[col * K + u * Wm + v] = tn;
    }
    const float s = ldexpf(tn, kTemplateScaleLog2);
    const __half hi = __float2half_rn(s);
    const float lo = s - __half2float(hi);
    thi[col * Kpad + k] = hi;
    if (t32p) t32p[col * Kpad + k] = s;
    if (tlo) tlo[col * Kpad + k] = __float2half_rn(lo);
    if (t8b) {  // fp8 copies for the correction MMAs: (B_hi / 64) and (B_lo * 64), see sir_ncc_tc.cu
      t8b[col * Kpad + k] = to_e4m3(__half2float(hi) * kFp8HiScale);
      t8l[col * Kpad + k] = to_e4m3(lo * kFp8LoScale);
    }
  }
}

// Warp-per-(map, channel) version of the template pack, same idea as gallery_pack_warp_kernel.  The kernel is bound by
// instruction issue, not by HBM (ncu: ~2 warp instructions per cycle and SM, DRAM < 50 %), so it is specialised per operand
// set (MODE 0: screening, hi + float32 taps; 1: hi + lo [+ t32]; 2: hi + the two fp8 companions), divides by multiplication
// and does the per-cell arithmetic in float32 (the mean and the energy are still accumulated in float64).
__device__ __forceinline__ int div_small(int n, unsigned magic) { return (int)(((unsigned)n * magic) >> 24); }  // n < 2^16, d < 2^8

template <int MODE>
__global__ void __launch_bounds__(256) template_pack_warp_kernel(const float* __restrict__ maps, long long planes, int C, int h, int w,
                                                                 int Hb, int Wb, int col0, int ncols_alloc, int row_align,
                                                                 __half* __restrict__ thi,
                                                                 __half* __restrict__ tlo, float* __restrict__ t32,
                                                                 uint8_t* __restrict__ t8b, uint8_t* __restrict__ t8l,
                                                                 float* __restrict__ t32p, const int* __restrict__ gather) {
  extern __shared__ float slab[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int Hm = h - 2 * kEdge, Wm = w - 2 * kEdge, K = Hm * Wm, HW = h * w;
  // K layout of the BUCKET shape Hb x Wb (== Hm x Wm for single-shape blocks); the true template sits
  // inside it so that its anchor (Hm/2, Wm/2) lands on the bucket's anchor (Hb/2, Wb/2)
  const int rowk = tpl_row_taps(Wb, row_align), Kpad = tpl_kpad_aligned(Hb, Wb, row_align);
  const int oy = Hb / 2 - Hm / 2, ox = Wb / 2 - Wm / 2;
  float* ch = slab + (size_t)wid * HW;
  const int c8 = rowk / 8, n8 = Kpad / 8, m8 = (Wm + 7) / 8;
  const unsigned magic_c8 = (1u << 24) / (unsigned)c8 + 1, magic_m8 = (1u << 24) / (unsigned)m8 + 1;
  for (long long pc = (long long)blockIdx.x * nw + wid; pc < planes; pc += (long long)gridDim.x * nw) {
    const int n = (int)(pc / C), c = (int)(pc - (long long)n * C);
    const float* src = maps + pc * HW;
    if (gather) {  // the variant (rotation / transposition) is applied on the way in: no variant map in HBM
      for (int i = lane; i < HW; i += 32) {
        const int gi = __ldg(gather + i);
        ch[i] = gi >= 0 ? __ldg(src + gi) : 0.0f;
      }
    } else {
      warp_load_plane(ch, src, HW, lane);
    }
    __syncwarp();
    double acc = 0.0;
    for (int o = lane; o < Hm * m8; o += 32) {
      const int u = div_small(o, magic_m8), v0 = (o - u * m8) * 8;
      const float* row = ch + (u + kEdge) * w + kEdge + v0;
      float part = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) part += (v0 + j < Wm) ? row[j] : 0.0f;
      acc += (double)part;
    }
    const float mean = (float)(warp_sum(acc) / (double)K);
    double e = 0.0;
    for (int o = lane; o < Hm * m8; o += 32) {
      const int u = div_small(o, magic_m8), v0 = (o - u * m8) * 8;
      const float* row = ch + (u + kEdge) * w + kEdge + v0;
      float part = 0.0f;  // eight squares in float32 (all positive: ~1e-7 relative), the 100+ partials in float64
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = (v0 + j < Wm) ? row[j] - mean : 0.0f;
        part = fmaf(z, z, part);
      }
      e += (double)part;
    }
    e = warp_sum(e);
    const double inv = e > 0.0 ? 1.0 / sqrt(e) : 0.0;
    const float inv_f = (float)inv;
    const float inv_scaled = (float)ldexp(inv, kTemplateScaleLog2);  // a power of two: same rounding as scaling afterwards
    const size_t col = (size_t)c * ncols_alloc + col0 + n;
    for (int o = lane; o < n8; o += 32) {
      const int ub = div_small(o, magic_c8), vb0 = (o - ub * c8) * 8;  // bucket coordinates of this 8-tap chunk
      const int u = ub - oy;
      const bool row_in = u >= 0 && u < Hm;
      const float* row = ch + (u + kEdge) * w + kEdge - ox + vb0;
      __align__(16) __half h8[8];
      __align__(16) float s8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int v = vb0 + j - ox;
        const bool in = row_in && v >= 0 && v < Wm;
        const float z = in ? row[j] - mean : 0.0f;
        s8[j] = z * inv_scaled;
        h8[j] = __float2half_rn(s8[j]);
        if (MODE == 1 && t32 && in) t32[col * K + u * Wm + v] = z * inv_f;
      }
      *reinterpret_cast<uint4*>(thi + col * Kpad + (size_t)o * 8) = *reinterpret_cast<const uint4*>(h8);
      if (MODE == 0) {
        float4* d = reinterpret_cast<float4*>(t32p + col * Kpad + (size_t)o * 8);
        d[0] = *reinterpret_cast<const float4*>(s8);
        d[1] = *reinterpret_cast<const float4*>(s8 + 4);
      } else if (MODE == 1) {
        if (!tlo) continue;
        __align__(16) __half l8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) l8[j] = __float2half_rn(s8[j] - __half2float(h8[j]));
        *reinterpret_cast<uint4*>(tlo + col * Kpad + (size_t)o * 8) = *reinterpret_cast<const uint4*>(l8);
      } else {  // fp8 copies for the correction MMAs: (B_hi / 64) and (B_lo * 64), see sir_ncc_tc.cu
        __align__(8) uint8_t b8[8];
        __align__(8) uint8_t q8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b8[j] = to_e4m3(__half2float(h8[j]) * kFp8HiScale);
          q8[j] = to_e4m3((s8[j] - __half2float(h8[j])) * kFp8LoScale);
        }
        *reinterpret_cast<uint2*>(t8b + col * Kpad + (size_t)o * 8) = *reinterpret_cast<const uint2*>(b8);
        *reinterpret_cast<uint2*>(t8l + col * Kpad + (size_t)o * 8) = *reinterpret_cast<const uint2*>(q8);
      }
    }
    __syncwarp();
  }
}

// fp8 companions of the packed gallery: a8 = e4m3(hi / 64), l8 = e4m3(lo * 64); rows padded to 16 cells
__global__ void __launch_bounds__(256) gallery_fp8_kernel(const __half* __restrict__ ghi, const __half* __restrict__ glo, int Hp,
                                                          int Wp, uint8_t* __restrict__ g8a, uint8_t* __restrict__ g8l) {
  const int WP = gal_pitch(Wp), WP8 = gal_pitch8(Wp);
  const size_t gc = blockIdx.x;
  for (int i = threadIdx.x; i < Hp * WP8; i += blockDim.x) {
    const int y = i / WP8, x = i - y * WP8;
    uint8_t a = 0, l = 0;
    if (x < Wp) {
      const size_t j = (gc * Hp + y) * WP + x;
      a = to_e4m3(__half2float(ghi[j]) * kFp8HiScale);
      l = to_e4m3(__half2float(glo[j]) * kFp8LoScale);
    }
    g8a[gc * Hp * WP8 + i] = a;
    g8l[gc * Hp * WP8 + i] = l;
  }
}

}  // namespace sir

// =========================================================================== C ABI launchers
#include <cmath>
#include <cstdlib>
#include <vector>

using namespace sir;

extern "C" int sir_gallery_pack(const float* d_gallery, int G, int C, int hg, int wg, uint16_t* d_ghi, uint16_t* d_glo,
                                int32_t* d_gexp, float* d_gz, void* stream) {
  return sir_gallery_pack_f32(d_gallery, G, C, hg, wg, d_ghi, d_glo, d_gexp, d_gz, nullptr, stream);
}

extern "C" int sir_gallery_pack_f32(const float* d_gallery, int G, int C, int hg, int wg, uint16_t* d_ghi, uint16_t* d_glo,
                                    int32_t* d_gexp, float* d_gz, float* d_g32, void* stream) {
  SIR_CHECK_ARG(d_gallery && d_ghi && d_glo && d_gexp, "sir_gallery_pack: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0, "sir_gallery_pack: empty gallery (G=%d C=%d)", G, C);
  SIR_CHECK_ARG(hg > 2 * kEdge && wg > 2 * kEdge, "sir_gallery_pack: map %dx%d vanishes after the 2-cell crop", hg, wg);
  const size_t slab = (size_t)hg * wg * sizeof(float);
  if (8 * slab <= 48 * 1024) {
    const long long planes = (long long)G * C;
    const unsigned blocks = (unsigned)std::min<long long>((planes + 7) / 8, 148 * 8);
    gallery_pack_warp_kernel<<<blocks, 256, 8 * slab, (cudaStream_t)stream>>>(d_gallery, planes, hg, wg, (__half*)d_ghi, (__half*)d_glo,
                                                                               d_gexp, d_gz, d_g32);
  } else {
    gallery_pack_kernel<<<(unsigned)((size_t)G * C), 256, 0, (cudaStream_t)stream>>>(
        d_gallery, C, hg, wg, (__half*)d_ghi, (__half*)d_glo, d_gexp, d_gz, d_g32);
  }
  SIR_LAUNCH_CHECK("gallery_pack_kernel");
  return SIR_OK;
}

extern "C" int sir_gallery_window_rnorm(const uint16_t* d_ghi, const uint16_t* d_glo, const float* d_gz, int G, int C,
                                        int Hp, int Wp, int Hm, int Wm, float* d_rnorm, void* stream) {
  SIR_CHECK_ARG((d_gz || (d_ghi && d_glo)) && d_rnorm, "sir_gallery_window_rnorm: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0 && Hm > 0 && Wm > 0, "sir_gallery_window_rnorm: bad shape");
  const size_t smem = 2 * (size_t)(Hp + 1) * (Wp + 1) * sizeof(double);
  SIR_CHECK_ARG(smem <= 227 * 1024, "sir_gallery_window_rnorm: map %dx%d too large for the smem SAT", Hp, Wp);
  static thread_local size_t configured_dev[64] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    SIR_CUDA(cudaFuncSetAttribute(window_rnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  window_rnorm_kernel<<<(unsigned)((size_t)G * C), 256, smem, (cudaStream_t)stream>>>(
      (const __half*)d_ghi, (const __half*)d_glo, d_gz, Hp, Wp, Hm, Wm, d_rnorm);
  SIR_LAUNCH_CHECK("window_rnorm_kernel");
  return SIR_OK;
}

extern "C" int sir_gallery_window_rnorm_multi(const uint16_t* d_ghi, const uint16_t* d_glo, int G, int C, int Hp, int Wp, int nshapes,
                                              const int* h_hm, const int* h_wm, float* const* h_out, void* stream) {
  SIR_CHECK_ARG(d_ghi && d_glo && h_hm && h_wm && h_out, "sir_gallery_window_rnorm_multi: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0 && nshapes > 0, "sir_gallery_window_rnorm_multi: bad shape");
  const size_t smem = 2 * (size_t)(Hp + 1) * (Wp + 1) * sizeof(double);
  SIR_CHECK_ARG(smem <= 227 * 1024, "sir_gallery_window_rnorm_multi: map %dx%d too large for the smem SAT", Hp, Wp);
  static thread_local size_t configured_dev[64] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    SIR_CUDA(cudaFuncSetAttribute(window_rnorm_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  for (int s0 = 0; s0 < nshapes; s0 += kMaxRnormShapes) {
    RnormShapes sh{};
    sh.n = std::min(kMaxRnormShapes, nshapes - s0);
    for (int i = 0; i < sh.n; ++i) {
      SIR_CHECK_ARG(h_hm[s0 + i] > 0 && h_wm[s0 + i] > 0 && h_out[s0 + i], "sir_gallery_window_rnorm_multi: bad shape entry %d", s0 + i);
      sh.hm[i] = h_hm[s0 + i];
      sh.wm[i] = h_wm[s0 + i];
      sh.out[i] = h_out[s0 + i];
    }
    window_rnorm_multi_kernel<<<(unsigned)((size_t)G * C), 256, smem, (cudaStream_t)stream>>>((const __half*)d_ghi, (const __half*)d_glo,
                                                                                              nullptr, Hp, Wp, sh);
    SIR_LAUNCH_CHECK("window_rnorm_multi_kernel");
  }
  return SIR_OK;
}

// Python's round(x, 15) == correctly rounded decimal with 15 fractional digits, re-parsed.
static double round15(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.15f", v);
  return strtod(buf, nullptr);
}
static long long fix16(double v) { return (long long)std::floor(v * 65536.0 + 0.5); }

// Pillow's rotate set-up (Image.rotate -> affine_fixed): mode + 16.16 fixed point coefficients for an h x w map.
static int rotate_coeffs(int h, int w, double angle, sir::RotateCoeffs* out) {
  sir::RotateCoeffs rc{};
  double a = std::fmod(angle, 360.0);
  if (a < 0) a += 360.0;  // Python's % on floats
  if (a == 0.0) rc.mode = 0;
  else if (a == 180.0) rc.mode = 1;
  else if (a == 90.0 && w == h) rc.mode = 2;
  else if (a == 270.0 && w == h) rc.mode = 3;
  else {
    rc.mode = 4;
    const double r = -(a * (M_PI / 180.0));
    // math.radians(a) is a * (pi / 180); keep the same association
    const double m0 = round15(std::cos(r)), m1 = round15(std::sin(r));
    const double m3 = round15(-std::sin(r)), m4 = round15(std::cos(r));
    const double cx = w / 2.0, cy = h / 2.0;
    const double m2 = m0 * (-cx) + m1 * (-cy) + 0.0 + cx;
    const double m5 = m3 * (-cx) + m4 * (-cy) + 0.0 + cy;
    rc.a0 = fix16(m0); rc.a1 = fix16(m1); rc.a3 = fix16(m3); rc.a4 = fix16(m4);
    rc.a2 = fix16(m2 + m0 * 0.5 + m1 * 0.5);
    rc.a5 = fix16(m5 + m3 * 0.5 + m4 * 0.5);
  }
  *out = rc;
  return SIR_OK;
}

extern "C" int sir_variant_rotate(const float* d_in, int N, int C, int h, int w, double angle, float* d_out,
                                  void* stream) {
  SIR_CHECK_ARG(d_in && d_out, "sir_variant_rotate: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0 && h > 0 && w > 0, "sir_variant_rotate: bad shape");
  SIR_CHECK_ARG(h < 32768 && w < 32768, "sir_variant_rotate: map too large for the 16.16 fixed point walk");
  RotateCoeffs rc{};
  int rcode = rotate_coeffs(h, w, angle, &rc);
  if (rcode) return rcode;
  const long long planes = (long long)N * C;
  const unsigned bx = (unsigned)ceil_div(h * w, 256);
  const unsigned by = (unsigned)std::min<long long>(planes, std::max(1u, 148u * 16u / bx));
  rotate_kernel<<<dim3(bx, by), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, h, w, planes, rc);
  SIR_LAUNCH_CHECK("rotate_kernel");
  return SIR_OK;
}

extern "C" int sir_maps_transpose(const float* d_in, int N, int C, int h, int w, float* d_out, void* stream) {
  SIR_CHECK_ARG(d_in && d_out && N > 0 && C > 0 && h > 0 && w > 0, "sir_maps_transpose: bad argument");
  const long long planes = (long long)N * C;
  const unsigned bx = (unsigned)ceil_div(h * w, 256);
  const unsigned by = (unsigned)std::min<long long>(planes, std::max(1u, 148u * 16u / bx));
  transpose_kernel<<<dim3(bx, by), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, h, w, planes);
  SIR_LAUNCH_CHECK("transpose_kernel");
  return SIR_OK;
}

namespace {
struct Coeffs {
  int ksize;
  std::vector<int> xmin, cnt;
  std::vector<double> kk;
};
double bicubic(double x) {
  const double a = -0.5;
  if (x < 0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return std::sin(x) / x;
}
double lanczos3(double x) { return (-3.0 <= x && x < 3.0) ? sinc_filter(x) * sinc_filter(x / 3) : 0.0; }  // truncated sinc
// Pillow Resample.c precompute_coeffs: bicubic (support 2.0) or, with `lanczos`, LANCZOS (support 3.0).
Coeffs precompute(int n_in, int n_out, bool lanczos = false) {
  Coeffs c;
  const double scale = (double)n_in / n_out;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = (lanczos ? 3.0 : 2.0) * fscale;
  c.ksize = (int)std::ceil(support) * 2 + 1;
  c.xmin.assign(n_out, 0);
  c.cnt.assign(n_out, 0);
  c.kk.assign((size_t)n_out * c.ksize, 0.0);
  const double ss = 1.0 / fscale;
  for (int xx = 0; xx < n_out; ++xx) {
    const double center = (xx + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > n_in) hi = n_in;
    const int n = hi - lo;
    double* k = &c.kk[(size_t)xx * c.ksize];
    double ww = 0.0;
    for (int i = 0; i < n; ++i) {
      const double wgt = lanczos ? lanczos3((i + lo - center + 0.5) * ss) : bicubic((i + lo - center + 0.5) * ss);
      k[i] = wgt;
      ww += wgt;
    }
    if (ww != 0.0)
      for (int i = 0; i < n; ++i) k[i] /= ww;
    c.xmin[xx] = lo;
    c.cnt[xx] = n;
  }
  return c;
}

// bytes of one pass's tap tables in the workspace: xmin, cnt (int per output index) + weights (double, ksize per index)
size_t pass_ws_bytes(int n_in, int n_out) {
  const double scale = (double)n_in / n_out, fscale = scale < 1.0 ? 1.0 : scale;
  const int ksize = (int)std::ceil(2.0 * fscale) * 2 + 1;
  return (size_t)round_up(2 * n_out * (int)sizeof(int), 16) + (size_t)n_out * ksize * sizeof(double);
}

int run_pass(const float* in, float* out, int planes, int h_in, int w_in, int h_out, int w_out, int axis, void* d_ws,
             cudaStream_t st) {
  const Coeffs c = precompute(axis ? w_in : h_in, axis ? w_out : h_out);
  const int n_out = axis ? w_out : h_out;
  int* d_xmin = reinterpret_cast<int*>(d_ws);
  int* d_cnt = d_xmin + n_out;
  double* d_kk = reinterpret_cast<double*>(reinterpret_cast<char*>(d_ws) + round_up(2 * n_out * (int)sizeof(int), 16));
  // pageable sources: the runtime stages them before returning, so the vectors may die here
  SIR_CUDA(cudaMemcpyAsync(d_xmin, c.xmin.data(), sizeof(int) * n_out, cudaMemcpyHostToDevice, st));
  SIR_CUDA(cudaMemcpyAsync(d_cnt, c.cnt.data(), sizeof(int) * n_out, cudaMemcpyHostToDevice, st));
  SIR_CUDA(cudaMemcpyAsync(d_kk, c.kk.data(), sizeof(double) * c.kk.size(), cudaMemcpyHostToDevice, st));
  const unsigned bx = (unsigned)ceil_div(h_out * w_out, 256);
  const unsigned by = (unsigned)std::min<long long>(planes, std::max(1u, 148u * 16u / bx));
  resample_kernel<<<dim3(bx, by), 256, 0, st>>>(in, out, planes, h_in, w_in, h_out, w_out, axis, d_xmin, d_cnt, d_kk, c.ksize);
  SIR_LAUNCH_CHECK("resample_kernel");
  return SIR_OK;
}
}  // namespace

extern "C" size_t sir_variant_resize_workspace_bytes(int h, int w, int h2, int w2) {
  if (h <= 0 || w <= 0 || h2 <= 0 || w2 <= 0) return 0;
  size_t n = 0;
  if (w2 != w) n += round_up((int)pass_ws_bytes(w, w2), 256);
  if (h2 != h) n += round_up((int)pass_ws_bytes(h, h2), 256);
  return n;
}

extern "C" int sir_variant_resize(const float* d_in, int N, int C, int h, int w, int h2, int w2, float* d_out,
                                  float* d_tmp, void* d_ws, size_t ws_bytes, void* stream) {
  SIR_CHECK_ARG(d_in && d_out, "sir_variant_resize: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0 && h > 0 && w > 0 && h2 > 0 && w2 > 0, "sir_variant_resize: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = N * C;
  if (h2 == h && w2 == w) {  // Image.resize with an unchanged size is a copy
    SIR_CUDA(cudaMemcpyAsync(d_out, d_in, sizeof(float) * (size_t)planes * h * w, cudaMemcpyDeviceToDevice, st));
    return SIR_OK;
  }
  SIR_CHECK_ARG(d_ws && ws_bytes >= sir_variant_resize_workspace_bytes(h, w, h2, w2) && (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0,
                "sir_variant_resize: workspace of %zu bytes (16-byte aligned) needed, see sir_variant_resize_workspace_bytes",
                sir_variant_resize_workspace_bytes(h, w, h2, w2));
  char* ws = reinterpret_cast<char*>(d_ws);
  if (w2 != w && h2 != h) {
    SIR_CHECK_ARG(d_tmp, "sir_variant_resize: two-pass resize needs d_tmp");
    int rc = run_pass(d_in, d_tmp, planes, h, w, h, w2, 1, ws, st);
    if (rc) return rc;
    return run_pass(d_tmp, d_out, planes, h, w2, h2, w2, 0, ws + round_up((int)pass_ws_bytes(w, w2), 256), st);
  }
  if (w2 != w) return run_pass(d_in, d_out, planes, h, w, h, w2, 1, ws, st);
  return run_pass(d_in, d_out, planes, h, w, h2, w, 0, ws, st);
}

namespace {
size_t pass_u8_ws_bytes(int n_in, int n_out) {
  const double scale = (double)n_in / n_out, fscale = scale < 1.0 ? 1.0 : scale;
  const int ksize = (int)std::ceil(3.0 * fscale) * 2 + 1;
  return (size_t)round_up((2 + ksize) * n_out * (int)sizeof(int), 256);
}
int run_pass_u8(const uint8_t* in, uint8_t* out, int planes, int h_in, int w_in, int h_out, int w_out, int ch, int axis, void* d_ws,
                cudaStream_t st) {
  const Coeffs c = precompute(axis ? w_in : h_in, axis ? w_out : h_out, true);
  const int n_out = axis ? w_out : h_out;
  std::vector<int> tab((size_t)(2 + c.ksize) * n_out);
  for (int i = 0; i < n_out; ++i) {
    tab[i] = c.xmin[i];
    tab[n_out + i] = c.cnt[i];
  }
  for (size_t i = 0; i < c.kk.size(); ++i) {  // normalize_coeffs_8bpc
    const double v = c.kk[i];
    tab[2 * (size_t)n_out + i] = v < 0 ? (int)(-0.5 + v * (double)(1 << 22)) : (int)(0.5 + v * (double)(1 << 22));
  }
  int* d_tab = reinterpret_cast<int*>(d_ws);
  SIR_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, st));  // pageable: staged before returning
  const unsigned bx = (unsigned)ceil_div(h_out * w_out * ch, 256);
  const unsigned by = (unsigned)std::min<long long>(planes, std::max(1u, 148u * 16u / bx));
  resample_u8_kernel<<<dim3(bx, by), 256, 0, st>>>(in, out, planes, h_in, w_in, h_out, w_out, ch, axis, d_tab, d_tab + n_out, d_tab + 2 * n_out,
                                                   c.ksize);
  SIR_LAUNCH_CHECK("resample_u8_kernel");
  return SIR_OK;
}
}  // namespace

extern "C" size_t sir_image_resize_workspace_bytes(int h, int w, int h2, int w2) {
  if (h <= 0 || w <= 0 || h2 <= 0 || w2 <= 0) return 0;
  return (w2 != w ? pass_u8_ws_bytes(w, w2) : 0) + (h2 != h ? pass_u8_ws_bytes(h, h2) : 0);
}

extern "C" int sir_image_resize_lanczos(const uint8_t* d_in, int N, int h, int w, int ch, int h2, int w2, uint8_t* d_out, uint8_t* d_tmp,
                                        void* d_ws, size_t ws_bytes, void* stream) {
  SIR_CHECK_ARG(d_in && d_out, "sir_image_resize_lanczos: null pointer");
  SIR_CHECK_ARG(N > 0 && h > 0 && w > 0 && h2 > 0 && w2 > 0 && ch >= 1 && ch <= 4, "sir_image_resize_lanczos: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (h2 == h && w2 == w) {
    SIR_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)N * h * w * ch, cudaMemcpyDeviceToDevice, st));
    return SIR_OK;
  }
  SIR_CHECK_ARG(d_ws && ws_bytes >= sir_image_resize_workspace_bytes(h, w, h2, w2) && (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0,
                "sir_image_resize_lanczos: workspace of %zu bytes (16-byte aligned) needed", sir_image_resize_workspace_bytes(h, w, h2, w2));
  char* ws = reinterpret_cast<char*>(d_ws);
  if (w2 != w && h2 != h) {  // horizontal pass first, 8-bit intermediate, as ImagingResample does
    SIR_CHECK_ARG(d_tmp, "sir_image_resize_lanczos: two-pass resize needs d_tmp [N][h][w2][ch]");
    int rc = run_pass_u8(d_in, d_tmp, N, h, w, h, w2, ch, 1, ws, st);
    if (rc) return rc;
    return run_pass_u8(d_tmp, d_out, N, h, w2, h2, w2, ch, 0, ws + pass_u8_ws_bytes(w, w2), st);
  }
  if (w2 != w) return run_pass_u8(d_in, d_out, N, h, w, h, w2, ch, 1, ws, st);
  return run_pass_u8(d_in, d_out, N, h, w, h2, w, ch, 0, ws, st);
}

extern "C" int sir_gallery_pitch(int Wp) { return Wp > 0 ? gal_pitch(Wp) : 0; }

static void launch_template_pack(const float* d_maps, int N, int C, int h, int w, int Hb, int Wb, int col0, int ncols_alloc, int row_align,
                                 __half* thi, __half* tlo, float* t32, uint8_t* t8b, uint8_t* t8l, cudaStream_t st, float* t32p = nullptr,
                                 const int* gather = nullptr) {
  const size_t slab = (size_t)h * w * sizeof(float);
  if (8 * slab <= 48 * 1024) {
    const long long planes = (long long)N * C;
    const unsigned blocks = (unsigned)std::min<long long>((planes + 7) / 8, 148 * 8);
    auto kern = t32p ? template_pack_warp_kernel<0> : (t8b ? template_pack_warp_kernel<2> : template_pack_warp_kernel<1>);
    kern<<<blocks, 256, 8 * slab, st>>>(d_maps, planes, C, h, w, Hb, Wb, col0, ncols_alloc, row_align, thi, tlo, t32, t8b, t8l, t32p, gather);
  } else {
    template_pack_kernel<<<(unsigned)((size_t)N * C), 128, 0, st>>>(d_maps, C, h, w, Hb, Wb, col0, ncols_alloc, row_align, thi, tlo, t32, t8b,
                                                                    t8l, t32p, gather);
  }
}

extern "C" int sir_gallery_pitch8(int Wp) { return Wp > 0 ? gal_pitch8(Wp) : 0; }

extern "C" int sir_gallery_pack_fp8c(const uint16_t* d_ghi, const uint16_t* d_glo, int G, int C, int Hp, int Wp, uint8_t* d_g8a,
                                     uint8_t* d_g8l, void* stream) {
  SIR_CHECK_ARG(d_ghi && d_glo && d_g8a && d_g8l, "sir_gallery_pack_fp8c: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0, "sir_gallery_pack_fp8c: bad shape");
  gallery_fp8_kernel<<<(unsigned)((size_t)G * C), 256, 0, (cudaStream_t)stream>>>((const __half*)d_ghi, (const __half*)d_glo, Hp, Wp,
                                                                                   d_g8a, d_g8l);
  SIR_LAUNCH_CHECK("gallery_fp8_kernel");
  return SIR_OK;
}

extern "C" int sir_template_kpad(int Hm, int Wm) { return (Hm > 0 && Wm > 0) ? tpl_kpad(Hm, Wm) : 0; }
extern "C" int sir_template_kpad_fp8c(int Hm, int Wm) { return (Hm > 0 && Wm > 0) ? tpl_kpad_aligned(Hm, Wm, 16) : 0; }

extern "C" int sir_template_pack_fp8c(const float* d_maps, int N, int C, int h, int w, int col0, int ncols_alloc, uint16_t* d_thi,
                                      uint8_t* d_t8b, uint8_t* d_t8l, void* stream) {
  SIR_CHECK_ARG(d_maps && d_thi && d_t8b && d_t8l, "sir_template_pack_fp8c: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0, "sir_template_pack_fp8c: empty input");
  SIR_CHECK_ARG(h > 2 * kEdge && w > 2 * kEdge, "sir_template_pack_fp8c: map %dx%d vanishes after the 2-cell crop", h, w);
  SIR_CHECK_ARG(col0 >= 0 && col0 + N <= ncols_alloc, "sir_template_pack_fp8c: columns [%d,%d) outside %d", col0, col0 + N, ncols_alloc);
  launch_template_pack(d_maps, N, C, h, w, h - 2 * kEdge, w - 2 * kEdge, col0, ncols_alloc, 16, (__half*)d_thi, nullptr, nullptr, d_t8b, d_t8l,
                       (cudaStream_t)stream);
  SIR_LAUNCH_CHECK("template_pack_kernel");
  return SIR_OK;
}

extern "C" int sir_template_pack(const float* d_maps, int N, int C, int h, int w, int col0, int ncols_alloc,
                                 uint16_t* d_thi, uint16_t* d_tlo, float* d_t32, void* stream) {
  SIR_CHECK_ARG(d_maps && d_thi && d_tlo, "sir_template_pack: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0, "sir_template_pack: empty input");
  SIR_CHECK_ARG(h > 2 * kEdge && w > 2 * kEdge, "sir_template_pack: map %dx%d vanishes after the 2-cell crop", h, w);
  SIR_CHECK_ARG(col0 >= 0 && col0 + N <= ncols_alloc, "sir_template_pack: columns [%d,%d) outside %d", col0, col0 + N,
                ncols_alloc);
  launch_template_pack(d_maps, N, C, h, w, h - 2 * kEdge, w - 2 * kEdge, col0, ncols_alloc, 8, (__half*)d_thi, (__half*)d_tlo, d_t32, nullptr,
                       nullptr, (cudaStream_t)stream);
  SIR_LAUNCH_CHECK("template_pack_kernel");
  return SIR_OK;
}

// Multi-shape column tiles: pack templates of true shape (h-4) x (w-4) into the K layout of a BUCKET shape
// Hb x Wb (>= the true shape), anchor on anchor, zeros elsewhere.  precision selects the row alignment and
// which companion operands are written (FP16X3: d_tlo; FP16_FP8C: d_t8b, d_t8l).
extern "C" int sir_template_pack_embed(const float* d_maps, int N, int C, int h, int w, int Hb, int Wb, int col0, int ncols_alloc,
                                       int precision, uint16_t* d_thi, uint16_t* d_tlo, uint8_t* d_t8b, uint8_t* d_t8l, void* stream) {
  SIR_CHECK_ARG(d_maps && d_thi, "sir_template_pack_embed: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0 && h > 2 * kEdge && w > 2 * kEdge, "sir_template_pack_embed: bad input shape");
  SIR_CHECK_ARG(Hb >= h - 2 * kEdge && Wb >= w - 2 * kEdge, "sir_template_pack_embed: bucket %dx%d smaller than template %dx%d", Hb, Wb,
                h - 2 * kEdge, w - 2 * kEdge);
  SIR_CHECK_ARG(col0 >= 0 && col0 + N <= ncols_alloc, "sir_template_pack_embed: columns [%d,%d) outside %d", col0, col0 + N, ncols_alloc);
  const bool fp8c = precision == SIR_PREC_FP16_FP8C;
  SIR_CHECK_ARG(fp8c ? (d_t8b && d_t8l) : (d_tlo != nullptr), "sir_template_pack_embed: missing companion operands for precision %d", precision);
  launch_template_pack(d_maps, N, C, h, w, Hb, Wb, col0, ncols_alloc, fp8c ? 16 : 8, (__half*)d_thi, fp8c ? nullptr : (__half*)d_tlo, nullptr,
                       fp8c ? d_t8b : nullptr, fp8c ? d_t8l : nullptr, (cudaStream_t)stream);
  SIR_LAUNCH_CHECK("template_pack_kernel");
  return SIR_OK;
}

// Index map of a rotation (+ optional transposition) for sir_template_pack_screen's d_gather.
extern "C" int sir_variant_index_map(int h, int w, double angle, int transpose, int32_t* d_map, void* stream) {
  SIR_CHECK_ARG(d_map && h > 0 && w > 0 && h < 32768 && w < 32768, "sir_variant_index_map: bad argument");
  RotateCoeffs rc{};
  int rcode = rotate_coeffs(h, w, angle, &rc);
  if (rcode) return rcode;
  rotate_index_map_kernel<<<ceil_div(h * w, 256), 256, 0, (cudaStream_t)stream>>>(d_map, h, w, rc, transpose ? 1 : 0);
  SIR_LAUNCH_CHECK("rotate_index_map_kernel");
  return SIR_OK;
}

// Screen + refine operands (SIR_PREC_FP16_REFINE): d_thi as sir_template_pack (rows padded to 8 taps) for the tensor-core
// screening pass and d_t32p [C][ncols_alloc][Kpad] float32 = (t - mean)/sqrt(E) * 2^10 in the same padded K layout for the
// exact re-evaluation.  Hb x Wb >= the true template shape selects a bucket layout (anchor on anchor, zeros elsewhere) as in
// sir_template_pack_embed; pass the true shape for a single-shape block.
extern "C" int sir_template_pack_screen(const float* d_maps, int N, int C, int h, int w, int Hb, int Wb, int col0, int ncols_alloc,
                                        uint16_t* d_thi, float* d_t32p, const int32_t* d_gather, void* stream) {
  SIR_CHECK_ARG(d_maps && d_thi && d_t32p, "sir_template_pack_screen: null pointer");
  SIR_CHECK_ARG(N > 0 && C > 0 && h > 2 * kEdge && w > 2 * kEdge, "sir_template_pack_screen: bad input shape");
  SIR_CHECK_ARG(Hb >= h - 2 * kEdge && Wb >= w - 2 * kEdge, "sir_template_pack_screen: bucket %dx%d smaller than template %dx%d", Hb, Wb,
                h - 2 * kEdge, w - 2 * kEdge);
  SIR_CHECK_ARG(col0 >= 0 && col0 + N <= ncols_alloc, "sir_template_pack_screen: columns [%d,%d) outside %d", col0, col0 + N, ncols_alloc);
  launch_template_pack(d_maps, N, C, h, w, Hb, Wb, col0, ncols_alloc, 8, (__half*)d_thi, nullptr, nullptr, nullptr, nullptr,
                       (cudaStream_t)stream, d_t32p, d_gather);
  SIR_LAUNCH_CHECK("template_pack_kernel");
  return SIR_OK;
}
