// Feature stage (K1-K3): the truncated CNN backbone of network.py:185-186,234-235 as hand-written
// sm_100a kernels, operator by operator.  Activations are float32 NHWC in HBM.
//
//   K1  convolution (groups=1): the tensor-core kernels live in sir_conv_tc.cu (implicit GEMM over fp16 hi/lo
//       NHWC planes).  Here: im2col_split_kernel = the split pass (float32 -> hi/lo planes with a per-tensor
//       power-of-two scale from a device-side running |max|, optional squeeze-excitation channel scale), which
//       for strided convolutions also gathers an explicit im2col matrix; conv_c3k3_kernel = the 3-channel stem
//       in float32 on the CUDA cores.
//   K2  depthwise k x k convolution + bias + activation (CUDA cores, HBM bound), the squeeze of a following
//       SqueezeExcitation fused as partial sums.
//   K3  squeeze-excitation MLP (two tiny FCs, SiLU, sigmoid); the channel scale is applied by the split pass
//       of the following 1x1 projection, i.e. fused into its operand load.
//
// Reference semantics: torchvision Conv2d / BatchNorm2d(eval) / SiLU / SqueezeExcitation /
// MaxPool2d as composed by torchvision.models.efficientnet / vgg (third party; the reference only
// selects and truncates them, network.py:121-186).  ToTensor + Normalize (network.py:51-87) are
// image_to_nhwc_kernel.
#include <cuda.h>
#include <cuda_fp16.h>

#include "sir_common.cuh"
#include "sir_ptx.cuh"

namespace sir {

// ------------------------------------------------------------------------------------------ misc
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return v / (1.0f + expf(-v));  // SiLU
  if (act == 2) return fmaxf(v, 0.0f);         // ReLU
  return v;
}

// uint8 image(s) -> normalised float32 NHWC with 3 channels (network.py:51-87: ToTensor, repeat to
// 3 channels for grayscale, Normalize(mean, std)).
__global__ void __launch_bounds__(256) image_to_nhwc_kernel(const uint8_t* __restrict__ img, int in_ch, size_t pixels,
                                                            float m0, float m1, float m2, float s0, float s1, float s2,
                                                            float* __restrict__ out, float* __restrict__ amax) {
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) {
    float v[3];
    if (in_ch == 1) {
      const float x = __fdiv_rn((float)img[i], 255.0f);
      v[0] = v[1] = v[2] = x;
    } else {
      v[0] = __fdiv_rn((float)img[3 * i], 255.0f);
      v[1] = __fdiv_rn((float)img[3 * i + 1], 255.0f);
      v[2] = __fdiv_rn((float)img[3 * i + 2], 255.0f);
    }
    v[0] = __fdiv_rn(v[0] - m0, s0);
    v[1] = __fdiv_rn(v[1] - m1, s1);
    v[2] = __fdiv_rn(v[2] - m2, s2);
    out[3 * i] = v[0];
    out[3 * i + 1] = v[1];
    out[3 * i + 2] = v[2];
    local = fmaxf(local, fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(amax, local);
}

// ------------------------------------------------------------------------------------------ CLAHE
// OpenCV CLAHE (modules/imgproc/src/clahe.cpp; network.py:108-111,197-208) for uint8 grayscale, bit exact:
// per tile a 256-bin histogram of the (REFLECT_101 extended) image, clipped and redistributed, its
// cumulative sum scaled to 0..255 is the tile's LUT; every pixel is the bilinear blend (float32, products
// and sums rounded separately, round half to even) of the LUT values of the four nearest tile centres.
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ img, int H, int W, int tiles_x, int th, int tw,
                                                        int clip, float lut_scale, uint8_t* __restrict__ lut) {
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int red[8];
  const int tile = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const uint8_t* src = img + (size_t)b * H * W;
  hist[t] = 0;
  __syncthreads();
  for (int i = t; i < th * tw; i += 256) {
    int y = ty * th + i / tw, x = tx * tw + i % tw;
    if (y >= H) y = 2 * H - 2 - y;  // BORDER_REFLECT_101 extension to a multiple of the grid
    if (x >= W) x = 2 * W - 2 - x;
    atomicAdd(&hist[src[(size_t)y * W + x]], 1);
  }
  __syncthreads();
  int hv = hist[t];
  if (clip > 0) {
    int over = max(hv - clip, 0);
    hv = min(hv, clip);
    over = warp_sum(over);
    if ((t & 31) == 0) red[t >> 5] = over;
    __syncthreads();
    int clipped = 0;
    for (int i = 0; i < 8; ++i) clipped += red[i];
    const int batch = clipped / 256, residual = clipped - batch * 256;
    hv += batch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      if (t % step == 0 && t / step < residual) hv += 1;
    }
  }
  // inclusive scan over the 256 bins
  scan[t] = hv;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int add = t >= o ? scan[t - o] : 0;
    __syncthreads();
    scan[t] += add;
    __syncthreads();
  }
  const int v = __float2int_rn(__fmul_rn((float)scan[t], lut_scale));
  lut[((size_t)b * gridDim.x + tile) * 256 + t] = (uint8_t)min(max(v, 0), 255);
}

// the equalised value of pixel (x, y) of image b: bilinear blend of the four nearest tiles' LUT entries for value v
__device__ __forceinline__ int clahe_interpolate(int v, int x, int y, int b, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                                                 const uint8_t* __restrict__ lut) {
  const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
  const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
  const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
  const int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
  tx1 = max(tx1, 0);
  ty1 = max(ty1, 0);
  const uint8_t* lb = lut + (size_t)b * tiles_x * tiles_y * 256 + v;
  const float l11 = lb[(ty1 * tiles_x + tx1) * 256], l12 = lb[(ty1 * tiles_x + tx2) * 256];
  const float l21 = lb[(ty2 * tiles_x + tx1) * 256], l22 = lb[(ty2 * tiles_x + tx2) * 256];
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  return min(max(__float2int_rn(res), 0), 255);
}

// CLAHE interpolation fused with ToTensor / grayscale repeat / Normalize: uint8 in, float32 NHWC(3) out.
__global__ void __launch_bounds__(256) clahe_apply_to_nhwc_kernel(const uint8_t* __restrict__ img, int B, int H, int W, int tiles_x,
                                                                  int tiles_y, float inv_tw, float inv_th, const uint8_t* __restrict__ lut,
                                                                  float m0, float m1, float m2, float s0, float s1, float s2,
                                                                  uint8_t* __restrict__ clahe_out, float* __restrict__ out,
                                                                  float* __restrict__ amax) {
  const size_t total = (size_t)B * H * W;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const size_t r = i / W;
    const int y = (int)(r % H), b = (int)(r / H);
    const int q = clahe_interpolate(img[i], x, y, b, tiles_x, tiles_y, inv_tw, inv_th, lut);
    if (clahe_out) clahe_out[i] = (uint8_t)q;
    const float g = __fdiv_rn((float)q, 255.0f);
    const float o0 = __fdiv_rn(g - m0, s0), o1 = __fdiv_rn(g - m1, s1), o2 = __fdiv_rn(g - m2, s2);
    out[3 * i] = o0;
    out[3 * i + 1] = o1;
    out[3 * i + 2] = o2;
    local = fmaxf(local, fmaxf(fabsf(o0), fmaxf(fabsf(o1), fabsf(o2))));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(amax, local);
}

// CLAHE of RGB prints (network.py:199-204): RGB -> LAB, CLAHE on L, LAB -> RGB.  OpenCV's 8-bit colour conversions are pure
// functions of the 24-bit pixel, so both are 2^24-entry tables (built once on the host by cv2.cvtColor itself, packed
// c0 | c1 << 8 | c2 << 16, indexed c0 << 16 | c1 << 8 | c2): bit exact by construction, one gather per pixel.
__global__ void __launch_bounds__(256) rgb_to_lab_kernel(const uint8_t* __restrict__ rgb, size_t pixels, const uint32_t* __restrict__ rgb2lab,
                                                         uint8_t* __restrict__ l_plane, uint16_t* __restrict__ ab_plane) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t key = ((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2];
    const uint32_t lab = __ldg(rgb2lab + key);
    l_plane[i] = (uint8_t)(lab & 0xffu);
    ab_plane[i] = (uint16_t)(lab >> 8);
  }
}

__global__ void __launch_bounds__(256) clahe_apply_rgb_to_nhwc_kernel(const uint8_t* __restrict__ l_plane, const uint16_t* __restrict__ ab_plane,
                                                                      int B, int H, int W, int tiles_x, int tiles_y, float inv_tw,
                                                                      float inv_th, const uint8_t* __restrict__ lut,
                                                                      const uint32_t* __restrict__ lab2rgb, float m0, float m1, float m2,
                                                                      float s0, float s1, float s2, uint8_t* __restrict__ rgb_out,
                                                                      float* __restrict__ out, float* __restrict__ amax) {
  const size_t total = (size_t)B * H * W;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const size_t r = i / W;
    const int y = (int)(r % H), b = (int)(r / H);
    const int q = clahe_interpolate(l_plane[i], x, y, b, tiles_x, tiles_y, inv_tw, inv_th, lut);
    const uint32_t ab = ab_plane[i];
    const uint32_t px = __ldg(lab2rgb + (((uint32_t)q << 16) | ((ab & 0xffu) << 8) | (ab >> 8)));
    const int c0 = px & 0xff, c1 = (px >> 8) & 0xff, c2 = (px >> 16) & 0xff;
    if (rgb_out) {
      rgb_out[3 * i] = (uint8_t)c0;
      rgb_out[3 * i + 1] = (uint8_t)c1;
      rgb_out[3 * i + 2] = (uint8_t)c2;
    }
    const float o0 = __fdiv_rn(__fdiv_rn((float)c0, 255.0f) - m0, s0), o1 = __fdiv_rn(__fdiv_rn((float)c1, 255.0f) - m1, s1),
                o2 = __fdiv_rn(__fdiv_rn((float)c2, 255.0f) - m2, s2);
    out[3 * i] = o0;
    out[3 * i + 1] = o1;
    out[3 * i + 2] = o2;
    local = fmaxf(local, fmaxf(fabsf(o0), fmaxf(fabsf(o1), fabsf(o2))));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(amax, local);
}

// ------------------------------------------------------------------------------------------ K1 (first layer)
// 3x3 convolution of the 3-channel input image (the backbone's stem, K = 27): too thin for the tensor cores, so it
// runs in float32 on the CUDA cores, one thread per output pixel, the 27 inputs in registers and the [27][Cout] weights
// in shared memory.  Optionally also writes the result as the next convolution's fp16 operand planes (see sir_feat_conv).
template <int ACT>
__global__ void __launch_bounds__(256) conv_c3k3_kernel(const float* __restrict__ in, const float* __restrict__ amax_in, int B, int H,
                                                        int W, int stride, int pad, int Ho, int Wo, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int Cout, float* __restrict__ out,
                                                        float* __restrict__ amax_out, __half* __restrict__ out_hi,
                                                        __half* __restrict__ out_lo, int* __restrict__ exp_out, float bound_mult,
                                                        float bound_add) {
  extern __shared__ float sw[];  // [27][Cout] weights, [Cout] bias
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias[i];
  __syncthreads();
  float oscale = 1.0f;
  if (out_hi) {
    const float bound = fmaf(*amax_in, bound_mult, bound_add);
    int e_out = 0;
    if (bound > 0.0f && isfinite(bound)) {
      int ex;
      (void)frexpf(bound, &ex);
      e_out = max(-100, min(100, 15 - ex));
    }
    oscale = ldexpf(1.0f, e_out);
    if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e_out;
  }
  const size_t total = (size_t)B * Ho * Wo;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    size_t r = i / Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    float x[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - pad + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - pad + kx;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const float* src = in + (((size_t)b * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) x[(ky * 3 + kx) * 3 + c] = ok ? __ldg(src + c) : 0.0f;
      }
    }
    for (int co = 0; co < Cout; co += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = sw[27 * Cout + co + j];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + k * Cout + co);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + k * Cout + co + 4);
        acc[0] = fmaf(x[k], w0.x, acc[0]); acc[1] = fmaf(x[k], w0.y, acc[1]); acc[2] = fmaf(x[k], w0.z, acc[2]); acc[3] = fmaf(x[k], w0.w, acc[3]);
        acc[4] = fmaf(x[k], w1.x, acc[4]); acc[5] = fmaf(x[k], w1.y, acc[5]); acc[6] = fmaf(x[k], w1.z, acc[6]); acc[7] = fmaf(x[k], w1.w, acc[7]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ACT == 1) acc[j] = silu_fast(acc[j]);
        if (ACT == 2) acc[j] = fmaxf(acc[j], 0.0f);
        local = fmaxf(local, fabsf(acc[j]));
      }
      if (out) {
        float4* dst = reinterpret_cast<float4*>(out + i * Cout + co);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (out_hi) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sv = acc[j] * oscale;
          hi[j] = __float2half_rn(sv);
          lo[j] = __float2half_rn(sv - __half2float(hi[j]));
        }
        *reinterpret_cast<uint4*>(out_hi + i * Cout + co) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out_lo + i * Cout + co) = *reinterpret_cast<const uint4*>(lo);
      }
    }
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax_out) atomic_max_nonneg(amax_out, local);
}

// ------------------------------------------------------------------------------------------ K1a
// im2col + split: A[m][k] = in[b][oy*s - pad + ky][ox*s - pad + kx][c] * chan_scale[b][c], k = (ky*kw + kx)*C + c,
// scaled by 2^e(amax_in) and split into fp16 hi/lo.  One thread produces 8 consecutive k (16 bytes).
__global__ void __launch_bounds__(256) im2col_split_kernel(const float* __restrict__ in, const float* __restrict__ amax_in,
                                                           int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                                                           int Ho, int Wo, const float* __restrict__ chan_scale, int Kp,
                                                           __half* __restrict__ ahi, __half* __restrict__ alo) {
  const int K = kh * kw * C;
  const int e = scale_exp_from_amax(*amax_in);
  const size_t M = (size_t)B * Ho * Wo;
  const int k8s = Kp / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M * k8s; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / k8s;
    const int k0 = (int)(i - m * k8s) * 8;
    const int b = (int)(m / ((size_t)Ho * Wo));
    const int r = (int)(m - (size_t)b * Ho * Wo);
    const int oy = r / Wo, ox = r - oy * Wo;
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      float v = 0.0f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int ky = tap / kw, kx = tap - ky * kw;
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          v = in[(((size_t)b * H + iy) * W + ix) * C + c];
          if (chan_scale) v *= chan_scale[(size_t)b * C + c];
        }
      }
      const float s = ldexpf(v, e);
      hi[j] = __float2half_rn(s);
      lo[j] = __float2half_rn(s - __half2float(hi[j]));
    }
    *reinterpret_cast<uint4*>(ahi + m * Kp + k0) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(alo + m * Kp + k0) = *reinterpret_cast<const uint4*>(lo);
  }
}

// Fast path for C % 8 == 0: the 8 values of a thread are 8 consecutive channels of ONE tap, i.e. 32
// contiguous bytes of the NHWC activation -> two 16-byte loads, no per-element index arithmetic.
__global__ void __launch_bounds__(256) im2col_split_c8_kernel(const float* __restrict__ in, const float* __restrict__ amax_in,
                                                              int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                                                              int Ho, int Wo, const float* __restrict__ chan_scale, int Kp,
                                                              __half* __restrict__ ahi, __half* __restrict__ alo) {
  const int K = kh * kw * C;
  const int e = scale_exp_from_amax(*amax_in);
  const size_t M = (size_t)B * Ho * Wo;
  const int k8s = Kp / 8, c8s = C / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M * k8s; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / k8s;
    const int k8 = (int)(i - m * k8s);
    const int k0 = k8 * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (k0 < K) {
      const int tap = k8 / c8s, c0 = (k8 - tap * c8s) * 8;
      const int ky = tap / kw, kx = tap - ky * kw;
      const int b = (int)(m / ((size_t)Ho * Wo));
      const int r = (int)(m - (size_t)b * Ho * Wo);
      const int oy = r / Wo, ox = r - oy * Wo;
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float4* src = reinterpret_cast<const float4*>(in + (((size_t)b * H + iy) * W + ix) * C + c0);
        const float4 a = __ldg(src), d = __ldg(src + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
        if (chan_scale) {
          const float4* sc = reinterpret_cast<const float4*>(chan_scale + (size_t)b * C + c0);
          const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1);
          v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
        }
      }
    }
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sv = ldexpf(v[j], e);
      hi[j] = __float2half_rn(sv);
      lo[j] = __float2half_rn(sv - __half2float(hi[j]));
    }
    *reinterpret_cast<uint4*>(ahi + m * Kp + k0) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(alo + m * Kp + k0) = *reinterpret_cast<const uint4*>(lo);
  }
}

// ------------------------------------------------------------------------------------------ K2
// depthwise k x k convolution, NHWC float32, weights [k][k][C], + bias, + activation
__global__ void __launch_bounds__(256) dwconv_kernel(const float* __restrict__ in, int B, int H, int W, int C, int k, int stride,
                                                     int pad, int Ho, int Wo, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int act, float* __restrict__ out,
                                                     float* __restrict__ amax) {
  const size_t total = (size_t)B * Ho * Wo * C;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float acc = 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(in[(((size_t)b * H + iy) * W + ix) * C + c], w[(ky * k + kx) * C + c], acc);
      }
    }
    const float o = act_apply(acc + bias[c], act);
    out[i] = o;
    local = fmaxf(local, fabsf(o));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// 4 channels per thread (C % 4 == 0): 16-byte loads of activations and weights
__global__ void __launch_bounds__(256) dwconv_c4_kernel(const float* __restrict__ in, int B, int H, int W, int C, int k, int stride,
                                                        int pad, int Ho, int Wo, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int act, float* __restrict__ out,
                                                        float* __restrict__ amax) {
  const int C4 = C / 4;
  const size_t total = (size_t)B * Ho * Wo * C4;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const float4 x = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + iy) * W + ix) * C + c));
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ky * k + kx) * C + c));
        acc.x = fmaf(x.x, ww.x, acc.x); acc.y = fmaf(x.y, ww.y, acc.y); acc.z = fmaf(x.z, ww.z, acc.z); acc.w = fmaf(x.w, ww.w, acc.w);
      }
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 o;
    o.x = act_apply(acc.x + bb.x, act); o.y = act_apply(acc.y + bb.y, act); o.z = act_apply(acc.z + bb.z, act); o.w = act_apply(acc.w + bb.w, act);
    *reinterpret_cast<float4*>(out + i * 4) = o;
    local = fmaxf(local, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// Depthwise 3x3 fast path (C % 4 == 0, pad 1, stride S): one thread owns 4 channels of one output column over a
// strip of kDwRows output rows and slides a 3-row register window down the strip, so every input value is
// fetched 3 times (its column neighbours) instead of 9, and the 9 weight vectors live in registers.  The
// squeeze of the following SqueezeExcitation (sum over all pixels) is fused: each thread writes the sum of its
// strip to pool_part[b][strip*Wo + ox][c]; the SE kernel adds the parts in a fixed order (deterministic).
constexpr int kDwRows = 8;
__device__ __forceinline__ float4 f4_fma(float4 x, float4 w, float4 a) {
  return make_float4(fmaf(x.x, w.x, a.x), fmaf(x.y, w.y, a.y), fmaf(x.z, w.z, a.z), fmaf(x.w, w.w, a.w));
}
template <int S, int ACT>
__global__ void __launch_bounds__(256) dwconv3_rows_kernel(const float* __restrict__ in, int B, int H, int W, int C, int Ho, int Wo,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ out, float* __restrict__ amax,
                                                           float* __restrict__ pool_part, const float* __restrict__ amax_in,
                                                           __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                           int* __restrict__ exp_out, float bound_mult, float bound_add) {
  const int C4 = C >> 2, strips = (Ho + kDwRows - 1) / kDwRows;
  float oscale = 1.0f;
  if (out_hi) {  // operand planes for the following projection, a-priori exponent as in sir_feat_conv
    const float bound = fmaf(*amax_in, bound_mult, bound_add);
    int e_out = 0;
    if (bound > 0.0f && isfinite(bound)) {
      int ex;
      (void)frexpf(bound, &ex);
      e_out = max(-100, min(100, 15 - ex));
    }
    oscale = ldexpf(1.0f, e_out);
    if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e_out;
  }
  // One block = tile_x adjacent output columns (one warp each) x 128 channels (4 per lane) x one strip of rows: the three
  // threads that need an input value sit in the same block, so two of the three fetches hit L1 instead of going to L2
  // (with a flat thread -> output mapping the column neighbours landed on other SMs and L2 carried 3.7x the tensor).
  const int tile_x = blockDim.x >> 5, x_tiles = (Wo + tile_x - 1) / tile_x, c_groups = (C4 + 31) >> 5;
  const size_t total = (size_t)B * strips * x_tiles * c_groups;
  float local = 0.0f;
  for (size_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int cg = (int)(item % c_groups);
    size_t r = item / c_groups;
    const int xt = (int)(r % x_tiles);
    r /= x_tiles;
    const int strip = (int)(r % strips), b = (int)(r / strips);
    const int c4 = cg * 32 + (threadIdx.x & 31), ox = xt * tile_x + (threadIdx.x >> 5);
    if (c4 >= C4 || ox >= Wo) continue;
    const int c = c4 * 4;
    // The kernel is instruction bound (ncu: issue slots 56 % busy at 21 % occupancy, two thirds of it integer work), so the row loop
    // carries no index arithmetic and no per-load predicates: a column outside the image is handled by ZEROING ITS WEIGHTS once per
    // thread (its loads then go to a clamped, valid address and contribute nothing), the three column pointers advance by one row
    // pitch per input row, and a row outside the image is one predicate for its three loads.
    const int ix0 = ox * S - 1;
    const bool v0 = ix0 >= 0, v2 = ix0 + 2 < W;  // the centre column ix0 + 1 always exists
    float4 wv[9];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const bool live = (t % 3 == 0) ? v0 : (t % 3 == 2) ? v2 : true;
      wv[t] = live ? __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C + c)) : zero4;
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    const int row_pitch = W * C;
    const int oy0 = strip * kDwRows, oy1 = min(Ho, oy0 + kDwRows);
    const float* img = in + (size_t)b * H * W * C + c;
    int iy = oy0 * S - 1;  // next input row to load
    const float* p0 = img + (ptrdiff_t)iy * row_pitch + (v0 ? ix0 : ix0 + 1) * C;
    const float* p1 = img + (ptrdiff_t)iy * row_pitch + (ix0 + 1) * C;
    const float* p2 = img + (ptrdiff_t)iy * row_pitch + (v2 ? ix0 + 2 : ix0 + 1) * C;
    auto load_next = [&](float4 (&row)[3]) {
      if (iy >= 0 && iy < H) {
        row[0] = __ldg(reinterpret_cast<const float4*>(p0));
        row[1] = __ldg(reinterpret_cast<const float4*>(p1));
        row[2] = __ldg(reinterpret_cast<const float4*>(p2));
      } else {
        row[0] = row[1] = row[2] = zero4;
      }
      ++iy;
      p0 += row_pitch;
      p1 += row_pitch;
      p2 += row_pitch;
    };
    const int out_pitch = Wo * C;
    size_t off = (((size_t)b * Ho + oy0) * Wo + ox) * C + c;  // running output offset
    float4 pool = zero4;
    auto emit = [&](const float4 (&a)[3], const float4 (&m)[3], const float4 (&z)[3]) {
      float4 acc = bb;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        acc = f4_fma(a[kx], wv[kx], acc);
        acc = f4_fma(m[kx], wv[3 + kx], acc);
        acc = f4_fma(z[kx], wv[6 + kx], acc);
      }
      float4 o;
      if (ACT == 1) {
        o = make_float4(silu_fast(acc.x), silu_fast(acc.y), silu_fast(acc.z), silu_fast(acc.w));
      } else if (ACT == 2) {
        o = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
      } else {
        o = acc;
      }
      if (out) *reinterpret_cast<float4*>(out + off) = o;
      if (out_hi) {
        const float s0 = o.x * oscale, s1 = o.y * oscale, s2 = o.z * oscale, s3 = o.w * oscale;
        const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
        const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(s0 - b01.x, s1 - b01.y), l23 = __floats2half2_rn(s2 - b23.x, s3 - b23.y);
        *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      off += out_pitch;
      pool.x += o.x; pool.y += o.y; pool.z += o.z; pool.w += o.w;
      local = fmaxf(local, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
    };
    if (S == 1) {
      // two output rows per iteration: the six loads of the two new input rows are in flight together
      float4 r0[3], r1[3], r2[3], r3[3];
      load_next(r0);
      load_next(r1);
      int oy = oy0;
      for (; oy + 1 < oy1; oy += 2) {
        load_next(r2);
        load_next(r3);
        emit(r0, r1, r2);
        emit(r1, r2, r3);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) { r0[kx] = r2[kx]; r1[kx] = r3[kx]; }
      }
      if (oy < oy1) {
        load_next(r2);
        emit(r0, r1, r2);
      }
    } else {
      float4 r0[3], r1[3], r2[3];
      load_next(r0);
      for (int oy = oy0; oy < oy1; ++oy) {
        load_next(r1);
        load_next(r2);
        emit(r0, r1, r2);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) r0[kx] = r2[kx];
      }
    }
    if (pool_part) *reinterpret_cast<float4*>(pool_part + ((size_t)b * strips * Wo + (size_t)strip * Wo + ox) * C + c) = pool;
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// sum over all pixels: [B][HW][C] -> part[B][1][C] (the one-part squeeze for producers without a fused pool)
__global__ void __launch_bounds__(256) pool_sum_kernel(const float* __restrict__ in, int HW, int C, float* __restrict__ out) {
  __shared__ float part[4][64];
  const int b = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), sub = threadIdx.x >> 6;
  float acc = 0.0f;
  if (c < C)
    for (int p = sub; p < HW; p += 4) acc += in[((size_t)b * HW + p) * C + c];
  part[sub][threadIdx.x & 63] = acc;
  __syncthreads();
  if (sub == 0 && c < C) out[(size_t)b * C + c] = (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
}

// ------------------------------------------------------------------------------------------ K3
// squeeze-excitation, one CTA per image: avg = sum of the pooled parts / HW (fixed order), then
// scale = sigmoid(W2 . silu(W1 . avg + b1) + b2).  w1 [S][C], w2t [S][C] (fc2 transposed: coalesced over c).
// avg[b][c] = (sum over parts, fixed order) / HW; one CTA per (image, 64-channel slab), 4 threads per channel
__global__ void __launch_bounds__(256) pool_reduce_kernel(const float* __restrict__ pool_part, int parts, float inv_hw, int C,
                                                          float* __restrict__ avg) {
  __shared__ float part[4][64];
  const int b = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), sub = threadIdx.x >> 6;
  const float* src = pool_part + (size_t)b * parts * C;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int q = sub;
    for (; q + 28 < parts; q += 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += src[(size_t)(q + 4 * u) * C + c];
    }
    for (; q < parts; q += 4) a[0] += src[(size_t)q * C + c];
  }
  part[sub][threadIdx.x & 63] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (sub == 0 && c < C)
    avg[(size_t)b * C + c] = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x])) * inv_hw;
}
__global__ void __launch_bounds__(1024) se_fc_kernel(const float* __restrict__ avg, int C, int S,
                                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2t, const float* __restrict__ b2,
                                                     float* __restrict__ scale) {
  extern __shared__ float sh[];  // [C] avg, [S] hidden
  float* savg = sh;
  float* hid = sh + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) savg[c] = avg[(size_t)b * C + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < S; j += nw) {
    float acc = 0.0f;
    for (int c = lane; c < C; c += 32) acc = fmaf(w1[(size_t)j * C + c], savg[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) hid[j] = act_apply(acc + b1[j], 1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b2[c];
    for (int j = 0; j < S; ++j) acc = fmaf(w2t[(size_t)j * C + c], hid[j], acc);
    scale[(size_t)b * C + c] = 1.0f / (1.0f + expf(-acc));
  }
}

// max pooling, NHWC
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ in, int B, int H, int W, int C, int k, int stride,
                                                      int pad, int Ho, int Wo, float* __restrict__ out, float* __restrict__ amax) {
  const size_t total = (size_t)B * Ho * Wo * C;
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float best = -INFINITY;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        best = fmaxf(best, in[(((size_t)b * H + iy) * W + ix) * C + c]);
      }
    }
    out[i] = best;
    local = fmaxf(local, fabsf(best));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// standalone per-channel affine + activation (BatchNorm2d / ReLU children that a block cut separates
// from their convolution, e.g. VGG features[:n] ending between Conv2d and ReLU)
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ in, size_t total, int C, int ld_in, int ld_out,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                         float* __restrict__ out, float* __restrict__ amax) {
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / C;
    const int c = (int)(i - row * C);
    float v = in[row * ld_in + c];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    v = act_apply(v, act);
    out[row * ld_out + c] = v;
    local = fmaxf(local, fabsf(v));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// average pooling (DenseNet transitions: AvgPool2d(2, 2)), NHWC, no padding
__global__ void __launch_bounds__(256) avgpool2d_kernel(const float* __restrict__ in, int B, int H, int W, int C, int k, int stride,
                                                        int Ho, int Wo, float* __restrict__ out, float* __restrict__ amax) {
  const size_t total = (size_t)B * Ho * Wo * C;
  const float inv = 1.0f / (float)(k * k);
  float local = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float acc = 0.0f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) acc += in[(((size_t)b * H + oy * stride + ky) * W + ox * stride + kx) * C + c];
    const float o = acc * inv;
    out[i] = o;
    local = fmaxf(local, fabsf(o));
  }
  local = warp_max(local);
  if ((threadIdx.x & 31) == 0 && amax) atomic_max_nonneg(amax, local);
}

// NHWC -> NCHW (the reference returns [C,h,w] maps, network.py:238-244)
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int HW, int C, float* __restrict__ out, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    size_t r = i / HW;
    const int c = (int)(r % C);
    const size_t b = r / C;
    out[i] = in[(b * HW + p) * C + c];
  }
}

// ------------------------------------------------------------------------------------------ host
namespace {
unsigned grid_for(size_t work, int per_block = 256) { return (unsigned)std::min<size_t>((work + per_block - 1) / per_block, 148 * 32); }
}  // namespace
}  // namespace sir

using namespace sir;

extern "C" int sir_feat_image_to_nhwc(const uint8_t* d_img, int B, int H, int W, int in_ch, const float* h_mean, const float* h_std,
                                      float* d_out, float* d_amax, void* stream) {
  SIR_CHECK_ARG(d_img && d_out && d_amax && h_mean && h_std, "sir_feat_image_to_nhwc: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && (in_ch == 1 || in_ch == 3), "sir_feat_image_to_nhwc: bad shape");
  const size_t pixels = (size_t)B * H * W;
  image_to_nhwc_kernel<<<grid_for(pixels), 256, 0, (cudaStream_t)stream>>>(d_img, in_ch, pixels, h_mean[0], h_mean[1], h_mean[2],
                                                                            h_std[0], h_std[1], h_std[2], d_out, d_amax);
  SIR_LAUNCH_CHECK("image_to_nhwc_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_clahe_to_nhwc(const uint8_t* d_img, int B, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                                      const float* h_mean, const float* h_std, uint8_t* d_lut, uint8_t* d_clahe_u8, float* d_out,
                                      float* d_amax, void* stream) {
  SIR_CHECK_ARG(d_img && d_lut && d_out && d_amax && h_mean && h_std, "sir_feat_clahe_to_nhwc: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 1 && W > 1 && tiles_x > 0 && tiles_y > 0 && tiles_x * tiles_y <= 65535, "sir_feat_clahe_to_nhwc: bad shape");
  const int eh = H % tiles_y == 0 && W % tiles_x == 0 ? H : H + (tiles_y - H % tiles_y);
  const int ew = H % tiles_y == 0 && W % tiles_x == 0 ? W : W + (tiles_x - W % tiles_x);
  const int th = eh / tiles_y, tw = ew / tiles_x;
  SIR_CHECK_ARG(eh - H < H && ew - W < W, "sir_feat_clahe_to_nhwc: image %dx%d too small for a %dx%d tile grid", H, W, tiles_y, tiles_x);
  const int area = th * tw;
  const float lut_scale = 255.0f / (float)area;
  int clip = 0;
  if (clip_limit > 0.0) clip = std::max((int)(clip_limit * area / 256), 1);
  cudaStream_t st = (cudaStream_t)stream;
  clahe_lut_kernel<<<dim3((unsigned)(tiles_x * tiles_y), (unsigned)B), 256, 0, st>>>(d_img, H, W, tiles_x, th, tw, clip, lut_scale, d_lut);
  SIR_LAUNCH_CHECK("clahe_lut_kernel");
  const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  clahe_apply_to_nhwc_kernel<<<grid_for((size_t)B * H * W), 256, 0, st>>>(d_img, B, H, W, tiles_x, tiles_y, inv_tw, inv_th, d_lut, h_mean[0],
                                                                          h_mean[1], h_mean[2], h_std[0], h_std[1], h_std[2], d_clahe_u8,
                                                                          d_out, d_amax);
  SIR_LAUNCH_CHECK("clahe_apply_to_nhwc_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_clahe_rgb_to_nhwc(const uint8_t* d_rgb, int B, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                                          const float* h_mean, const float* h_std, const uint32_t* d_rgb2lab, const uint32_t* d_lab2rgb,
                                          uint8_t* d_l_plane, uint16_t* d_ab_plane, uint8_t* d_lut, uint8_t* d_rgb_out, float* d_out,
                                          float* d_amax, void* stream) {
  SIR_CHECK_ARG(d_rgb && d_rgb2lab && d_lab2rgb && d_l_plane && d_ab_plane && d_lut && d_out && d_amax && h_mean && h_std,
                "sir_feat_clahe_rgb_to_nhwc: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 1 && W > 1 && tiles_x > 0 && tiles_y > 0 && tiles_x * tiles_y <= 65535, "sir_feat_clahe_rgb_to_nhwc: bad shape");
  const int eh = H % tiles_y == 0 && W % tiles_x == 0 ? H : H + (tiles_y - H % tiles_y);
  const int ew = H % tiles_y == 0 && W % tiles_x == 0 ? W : W + (tiles_x - W % tiles_x);
  const int th = eh / tiles_y, tw = ew / tiles_x;
  SIR_CHECK_ARG(eh - H < H && ew - W < W, "sir_feat_clahe_rgb_to_nhwc: image %dx%d too small for a %dx%d tile grid", H, W, tiles_y, tiles_x);
  const int area = th * tw;
  const float lut_scale = 255.0f / (float)area;
  int clip = 0;
  if (clip_limit > 0.0) clip = std::max((int)(clip_limit * area / 256), 1);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pixels = (size_t)B * H * W;
  rgb_to_lab_kernel<<<grid_for(pixels), 256, 0, st>>>(d_rgb, pixels, d_rgb2lab, d_l_plane, d_ab_plane);
  SIR_LAUNCH_CHECK("rgb_to_lab_kernel");
  clahe_lut_kernel<<<dim3((unsigned)(tiles_x * tiles_y), (unsigned)B), 256, 0, st>>>(d_l_plane, H, W, tiles_x, th, tw, clip, lut_scale, d_lut);
  SIR_LAUNCH_CHECK("clahe_lut_kernel");
  const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  clahe_apply_rgb_to_nhwc_kernel<<<grid_for(pixels), 256, 0, st>>>(d_l_plane, d_ab_plane, B, H, W, tiles_x, tiles_y, inv_tw, inv_th, d_lut,
                                                                   d_lab2rgb, h_mean[0], h_mean[1], h_mean[2], h_std[0], h_std[1], h_std[2],
                                                                   d_rgb_out, d_out, d_amax);
  SIR_LAUNCH_CHECK("clahe_apply_rgb_to_nhwc_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_im2col_split(const float* d_in, const float* d_amax_in, int B, int H, int W, int C, int kh, int kw, int stride,
                                     int pad, const float* d_chan_scale, int Kp, uint16_t* d_ahi, uint16_t* d_alo, void* stream) {
  SIR_CHECK_ARG(d_in && d_amax_in && d_ahi && d_alo, "sir_feat_im2col_split: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "sir_feat_im2col_split: bad shape");
  SIR_CHECK_ARG(Kp % 8 == 0 && Kp >= kh * kw * C, "sir_feat_im2col_split: Kp=%d must be a multiple of 8 and >= %d", Kp, kh * kw * C);
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  SIR_CHECK_ARG(Ho > 0 && Wo > 0, "sir_feat_im2col_split: empty output");
  const size_t work = (size_t)B * Ho * Wo * (Kp / 8);
  if (C % 8 == 0 && ((uintptr_t)d_in & 15) == 0 && (!d_chan_scale || ((uintptr_t)d_chan_scale & 15) == 0))
    im2col_split_c8_kernel<<<grid_for(work), 256, 0, (cudaStream_t)stream>>>(d_in, d_amax_in, B, H, W, C, kh, kw, stride, pad, Ho, Wo,
                                                                              d_chan_scale, Kp, (__half*)d_ahi, (__half*)d_alo);
  else
    im2col_split_kernel<<<grid_for(work), 256, 0, (cudaStream_t)stream>>>(d_in, d_amax_in, B, H, W, C, kh, kw, stride, pad, Ho, Wo,
                                                                           d_chan_scale, Kp, (__half*)d_ahi, (__half*)d_alo);
  SIR_LAUNCH_CHECK("im2col_split_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_conv_c3k3(const float* d_in, const float* d_amax_in, int B, int H, int W, int stride, int pad, const float* d_w,
                                  const float* d_bias, int Cout, int act, float* d_out, float* d_amax_out, uint16_t* d_out_hi,
                                  uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult, float bound_add, void* stream) {
  SIR_CHECK_ARG(d_in && d_amax_in && d_w && d_bias && (d_out || d_out_hi), "sir_feat_conv_c3k3: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && stride > 0 && pad >= 0 && Cout > 0 && Cout % 8 == 0 && Cout <= 256, "sir_feat_conv_c3k3: bad shape");
  SIR_CHECK_ARG(act >= 0 && act <= 2, "sir_feat_conv_c3k3: unknown activation %d", act);
  SIR_CHECK_ARG(!d_out_hi || (d_out_lo && d_exp_out), "sir_feat_conv_c3k3: operand planes need d_out_lo and d_exp_out");
  SIR_CHECK_ARG((((uintptr_t)d_out | (uintptr_t)d_out_hi | (uintptr_t)d_out_lo) & 15) == 0, "sir_feat_conv_c3k3: outputs must be 16-byte aligned");
  const int Ho = (H + 2 * pad - 3) / stride + 1, Wo = (W + 2 * pad - 3) / stride + 1;
  SIR_CHECK_ARG(Ho > 0 && Wo > 0, "sir_feat_conv_c3k3: empty output");
  const size_t smem = (size_t)28 * Cout * 4;
  const unsigned grid = grid_for((size_t)B * Ho * Wo);
  cudaStream_t st = (cudaStream_t)stream;
#define SIR_C3(A_) conv_c3k3_kernel<A_><<<grid, 256, smem, st>>>(d_in, d_amax_in, B, H, W, stride, pad, Ho, Wo, d_w, d_bias, Cout, d_out, d_amax_out, \
                                                                   (__half*)d_out_hi, (__half*)d_out_lo, d_exp_out, bound_mult, bound_add)
  if (act == 0) SIR_C3(0); else if (act == 1) SIR_C3(1); else SIR_C3(2);
#undef SIR_C3
  SIR_LAUNCH_CHECK("conv_c3k3_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_dwconv_pool_parts(int k, int stride, int C, int Ho, int Wo) {
  return (k == 3 && (stride == 1 || stride == 2) && C % 4 == 0) ? ceil_div(Ho, kDwRows) * Wo : 1;
}

extern "C" int sir_feat_dwconv(const float* d_in, int B, int H, int W, int C, int k, int stride, int pad, const float* d_w,
                               const float* d_bias, int act, float* d_out, float* d_amax_out, float* d_pool_part, const float* d_amax_in,
                               uint16_t* d_out_hi, uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult, float bound_add,
                               void* stream) {
  SIR_CHECK_ARG(d_in && d_w && d_bias && (d_out || d_out_hi), "sir_feat_dwconv: null pointer");
  SIR_CHECK_ARG(!d_out_hi || (d_out_lo && d_exp_out && d_amax_in && k == 3 && pad == 1 && (stride == 1 || stride == 2) && C % 8 == 0 &&
                              (((uintptr_t)d_out_hi | (uintptr_t)d_out_lo) & 15) == 0 && bound_mult >= 0.0f && bound_add >= 0.0f),
                "sir_feat_dwconv: operand planes need the 3x3 fast path (pad 1, stride 1|2, C %% 8 == 0), d_out_lo, d_exp_out, d_amax_in");
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0, "sir_feat_dwconv: bad shape");
  SIR_CHECK_ARG(act >= 0 && act <= 2, "sir_feat_dwconv: unknown activation %d", act);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  SIR_CHECK_ARG(Ho > 0 && Wo > 0, "sir_feat_dwconv: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = C % 4 == 0 && (((uintptr_t)d_in | (uintptr_t)d_w | (uintptr_t)d_bias | (uintptr_t)d_out | (uintptr_t)d_pool_part) & 15) == 0;
  if (aligned && k == 3 && pad == 1 && (stride == 1 || stride == 2)) {
    // columns per block: 4..8 warps, whichever leaves the fewest idle column slots (ties -> the wider tile)
    int tile_x = 8;
    for (int t = 8, best = 1 << 30; t >= 4; --t) {
      const int slots = ceil_div(Wo, t) * t;
      if (slots < best) {
        best = slots;
        tile_x = t;
      }
    }
    const size_t items = (size_t)B * ceil_div(Ho, kDwRows) * ceil_div(Wo, tile_x) * ceil_div(C / 4, 32);
    const unsigned grid = (unsigned)std::min<size_t>(items, 148 * 8);  // persistent blocks: few atomics on the running |max|
#define SIR_DW3(S_, A_)                                                                                                          \
  dwconv3_rows_kernel<S_, A_><<<grid, tile_x * 32, 0, st>>>(d_in, B, H, W, C, Ho, Wo, d_w, d_bias, d_out, d_amax_out, d_pool_part, d_amax_in, \
                                                    (__half*)d_out_hi, (__half*)d_out_lo, d_exp_out, bound_mult, bound_add)
    if (stride == 1) {
      if (act == 0) SIR_DW3(1, 0); else if (act == 1) SIR_DW3(1, 1); else SIR_DW3(1, 2);
    } else {
      if (act == 0) SIR_DW3(2, 0); else if (act == 1) SIR_DW3(2, 1); else SIR_DW3(2, 2);
    }
#undef SIR_DW3
    SIR_LAUNCH_CHECK("dwconv3_rows_kernel");
    return SIR_OK;
  }
  SIR_CHECK_ARG(d_out && !d_out_hi, "sir_feat_dwconv: this shape takes the generic path, which writes float32 only");
  if (aligned)
    dwconv_c4_kernel<<<grid_for((size_t)B * Ho * Wo * (C / 4)), 256, 0, st>>>(d_in, B, H, W, C, k, stride, pad, Ho, Wo, d_w, d_bias, act, d_out,
                                                                             d_amax_out);
  else
    dwconv_kernel<<<grid_for((size_t)B * Ho * Wo * C), 256, 0, st>>>(d_in, B, H, W, C, k, stride, pad, Ho, Wo, d_w, d_bias, act, d_out, d_amax_out);
  SIR_LAUNCH_CHECK("dwconv_kernel");
  if (d_pool_part) {  // one part per image
    pool_sum_kernel<<<dim3((unsigned)ceil_div(C, 64), (unsigned)B), 256, 0, st>>>(d_out, Ho * Wo, C, d_pool_part);
    SIR_LAUNCH_CHECK("pool_sum_kernel");
  }
  return SIR_OK;
}

extern "C" int sir_feat_pool_sum(const float* d_in, int B, int HW, int C, float* d_pool_part, void* stream) {
  SIR_CHECK_ARG(d_in && d_pool_part && B > 0 && HW > 0 && C > 0, "sir_feat_pool_sum: bad argument");
  pool_sum_kernel<<<dim3((unsigned)ceil_div(C, 64), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(d_in, HW, C, d_pool_part);
  SIR_LAUNCH_CHECK("pool_sum_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_se_scale(const float* d_pool_part, int B, int parts, int HW, int C, int S, const float* d_w1, const float* d_b1,
                                 const float* d_w2t, const float* d_b2, float* d_avg, float* d_scale, void* stream) {
  SIR_CHECK_ARG(d_pool_part && d_w1 && d_b1 && d_w2t && d_b2 && d_avg && d_scale, "sir_feat_se_scale: null pointer");
  const size_t se_smem = (size_t)(C + S) * 4;
  SIR_CHECK_ARG(B > 0 && parts > 0 && HW > 0 && C > 0 && S > 0 && se_smem <= 48 * 1024, "sir_feat_se_scale: bad shape");
  pool_reduce_kernel<<<dim3((unsigned)ceil_div(C, 64), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(d_pool_part, parts, 1.0f / (float)HW, C, d_avg);
  SIR_LAUNCH_CHECK("pool_reduce_kernel");
  se_fc_kernel<<<B, 1024, se_smem, (cudaStream_t)stream>>>(d_avg, C, S, d_w1, d_b1, d_w2t, d_b2, d_scale);
  SIR_LAUNCH_CHECK("se_fc_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_maxpool(const float* d_in, int B, int H, int W, int C, int k, int stride, int pad, float* d_out,
                                float* d_amax_out, void* stream) {
  SIR_CHECK_ARG(d_in && d_out, "sir_feat_maxpool: null pointer");
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0, "sir_feat_maxpool: bad shape");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  SIR_CHECK_ARG(Ho > 0 && Wo > 0, "sir_feat_maxpool: empty output");
  maxpool_kernel<<<grid_for((size_t)B * Ho * Wo * C), 256, 0, (cudaStream_t)stream>>>(d_in, B, H, W, C, k, stride, pad, Ho, Wo, d_out,
                                                                                       d_amax_out);
  SIR_LAUNCH_CHECK("maxpool_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_affine_act(const float* d_in, long long rows, int C, int ld_in, int ld_out, const float* d_scale,
                                   const float* d_shift, int act, float* d_out, float* d_amax_out, void* stream) {
  SIR_CHECK_ARG(d_in && d_out && rows > 0 && C > 0 && ld_in >= C && ld_out >= C && (!d_scale == !d_shift), "sir_feat_affine_act: bad argument");
  const size_t total = (size_t)rows * C;
  affine_act_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_in, total, C, ld_in, ld_out, d_scale, d_shift, act, d_out, d_amax_out);
  SIR_LAUNCH_CHECK("affine_act_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_avgpool2d(const float* d_in, int B, int H, int W, int C, int k, int stride, float* d_out, float* d_amax_out,
                                  void* stream) {
  SIR_CHECK_ARG(d_in && d_out && B > 0 && H >= k && W >= k && C > 0 && k > 0 && stride > 0, "sir_feat_avgpool2d: bad argument");
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  avgpool2d_kernel<<<grid_for((size_t)B * Ho * Wo * C), 256, 0, (cudaStream_t)stream>>>(d_in, B, H, W, C, k, stride, Ho, Wo, d_out, d_amax_out);
  SIR_LAUNCH_CHECK("avgpool2d_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_nhwc_to_nchw(const float* d_in, int B, int HW, int C, float* d_out, void* stream) {
  SIR_CHECK_ARG(d_in && d_out && B > 0 && HW > 0 && C > 0, "sir_feat_nhwc_to_nchw: bad argument");
  const size_t total = (size_t)B * HW * C;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_in, HW, C, d_out, total);
  SIR_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return SIR_OK;
}
