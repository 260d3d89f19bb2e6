// Feature stage K1: Conv2d (groups 1) as an implicit GEMM on the tcgen05 tensor cores.
//
//   out[b][oy][ox][n] = act( sum_{ky,kx,c} x[b][s*oy+ky-pad][s*ox+kx-pad][c] * w[n][ky][kx][c] + bias[n] ) (+ residual)
//
// (torchvision Conv2d + folded BatchNorm + SiLU/ReLU as composed by network.py:121-186.)  The
// activation arrives as two fp16 NHWC planes (hi + lo = the float32 value times a power of two,
// written by the split pass), the weights as two fp16 [N][taps*Cp] matrices.  No im2col matrix
// exists anywhere: the A operand of output patch (TH x TW pixels = 128 rows) and tap (ky,kx) is the
// same patch of the input shifted by (ky-pad, kx-pad), which one 4-D TMA box fetches straight into
// the swizzled K-major layout tcgen05.mma reads; rows outside the image are zero filled by the TMA
// unit, which is the convolution's zero padding.  A 1x1 convolution (and a plain GEMM over an
// explicit [M][K] matrix) is the same kernel with a 128 x 1 "patch" over a 1 x M image.
//
// Persistent, warp specialised, one CTA per SM:
//   warp 0      TMA producer: per K step (tap, 16/32-channel chunk) A_hi, A_lo, B_hi, B_lo into a ring
//   warp 1      TMEM owner + MMA issuer: hi*hi + lo*hi + hi*lo (fp32-grade), accumulators double
//               buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of i+1
//   warps 2-13  epilogue: tcgen05.ld 16 columns, transpose through shared memory so that global
//               stores / residual loads are row contiguous, un-scale + bias + activation + residual,
//               running max |out| for the next layer's scaling
// Bound: tensor pipe for wide layers (three MMAs per algorithmic MAC), L2->smem operand feed otherwise.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "sir_common.cuh"
#include "sir_ptx.cuh"

namespace sir {

constexpr int kConvEpiParts = 3;                              // epilogue warps per TMEM lane quarter
constexpr int kConvEpiWarps = 4 * kConvEpiParts;
constexpr int kConvThreads = 64 + 32 * kConvEpiWarps;         // warp 0 TMA, warp 1 MMA, then the epilogue warps
constexpr int kConvBM = 128;
constexpr int kConvMaxStages = 8;
constexpr uint32_t kConvAccStride = 256;  // TMEM columns between the two accumulator buffers

struct ConvParams {
  int Ho, Wo;             // output grid of one image (1 x M for flat GEMMs)
  int N, BN, n_tiles_n;
  int taps, kw, pad;
  int stride;             // spatial stride (1 or 2 ...): the patch load walks the input with TMA element strides
  int chunks, bk;         // K step = bk channels of one tap; chunks = Cp / bk
  int tw_log2, TH;        // patch = TH x (1 << tw_log2) pixels = 128 rows
  int tiles_x, tiles_y;
  int total_tiles;
  int stages;
  int w_exp;
  const uint8_t* wpack;   // weights as per-stage shared-memory images: [n_tile][k_step][hi | lo][BN][granule], swizzled
  long long wpack_image_bytes;  // 0: one weight set; else one set per image (squeeze-excitation scale folded into the weights)
  const float* amax_in;
  const int* exp_in;      // exponent the input planes were written with; NULL: derived from amax_in (split pass)
  const float* bias;
  const float* residual;
  float* out;             // float32 output (may be NULL when only the operand planes are wanted)
  float* amax_out;
  int ldc;
  __half* out_hi;         // optional: the output again as fp16 hi/lo operand planes [pixels][N] for the next convolution
  __half* out_lo;
  int* exp_out;           // exponent used for those planes (written by CTA 0)
  float bound_mult, bound_add;  // |out| <= amax_in * bound_mult + bound_add (+ amax_res): a-priori bound -> plane exponent
  const float* amax_res;
};

// Input-plane exponent and the un-scale factor of the accumulator.
__device__ __forceinline__ float conv_unscale(const ConvParams& p) {
  const int e_in = p.exp_in ? *p.exp_in : scale_exp_from_amax(*p.amax_in);
  return ldexpf(1.0f, -(e_in + p.w_exp));
}
// Exponent e with bound * 2^e < 2^15 (fp16 planes cannot overflow); the bound only uses values final before the launch.
__device__ __forceinline__ int conv_plane_exp(const ConvParams& p) {
  const float bound = fmaf(*p.amax_in, p.bound_mult, p.bound_add) + (p.amax_res ? *p.amax_res : 0.0f);
  if (!(bound > 0.0f) || !isfinite(bound)) return 0;
  int ex;
  (void)frexpf(bound, &ex);
  return max(-100, min(100, 15 - ex));
}

template <int ACT>
__device__ __forceinline__ float conv_act(float v) {
  if (ACT == 1) return silu_fast(v);
  if (ACT == 2) return fmaxf(v, 0.0f);                    // ReLU
  return v;
}

// Epilogue of one 128 x BN accumulator tile: this warp owns 32 TMEM lanes (rows) and every kConvEpiParts-th 16-column chunk
// (`half` = which of them).
// pix[ps]: output pixel index of row (ps*8 + lane/4) of the warp's 32 rows, or -1 if outside the image.
// One chunk: 32 rows x 16 columns arrive row-per-lane from TMEM, are transposed through shared memory (XOR swizzle,
// conflict free both ways) so that each lane then owns 4 consecutive columns of 4 rows: stores and residual loads are
// row contiguous (64 B per row, full sectors).
template <int ACT>
__device__ __forceinline__ void conv_epilogue_chunk(const ConvParams& p, const uint32_t (&v)[16], float* xp, int lane, int n,
                                                    const long long (&pix)[4], float unscale, float oscale, float& local_max) {
  const int rs = lane >> 2, c4 = lane & 3;
  const bool live = n < p.N;  // N % 4 == 0: a lane's 4 columns are all inside or all outside
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  // residual rows: all four loads are issued before the transposition so that their (DRAM) latencies overlap each other and
  // the shared-memory round trip; loading each right before its use left the epilogue stalled on them half of its time
  float4 rr[4];
  if (p.residual) {
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      rr[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && pix[ps] >= 0) rr[ps] = *reinterpret_cast<const float4*>(p.residual + pix[ps] * p.ldc + n);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(xp + lane * 16 + 4 * (j ^ ((lane >> 1) & 3))) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  if (!live) return;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    if (pix[ps] < 0) continue;
    const long long row_off = pix[ps] * p.ldc + n;
    const int row = ps * 8 + rs;
    const float4 a = *reinterpret_cast<const float4*>(xp + row * 16 + 4 * (c4 ^ ((row >> 1) & 3)));
    float4 o = make_float4(conv_act<ACT>(fmaf(a.x, unscale, bb.x)), conv_act<ACT>(fmaf(a.y, unscale, bb.y)),
                           conv_act<ACT>(fmaf(a.z, unscale, bb.z)), conv_act<ACT>(fmaf(a.w, unscale, bb.w)));
    if (p.residual) {
      o.x += rr[ps].x; o.y += rr[ps].y; o.z += rr[ps].z; o.w += rr[ps].w;
    }
    if (p.out) *reinterpret_cast<float4*>(p.out + row_off) = o;
    if (p.out_hi) {  // the same values as the next convolution's operand planes
      const float s0 = o.x * oscale, s1 = o.y * oscale, s2 = o.z * oscale, s3 = o.w * oscale;
      const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
      const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(s0 - b01.x, s1 - b01.y), l23 = __floats2half2_rn(s2 - b23.x, s3 - b23.y);
      const long long so = pix[ps] * p.N + n;
      *reinterpret_cast<uint2*>(p.out_hi + so) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(p.out_lo + so) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
    local_max = fmaxf(local_max, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
  }
}

template <int ACT>
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, uint32_t tacc, float* xp, int lane, int half, int nt,
                                                   const long long (&pix)[4], float unscale, float oscale, float& local_max) {
  const int n_chunks = p.BN >> 4, n_base = nt * p.BN + (lane & 3) * 4;
  // two register sets: the TMEM load of the next chunk is in flight while the current one is processed
  uint32_t va[16], vb[16];
  if (half < n_chunks) ptx::tmem_ld_32x16(tacc + half * 16, va);
  for (int ci = half; ci < n_chunks; ci += 2 * kConvEpiParts) {
    ptx::tmem_ld_wait();
    if (ci + kConvEpiParts < n_chunks) ptx::tmem_ld_32x16(tacc + (ci + kConvEpiParts) * 16, vb);
    conv_epilogue_chunk<ACT>(p, va, xp, lane, n_base + ci * 16, pix, unscale, oscale, local_max);
    if (ci + kConvEpiParts < n_chunks) {
      ptx::tmem_ld_wait();
      if (ci + 2 * kConvEpiParts < n_chunks) ptx::tmem_ld_32x16(tacc + (ci + 2 * kConvEpiParts) * 16, va);
      conv_epilogue_chunk<ACT>(p, vb, xp, lane, n_base + (ci + kConvEpiParts) * 16, pix, unscale, oscale, local_max);
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_bytes = (uint32_t)kConvBM * p.bk * 2, b_bytes = (uint32_t)p.BN * p.bk * 2;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const uint32_t ring_bytes = stage_bytes * p.stages;
  const uint32_t bar0 = base + ring_bytes;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (kConvMaxStages + i); };
  auto bar_acc_full = [&](int i) { return bar0 + 8u * (2 * kConvMaxStages + i); };
  auto bar_acc_empty = [&](int i) { return bar0 + 8u * (2 * kConvMaxStages + 2 + i); };
  const uint32_t slot_off = ring_bytes + 8u * (2 * kConvMaxStages + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + slot_off);
  float* xpose = reinterpret_cast<float*>(base_ptr + slot_off + 16);  // 8 warps x 32 rows x 16 floats

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(bar_full(i), 1);
      ptx::mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_acc_full(i), 1);
      ptx::mbar_init(bar_acc_empty(i), kConvEpiWarps);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tm_xhi);
    ptx::prefetch_tmap(&tm_xlo);
  }
  if (warp == 1) ptx::tmem_alloc<512>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int TW = 1 << p.tw_log2;
  const int k_iters = p.taps * p.chunks;

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ring position as running counters: a runtime modulo / division per K step costs more than the step's MMAs
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
        const int px = mt % p.tiles_x, r1 = mt / p.tiles_x;
        const int py = r1 % p.tiles_y, b = r1 / p.tiles_y;
        const int x0 = px * TW * p.stride - p.pad, y0 = py * p.TH * p.stride - p.pad;
        // The weight stage is ONE contiguous bulk copy of a pre-swizzled image instead of a tensor-map load of hundreds of
        // 32 / 64-byte rows (measured: 1 - 5 % faster per layer, and one instruction per stage in the producer).
        const uint8_t* wsrc = p.wpack + (size_t)b * p.wpack_image_bytes + (size_t)nt * k_iters * (2 * b_bytes);
        int tap_y = 0, tap_x = 0, chunk = 0;
        for (int ks = 0; ks < k_iters; ++ks) {
          ptx::mbar_wait(bar_empty(slot), phase ^ 1);
          ptx::mbar_arrive_expect_tx(bar_full(slot), stage_bytes);
          const uint32_t dst = base + slot * stage_bytes;
          const int c0 = chunk * p.bk;
          ptx::tma_load_4d(dst, &tm_xhi, bar_full(slot), c0, x0 + tap_x, y0 + tap_y, b);
          ptx::tma_load_4d(dst + a_bytes, &tm_xlo, bar_full(slot), c0, x0 + tap_x, y0 + tap_y, b);
          ptx::bulk_load(dst + 2 * a_bytes, wsrc + (size_t)ks * (2 * b_bytes), 2 * b_bytes, bar_full(slot));
          if (++chunk == p.chunks) {
            chunk = 0;
            if (++tap_x == p.kw) {
              tap_x = 0;
              ++tap_y;
            }
          }
          if (++slot == (uint32_t)p.stages) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(kConvBM, p.BN);
      const uint32_t sbo = (uint32_t)p.bk * 16, layout = p.bk == 32 ? 4u : 6u;  // 8 rows of 64 B / 32 B
      const int k16s = p.bk / 16;
      // descriptors of stage 0; a stage / operand / K16 step further is an addition to the 16-byte address field
      const uint64_t desc0 = ptx::make_smem_desc(base, 16, sbo, layout);
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4, b16 = b_bytes >> 4;
      uint32_t slot = 0, phase = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        ptx::mbar_wait(bar_acc_empty(buf), ((lt >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * kConvAccStride;
        uint32_t accumulate = 0;
        for (int ks = 0; ks < k_iters; ++ks) {
          ptx::mbar_wait(bar_full(slot), phase);
          ptx::tc_fence_after();
          uint64_t da_hi = desc0 + slot * stage16;
          for (int kk = 0; kk < k16s; ++kk, da_hi += 2) {  // +32 B inside the swizzled row
            const uint64_t da_lo = da_hi + a16, db_hi = da_hi + 2 * a16, db_lo = db_hi + b16;
            ptx::mma_f16_ss(acc, da_hi, db_hi, idesc, accumulate);
            ptx::mma_f16_ss(acc, da_lo, db_hi, idesc, 1);
            ptx::mma_f16_ss(acc, da_hi, db_lo, idesc, 1);
            accumulate = 1;
          }
          ptx::tc_commit(bar_empty(slot));
          if (++slot == (uint32_t)p.stages) {
            slot = 0;
            phase ^= 1;
          }
        }
        ptx::tc_commit(bar_acc_full(buf));
      }
    }
  } else {
    // epilogue warps: TMEM lane quarter = warp % 4, the kConvEpiParts warps of a quarter take every kConvEpiParts-th 16-column chunk
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* xp = xpose + (warp - 2) * (32 * 16);
    const int rs = lane >> 2;
    const float unscale = conv_unscale(p);
    float oscale = 1.0f;
    if (p.out_hi) {
      const int e_out = conv_plane_exp(p);
      oscale = ldexpf(1.0f, e_out);
      if (blockIdx.x == 0 && threadIdx.x == 64) *p.exp_out = e_out;
    }
    float local_max = 0.0f;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
      const int px = mt % p.tiles_x, r1 = mt / p.tiles_x;
      const int py = r1 % p.tiles_y, b = r1 / p.tiles_y;
      long long pix[4];  // output pixel of the 4 rows this lane stores per chunk; -1 = outside the image
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int r = q * 32 + ps * 8 + rs;
        const int oy = py * p.TH + (r >> p.tw_log2), ox = px * TW + (r & (TW - 1));
        pix[ps] = (oy < p.Ho && ox < p.Wo) ? ((long long)b * p.Ho + oy) * p.Wo + ox : -1;
      }
      ptx::mbar_wait(bar_acc_full(buf), (lt >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + buf * kConvAccStride + ((uint32_t)(q * 32) << 16);
      conv_epilogue_tile<ACT>(p, tacc, xp, lane, half, nt, pix, unscale, oscale, local_max);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_acc_empty(buf));
    }
    local_max = warp_max(local_max);
    if (lane == 0 && p.amax_out) atomic_max_nonneg(p.amax_out, local_max);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Halo variant for k x k kernels (stride 1): the A operand of ALL taps comes from one shared-memory copy of the
// patch plus its halo.  conv_tc_kernel fetches the shifted patch once per tap, i.e. kh*kw times from L2; here one 5-D TMA box per 32-channel chunk lands the (16+kh-1) x (8+kw-1) pixel
// halo as [slab of 8 channels][y][x][8 ch], which is the un-swizzled K-major core-matrix layout: 8 consecutive
// pixels of a patch row are the 8 rows of a core matrix (16 B apart), the next patch row is SBO = one halo row
// further, the next 8 channels LBO = one slab further, and tap (ky,kx) is nothing but a different start address.
// Narrow layers (N <= 64) process two patches (32 x 8 pixels) per weight stage.
constexpr int kHaloTW = 8, kHaloTH = 16;  // patch = 16 rows x 8 columns of output pixels
constexpr int kHaloAStages = 3;
constexpr int kHaloMaxBStages = 12;

struct HaloParams {
  ConvParams c;          // Ho, Wo, N, BN, n_tiles_n, taps, kw, pad, w_exp, pointers, ldc; tiles_x/tiles_y in super-tiles
  int np;                // patches per tile (1 or 2), stacked in y
  int hw, hh;            // halo width / height in pixels
  int chunks32;          // 32-channel chunks of the (16-padded) input channels
  int cp16;              // input channels padded to 16
  int b_stages;
};

template <int ACT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo, const HaloParams hp) {
  const ConvParams& p = hp.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t slab_bytes = (uint32_t)hp.hw * hp.hh * 16;           // one 8-channel slab of one halo
  const uint32_t plane_bytes = 4 * slab_bytes;                        // 32 channels
  const uint32_t a_stage = (uint32_t)hp.np * 2 * plane_bytes;         // np halos x (hi, lo)
  const uint32_t a_stage_al = (a_stage + 1023u) & ~1023u;
  const uint32_t b_half = (uint32_t)p.BN * 32, b_stage = 2 * b_half;  // BN rows x 16 channels x 2 B, hi + lo
  const uint32_t b_base = base + kHaloAStages * a_stage_al;
  const uint32_t bar0 = b_base + hp.b_stages * b_stage;
  auto bar_a_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_a_empty = [&](int i) { return bar0 + 8u * (kHaloAStages + i); };
  auto bar_b_full = [&](int i) { return bar0 + 8u * (2 * kHaloAStages + i); };
  auto bar_b_empty = [&](int i) { return bar0 + 8u * (2 * kHaloAStages + kHaloMaxBStages + i); };
  auto bar_acc_full = [&](int i) { return bar0 + 8u * (2 * kHaloAStages + 2 * kHaloMaxBStages + i); };
  auto bar_acc_empty = [&](int i) { return bar0 + 8u * (2 * kHaloAStages + 2 * kHaloMaxBStages + 2 + i); };
  const uint32_t slot_off = (bar0 - base) + 8u * (2 * kHaloAStages + 2 * kHaloMaxBStages + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + slot_off);
  float* xpose = reinterpret_cast<float*>(base_ptr + slot_off + 16);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kHaloAStages; ++i) {
      ptx::mbar_init(bar_a_full(i), 1);
      ptx::mbar_init(bar_a_empty(i), 1);
    }
    for (int i = 0; i < hp.b_stages; ++i) {
      ptx::mbar_init(bar_b_full(i), 1);
      ptx::mbar_init(bar_b_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_acc_full(i), 1);
      ptx::mbar_init(bar_acc_empty(i), kConvEpiWarps);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tm_xhi);
    ptx::prefetch_tmap(&tm_xlo);
  }
  if (warp == 1) ptx::tmem_alloc<512>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& nt, int& px, int& py, int& b) {
    nt = tile % p.n_tiles_n;
    const int mt = tile / p.n_tiles_n;
    px = mt % p.tiles_x;
    const int r1 = mt / p.tiles_x;
    py = r1 % p.tiles_y;
    b = r1 / p.tiles_y;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      // items = (tile, chunk) in order; the halo of item i+1 is requested before the weight stages of item i
      uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;  // running ring positions (no runtime division per step)
      auto load_halo = [&](int tile, int chunk) {
        int nt, px, py, b;
        decode(tile, nt, px, py, b);
        const uint32_t slot = a_slot;
        ptx::mbar_wait(bar_a_empty(slot), a_phase ^ 1);
        ptx::mbar_arrive_expect_tx(bar_a_full(slot), a_stage);
        const uint32_t dst = base + slot * a_stage_al;
        for (int q = 0; q < hp.np; ++q) {
          const int x0 = px * kHaloTW - p.pad, y0 = (py * hp.np + q) * kHaloTH - p.pad;
          ptx::tma_load_5d(dst + q * 2 * plane_bytes, &tm_xhi, bar_a_full(slot), 0, x0, y0, chunk * 4, b);
          ptx::tma_load_5d(dst + q * 2 * plane_bytes + plane_bytes, &tm_xlo, bar_a_full(slot), 0, x0, y0, chunk * 4, b);
        }
        if (++a_slot == kHaloAStages) {
          a_slot = 0;
          a_phase ^= 1;
        }
      };
      int tile = blockIdx.x;
      if (tile < p.total_tiles) load_halo(tile, 0);
      for (; tile < p.total_tiles; tile += gridDim.x) {
        const uint8_t* wsrc = p.wpack + (size_t)(tile % p.n_tiles_n) * (p.taps * (hp.cp16 >> 4)) * b_stage;
        for (int chunk = 0; chunk < hp.chunks32; ++chunk) {
          if (chunk + 1 < hp.chunks32)
            load_halo(tile, chunk + 1);
          else if (tile + (int)gridDim.x < p.total_tiles)
            load_halo(tile + gridDim.x, 0);
          const int k16s = min(2, (hp.cp16 - chunk * 32) >> 4);
          for (int tap = 0; tap < p.taps; ++tap) {
            for (int kk = 0; kk < k16s; ++kk) {
              const uint32_t slot = b_slot;
              ptx::mbar_wait(bar_b_empty(slot), b_phase ^ 1);
              ptx::mbar_arrive_expect_tx(bar_b_full(slot), b_stage);
              const int k16 = tap * (hp.cp16 >> 4) + chunk * 2 + kk;
              ptx::bulk_load(b_base + slot * b_stage, wsrc + (size_t)k16 * b_stage, b_stage, bar_b_full(slot));
              if (++b_slot == (uint32_t)hp.b_stages) {
                b_slot = 0;
                b_phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(kConvBM, p.BN);
      const uint32_t a_sbo = (uint32_t)hp.hw * 16, a_lbo = slab_bytes;
      // stage-0 descriptors; other stages / planes / taps / K16 steps are additions to the 16-byte address field
      const uint64_t da0 = ptx::make_smem_desc(base, a_lbo, a_sbo, 0);
      const uint64_t db0 = ptx::make_smem_desc(b_base, 16, 256, 6);
      const uint32_t plane16 = plane_bytes >> 4, slab2_16 = (2 * slab_bytes) >> 4, a_stage16 = a_stage_al >> 4;
      const uint32_t b_stage16 = b_stage >> 4, b_half16 = b_half >> 4;
      uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        ptx::mbar_wait(bar_acc_empty(buf), ((lt >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * kConvAccStride;
        uint32_t accumulate = 0;
        for (int chunk = 0; chunk < hp.chunks32; ++chunk) {
          ptx::mbar_wait(bar_a_full(a_slot), a_phase);
          ptx::tc_fence_after();
          const uint64_t da_stage = da0 + a_slot * a_stage16;
          const int k16s = min(2, (hp.cp16 - chunk * 32) >> 4);
          int kx = 0;
          uint32_t tap16 = 0;  // (ky * hw + kx) in 16-byte units
          for (int tap = 0; tap < p.taps; ++tap) {
            for (int kk = 0; kk < k16s; ++kk) {
              ptx::mbar_wait(bar_b_full(b_slot), b_phase);
              ptx::tc_fence_after();
              const uint64_t db_hi = db0 + b_slot * b_stage16, db_lo = db_hi + b_half16;
              uint64_t da_hi = da_stage + tap16 + kk * slab2_16;
              for (int q = 0; q < hp.np; ++q, da_hi += 2 * plane16) {
                const uint64_t da_lo = da_hi + plane16;
                const uint32_t d = acc + q * p.BN;
                ptx::mma_f16_ss(d, da_hi, db_hi, idesc, accumulate);
                ptx::mma_f16_ss(d, da_lo, db_hi, idesc, 1);
                ptx::mma_f16_ss(d, da_hi, db_lo, idesc, 1);
              }
              accumulate = 1;
              ptx::tc_commit(bar_b_empty(b_slot));
              if (++b_slot == (uint32_t)hp.b_stages) {
                b_slot = 0;
                b_phase ^= 1;
              }
            }
            ++tap16;
            if (++kx == p.kw) {
              kx = 0;
              tap16 += hp.hw - p.kw;
            }
          }
          ptx::tc_commit(bar_a_empty(a_slot));
          if (++a_slot == kHaloAStages) {
            a_slot = 0;
            a_phase ^= 1;
          }
        }
        ptx::tc_commit(bar_acc_full(buf));
      }
    }
  } else {
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    float* xp = xpose + (warp - 2) * (32 * 16);
    const int rs = lane >> 2;
    const float unscale = conv_unscale(p);
    float oscale = 1.0f;
    if (p.out_hi) {
      const int e_out = conv_plane_exp(p);
      oscale = ldexpf(1.0f, e_out);
      if (blockIdx.x == 0 && threadIdx.x == 64) *p.exp_out = e_out;
    }
    float local_max = 0.0f;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      int nt, px, py, b;
      decode(tile, nt, px, py, b);
      ptx::mbar_wait(bar_acc_full(buf), (lt >> 1) & 1);
      ptx::tc_fence_after();
      for (int q = 0; q < hp.np; ++q) {
        long long pix[4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int r = q4 * 32 + ps * 8 + rs;
          const int oy = (py * hp.np + q) * kHaloTH + (r >> 3), ox = px * kHaloTW + (r & 7);
          pix[ps] = (oy < p.Ho && ox < p.Wo) ? ((long long)b * p.Ho + oy) * p.Wo + ox : -1;
        }
        const uint32_t tacc = tmem_base + buf * kConvAccStride + q * p.BN + ((uint32_t)(q4 * 32) << 16);
        conv_epilogue_tile<ACT>(p, tacc, xp, lane, half, nt, pix, unscale, oscale, local_max);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_acc_empty(buf));
    }
    local_max = warp_max(local_max);
    if (lane == 0 && p.amax_out) atomic_max_nonneg(p.amax_out, local_max);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn conv_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// bk: 32 -> 64-byte swizzle, 16 -> 32-byte swizzle, 0 -> no swizzle
int encode(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, int bk,
           const char* what, int spatial_stride = 1) {
  EncodeTiledFn fn = conv_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SIR_E_CUDA;
  }
  cuuint32_t estr[5] = {1, (cuuint32_t)spatial_stride, (cuuint32_t)spatial_stride, 1, 1};  // dims 1, 2 = x, y of the 4-D patch map
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : bk == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return SIR_E_CUDA;
  }
  return SIR_OK;
}
int sm_count() {
  static int per_dev[64] = {};
  int& n = per_dev[current_device_slot()];
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
}  // namespace

// Column tile: multiple of 16 up to 256 that wastes the fewest padded columns (ties -> the wider tile).
int conv_tile_n(int N) {
  int best = 16;
  long best_cols = -1;
  for (int bn = 16; bn <= 256; bn += 16) {
    const long cols = (long)ceil_div(N, bn) * bn;
    if (best_cols < 0 || cols < best_cols || (cols == best_cols && bn > best)) {
      best = bn;
      best_cols = cols;
    }
  }
  return best;
}

}  // namespace sir

namespace sir {
namespace {
// Column tile for the halo kernel: np patches share a tile, so np * BN accumulator columns per TMEM buffer (<= 256).
int halo_tile_n(int N, int np) {
  int best = 16;
  long best_cols = -1;
  for (int bn = 16; bn * np <= 256; bn += 16) {
    const long cols = (long)ceil_div(N, bn) * bn;
    if (best_cols < 0 || cols < best_cols || (cols == best_cols && bn > best)) {
      best = bn;
      best_cols = cols;
    }
  }
  return best;
}
struct HaloPlan {
  HaloParams hp;
  size_t smem;
  bool ok;
};
HaloPlan plan_halo(int B, int H, int W, int C, int kh, int kw, int pad, int N, int cp16) {
  HaloPlan pl{};
  HaloParams& hp = pl.hp;
  ConvParams& p = hp.c;
  p.Ho = H + 2 * pad - kh + 1;
  p.Wo = W + 2 * pad - kw + 1;
  // Measured: one patch per tile with the widest N wins for N > 64 (a 128 x N x 16 MMA costs max(61, 0.51 N) cycles, so two patches
  // with half the N double the MMA time although they halve the weight bytes); narrow layers (N <= 64) pair patches.
  static const char* np_env = getenv("SIR_CONV_HALO_NP");
  hp.np = np_env ? atoi(np_env) : (N <= 64 ? 2 : 1);
  if (p.Ho <= kHaloTH) hp.np = 1;
  p.N = N;
  p.BN = halo_tile_n(N, hp.np);
  p.n_tiles_n = ceil_div(N, p.BN);
  p.taps = kh * kw;
  p.kw = kw;
  p.pad = pad;
  hp.hw = kHaloTW + kw - 1;
  hp.hh = kHaloTH + kh - 1;
  hp.cp16 = cp16;
  hp.chunks32 = ceil_div(cp16, 32);
  p.tiles_x = ceil_div(p.Wo, kHaloTW);
  p.tiles_y = ceil_div(p.Ho, kHaloTH * hp.np);
  const long long total = (long long)B * p.tiles_x * p.tiles_y * p.n_tiles_n;
  const uint32_t a_stage = ((uint32_t)hp.np * 2 * 4 * hp.hw * hp.hh * 16 + 1023u) & ~1023u;
  const uint32_t b_stage = (uint32_t)p.BN * 64;
  const uint32_t tail = 8u * (2 * kHaloAStages + 2 * kHaloMaxBStages + 4) + 16 + kConvEpiWarps * 32 * 16 * 4;
  const long long room = 220ll * 1024 - 1024 - tail - (long long)kHaloAStages * a_stage;
  hp.b_stages = (int)std::min<long long>(kHaloMaxBStages, room / b_stage);
  pl.ok = total < (1ll << 31) && hp.b_stages >= 4 && hp.hw * 16 < (1 << 18) && 4 * hp.hw * hp.hh * 16 < (1 << 18);
  p.total_tiles = (int)total;
  pl.smem = 1024 + (size_t)kHaloAStages * a_stage + (size_t)hp.b_stages * b_stage + tail;
  return pl;
}
}  // namespace
}  // namespace sir

namespace sir {
namespace {
// Everything the launch needs that depends only on the shapes: which kernel, its tiles, its shared memory.
struct ConvPlan {
  bool halo;
  ConvParams p;    // the 4-D patch kernel
  HaloPlan hpl;    // the halo kernel (valid when halo)
  size_t smem;
  int gB, gH, gW;  // image geometry seen by the patch kernel (a 1x1 convolution is one 1 x M row)
  int Kp;
  int BN, granule, n_tiles_n, k_steps;  // layout of the packed weights
};

int make_plan(int B, int H, int W, int C, int kh, int kw, int pad, int stride, int bk, int N, int per_image, ConvPlan* out) {
  SIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && kh > 0 && kw > 0 && pad >= 0 && N > 0,
                "sir_feat_conv: bad shape B=%d H=%d W=%d C=%d (C must be a multiple of 8) k=%dx%d N=%d", B, H, W, C, kh, kw, N);
  SIR_CHECK_ARG(bk == 16 || bk == 32, "sir_feat_conv: bk must be 16 or 32, got %d", bk);
  SIR_CHECK_ARG(N % 4 == 0, "sir_feat_conv: N = %d must be a multiple of 4", N);
  SIR_CHECK_ARG(stride >= 1 && stride <= 8, "sir_feat_conv: stride %d not supported", stride);
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  SIR_CHECK_ARG(H + 2 * pad >= kh && W + 2 * pad >= kw && Ho > 0 && Wo > 0, "sir_feat_conv: empty output");
  ConvPlan& pl = *out;
  pl = ConvPlan{};
  ConvParams& p = pl.p;
  p.N = N;
  p.BN = conv_tile_n(N);
  p.n_tiles_n = ceil_div(N, p.BN);
  p.taps = kh * kw;
  p.kw = kw;
  p.stride = stride;
  p.pad = pad;
  p.bk = bk;
  p.chunks = ceil_div(C, bk);
  pl.Kp = p.taps * p.chunks * bk;
  pl.gB = B;
  pl.gH = H;
  pl.gW = W;
  SIR_CHECK_ARG(!per_image || (p.taps == 1 && pad == 0 && stride == 1), "sir_feat_conv: per-image weights need a 1x1 convolution");
  if (p.taps == 1 && pad == 0 && stride == 1) {  // flat rows; with per-image weights one row per image so that no tile straddles two images
    SIR_CHECK_ARG((long long)B * H * W < (1ll << 31), "sir_feat_conv: too many rows");
    pl.gW = per_image ? H * W : B * H * W;
    pl.gH = 1;
    pl.gB = per_image ? B : 1;
  }
  p.Ho = (pl.gH + 2 * pad - kh) / stride + 1;
  p.Wo = (pl.gW + 2 * pad - kw) / stride + 1;
  long long best_tiles = -1;
  for (int l2 = 7; l2 >= 0; --l2) {  // patch = (128 >> l2) rows x (1 << l2) columns; prefer wide patches on ties
    const int tw = 1 << l2, th = kConvBM >> l2;
    const long long tiles = (long long)ceil_div(p.Ho, th) * ceil_div(p.Wo, tw);
    if (best_tiles < 0 || tiles < best_tiles) {
      best_tiles = tiles;
      p.tw_log2 = l2;
      p.TH = th;
    }
  }
  p.tiles_x = ceil_div(p.Wo, 1 << p.tw_log2);
  p.tiles_y = ceil_div(p.Ho, p.TH);
  const long long total = (long long)pl.gB * p.tiles_x * p.tiles_y * p.n_tiles_n;
  SIR_CHECK_ARG(total < (1ll << 31), "sir_feat_conv: too many tiles");
  p.total_tiles = (int)total;
  const uint32_t stage_bytes = (uint32_t)(2 * kConvBM + 2 * p.BN) * bk * 2;
  const uint32_t tail = 8u * (2 * kConvMaxStages + 4) + 16 + kConvEpiWarps * 32 * 16 * 4;
  p.stages = std::min<int>(kConvMaxStages, (int)((220u * 1024 - 1024 - tail) / stage_bytes));
  SIR_CHECK_ARG(p.stages >= 2, "sir_feat_conv: tile does not fit shared memory");
  pl.smem = 1024 + (size_t)p.stages * stage_bytes + tail;
  pl.BN = p.BN;
  pl.granule = bk;
  pl.n_tiles_n = p.n_tiles_n;
  pl.k_steps = p.taps * p.chunks;
  if (p.taps > 1 && stride == 1) {
    pl.hpl = plan_halo(B, H, W, C, kh, kw, pad, N, p.chunks * bk);
    // activation-side TMA rows per output row and K16 step: the patch kernel refetches the patch for every tap,
    // the halo kernel fetches patch + halo once per 32-channel chunk (16-byte rows)
    static const char* force = getenv("SIR_CONV_HALO");  // "0" / "1" force the choice (testing)
    pl.halo = pl.hpl.ok && (force ? force[0] == '1' : true);
    if (pl.halo) {
      pl.BN = pl.hpl.hp.c.BN;
      pl.granule = 16;
      pl.n_tiles_n = pl.hpl.hp.c.n_tiles_n;
      pl.k_steps = p.taps * (pl.hpl.hp.cp16 >> 4);
      pl.smem = pl.hpl.smem;
    }
  }
  return SIR_OK;
}

// weights [rows][Kp] hi/lo -> per-stage shared-memory images [n_tile][k_step][hi | lo][BN][granule] with the 32- / 64-byte
// swizzle the UMMA descriptors expect (16-byte chunk index XOR row bits), rows beyond n_rows read as zero
__global__ void __launch_bounds__(256) pack_weights_kernel(const __half* __restrict__ whi, const __half* __restrict__ wlo, int n_rows, int Kp,
                                                           int BN, int g, int n_tiles_n, int k_steps, __half* __restrict__ out) {
  const int chunks = g / 8;  // 16-byte chunks per row
  const size_t total = (size_t)n_tiles_n * k_steps * 2 * BN * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % chunks);
    size_t r = i / chunks;
    const int row = (int)(r % BN);
    r /= BN;
    const int plane = (int)(r & 1);
    r >>= 1;
    const int ks = (int)(r % k_steps), nt = (int)(r / k_steps);
    const int n = nt * BN + row, k = ks * g + c * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < n_rows && k < Kp) v = *reinterpret_cast<const uint4*>((plane ? wlo : whi) + (size_t)n * Kp + k);
    const int cs = g == 16 ? (c ^ ((row >> 2) & 1)) : (c ^ ((row >> 1) & 3));
    *reinterpret_cast<uint4*>(out + ((((size_t)nt * k_steps + ks) * 2 + plane) * BN + row) * g + cs * 8) = v;
  }
}
// per-image weights W_b[n][k] = W[n][k] * scale[b][channel(k)] (SqueezeExcitation folded into the projection), re-split into
// hi/lo and written in the packed layout, one set per image
__global__ void __launch_bounds__(256) scale_pack_weights_kernel(const __half* __restrict__ whi, const __half* __restrict__ wlo, int n_rows,
                                                                 int Kp, int Cp, int C, const float* __restrict__ scale, int B, int BN, int g,
                                                                 int n_tiles_n, int k_steps, __half* __restrict__ out) {
  const int chunks = g / 8;
  const size_t per_image = (size_t)n_tiles_n * k_steps * 2 * BN * chunks, total = per_image / 2 * B;  // one thread writes hi and lo
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % chunks);
    size_t r = i / chunks;
    const int row = (int)(r % BN);
    r /= BN;
    const int ks = (int)(r % k_steps);
    r /= k_steps;
    const int nt = (int)(r % n_tiles_n), b = (int)(r / n_tiles_n);
    const int n = nt * BN + row, k = ks * g + c * 8;
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hi[j] = lo[j] = __float2half_rn(0.0f);
    if (n < n_rows && k < Kp) {
      const uint4 vh = *reinterpret_cast<const uint4*>(whi + (size_t)n * Kp + k), vl = *reinterpret_cast<const uint4*>(wlo + (size_t)n * Kp + k);
      const __half* ph = reinterpret_cast<const __half*>(&vh);
      const __half* pl = reinterpret_cast<const __half*>(&vl);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = (k + j) % Cp;
        const float sc = ch < C ? scale[(size_t)b * C + ch] : 0.0f;
        const float v = (__half2float(ph[j]) + __half2float(pl[j])) * sc;
        hi[j] = __float2half_rn(v);
        lo[j] = __float2half_rn(v - __half2float(hi[j]));
      }
    }
    const int cs = g == 16 ? (c ^ ((row >> 2) & 1)) : (c ^ ((row >> 1) & 3));
    __half* dst = out + (size_t)b * per_image * 8 + ((((size_t)nt * k_steps + ks) * 2) * BN + row) * g + cs * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + (size_t)BN * g) = *reinterpret_cast<const uint4*>(lo);
  }
}
}  // namespace
}  // namespace sir

using namespace sir;

extern "C" int sir_feat_conv_scale_weights(const uint16_t* d_whi, const uint16_t* d_wlo, int n_rows, int Kp, int Cp, int C,
                                           const float* d_scale, int B, int tile_n, int granule, uint8_t* d_pack, void* stream) {
  SIR_CHECK_ARG(d_whi && d_wlo && d_scale && d_pack && n_rows > 0 && Kp > 0 && Kp % 8 == 0 && B > 0 && C > 0 && Cp >= C && Kp % Cp == 0,
                "sir_feat_conv_scale_weights: bad argument");
  SIR_CHECK_ARG(tile_n >= 16 && tile_n <= 256 && tile_n % 16 == 0 && (granule == 16 || granule == 32) && Kp % granule == 0,
                "sir_feat_conv_scale_weights: bad tile (tile_n %d, granule %d, Kp %d)", tile_n, granule, Kp);
  SIR_CHECK_ARG((((uintptr_t)d_whi | (uintptr_t)d_wlo | (uintptr_t)d_pack) & 15) == 0, "sir_feat_conv_scale_weights: pointers must be 16-byte aligned");
  const int n_tiles_n = ceil_div(n_rows, tile_n), k_steps = Kp / granule;
  const size_t total = (size_t)B * n_tiles_n * k_steps * tile_n * (granule / 8);
  scale_pack_weights_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)d_whi, (const __half*)d_wlo, n_rows, Kp, Cp, C, d_scale, B, tile_n, granule, n_tiles_n, k_steps, (__half*)d_pack);
  SIR_LAUNCH_CHECK("scale_pack_weights_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_conv_tile_n(int N) { return N > 0 ? conv_tile_n(N) : 0; }

extern "C" int sir_feat_conv_plan(int B, int H, int W, int C, int kh, int kw, int pad, int stride, int bk, int N, int per_image,
                                  int* tile_n, int* granule, long long* pack_bytes) {
  ConvPlan pl;
  const int rc = make_plan(B, H, W, C, kh, kw, pad, stride, bk, N, per_image, &pl);
  if (rc) return rc;
  if (tile_n) *tile_n = pl.BN;
  if (granule) *granule = pl.granule;
  if (pack_bytes) *pack_bytes = (long long)pl.n_tiles_n * pl.k_steps * 2 * pl.BN * pl.granule * 2;
  return SIR_OK;
}

extern "C" int sir_feat_conv_pack_weights(const uint16_t* d_whi, const uint16_t* d_wlo, int n_rows, int Kp, int tile_n, int granule,
                                          uint8_t* d_pack, void* stream) {
  SIR_CHECK_ARG(d_whi && d_wlo && d_pack && n_rows > 0 && Kp > 0 && Kp % 8 == 0, "sir_feat_conv_pack_weights: bad argument");
  SIR_CHECK_ARG(tile_n >= 16 && tile_n <= 256 && tile_n % 16 == 0 && (granule == 16 || granule == 32) && Kp % granule == 0,
                "sir_feat_conv_pack_weights: bad tile (tile_n %d, granule %d, Kp %d)", tile_n, granule, Kp);
  SIR_CHECK_ARG((((uintptr_t)d_whi | (uintptr_t)d_wlo | (uintptr_t)d_pack) & 15) == 0, "sir_feat_conv_pack_weights: pointers must be 16-byte aligned");
  const int n_tiles_n = ceil_div(n_rows, tile_n), k_steps = Kp / granule;
  const size_t total = (size_t)n_tiles_n * k_steps * 2 * tile_n * (granule / 8);
  pack_weights_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)d_whi, (const __half*)d_wlo, n_rows, Kp, tile_n, granule, n_tiles_n, k_steps, (__half*)d_pack);
  SIR_LAUNCH_CHECK("pack_weights_kernel");
  return SIR_OK;
}

extern "C" int sir_feat_conv(const uint16_t* d_xhi, const uint16_t* d_xlo, const float* d_amax_in, int B, int H, int W, int C, int kh,
                             int kw, int pad, int stride, int bk, const uint8_t* d_wpack, int pack_tile_n, int pack_granule, int per_image, int N,
                             int w_exp,
                             const float* d_bias, const float* d_residual, int act, float* d_out, int ldc, float* d_amax_out,
                             const int32_t* d_exp_in, uint16_t* d_out_hi, uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult,
                             float bound_add, const float* d_amax_res, void* stream) {
  SIR_CHECK_ARG(d_xhi && d_xlo && d_wpack && d_amax_in && d_bias && (d_out || d_out_hi), "sir_feat_conv: null pointer");
  SIR_CHECK_ARG(!d_out_hi || (d_out_lo && d_exp_out && N % 8 == 0 && ((uintptr_t)d_out_hi & 15) == 0 && ((uintptr_t)d_out_lo & 15) == 0 &&
                              bound_mult >= 0.0f && bound_add >= 0.0f),
                "sir_feat_conv: operand-plane output needs d_out_lo, d_exp_out, N %% 8 == 0 and a non-negative bound");
  SIR_CHECK_ARG(act >= 0 && act <= 2, "sir_feat_conv: unknown activation %d", act);
  ConvPlan pl;
  int rc = make_plan(B, H, W, C, kh, kw, pad, stride, bk, N, per_image, &pl);
  if (rc) return rc;
  SIR_CHECK_ARG(ldc >= N, "sir_feat_conv: ldc %d < N %d", ldc, N);
  SIR_CHECK_ARG(pack_tile_n == pl.BN && pack_granule == pl.granule,
                "sir_feat_conv: weights packed for tile_n %d / granule %d, this shape needs %d / %d (sir_feat_conv_plan)", pack_tile_n,
                pack_granule, pl.BN, pl.granule);
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  SIR_CHECK_ARG((long long)B * Ho * Wo * (long long)ldc < (1ll << 40), "sir_feat_conv: output too large");
  SIR_CHECK_ARG(d_out || N % 4 == 0, "sir_feat_conv: plane-only output needs N %% 4 == 0");
  SIR_CHECK_ARG(!d_residual || !d_out_hi || d_amax_res, "sir_feat_conv: residual + operand planes need d_amax_res");
  SIR_CHECK_ARG(((uintptr_t)d_bias & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && ldc % 4 == 0 && (!d_residual || ((uintptr_t)d_residual & 15) == 0) &&
                    ((uintptr_t)d_xhi & 15) == 0 && ((uintptr_t)d_xlo & 15) == 0 && ((uintptr_t)d_wpack & 15) == 0,
                "sir_feat_conv: operands must be 16-byte aligned and ldc a multiple of 4");
  ConvParams& p = pl.halo ? pl.hpl.hp.c : pl.p;
  p.wpack = d_wpack;
  p.wpack_image_bytes = per_image ? (long long)pl.n_tiles_n * pl.k_steps * 2 * pl.BN * pl.granule * 2 : 0;
  p.w_exp = w_exp;
  p.amax_in = d_amax_in;
  p.exp_in = d_exp_in;
  p.bias = d_bias;
  p.residual = d_residual;
  p.out = d_out;
  p.amax_out = d_amax_out;
  p.ldc = ldc;
  p.out_hi = (__half*)d_out_hi;
  p.out_lo = (__half*)d_out_lo;
  p.exp_out = d_exp_out;
  p.bound_mult = bound_mult;
  p.bound_add = bound_add;
  p.amax_res = d_residual ? d_amax_res : nullptr;
  cudaStream_t st = (cudaStream_t)stream;

  if (pl.halo) {
    HaloParams& hp = pl.hpl.hp;
    CUtensorMap hxh, hxl;
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)hp.hw, (cuuint32_t)hp.hh, 4, 1};
    rc = encode(&hxh, d_xhi, 5, dims, strides, box, 0, "activation hi (halo)");
    if (rc) return rc;
    rc = encode(&hxl, d_xlo, 5, dims, strides, box, 0, "activation lo (halo)");
    if (rc) return rc;
    static thread_local bool halo_configured_dev[64][3] = {};
    bool* halo_configured = halo_configured_dev[current_device_slot()];
    const void* hfn = act == 0 ? (const void*)conv_halo_kernel<0> : act == 1 ? (const void*)conv_halo_kernel<1> : (const void*)conv_halo_kernel<2>;
    if (!halo_configured[act]) {
      SIR_CUDA(cudaFuncSetAttribute(hfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024)));
      halo_configured[act] = true;
    }
    const unsigned hgrid = (unsigned)std::min<long long>(hp.c.total_tiles, sm_count());
    if (act == 0)
      conv_halo_kernel<0><<<hgrid, kConvThreads, pl.smem, st>>>(hxh, hxl, hp);
    else if (act == 1)
      conv_halo_kernel<1><<<hgrid, kConvThreads, pl.smem, st>>>(hxh, hxl, hp);
    else
      conv_halo_kernel<2><<<hgrid, kConvThreads, pl.smem, st>>>(hxh, hxl, hp);
    SIR_LAUNCH_CHECK("conv_halo_kernel");
    return SIR_OK;
  }

  CUtensorMap txh, txl;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)pl.gW, (cuuint64_t)pl.gH, (cuuint64_t)pl.gB};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)pl.gW * C * 2, (cuuint64_t)pl.gH * pl.gW * C * 2};
  // strided convolution: the box spans stride * (patch extent) input pixels and the TMA unit picks every stride-th one
  // (element strides; verified with tools/tma_stride_probe.cu: origin in input coordinates, dense patch in shared memory)
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)((1 << p.tw_log2) * stride), (cuuint32_t)(p.TH * stride), 1};
  SIR_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "sir_feat_conv: stride %d too large for the patch", stride);
  rc = encode(&txh, d_xhi, 4, dims, strides, box, bk, "activation hi", stride);
  if (rc) return rc;
  rc = encode(&txl, d_xlo, 4, dims, strides, box, bk, "activation lo", stride);
  if (rc) return rc;
  static thread_local bool configured_dev[64][3] = {};
  bool* configured = configured_dev[current_device_slot()];
  const void* fn = act == 0 ? (const void*)conv_tc_kernel<0> : act == 1 ? (const void*)conv_tc_kernel<1> : (const void*)conv_tc_kernel<2>;
  if (!configured[act]) {
    SIR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024)));
    configured[act] = true;
  }
  const unsigned grid = (unsigned)std::min<long long>(p.total_tiles, sm_count());
  if (act == 0)
    conv_tc_kernel<0><<<grid, kConvThreads, pl.smem, st>>>(txh, txl, p);
  else if (act == 1)
    conv_tc_kernel<1><<<grid, kConvThreads, pl.smem, st>>>(txh, txl, p);
  else
    conv_tc_kernel<2><<<grid, kConvThreads, pl.smem, st>>>(txh, txl, p);
  SIR_LAUNCH_CHECK("conv_tc_kernel");
  return SIR_OK;
}
