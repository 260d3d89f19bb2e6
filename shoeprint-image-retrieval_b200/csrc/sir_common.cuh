// Shared helpers for libsir (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "sir.h"

namespace sir {

void set_error(const char* fmt, ...);

#define SIR_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::sir::set_error(__VA_ARGS__);    \
      return SIR_E_ARG;                 \
    }                                   \
  } while (0)

#define SIR_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::sir::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SIR_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

#define SIR_LAUNCH_CHECK(name)                                                              \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      ::sir::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));           \
      return SIR_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

// cudaFuncSetAttribute state is per DEVICE: call sites cache "already opted in" per (thread, device), so a host thread
// that drives a second GPU opts in there too.
inline int current_device_slot() {
  int dev = 0;
  (void)cudaGetDevice(&dev);
  return dev & 63;
}

constexpr int kEdge = 2;           // similarity.py:92-93 crop
constexpr int kNormChunkCols = 8;  // multi-shape column tiles: columns that share one window-norm table (sir_ncc_norm_chunk)
constexpr int kTemplateScaleLog2 = 10;  // packed templates are (t-mean)/sqrt(E) * 2^10
constexpr int kGalleryPeakLog2 = 10;    // packed gallery channels have max|v| in [2^9, 2^10)
// fp8 (e4m3: 2^-9 .. 448) companions of the fp16 operands for the correction MMAs: hi / 64 (peak 16) and
// lo * 64 (|lo| <= 0.5 -> 32).  The two factors cancel in every product; 64 centres both parts in the
// e4m3 range so that cells down to ~1/8000 of the channel peak still get their correction term.
constexpr float kFp8HiScale = 1.0f / 64.0f;
constexpr float kFp8LoScale = 64.0f;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Row pitch (cells) of the packed fp16 gallery operands: rows padded to 8 cells = 16 bytes so a TMA
// tensor map can walk them (global strides must be multiples of 16 bytes).
__host__ __device__ inline int gal_pitch(int Wp) { return round_up(Wp, 8); }

// Number of 8-tap chunks per template row and padded K (multiple of 32) -- see sir_template_kpad.
__host__ __device__ inline int tpl_chunks_per_row(int Wm) { return ceil_div(Wm, 8); }
__host__ __device__ inline int tpl_kpad(int Hm, int Wm) { return round_up(Hm * tpl_chunks_per_row(Wm) * 8, 32); }
// The fp8-corrected mode pads template rows to 16 taps (one 16-byte fp8 core-matrix row).
__host__ __device__ inline int tpl_row_taps(int Wm, int row_align) { return round_up(Wm, row_align); }
__host__ __device__ inline int tpl_kpad_aligned(int Hm, int Wm, int row_align) { return round_up(Hm * tpl_row_taps(Wm, row_align), 32); }
__host__ __device__ inline int gal_pitch8(int Wp) { return round_up(Wp, 16); }  // 1-byte operands: 16 cells = 16 bytes

// SiLU for the fused epilogues: v / (1 + e^-v) with ex2.approx and rcp.approx (2 MUFU + 3 FP32 ops, relative error ~2^-22;
// 1 + e^-v >= 1, so the reciprocal needs none of the range handling __fdividef carries; v -> -inf gives -0).
__device__ __forceinline__ float silu_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + __expf(-v)));
  return v * r;
}

// Feature stage: exponent e with amax * 2^e in [2^9, 2^10); 0 for an all-zero tensor.
__device__ __forceinline__ int scale_exp_from_amax(float amax) {
  if (!(amax > 0.0f) || !isfinite(amax)) return 0;
  int ex;
  (void)frexpf(amax, &ex);
  return kGalleryPeakLog2 - ex;
}

// warp / block reductions ---------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of doubles; `scratch` must hold 32 doubles. All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += scratch[i];  // same order in every thread
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = scratch[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, scratch[i]);
  return r;
}

// Non-negative floats order like their bit patterns: max with a plain integer atomic.
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  if (v > 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

}  // namespace sir
