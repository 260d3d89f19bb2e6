// Probe: what does cuTensorMapEncodeTiled's elementStrides do?  A 4-D fp16 tensor [B=1][H=12][W=20][C=16] holds the value
// 100*y + x in every channel.  A box is loaded with element strides {1,2,2,1} under two guesses for boxDim (traversed extent vs
// elements written); shared memory is pre-filled with -1 and dumped after a delay, so the probe does not depend on the byte
// count the barrier would expect.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I shoeprint-image-retrieval_b200/csrc -o tools/tma_stride_probe_bin tools/tma_stride_probe.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "sir_ptx.cuh"
using namespace sir;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int y0, float* out, int n_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_storage;
  const uint32_t bar = ptx::smem_u32(&bar_storage);
  __half* sh = reinterpret_cast<__half*>(smem);
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) sh[i] = __float2half(-1.0f);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, 1 << 19);  // never completes; we only look at what lands
    ptx::tma_load_4d(ptx::smem_u32(smem), &tm, bar, 0, x0, y0, 0);
  }
  for (int i = 0; i < 2000; ++i) __nanosleep(100);
  __syncthreads();
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = __half2float(sh[i]);
}

int main() {
  const int H = 12, W = 20, C = 16;
  __half* h = new __half[H * W * C];
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = __float2half((float)(100 * y + x));
  __half* d;
  cudaMalloc(&d, H * W * C * 2);
  cudaMemcpy(d, h, H * W * C * 2, cudaMemcpyHostToDevice);
  float* d_out;
  const int n_out = 4096;
  cudaMalloc(&d_out, n_out * 4);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  Fn fn = (Fn)ptr;
  for (int guess = 0; guess < 2; ++guess) {
    const int TW = 4, TH = 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(guess == 0 ? 2 * TW : TW), (cuuint32_t)(guess == 0 ? 2 * TH : TH), 1};
    cuuint32_t estr[4] = {1, 2, 2, 1};
    CUtensorMap tm;
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("guess %d (boxDim = {%u,%u,%u,1}, elementStrides {1,2,2,1}): encode -> %d\n", guess, box[0], box[1], box[2], (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int origin = 0; origin < 2; ++origin) {
      const int x0 = origin ? -1 : 3, y0 = origin ? -1 : 2;
      probe<<<1, 128, 16384>>>(tm, x0, y0, d_out, n_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      float o[4096];
      cudaMemcpy(o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
      int written = 0;
      for (int i = 0; i < n_out; ++i) written += o[i] != -1.0f;
      printf("  origin (x0=%d, y0=%d): %d fp16 written = %d pixels; first pixels (channel 0):", x0, y0, written, written / C);
      for (int pix = 0; pix < 12 && pix * C < n_out; ++pix) printf(" %g", o[pix * C]);
      printf("\n");
    }
  }
  return 0;
}
