// Development probe: which 3-D fp16 tiled TMA boxes does the hardware accept?
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../shoeprint-image-retrieval_b200/csrc/sir_ptx.cuh"
using namespace sir;
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int y0, int z0, int bytes, float* out, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = ptx::smem_u32(&bar);
  for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<__half*>(smem)[i] = __float2half(-7.0f);
  if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(b, bytes);
    ptx::tma_load_3d(ptx::smem_u32(smem), &tm, b, x0, y0, z0);
  }
  ptx::mbar_wait(b, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = __half2float(reinterpret_cast<__half*>(smem)[i]);
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), pitch = atoi(argv[4]);
  int bx = atoi(argv[5]), by = atoi(argv[6]), x0 = atoi(argv[7]), y0 = atoi(argv[8]), z0 = atoi(argv[9]);
  int promo = argc > 10 ? atoi(argv[10]) : 1;
  std::vector<__half> h((size_t)P * H * pitch);
  for (int z = 0; z < P; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[((size_t)z * H + y) * pitch + x] = __float2half(z * 100 + y * 10 + x * 0.5f);
  __half* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* ptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  CUtensorMap tm; cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}; cuuint64_t st[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * H};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
  CUresult r = ((Fn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d first host vals %f %f\n", (int)r, __half2float(h[1]), __half2float(h[pitch + 1]));
  int n = bx * by; float* out; cudaMalloc(&out, n * 4);
  probe<<<1, 64, 40 * 1024>>>(tm, x0, y0, z0, n * 2, out, n);
  cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    for (int y = 0; y < by && y < 4; ++y) { for (int x = 0; x < bx && x < 12; ++x) printf("%6.1f ", o[y * bx + x]); printf("\n"); } }
  return 0;
}
