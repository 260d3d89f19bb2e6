// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of N, operand layouts
// and accumulator dependence.  One thread per CTA issues a chain of MMAs on shared-memory operands that are already
// resident (contents irrelevant), commits, waits; all SMs run the same thing.  Build and run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I shoeprint-image-retrieval_b200/csrc -o /tmp/mma_probe tools/mma_probe.cu && /tmp/mma_probe
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "sir_ptx.cuh"
using namespace sir;

struct Cfg {
  int N;
  int a_layout, a_lbo, a_sbo;  // layout type (0 none, 2 128B, 4 64B, 6 32B)
  int b_layout, b_lbo, b_sbo;
  int a_kstep16, b_kstep16;    // descriptor address increment (16-byte units) between consecutive K16 steps
  int n_acc;                   // accumulators used round-robin (1 = every MMA depends on the previous one)
  int chain;                   // MMAs per commit
  int rounds;
};

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar_storage;
  __shared__ uint32_t tmem_slot;
  const uint32_t bar = ptx::smem_u32(&bar_storage);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(ptx::smem_u32(&tmem_slot));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(128, c.N);
    const uint64_t da0 = ptx::make_smem_desc(base, c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t db0 = ptx::make_smem_desc(base + 96 * 1024, c.b_lbo, c.b_sbo, c.b_layout);
    uint32_t phase = 0;
    long long best = 1ll << 62;
    for (int r = 0; r < c.rounds; ++r) {
      const long long t0 = clock64();
      const uint32_t acc_delta = c.n_acc == 2 ? (uint32_t)c.N : 0u;  // alternate two accumulators (or stay on one)
      uint32_t acc = tmem;
#pragma unroll 4
      for (int i = 0; i < c.chain; ++i) {
        const int k = i & 3;  // cycle over 4 K16 steps of the resident tiles
        ptx::mma_f16_ss(acc, da0 + (uint64_t)(k * c.a_kstep16), db0 + (uint64_t)(k * c.b_kstep16), idesc, 1);
        acc = (acc == tmem) ? tmem + acc_delta : tmem;
      }
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  struct Named { const char* name; int al, albo, asbo, ak; };
  const Named as[] = {
      {"A none  SBO160 LBO2880 (halo)", 0, 2880, 160, 2 * 2880 / 16},
      {"A none  SBO128 LBO2048 (dense)", 0, 2048, 128, 2 * 2048 / 16},
      {"A sw32  [128][16]", 6, 16, 256, 4096 / 16},
      {"A sw64  [128][32]", 4, 16, 512, 2},
      {"A sw128 [128][64]", 2, 16, 1024, 2},
  };
  struct NamedB { const char* name; int bl, blbo, bsbo; int per_row_bytes; };
  const NamedB bs[] = {{"B sw32 ", 6, 16, 256, 32}, {"B sw64 ", 4, 16, 512, 64}, {"B sw128", 2, 16, 1024, 128}};
  const int Ns[] = {32, 96, 192, 256};
  printf("cycles per MMA (M=128, K=16; math = N/4 cycles), chain of 96 MMAs, best of 5 rounds, median over 148 CTAs\n");
  for (const Named& a : as)
    for (const NamedB& b : bs) if ((&a - as) == 0 || (&a - as) == 3 || (&b - bs) == 1)
      for (int N : Ns)
        for (int n_acc : {1, 2}) {
          if (n_acc * N > 512) continue;
          Cfg c{};
          c.N = N;
          c.a_layout = a.al; c.a_lbo = a.albo; c.a_sbo = a.asbo; c.a_kstep16 = a.ak;
          c.b_layout = b.bl; c.b_lbo = b.blbo; c.b_sbo = b.bsbo;
          c.b_kstep16 = b.per_row_bytes == 32 ? (N * 32) / 16 : 2;  // sw32: next K16 = next tile; sw64/128: +32 B in the row
          c.n_acc = n_acc; c.chain = 96; c.rounds = 5;
          probe<<<148, 128, 210 * 1024>>>(c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s | %s N=%d: %s\n", a.name, b.name, N, cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          // median
          for (int i = 0; i < 148; ++i) for (int j = i + 1; j < 148; ++j) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
          printf("%-32s | %s | N=%3d | accs=%d | %6.1f cyc/MMA (math %5.1f)\n", a.name, b.name, N, n_acc, (double)h[74] / c.chain, N / 4.0);
        }
  return 0;
}
