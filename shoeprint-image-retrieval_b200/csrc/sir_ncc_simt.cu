// K7 (check path): fp32 CUDA-core evaluation of the correlation score, literal windowed form.
//
//   s[g][n] = (1/C) max_{y,x} sum_c rnorm[g][c][y,x] * sum_{u,v} t[n][c][u,v] * gz[g][c][y+u-a][x+v-b]
//
// (similarity.py:53-55, 68, 106-108).  One CTA per (gallery, column).  Not the fast path: it
// exists as an independent fp32 evaluation of the same quantity the tcgen05 kernel produces, and
// serves SIR_PREC_FP32_SIMT.  Compute bound on the FMA pipe; shared memory holds one zero padded
// gallery channel and one template channel.
#include "sir_common.cuh"

namespace sir {

constexpr int kSimtThreads = 256;
constexpr int kSimtPosPerThread = 4;  // positions handled per thread per pass

__global__ void __launch_bounds__(kSimtThreads) ncc_simt_kernel(const float* __restrict__ gz, const float* __restrict__ rnorm,
                                                                int C, int Hp, int Wp, const float* __restrict__ t32,
                                                                int ncols, int ncols_alloc, int Hm, int Wm,
                                                                const int32_t* __restrict__ col2probe,
                                                                float* __restrict__ scores, int score_ld, int g0) {
  extern __shared__ float smem[];
  __shared__ float fred[32];
  const int PH = Hp + Hm - 1, PW = Wp + Wm - 1;
  float* sg = smem;            // [PH][PW] zero padded gallery channel
  float* st = smem + PH * PW;  // [Hm][Wm]
  const int a = Hm / 2, b = Wm / 2, M = Hp * Wp, K = Hm * Wm;
  const int g = blockIdx.x / ncols, n = blockIdx.x - g * ncols;

  float best = 0.0f;
  constexpr int kChunk = kSimtThreads * kSimtPosPerThread;
  for (int m0 = 0; m0 < M; m0 += kChunk) {
    float total[kSimtPosPerThread];
#pragma unroll
    for (int j = 0; j < kSimtPosPerThread; ++j) total[j] = 0.0f;
    for (int c = 0; c < C; ++c) {
      __syncthreads();
      const float* gsrc = gz + ((size_t)g * C + c) * M;
      for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) {
        const int y = i / PW - a, x = i % PW - b;
        sg[i] = (y >= 0 && y < Hp && x >= 0 && x < Wp) ? gsrc[y * Wp + x] : 0.0f;
      }
      const float* tsrc = t32 + ((size_t)c * ncols_alloc + n) * K;
      for (int i = threadIdx.x; i < K; i += blockDim.x) st[i] = tsrc[i];
      __syncthreads();
      const float* rsrc = rnorm + ((size_t)g * C + c) * M;
#pragma unroll
      for (int j = 0; j < kSimtPosPerThread; ++j) {
        const int m = m0 + j * kSimtThreads + threadIdx.x;
        if (m < M) {
          const int y = m / Wp, x = m - y * Wp;
          float acc = 0.0f;
          for (int u = 0; u < Hm; ++u) {
            const float* grow = sg + (y + u) * PW + x;
            const float* trow = st + u * Wm;
            for (int v = 0; v < Wm; ++v) acc = fmaf(trow[v], grow[v], acc);
          }
          total[j] = fmaf(rsrc[m], acc, total[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kSimtPosPerThread; ++j)
      if (m0 + j * kSimtThreads + threadIdx.x < M) best = fmaxf(best, total[j]);
  }
  best = block_max(best, fred);
  if (threadIdx.x == 0) atomic_max_nonneg(&scores[(size_t)col2probe[n] * score_ld + g0 + g], best / (float)C);
}

// Full correlation surface of ONE (template, image) channel pair: out[y][x] = rnorm[y][x] * sum t*g.
// Serves the public helper normxcorr() (similarity.py:26-72); never used by the matching path, which
// keeps the surface on chip.
__global__ void __launch_bounds__(256) ncc_surface_kernel(const float* __restrict__ gz, const float* __restrict__ rnorm, int Hp, int Wp,
                                                          const float* __restrict__ t32, int Hm, int Wm, float* __restrict__ out) {
  const int a = Hm / 2, b = Wm / 2;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Hp * Wp; m += gridDim.x * blockDim.x) {
    const int y = m / Wp, x = m - y * Wp;
    float acc = 0.0f;
    for (int u = 0; u < Hm; ++u) {
      const int iy = y + u - a;
      if (iy < 0 || iy >= Hp) continue;
      for (int v = 0; v < Wm; ++v) {
        const int ix = x + v - b;
        if (ix >= 0 && ix < Wp) acc = fmaf(t32[u * Wm + v], gz[iy * Wp + ix], acc);
      }
    }
    out[m] = acc * rnorm[m];
  }
}

int launch_ncc_simt(const float* d_gz, const float* d_rnorm, int G, int C, int Hp, int Wp, const float* d_t32, int ncols,
                    int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                    cudaStream_t st) {
  SIR_CHECK_ARG(d_gz && d_t32, "sir_ncc_scores(FP32_SIMT): needs d_gz and d_t32");
  const size_t smem = sizeof(float) * ((size_t)(Hp + Hm - 1) * (Wp + Wm - 1) + (size_t)Hm * Wm);
  SIR_CHECK_ARG(smem <= 227 * 1024, "sir_ncc_scores(FP32_SIMT): maps too large for shared memory");
  static thread_local size_t configured_dev[64] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    SIR_CUDA(cudaFuncSetAttribute(ncc_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const size_t blocks = (size_t)G * ncols;
  SIR_CHECK_ARG(blocks < (1ull << 31), "sir_ncc_scores(FP32_SIMT): too many pairs for one launch");
  ncc_simt_kernel<<<(unsigned)blocks, kSimtThreads, smem, st>>>(d_gz, d_rnorm, C, Hp, Wp, d_t32, ncols, ncols_alloc, Hm,
                                                                 Wm, d_col2probe, d_scores, score_ld, g0);
  SIR_LAUNCH_CHECK("ncc_simt_kernel");
  return SIR_OK;
}

}  // namespace sir

extern "C" int sir_ncc_surface(const float* d_gz, const float* d_rnorm, int Hp, int Wp, const float* d_t32, int Hm, int Wm, float* d_out,
                               void* stream) {
  SIR_CHECK_ARG(d_gz && d_rnorm && d_t32 && d_out, "sir_ncc_surface: null pointer");
  SIR_CHECK_ARG(Hp > 0 && Wp > 0 && Hm > 0 && Wm > 0, "sir_ncc_surface: bad shape");
  sir::ncc_surface_kernel<<<sir::ceil_div(Hp * Wp, 256), 256, 0, (cudaStream_t)stream>>>(d_gz, d_rnorm, Hp, Wp, d_t32, Hm, Wm, d_out);
  SIR_LAUNCH_CHECK("ncc_surface_kernel");
  return SIR_OK;
}
