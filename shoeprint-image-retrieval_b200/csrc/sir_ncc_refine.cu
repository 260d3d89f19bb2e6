// K7r: exact re-evaluation of the screening candidates (second half of SIR_PREC_FP16_REFINE).
//
// sir_ncc_screen runs the whole probe x gallery correlation (similarity.py:53-55,100-108,357-367) with plain fp16
// operands: one MMA per K step, but only ~2e-4 relative.  What the reference wants per pair is ONE number,
// max over positions and variants (similarity.py:106-108,365-367), so the exact arithmetic is only needed where the
// maximum can be: every (column, gallery, patch) record whose screened maximum lies within the candidate margin of
// the pair's screened maximum names the rows that are within the margin of it, and this kernel evaluates
//
//     s(n, g, y, x) = (1/C) sum_c rnorm_c[g][y,x] * sum_{u,v} (t_hi + t_lo)[n][c][u,v] * (g_hi + g_lo)[g][c][y+u-a][x+v-b]
//
// for exactly those positions in float32 (the hi + lo pairs carry 22 significant bits) and max-reduces the result
// into d_scores.  The margin covers twice the screening error, so the position of the true maximum is always among
// the candidates; if it ever were not, the result would still be an exactly evaluated correlation value within
// 2 * (screening error) of the true maximum.
//
// One CTA owns a tile of TN columns x TG gallery prints.  Per channel it stages the tile's template columns and
// gallery planes in shared memory as float32 (each operand byte is read once per tile, not once per candidate);
// a warp takes one candidate at a time, lanes spread over (template row, tap), zero rows of the "same" padding are
// skipped.  The work list is built without atomics (block scan over per-thread counts), so every run evaluates the
// same candidates in the same order.
#include <algorithm>

#include "sir_common.cuh"

namespace sir {

constexpr int kRefThreads = 256;
constexpr int kRefWarps = kRefThreads / 32;

struct RefineParams {
  const __half* ghi;
  const __half* glo;
  const float* rnorm;
  const float* const* rnorm_tab;
  const __half* thi;
  const __half* tlo;
  const int32_t* col2probe;
  const float* approx;
  float* scores;
  const uint2* rec;
  unsigned long long* stats;  // optional: [0] positions evaluated, [1] records that listed more than 3 rows, [2] tiles with work
  int G, C, Hp, Wp, WP, Hb, Wb, rowk, Kpad, ncols, ncols_alloc, npx, NP, score_ld, g0;
  int TN, TG, cap, tiles_g;
  float tau_rel, tau_abs, inv_scale;
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  int base = 0, sum = 0;
  for (int i = 0; i < kRefWarps; ++i) {
    if (i < wid) base += scratch[i];
    sum += scratch[i];
  }
  __syncthreads();
  *total = sum;
  return base + inc - v;
}

__global__ void __launch_bounds__(kRefThreads) ncc_refine_kernel(const RefineParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int PG = p.Hp * p.WP;  // cells of one packed gallery plane
  float* tpl = reinterpret_cast<float*>(smem_raw);                 // [TN][Kpad]
  float* gal = tpl + (size_t)p.TN * p.Kpad;                        // [TG][PG]
  uint2* list = reinterpret_cast<uint2*>(gal + (size_t)p.TG * PG);  // [cap] (j | i << 8, y | x << 16)
  float* acc = reinterpret_cast<float*>(list + p.cap);             // [cap]
  int* flags = reinterpret_cast<int*>(acc + p.cap);                // [TN + TG]
  int* scratch = flags + p.TN + p.TG;                              // [kRefWarps]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = (blockIdx.x / p.tiles_g) * p.TN, gt0 = (blockIdx.x % p.tiles_g) * p.TG;
  const int M = p.Hp * p.Wp;
  const int items = p.TN * p.TG * p.NP;

  // candidate rows of one record: 0 when the record cannot hold the pair's maximum
  auto expand = [&](int item, uint32_t* info_out, int* py_out, int* px_out, int* j_out, int* i_out) -> int {
    const int pidx = item % p.NP, pair = item / p.NP;
    const int i = pair % p.TG, j = pair / p.TG;
    const int n = n0 + j, g = gt0 + i;
    if (n >= p.ncols || g >= p.G) return 0;
    const float a = __ldg(p.approx + (size_t)__ldg(p.col2probe + n) * p.score_ld + p.g0 + g);
    if (!(a > 0.0f)) return 0;  // nothing positive was screened for this pair: the score stays at the 0 floor (similarity.py:355)
    const uint2 r = __ldg(p.rec + ((size_t)n * p.G + g) * p.NP + pidx);
    const float m = __uint_as_float(r.x);
    if (!(m >= a - (p.tau_rel * a + p.tau_abs)) || m == 0.0f) return 0;  // m == 0: an all-zero (flat / padding) template column
    *info_out = r.y;
    *py_out = pidx / p.npx;
    *px_out = pidx % p.npx;
    *j_out = j;
    *i_out = i;
    const int cnt = (int)(r.y >> 24);
    if (cnt <= 3) return cnt;
    return min(16, p.Hp - 16 * *py_out) * min(8, p.Wp - 8 * *px_out);  // every valid position of the patch
  };

  int mine = 0;
  for (int item = tid; item < items; item += kRefThreads) {
    uint32_t info;
    int py, px, j, i;
    mine += expand(item, &info, &py, &px, &j, &i);
  }
  int total = 0;
  const int base = block_exclusive_scan(mine, scratch, &total);
  if (total == 0) return;
  if (p.stats && tid == 0) {
    atomicAdd(p.stats + 0, (unsigned long long)total);
    atomicAdd(p.stats + 2, 1ull);
  }

  const int a = p.Hb / 2, b = p.Wb / 2;
  // lanes over (template row, tap): rows of up to 32 taps share a pass when they divide the warp
  const int vl = (p.rowk <= 32 && 32 % p.rowk == 0) ? p.rowk : 32;
  const int rpp = 32 / vl, lv = lane % vl, lr = lane / vl;

  for (int r0 = 0; r0 < total; r0 += p.cap) {
    const int nl = min(p.cap, total - r0);
    // ---- this round's slice of the work list, in scan order
    if (base < r0 + p.cap && base + mine > r0) {
      int o = base;
      for (int item = tid; item < items; item += kRefThreads) {
        uint32_t info;
        int py, px, j, i;
        const int cnt = expand(item, &info, &py, &px, &j, &i);
        if (cnt == 0) continue;
        if ((int)(info >> 24) <= 3) {
          for (int k = 0; k < cnt; ++k, ++o) {
            if (o < r0 || o >= r0 + p.cap) continue;
            const int row = (info >> (8 * k)) & 0xff;
            list[o - r0] = make_uint2((uint32_t)j | ((uint32_t)i << 8), (uint32_t)(16 * py + (row >> 3)) | ((uint32_t)(8 * px + (row & 7)) << 16));
          }
        } else {
          if (p.stats && r0 == 0) atomicAdd(p.stats + 1, 1ull);
          for (int row = 0; row < 128; ++row) {
            const int y = 16 * py + (row >> 3), x = 8 * px + (row & 7);
            if (y >= p.Hp || x >= p.Wp) continue;
            if (o >= r0 && o < r0 + p.cap) list[o - r0] = make_uint2((uint32_t)j | ((uint32_t)i << 8), (uint32_t)y | ((uint32_t)x << 16));
            ++o;
          }
        }
      }
    }
    for (int k = tid; k < p.TN + p.TG; k += kRefThreads) flags[k] = 0;
    __syncthreads();
    for (int e = tid; e < nl; e += kRefThreads) {
      acc[e] = 0.0f;
      flags[list[e].x & 0xff] = 1;
      flags[p.TN + (list[e].x >> 8)] = 1;
    }
    __syncthreads();

    for (int c = 0; c < p.C; ++c) {
      // ---- stage the flagged template columns and gallery planes of this channel as float32
      const int tch = p.Kpad / 8, gch = PG / 8;
      for (int idx = tid; idx < p.TN * tch; idx += kRefThreads) {
        const int j = idx / tch, ch = idx - j * tch;
        if (!flags[j]) continue;
        const size_t off = ((size_t)c * p.ncols_alloc + n0 + j) * p.Kpad + (size_t)ch * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.thi + off));
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(p.tlo + off));
        const __half2* hh = reinterpret_cast<const __half2*>(&h);
        const __half2* ll = reinterpret_cast<const __half2*>(&l);
        float4 o0, o1;
        float2 x0 = __half22float2(hh[0]), y0 = __half22float2(ll[0]);
        float2 x1 = __half22float2(hh[1]), y1 = __half22float2(ll[1]);
        o0 = make_float4(x0.x + y0.x, x0.y + y0.y, x1.x + y1.x, x1.y + y1.y);
        x0 = __half22float2(hh[2]); y0 = __half22float2(ll[2]);
        x1 = __half22float2(hh[3]); y1 = __half22float2(ll[3]);
        o1 = make_float4(x0.x + y0.x, x0.y + y0.y, x1.x + y1.x, x1.y + y1.y);
        float4* dst = reinterpret_cast<float4*>(tpl + (size_t)j * p.Kpad + (size_t)ch * 8);
        dst[0] = o0;
        dst[1] = o1;
      }
      for (int idx = tid; idx < p.TG * gch; idx += kRefThreads) {
        const int i = idx / gch, ch = idx - i * gch;
        if (!flags[p.TN + i]) continue;
        const size_t off = ((size_t)(gt0 + i) * p.C + c) * PG + (size_t)ch * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.ghi + off));
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(p.glo + off));
        const __half2* hh = reinterpret_cast<const __half2*>(&h);
        const __half2* ll = reinterpret_cast<const __half2*>(&l);
        float4 o0, o1;
        float2 x0 = __half22float2(hh[0]), y0 = __half22float2(ll[0]);
        float2 x1 = __half22float2(hh[1]), y1 = __half22float2(ll[1]);
        o0 = make_float4(x0.x + y0.x, x0.y + y0.y, x1.x + y1.x, x1.y + y1.y);
        x0 = __half22float2(hh[2]); y0 = __half22float2(ll[2]);
        x1 = __half22float2(hh[3]); y1 = __half22float2(ll[3]);
        o1 = make_float4(x0.x + y0.x, x0.y + y0.y, x1.x + y1.x, x1.y + y1.y);
        float4* dst = reinterpret_cast<float4*>(gal + (size_t)i * PG + (size_t)ch * 8);
        dst[0] = o0;
        dst[1] = o1;
      }
      __syncthreads();
      // ---- one warp per candidate position
      for (int e = warp; e < nl; e += kRefWarps) {
        const uint2 en = list[e];
        const int j = en.x & 0xff, i = en.x >> 8, y = en.y & 0xffff, x = en.y >> 16;
        const int u_lo = max(0, a - y), u_hi = min(p.Hb, p.Hp + a - y);  // template rows that meet the map
        const float* T = tpl + (size_t)j * p.Kpad;
        const float* Gs = gal + (size_t)i * PG + (y - a) * p.WP + (x - b);
        float part = 0.0f;
        for (int v0 = 0; v0 < p.rowk; v0 += vl) {
          const int v = v0 + lv, gx = x + v - b;
          if (v < p.rowk && gx >= 0 && gx < p.Wp) {
            const float* tp = T + (u_lo + lr) * p.rowk + v;
            const float* gp = Gs + (u_lo + lr) * p.WP + v;
#pragma unroll 4
            for (int u = u_lo + lr; u < u_hi; u += rpp, tp += rpp * p.rowk, gp += rpp * p.WP) part = fmaf(*tp, *gp, part);
          }
        }
        part = warp_sum(part);
        if (lane == 0) {
          const int n = n0 + j, g = gt0 + i;
          const float* table = p.rnorm_tab ? p.rnorm_tab[n >> 4] : p.rnorm;
          acc[e] = fmaf(part, __ldg(table + ((size_t)g * p.C + c) * M + y * p.Wp + x), acc[e]);
        }
      }
      __syncthreads();
    }
    for (int e = tid; e < nl; e += kRefThreads) {
      const int n = n0 + (list[e].x & 0xff), g = gt0 + (list[e].x >> 8);
      atomic_max_nonneg(&p.scores[(size_t)p.col2probe[n] * p.score_ld + p.g0 + g], acc[e] * p.inv_scale);
    }
    __syncthreads();
  }
}

}  // namespace sir

using namespace sir;

extern "C" int sir_ncc_refine(const uint16_t* d_ghi, const uint16_t* d_glo, const float* d_rnorm, const float* const* d_rnorm_tab, int G, int C,
                              int Hp, int Wp, const uint16_t* d_thi, const uint16_t* d_tlo, int ncols, int ncols_alloc, int Hb, int Wb,
                              const int32_t* d_col2probe, const float* d_approx, float* d_scores, int score_ld, int g0, float tau_rel,
                              float tau_abs, const void* d_rec, unsigned long long* d_stats, void* stream) {
  SIR_CHECK_ARG((d_rnorm != nullptr) != (d_rnorm_tab != nullptr), "sir_ncc_refine: give d_rnorm or d_rnorm_tab, not both");
  SIR_CHECK_ARG(d_ghi && d_glo && d_thi && d_tlo && d_col2probe && d_approx && d_scores && d_rec, "sir_ncc_refine: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0 && Hb > 0 && Wb > 0 && ncols > 0 && ncols <= ncols_alloc, "sir_ncc_refine: bad shape");
  SIR_CHECK_ARG(score_ld >= g0 + G, "sir_ncc_refine: score row (%d) shorter than g0+G (%d)", score_ld, g0 + G);
  SIR_CHECK_ARG(Hp < 65536 && Wp < 65536, "sir_ncc_refine: map too large");
  RefineParams p{};
  p.ghi = (const __half*)d_ghi;
  p.glo = (const __half*)d_glo;
  p.rnorm = d_rnorm;
  p.rnorm_tab = d_rnorm_tab;
  p.thi = (const __half*)d_thi;
  p.tlo = (const __half*)d_tlo;
  p.col2probe = d_col2probe;
  p.approx = d_approx;
  p.scores = d_scores;
  p.rec = (const uint2*)d_rec;
  p.stats = d_stats;
  p.G = G; p.C = C; p.Hp = Hp; p.Wp = Wp; p.WP = gal_pitch(Wp); p.Hb = Hb; p.Wb = Wb;
  p.rowk = tpl_row_taps(Wb, 8);
  p.Kpad = tpl_kpad(Hb, Wb);
  p.ncols = ncols; p.ncols_alloc = ncols_alloc;
  p.npx = ceil_div(Wp, 8);
  p.NP = ceil_div(Hp, 16) * p.npx;
  p.score_ld = score_ld; p.g0 = g0;
  p.tau_rel = tau_rel; p.tau_abs = tau_abs;
  p.inv_scale = 1.0f / ((float)C * (float)(1 << kTemplateScaleLog2));
  p.cap = 1024;
  // tile: as many (column, gallery) pairs per CTA as the staged planes allow (every operand byte is then read once
  // per tile); columns get the larger share because consecutive CTAs walk the gallery tiles of one column tile
  const size_t budget = 200 * 1024, fixed = (size_t)p.cap * 12 + 4 * (32 + 32 + kRefWarps) + 64;
  const size_t tbytes = (size_t)p.Kpad * 4, gbytes = (size_t)Hp * p.WP * 4;
  int best_tn = 0, best_tg = 0;
  for (int tn = 32; tn >= 1; tn >>= 1)
    for (int tg = 32; tg >= 1; tg >>= 1)
      if (fixed + tn * tbytes + tg * gbytes <= budget && (tn * tg > best_tn * best_tg || (tn * tg == best_tn * best_tg && tn > best_tn))) {
        best_tn = tn;
        best_tg = tg;
      }
  SIR_CHECK_ARG(best_tn > 0, "sir_ncc_refine: template %dx%d / map %dx%d do not fit shared memory", Hb, Wb, Hp, Wp);
  p.TN = std::min(best_tn, round_up(ncols, 1));
  p.TG = best_tg;
  p.tiles_g = ceil_div(G, p.TG);
  const size_t smem = fixed + p.TN * tbytes + p.TG * gbytes;
  int dev = 0;
  SIR_CUDA(cudaGetDevice(&dev));
  static thread_local size_t configured[16] = {0};
  if (smem > 48 * 1024 && (dev >= 16 || smem > configured[dev])) {
    SIR_CUDA(cudaFuncSetAttribute(ncc_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < 16) configured[dev] = smem;
  }
  const long long blocks = (long long)ceil_div(ncols, p.TN) * p.tiles_g;
  SIR_CHECK_ARG(blocks < (1ll << 31), "sir_ncc_refine: too many tiles");
  ncc_refine_kernel<<<(unsigned)blocks, kRefThreads, smem, (cudaStream_t)stream>>>(p);
  SIR_LAUNCH_CHECK("ncc_refine_kernel");
  return SIR_OK;
}
