// K7r: exact re-evaluation of the screening candidates (second half of SIR_PREC_FP16_REFINE).
//
// sir_ncc_screen runs the whole probe x gallery correlation (similarity.py:53-55,100-108,357-367) with plain fp16
// operands: one MMA per K step, but only ~2e-4 relative.  What the reference wants per pair is ONE number,
// max over positions and variants (similarity.py:106-108,365-367), so the exact arithmetic is only needed where the
// maximum can be: every (column, gallery, patch) record whose screened maximum lies within the candidate margin of
// the pair's screened maximum names the rows that are within the margin of it, and this kernel evaluates
//
//     s(n, g, y, x) = (1/C) sum_c rnorm_c[g][y,x] * sum_{u,v} t32[n][c][u,v] * g32[g][c][y+u-a][x+v-b]
//
// for exactly those positions in float32 (t32, g32: the float32 values the fp16 operands were rounded from) and
// max-reduces the result into d_scores.  The margin covers twice the screening error, so the position of the true
// maximum is always among the candidates; if it ever were not, the result would still be an exactly evaluated
// correlation value within 2 * (screening error) of the true maximum.
//
// Every position needs a template plane and a gallery plane per channel (~6 KB for ~700 multiply-adds), so what
// the kernel has to organise is operand reuse.  One CTA owns TN columns x TGB gallery prints.  Per channel the
// template planes of its columns that have work are staged ONCE (double buffered over channels) and the gallery
// planes stream past them in sub-chunks of TGS prints through a small ring -- plain bulk copies (cp.async.bulk, one
// per plane, issued by the lanes of a producer warp), full/empty mbarriers, no block barrier in the loop.  With
// several variants per probe only about one (column, gallery) cell in ten has a candidate, so a gallery plane is
// shared by ~TN/10 positions and a template plane by all its column's positions; with one variant every cell has
// one.  Sixteen consumer warps take one position at a time, lanes spread over (template row, tap); template rows
// that only meet the "same"-mode zero padding are skipped.  The work list is built without atomics (block scan over
// per-thread counts, in gallery-major order so that a sub-chunk's positions are contiguous): every run evaluates the
// same candidates in the same order.
#include <algorithm>
#include <cstdlib>

#include "sir_common.cuh"
#include "sir_ptx.cuh"

namespace sir {

// consumer warps (one candidate position at a time each) + one producer warp issuing the bulk copies: 16 consumers when
// few (column, gallery) cells have a candidate (several variants per probe: a step holds about one position per warp and
// more warps only add barrier traffic), 31 when every cell has one (one variant per probe: 79 -> 65 ms per 2.08 M positions)
constexpr int kRefMaxThreads = 1024;
constexpr int kRefMaxStages = 4;                   // gallery sub-chunk ring (depth chosen by the host)
constexpr int kRefMaxSub = 128;                    // sub-chunks per tile
constexpr int kRefMaxTgb = 2048;                   // gallery prints per tile

struct RefineParams {
  const float* g32;
  const float* t32;
  const float* rnorm;
  const float* const* rnorm_tab;
  const int32_t* col2probe;
  const float* approx;
  float* scores;
  const uint2* rec;
  unsigned long long* stats;  // optional: [0] positions evaluated, [1] records that listed more than 3 rows, [2] tiles with work
  int G, C, Hp, Wp, WP, Hb, Wb, rowk, Kpad, ncols, ncols_alloc, npx, NP, score_ld, g0;
  int TN, TGS, TGB, NS, ST, cap, tiles_g;
  int vl, ru, rv;  // lane mapping: vl lanes along a template row, the 32 / vl lane groups split ru ways over rows, rv ways over row chunks
  float tau_rel, tau_abs, inv_scale;
};

template <int kRefThreads>
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
  for (int i = 0; i < kRefThreads / 32; ++i) {
    if (i < wid) base += scratch[i];
    sum += scratch[i];
  }
  __syncthreads();
  *total = sum;
  return base + inc - v;
}

// TS / GS: the lane's row stride in the template / gallery plane (ru * rowk, ru * WP floats) as compile-time constants for
// the two shapes the benchmarks live on (0 = run-time strides): the dense case is issue bound at ~5 instructions per
// multiply-add with run-time strides, constant strides turn the address arithmetic into immediate offsets.
template <int kRefWarps, int TS, int GS>
__global__ void __launch_bounds__(32 * (kRefWarps + 1)) ncc_refine_kernel(const RefineParams p) {
  constexpr int kRefThreads = 32 * (kRefWarps + 1);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int PG = p.Hp * p.WP;                                      // cells of one packed gallery plane
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);           // full_g[ST], empty_g[ST], empty_t[2], full_t[2]
  float* tplbuf = reinterpret_cast<float*>(smem_raw + 128);         // [2][TN][Kpad]   templates of channel c in buffer c & 1
  float* galbuf = tplbuf + 2 * (size_t)p.TN * p.Kpad;               // [ST][TGS][PG]
  uint2* list = reinterpret_cast<uint2*>(galbuf + (size_t)p.ST * p.TGS * PG);  // [cap] (j | i << 8, y | x << 16)
  float* acc = reinterpret_cast<float*>(list + p.cap);              // [cap]
  int* flags = reinterpret_cast<int*>(acc + p.cap);                 // [TN + TGB]
  int* seg = flags + p.TN + p.TGB;                                  // [NS + 1] first list entry of every sub-chunk
  int* scratch = seg + p.NS + 1;                                    // [warps]

  const int kRefStages = p.ST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = (blockIdx.x / p.tiles_g) * p.TN, gt0 = (blockIdx.x % p.tiles_g) * p.TGB;
  const int M = p.Hp * p.Wp;
  const int items = p.TGB * p.TN * p.NP;  // (gallery i, column j, patch), gallery major
  const int ipt = (items + kRefThreads - 1) / kRefThreads, item0 = min(items, tid * ipt), item1 = min(items, item0 + ipt);

  // candidate rows of one record: 0 when the record cannot hold the pair's maximum
  auto expand = [&](int item, uint32_t* info_out, int* py_out, int* px_out, int* j_out, int* i_out) -> int {
    const int pidx = item % p.NP, pair = item / p.NP;
    const int j = pair % p.TN, i = pair / p.TN;
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
  for (int item = item0; item < item1; ++item) {
    uint32_t info;
    int py, px, j, i;
    mine += expand(item, &info, &py, &px, &j, &i);
  }
  int total = 0;
  const int base = block_exclusive_scan<kRefThreads>(mine, scratch, &total);
  if (total == 0) return;
  if (tid == 0) {
    for (int st = 0; st < kRefStages; ++st) {
      ptx::mbar_init(ptx::smem_u32(bars + st), 1);                      // full: the producer's expect_tx arrival + the bytes
      ptx::mbar_init(ptx::smem_u32(bars + kRefStages + st), kRefWarps);  // empty: one arrival per consumer warp
    }
    ptx::mbar_init(ptx::smem_u32(bars + 2 * kRefStages), kRefWarps);     // template buffers free again
    ptx::mbar_init(ptx::smem_u32(bars + 2 * kRefStages + 1), kRefWarps);
    ptx::mbar_init(ptx::smem_u32(bars + 2 * kRefStages + 2), 1);             // template buffers filled (own barriers: the planes of
    ptx::mbar_init(ptx::smem_u32(bars + 2 * kRefStages + 3), 1);             // channel c + 1 travel while channel c is being consumed)
    ptx::fence_barrier_init();
    if (p.stats) {
      atomicAdd(p.stats + 0, (unsigned long long)total);
      atomicAdd(p.stats + 2, 1ull);
    }
  }

  const int a = p.Hb / 2, b = p.Wb / 2;
  // lanes over (template row, tap): vl lanes along a row; the 32 / vl lane groups are dealt over rows (ru) and over the
  // vl-tap chunks of a row (rv), whichever split gives the longest inner loop for this shape (chosen by the host)
  const int vl = p.vl, lv = lane % vl, grp = lane / vl, gu = grp % p.ru, gv = grp / p.ru, nchunk = p.rowk / vl;
  const uint32_t tbytes = (uint32_t)p.Kpad * 4u, gbytes = (uint32_t)PG * 4u;
  uint32_t gst = 0, gpar = 0;  // gallery ring: current stage and the parity of its barriers (both roles step them alike)
  uint32_t cit = 0;  // channels so far        (template buffer = cit & 1, barrier parity = (cit >> 1) & 1)

  for (int r0 = 0; r0 < total; r0 += p.cap) {
    const int nl = min(p.cap, total - r0);
    // ---- this round's slice of the work list, in scan (= gallery-major) order
    if (base < r0 + p.cap && base + mine > r0) {
      int o = base;
      for (int item = item0; item < item1; ++item) {
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
    for (int k = tid; k < p.TN + p.TGB; k += kRefThreads) flags[k] = 0;
    for (int k = tid; k <= p.NS; k += kRefThreads) seg[k] = -1;
    __syncthreads();
    for (int e = tid; e < nl; e += kRefThreads) {
      acc[e] = 0.0f;
      const int i = list[e].x >> 8;
      flags[list[e].x & 0xff] = 1;
      flags[p.TN + i] = 1;
      const int s = i / p.TGS;
      if (e == 0 || (int)(list[e - 1].x >> 8) / p.TGS != s) seg[s] = e;  // the list is gallery major: first entry of sub-chunk s
    }
    __syncthreads();
    if (tid == 0) {  // empty sub-chunks start where the next one starts
      seg[p.NS] = nl;
      for (int s = p.NS - 1; s >= 0; --s)
        if (seg[s] < 0) seg[s] = seg[s + 1];
    }
    __syncthreads();

    if (warp == kRefWarps) {
      // ---- producer warp: per channel the flagged template planes once, then the flagged gallery planes sub-chunk by sub-chunk.
      // The template planes of channel c + 1 are requested at the START of channel c (their buffer was released at the end
      // of channel c - 1), so that burst (TN planes) never sits in front of a gallery step.
      uint32_t tpl_bytes = 0;
      for (int j = 0; j < p.TN; ++j) tpl_bytes += flags[j] ? tbytes : 0u;
      auto request_templates = [&](int c, uint32_t cc) {  // cc: channels so far, this one included in the count of its buffer
        const uint32_t buf = cc & 1u, full_t = ptx::smem_u32(bars + 2 * kRefStages + 2 + buf);
        ptx::mbar_wait(ptx::smem_u32(bars + 2 * kRefStages + buf), ((cc >> 1) & 1u) ^ 1u);  // the consumers are done with this buffer
        if (lane == 0) {
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive_expect_tx(full_t, tpl_bytes);
        }
        __syncwarp();
        float* tdst = tplbuf + buf * (size_t)p.TN * p.Kpad;
        for (int j = lane; j < p.TN; j += 32)
          if (flags[j]) ptx::bulk_load(ptx::smem_u32(tdst + (size_t)j * p.Kpad), p.t32 + ((size_t)c * p.ncols_alloc + n0 + j) * p.Kpad, tbytes, full_t);
      };
      request_templates(0, cit);
      for (int c = 0; c < p.C; ++c) {
        if (c + 1 < p.C) request_templates(c + 1, cit + (uint32_t)c + 1u);
        for (int s = 0; s < p.NS; ++s) {
          if (seg[s + 1] == seg[s]) continue;  // no position in this sub-chunk
          const uint32_t full = ptx::smem_u32(bars + gst);
          ptx::mbar_wait(ptx::smem_u32(bars + kRefStages + gst), gpar ^ 1u);  // consumers are done with this gallery stage
          float* gdst = galbuf + (size_t)gst * p.TGS * PG;
          uint32_t bytes = 0u;
          for (int k = 0; k < p.TGS; ++k) bytes += flags[p.TN + s * p.TGS + k] ? gbytes : 0u;
          if (lane == 0) {
            ptx::fence_proxy_async_smem();  // the consumers' generic reads of these buffers precede the async writes
            ptx::mbar_arrive_expect_tx(full, bytes);
          }
          __syncwarp();
          for (int k = lane; k < p.TGS; k += 32) {
            const int i = s * p.TGS + k;
            if (flags[p.TN + i]) ptx::bulk_load(ptx::smem_u32(gdst + (size_t)k * PG), p.g32 + ((size_t)(gt0 + i) * p.C + c) * PG, gbytes, full);
          }
          if (++gst == (uint32_t)kRefStages) {
            gst = 0;
            gpar ^= 1u;
          }
        }
      }
      cit += (uint32_t)p.C;
    } else {
      // ---- consumer warps: one candidate position at a time
      // window norm of list entry e in channel c (0 past the end of the sub-chunk)
      auto norm_of = [&](int e, int e1, int c) -> float {
        if (e >= e1) return 0.0f;
        const uint2 en = list[e];
        const float* table = p.rnorm_tab ? p.rnorm_tab[(n0 + (int)(en.x & 0xff)) / kNormChunkCols] : p.rnorm;
        return __ldg(table + ((size_t)(gt0 + (int)(en.x >> 8)) * p.C + c) * M + (en.y & 0xffff) * p.Wp + (en.y >> 16));
      };
      // The norms of a step are gathered ONE STEP AHEAD (lane t holds the norm of the warp's t-th position of the step): a
      // step holds about one position per warp, so a gather issued inside the step would put a global-load latency on
      // every one of the C * NS steps.
      auto next_step = [&](int& c, int& s) {  // the next (channel, sub-chunk) with positions; c == p.C when there is none
        for (++s;; ++s) {
          if (s >= p.NS) {
            s = 0;
            if (++c >= p.C) return;
          }
          if (seg[s + 1] != seg[s]) return;
        }
      };
      int c = 0, s = -1;
      next_step(c, s);
      float rnv = c < p.C ? norm_of(seg[s] + warp + kRefWarps * lane, seg[s + 1], c) : 0.0f;
      int c_seen = 0;  // channels whose template buffer this warp has released
      int c_ready = -1;  // channel whose template planes this warp has waited for
      while (c < p.C) {
        int cn = c, sn = s;
        next_step(cn, sn);
        const float rnv_next = cn < p.C ? norm_of(seg[sn] + warp + kRefWarps * lane, seg[sn + 1], cn) : 0.0f;
        const uint32_t cc = cit + (uint32_t)c;
        const float* tpl = tplbuf + (cc & 1u) * (size_t)p.TN * p.Kpad;
        const int e0 = seg[s], e1 = seg[s + 1];
        if (c != c_ready) {  // first step of a channel: its template planes have landed
          ptx::mbar_wait(ptx::smem_u32(bars + 2 * kRefStages + 2 + (cc & 1u)), (cc >> 1) & 1u);
          c_ready = c;
        }
        const uint32_t st = gst;
        ptx::mbar_wait(ptx::smem_u32(bars + st), gpar);
        const float* gal = galbuf + (size_t)st * p.TGS * PG;
        int t = 0;
        for (int e = e0 + warp; e < e1; e += kRefWarps, ++t) {
          const uint2 en = list[e];
          const int j = en.x & 0xff, i = (int)(en.x >> 8) - s * p.TGS, y = en.y & 0xffff, x = en.y >> 16;
          if ((t & 31) == 0 && t) rnv = norm_of(e + kRefWarps * lane, e1, c);  // more than 32 positions of this warp in one sub-chunk
          const int u_lo = max(0, a - y), u_hi = min(p.Hb, p.Hp + a - y);  // template rows that meet the map
          const float* T = tpl + (size_t)j * p.Kpad;
          const float* Gs = gal + (size_t)i * PG + (y - a) * p.WP + (x - b);
          float part0 = 0.0f, part1 = 0.0f;
          const int ts = TS ? TS : p.ru * p.rowk, gs = GS ? GS : p.ru * p.WP;
          for (int ch = gv; ch < nchunk; ch += p.rv) {
            const int v = ch * vl + lv, gx = x + v - b;
            if (gx >= 0 && gx < p.Wp) {
              const float* tp = T + (u_lo + gu) * p.rowk + v;
              const float* gp = Gs + (u_lo + gu) * p.WP + v;
              int u = u_lo + gu;
              for (; u + 3 * p.ru < u_hi; u += 4 * p.ru, tp += 4 * ts, gp += 4 * gs) {
                part0 = fmaf(tp[0], gp[0], part0);
                part1 = fmaf(tp[ts], gp[gs], part1);
                part0 = fmaf(tp[2 * ts], gp[2 * gs], part0);
                part1 = fmaf(tp[3 * ts], gp[3 * gs], part1);
              }
              for (; u < u_hi; u += p.ru, tp += ts, gp += gs) part0 = fmaf(*tp, *gp, part0);
            }
          }
          const float part = warp_sum(part0 + part1);
          const float rn = __shfl_sync(0xffffffffu, rnv, t & 31);
          if (lane == 0) acc[e] = fmaf(part, rn, acc[e]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(bars + kRefStages + st));
        if (++gst == (uint32_t)kRefStages) {
          gst = 0;
          gpar ^= 1u;
        }
        if (cn != c) {  // last sub-chunk of channel c: its templates are no longer needed (nor those of skipped-over channels)
          __syncwarp();
          for (; c_seen < min(cn, p.C); ++c_seen)
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(bars + 2 * kRefStages + ((cit + (uint32_t)c_seen) & 1u)));
        }
        c = cn;
        s = sn;
        rnv = rnv_next;
      }
      cit += (uint32_t)p.C;
    }
    __syncthreads();
    for (int e = tid; e < nl; e += kRefThreads) {
      const int n = n0 + (list[e].x & 0xff), g = gt0 + (list[e].x >> 8);
      atomic_max_nonneg(&p.scores[(size_t)p.col2probe[n] * p.score_ld + p.g0 + g], acc[e] * p.inv_scale);
    }
    __syncthreads();
  }
}

}  // namespace sir

using namespace sir;

extern "C" int sir_ncc_refine(const float* d_g32, const float* d_rnorm, const float* const* d_rnorm_tab, int G, int C, int Hp, int Wp,
                              const float* d_t32p, int ncols, int ncols_alloc, int Hb, int Wb, const int32_t* d_col2probe,
                              const float* d_approx, float* d_scores, int score_ld, int g0, float tau_rel, float tau_abs, const void* d_rec,
                              int variants_hint, unsigned long long* d_stats, void* stream) {
  SIR_CHECK_ARG((d_rnorm != nullptr) != (d_rnorm_tab != nullptr), "sir_ncc_refine: give d_rnorm or d_rnorm_tab, not both");
  SIR_CHECK_ARG(d_g32 && d_t32p && d_col2probe && d_approx && d_scores && d_rec, "sir_ncc_refine: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0 && Hb > 0 && Wb > 0 && ncols > 0 && ncols <= ncols_alloc, "sir_ncc_refine: bad shape");
  SIR_CHECK_ARG(score_ld >= g0 + G, "sir_ncc_refine: score row (%d) shorter than g0+G (%d)", score_ld, g0 + G);
  SIR_CHECK_ARG(Hp < 65536 && Wp < 65536, "sir_ncc_refine: map too large");
  SIR_CHECK_ARG((reinterpret_cast<uintptr_t>(d_g32) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_t32p) & 15) == 0,
                "sir_ncc_refine: operands must be 16-byte aligned");
  RefineParams p{};
  p.g32 = d_g32;
  p.t32 = d_t32p;
  p.rnorm = d_rnorm;
  p.rnorm_tab = d_rnorm_tab;
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
  // Tile: TN resident template columns (two channel buffers) + a ring of kRefStages x TGS gallery planes.  More columns =
  // every gallery plane serves more positions; TGS only has to keep a step long enough to hide the copies.
  const size_t budget = 220 * 1024;
  const size_t tbytes = (size_t)p.Kpad * 4, gbytes = (size_t)Hp * p.WP * 4;
  const size_t fixed = 128 + (size_t)p.cap * 12 + 4 * (size_t)(32 + kRefMaxTgb + kRefMaxSub + 1 + kRefMaxThreads / 32) + 64;
  p.TN = 0;
  p.ST = 2;
  // time follows the number of (channel, sub-chunk) steps: 16 prints per sub-chunk when that still leaves 16 resident columns
  // (configs[1]: 45.5 -> 41.6 ms), else 8 and down
  int tgs_first = 16;
  if (const char* env = getenv("SIR_REFINE_STAGES")) p.ST = std::max(2, std::min(kRefMaxStages, atoi(env)));
  if (const char* env = getenv("SIR_REFINE_TGS")) tgs_first = std::max(1, std::min(32, atoi(env)));
  const int kRefStages = p.ST;
  for (int tgs = tgs_first; tgs >= 1 && p.TN == 0; tgs >>= 1) {
    if (fixed + kRefStages * tgs * gbytes + 2 * tbytes > budget) continue;
    int tn = (int)std::min<size_t>(32, (budget - fixed - kRefStages * tgs * gbytes) / (2 * tbytes));
    if (const char* env = getenv("SIR_REFINE_TN")) tn = std::max(1, std::min(tn, atoi(env)));
    if (tn >= std::min(tgs == 16 ? 16 : 8, ncols) || tgs == 1) {
      p.TN = tn >= 8 ? tn / 8 * 8 : tn;
      p.TGS = tgs;
    }
  }
  SIR_CHECK_ARG(p.TN > 0, "sir_ncc_refine: template %dx%d / map %dx%d do not fit shared memory", Hb, Wb, Hp, Wp);
  p.TN = std::min(p.TN, ncols);
  // gallery prints per tile: as many as keep the expected work list (about 1.3 positions per (probe, gallery) pair, i.e.
  // 1.3 / variants per (column, gallery) cell) inside one round of `cap` entries
  const int variants = std::max(1, variants_hint);
  long long tgb = (long long)p.cap * variants * 10 / (13LL * p.TN);
  tgb = std::min<long long>({tgb, (long long)p.TGS * kRefMaxSub, (long long)kRefMaxTgb, (long long)round_up(G, p.TGS)});
  tgb = std::max<long long>(tgb, p.TGS);
  p.TGB = (int)(tgb / p.TGS * p.TGS);
  p.NS = p.TGB / p.TGS;
  p.tiles_g = ceil_div(G, p.TGB);
  // lane mapping: the inner loop walks template rows; estimated instructions per (position, channel) for a split of the lane
  // groups over rows (ru) and row chunks (rv): chunk rounds x (set-up + rows per lane x ~3.3)
  p.vl = (p.rowk % 32 == 0) ? 32 : (p.rowk % 16 == 0) ? 16 : 8;
  {
    const int groups = 32 / p.vl, nchunk = p.rowk / p.vl;
    double best = 1e30;
    for (int ru = 1; ru <= groups; ru <<= 1) {
      const int rv = groups / ru;
      const double cost = ceil_div(nchunk, rv) * (15.0 + ceil_div(Hb, ru) * 3.3);
      if (cost < best) {
        best = cost;
        p.ru = ru;
        p.rv = rv;
      }
    }
  }
  const size_t smem = 128 + 2 * p.TN * tbytes + (size_t)kRefStages * p.TGS * gbytes + (size_t)p.cap * 12 +
                      4 * (size_t)(p.TN + p.TGB + p.NS + 1 + kRefMaxThreads / 32) + 64;
  SIR_CHECK_ARG(smem <= 227 * 1024, "sir_ncc_refine: shared-memory plan overflow (%zu bytes)", smem);
  const long long blocks = (long long)ceil_div(ncols, p.TN) * p.tiles_g;
  SIR_CHECK_ARG(blocks < (1ll << 31), "sir_ncc_refine: too many tiles");
  const int ts = p.ru * p.rowk, gs = p.ru * p.WP;
  auto launch = [&](auto kernel, int warps) -> int {
    SIR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    kernel<<<(unsigned)blocks, 32 * (warps + 1), smem, (cudaStream_t)stream>>>(p);
    return SIR_OK;
  };
  int rc;
  if (variants <= 2) {
    if (ts == 56 && gs == 56) rc = launch(ncc_refine_kernel<31, 56, 56>, 31);
    else if (ts == 32 && gs == 32) rc = launch(ncc_refine_kernel<31, 32, 32>, 31);
    else rc = launch(ncc_refine_kernel<31, 0, 0>, 31);
  } else {
    if (ts == 56 && gs == 56) rc = launch(ncc_refine_kernel<16, 56, 56>, 16);
    else if (ts == 32 && gs == 32) rc = launch(ncc_refine_kernel<16, 32, 32>, 16);
    else rc = launch(ncc_refine_kernel<16, 0, 0>, 16);
  }
  if (rc) return rc;
  SIR_LAUNCH_CHECK("ncc_refine_kernel");
  return SIR_OK;
}
