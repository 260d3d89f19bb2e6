// K7: probe x gallery normalised cross-correlation on tcgen05 tensor cores (sm_100a).
//
// For one template shape (Hm x Wm) the reference's Q*G*V get_similarity calls
// (similarity.py:75-108, 357-367) become, per channel c, the contraction
//
//     num_c[(g,y,x), n] = sum_{u,v} G_c[g][y+u-a][x+v-b] * T_c[n][u][v]        (similarity.py:53-55)
//
// with M = gallery positions, N = packed probe variants (columns), K = template taps.  The per
// channel, per position normalisation 1/sqrt(D_c[y,x]) (similarity.py:57-68; 1/sqrt(E) is already
// folded into T) is a ROW scale of the accumulator, so channels are separate K segments: the MMA
// warp accumulates one channel into TMEM, the epilogue warps fold it into a register-resident
// running sum  total += rnorm_c[row] * acc  while the next channel is being multiplied into the
// other TMEM buffer.  After the last channel the epilogue takes the max over the tile's valid
// positions and atomically maxes it into scores[probe][gallery]: the correlation surface never
// reaches HBM, and neither does an im2col matrix.
//
// Operand A (gallery, Toeplitz): the im2col matrix A[(y,x),(u,v)] = g[y+u-a][x+v-b] is never
// materialised.  Generator warps expand the compact fp16 channel into "shifted entries"
//     E[r][i] = 8 consecutive cells g[row r][col i .. i+7]      (16 bytes each, zero filled)
// and a NO-SWIZZLE K-major UMMA descriptor with  SBO = one E row,  LBO = 8 entries  addresses
// them so that descriptor row (ml, mh) and K chunk (u, ch) land on entry (mh+u, ml+8ch): core
// matrices overlap in shared memory (8x replication instead of Hm*Wm x), which the tensor core
// does not mind because it only reads.  A tile's 128 rows are a 16 (y) x 8 (x) patch of positions.
//
// Operand B (templates): dense [column][K] fp16, streamed by TMA (64B swizzle, 32 taps per stage)
// through a ring of shared-memory stages.
//
// Precision: fp16 operands carry an 11-bit significand; the 1e-4 relative score tolerance needs
// more, so operands are split hi + lo and each K step issues three MMAs (hi*hi, lo*hi, hi*lo)
// into the same fp32 accumulator (SIR_PREC_FP16X3).  SIR_PREC_FP16X1 issues only hi*hi.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over work units = (column tile, gallery, patch)):
//   warp 0      TMA producer for B
//   warp 1      TMEM owner + MMA issuer (one elected lane)
//   warps 2-3   generators: stage the channel rows, build E, fence.proxy.async, signal
//   warps 4-11  epilogue: tcgen05.ld, row-scale FMA into 128 registers/thread, final max
#include <cuda.h>

#include <cstdlib>

#include "sir_common.cuh"
#include "sir_ptx.cuh"

namespace sir {

constexpr int kTcThreads = 384;
constexpr int kTileM = 128;      // 16 x 8 positions
constexpr int kTileN = 256;      // packed columns per tile
// taps per B stage = 16 * SPS (K16 steps per stage, a template parameter of the kernel): 32 taps (one 64B-swizzle row of
// fp16) for the split-precision modes, 64 taps (128B swizzle) for the one-pass fp16 modes -- there a stage of 32 taps is
// only two MMAs and the single issuing thread, not the tensor pipe, paced the kernel (ncu: pipe 55% busy, the issue
// loop ~100 instructions per stage)
constexpr int kGenThreads = 64;
constexpr int kEpiWarps = 8;
constexpr int kMaxBStages = 16;
constexpr uint32_t kTmemCols = 512;
__host__ __device__ constexpr uint32_t b_half_bytes(int sps) { return kTileN * 16 * sps * 2; }  // one operand half of a stage: 16 / 32 KB

struct TcParams {
  const __half* ghi;
  const __half* glo;
  const float* rnorm;
  // multi-shape column tiles: one window-norm table per kNormChunkCols-column chunk (32 per 256-column tile), NULL
  // when every column of the launch has the same true template shape
  const float* const* rnorm_tab;
  const int32_t* col2probe;
  float* scores;
  int score_ld, g0;
  int G, C, Hp, Wp, Hm, Wm;
  int Gp;          // G rounded up to a multiple of the CTA group size (the extra unit repeats gallery G-1)
  int ncols;
  int nkc;         // 8-tap chunks per template row
  int nsteps;      // K16 steps per channel = ceil(Hm*nkc/2)
  int nkstages;    // B stages per channel = ceil(Kpad / (16 * sps))
  int sps;         // K16 steps per B stage: 2 or 4 (== the kernel's SPS)
  int npy, npx;    // 16x8 patches over the gallery position grid
  int ntiles_n;
  long long nunits;
  int seg_stages;  // B stages per E segment
  int nseg;
  int nbstages;    // B ring depth
  int passes;      // 3: fp16 hi/lo split, 3 fp16 MMAs per K16 step; 1: hi only; 2: hi*hi in fp16 + the two
                   //    correction products (lo*hi, hi*lo) in fp8 e4m3 at twice the MMA rate ("fp8c")
  int nhalf;       // operand arrays per E buffer / staging buffer: 1, 2 or 3
  uint32_t gs8_bytes;  // fp8c: bytes of one staged 1-byte window
  float out_scale; // 1 / (C * 2^kTemplateScaleLog2)
  // shared memory carve-up (byte offsets from the 1024-aligned base)
  uint32_t off_b, off_e, off_gs, off_cm, off_bar;
  uint32_t e_half_bytes;   // one E buffer, one half (hi or lo)
  uint32_t gs_half_elems;  // staging cells per half per buffer >= gs_rows * (Pe + 16)
  int gs_rows;             // rows of the staging TMA box (max rows any segment needs)
  int gs_bufs;             // 2: stage one segment ahead; 1: no lookahead (very wide templates)
  // screening mode (SIR_PREC_FP16_REFINE): `scores` is the APPROXIMATE pair maximum and every work unit also leaves
  // a record (tile maximum, rows within the candidate margin of it) for sir_ncc_refine; NULL otherwise
  uint2* rec;
  float tau_rel, tau_abs;  // candidate margin tau(m) = tau_rel * |m| + tau_abs, in accumulator units
  uint32_t off_mask;       // [4][256] candidate-row masks next to the column maxima
  // multi-shape column tiles: rows [x, y) of the bucket's K layout that hold a non-zero tap in ANY column of tile nt (the
  // templates sit anchor on anchor, so the range always contains the anchor row Hm/2); NULL = all Hm rows
  const int2* tile_rows;
};

struct Seg {
  int st0, st1;   // B stage range
  int u_first;    // first template row touched
  int rows;       // E rows to build
};

// K range of one work unit.  Template rows whose 16 image rows all fall outside the gallery map
// multiply zeros only ("same"-mode padding): they are skipped by every role, stage-aligned.
struct KRange {
  int ks_lo, ks_hi;  // K16 steps with a non-zero A operand
  int st_lo, st_hi;  // B stages covering them
  int nseg;
};

__device__ __forceinline__ KRange k_range(const TcParams& p, int py, int nt) {
  const int a = p.Hm / 2;
  int u_lo = max(0, a - 16 * py - 15);
  int u_hi = min(p.Hm, p.Hp + a - 16 * py);
  if (p.tile_rows) {  // rows that are zero in every column of this tile are skipped like the "same" padding
    const int2 tr = p.tile_rows[nt];
    u_lo = max(u_lo, tr.x);
    u_hi = min(u_hi, tr.y);
  }
  KRange r;
  r.ks_lo = (u_lo * p.nkc) / 2;
  r.ks_hi = min(p.nsteps, (u_hi * p.nkc + 1) / 2);
  r.st_lo = r.ks_lo / p.sps;
  r.st_hi = (r.ks_hi + p.sps - 1) / p.sps;
  r.nseg = (r.st_hi - r.st_lo + p.seg_stages - 1) / p.seg_stages;
  return r;
}

// Columns the MMA of column tile `nt` really needs (the last tile of a block is usually narrower):
// a multiple of 32 so that a CTA pair splits it into two 16-row aligned halves.
__device__ __forceinline__ int tile_cols(const TcParams& p, int nt) { return min(kTileN, round_up(p.ncols - nt * kTileN, 32)); }

__device__ __forceinline__ Seg seg_geometry(const TcParams& p, const KRange& kr, int sg) {
  Seg s;
  s.st0 = kr.st_lo + sg * p.seg_stages;
  s.st1 = min(s.st0 + p.seg_stages, kr.st_hi);
  const int t_first = 2 * p.sps * s.st0;  // in 8-tap chunks: two per K16 step
  const int t_last = min(2 * p.sps * s.st1, 2 * kr.ks_hi) - 1;
  s.u_first = t_first / p.nkc;
  s.rows = 16 + (t_last / p.nkc - s.u_first);
  return s;
}

// CG = 1: every CTA works alone.  CG = 2: the two CTAs of a cluster form a CTA pair -- each runs the
// whole pipeline on its own work unit (same column tile and patch, neighbouring galleries) but the
// leader issues one tcgen05.mma.cta_group::2 (M = 256) for both, and each CTA streams only its half
// of the template columns: the B traffic per SM and the B footprint in shared memory halve.
template <int CG, int SPS>
__global__ void __launch_bounds__(kTcThreads, 1)
ncc_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_x,
              const __grid_constant__ CUtensorMap tm_ghi, const __grid_constant__ CUtensorMap tm_glo,
              const __grid_constant__ CUtensorMap tm_gx, const TcParams p) {
  // operand maps by mode   fp16x3: tm_lo = template lo (f16), tm_glo = gallery lo (f16), tm_x / tm_gx unused
  //                        fp8c  : tm_lo = template lo*4 (e4m3), tm_x = template hi/4 (e4m3),
  //                                tm_glo = gallery lo*4 (e4m3), tm_gx = gallery hi/4 (e4m3)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int kStageK = 16 * SPS;
  constexpr uint32_t kBHalfCta = b_half_bytes(SPS) / CG;                   // this CTA's share of one operand half
  const uint32_t stage_bytes = kBHalfCta * (p.passes == 1 ? 1 : 2);       // per CTA (fp8c: 1 + 1/2 + 1/2)
  const uint32_t e_buf_bytes = p.nhalf * p.e_half_bytes;
  const uint32_t bar0 = base + p.off_bar;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (kMaxBStages + i); };
  auto bar_efull = [&](int i) { return bar0 + 8u * (2 * kMaxBStages + i); };
  auto bar_eempty = [&](int i) { return bar0 + 8u * (2 * kMaxBStages + 2 + i); };
  auto bar_accfull = [&](int i) { return bar0 + 8u * (2 * kMaxBStages + 4 + i); };
  auto bar_accempty = [&](int i) { return bar0 + 8u * (2 * kMaxBStages + 6 + i); };
  auto bar_gsfull = [&](int i) { return bar0 + 8u * (2 * kMaxBStages + 8 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bar + 8u * (2 * kMaxBStages + 10));

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nbstages; ++i) {
      ptx::mbar_init(bar_full(i), 1);
      ptx::mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_efull(i), kGenThreads * CG);   // the leader also collects the peer's arrivals
      ptx::mbar_init(bar_eempty(i), 1);
      ptx::mbar_init(bar_accfull(i), 1);
      ptx::mbar_init(bar_accempty(i), kEpiWarps * CG);
      ptx::mbar_init(bar_gsfull(i), 1);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
    ptx::prefetch_tmap(&tm_ghi);
    ptx::prefetch_tmap(&tm_glo);
    if (p.passes == 2) {
      ptx::prefetch_tmap(&tm_x);
      ptx::prefetch_tmap(&tm_gx);
    }
  }
  if (warp == 1) {
    if constexpr (CG == 2) ptx::tmem_alloc_2cta(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    else ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  // The E buffers start from zeros: an MMA over a partial stage (the K32 fp8 corrections, the odd last K16 step) reads
  // entry rows no generator pass of this segment has written.  Their B taps are zero, but whatever an earlier kernel
  // left in shared memory may decode to NaN / Inf (0x7F / 0xFF as e4m3), and NaN * 0 poisons the position.
  for (uint32_t o = p.off_e + 16u * threadIdx.x; o < p.off_gs; o += 16u * blockDim.x)
    *reinterpret_cast<uint4*>(base_ptr + o) = make_uint4(0u, 0u, 0u, 0u);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx::cluster_sync();  // the peer's barriers are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work units are dealt round-robin (unit = blockIdx.x + k * gridDim.x) in the order
  // (column tile, patch, gallery): at any moment all CTAs work on the same column tile (one B stream
  // shared through L2) and on the same patch row (equal cost once zero rows are skipped).
  const int NP = p.npy * p.npx;
  const long long per_tile = (long long)p.Gp * NP;
  const int Pe = 8 * p.nkc;  // entries per E row
  const int a = p.Hm / 2, b = p.Wm / 2;
  const int M = p.Hp * p.Wp;

  // register budget: 168/thread at launch -> 88 for the producer warpgroup, 208 for the epilogue
  if (warp < 4) {
  ptx::setmaxnreg_dec<88>();
  if (warp == 0) {
    // ================================================================== TMA producer (B ring)
    if (ptx::elect_one()) {
      uint32_t slot = 0, par = 0;
      for (long long unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const int nt = (int)(unit / per_tile);
        const KRange kr = k_range(p, (int)((unit % per_tile) / p.Gp) / p.npx, nt);
        for (int c = 0; c < p.C; ++c) {
          for (int st = kr.st_lo; st < kr.st_hi; ++st) {
            ptx::mbar_wait(bar_empty(slot), par ^ 1);
            const uint32_t dst = base + p.off_b + slot * stage_bytes;
            if constexpr (CG == 2) {
              // both CTAs' loads complete on the leader's barrier; only the leader arms it
              if (leader) ptx::mbar_arrive_expect_tx(bar_full(slot), 2 * stage_bytes);
              const int col = nt * kTileN + (int)cta_rank * (tile_cols(p, nt) / 2);
              ptx::tma_load_3d_2sm(dst, &tm_hi, bar_full(slot), st * kStageK, col, c);
              if (p.passes == 3) ptx::tma_load_3d_2sm(dst + kBHalfCta, &tm_lo, bar_full(slot), st * kStageK, col, c);
              if (p.passes == 2) {
                ptx::tma_load_3d_2sm(dst + kBHalfCta, &tm_x, bar_full(slot), st * kStageK, col, c);
                ptx::tma_load_3d_2sm(dst + kBHalfCta + kBHalfCta / 2, &tm_lo, bar_full(slot), st * kStageK, col, c);
              }
            } else {
              ptx::mbar_arrive_expect_tx(bar_full(slot), stage_bytes);
              ptx::tma_load_3d(dst, &tm_hi, bar_full(slot), st * kStageK, nt * kTileN, c);
              if (p.passes == 3) ptx::tma_load_3d(dst + kBHalfCta, &tm_lo, bar_full(slot), st * kStageK, nt * kTileN, c);
              if (p.passes == 2) {
                ptx::tma_load_3d(dst + kBHalfCta, &tm_x, bar_full(slot), st * kStageK, nt * kTileN, c);
                ptx::tma_load_3d(dst + kBHalfCta + kBHalfCta / 2, &tm_lo, bar_full(slot), st * kStageK, nt * kTileN, c);
              }
            }
            if (++slot == (uint32_t)p.nbstages) {
              slot = 0;
              par ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (leader && ptx::elect_one()) {
      uint32_t idesc = 0;
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if constexpr (CG == 2) ptx::mma_f16_ss_2cta(d, da, db, idesc, acc);
        else ptx::mma_f16_ss(d, da, db, idesc, acc);
      };
      auto mma8 = [&](uint32_t d, uint64_t da, uint64_t db) {  // e4m3 x e4m3, K = 32, accumulate
        if constexpr (CG == 2) ptx::mma_f8_ss_2cta(d, da, db, idesc, 1);
        else ptx::mma_f8_ss(d, da, db, idesc, 1);
      };
      auto commit = [&](uint32_t bar) {
        if constexpr (CG == 2) ptx::tc_commit_2cta(bar, 0b11);
        else ptx::tc_commit(bar);
      };
      // The issuing thread is a single scalar instruction stream: everything it needs per MMA is kept
      // as a running 32-bit descriptor word (start address >> 4 in the low 14 bits) so that one stage
      // costs a barrier wait, a few adds and the MMAs themselves.
      //   A: no swizzle, rows of a core matrix 16 B apart (consecutive entries = consecutive x), 8-row
      //      groups one E row apart (SBO), the two K chunks of a step 8 entries apart (LBO = 128 B).
      //   B: 64B swizzle, 8-row groups 512 B apart; a K16 sub-step is +32 B inside the swizzle row.
      const uint64_t a_desc_hi = ((uint64_t)((16u * Pe) >> 4) | (1ull << 14)) << 32;                 // SBO | version
      // B: 32-tap stages are 64-byte rows (64B swizzle, 8-row groups 512 B apart), 64-tap stages 128-byte rows (128B swizzle,
      //    groups 1024 B apart); a K16 sub-step is +32 B inside the swizzle row either way
      const uint64_t b_desc_hi = SPS == 2 ? ((uint64_t)(512u >> 4) | (1ull << 14) | (4ull << 29)) << 32      // SBO | version | SW64
                                          : ((uint64_t)(1024u >> 4) | (1ull << 14) | (2ull << 29)) << 32;    // SBO | version | SW128
      constexpr uint32_t a_lbo = (128u >> 4) << 16, b_lbo = (16u >> 4) << 16;
      auto a_desc = [&](uint32_t addr) { return a_desc_hi | a_lbo | ((addr >> 4) & 0x3FFFu); };
      auto b_desc = [&](uint32_t addr) { return b_desc_hi | b_lbo | ((addr >> 4) & 0x3FFFu); };
      // fp8 operands: A entries hold 16 one-byte cells, the two K chunks of a K32 step are 16 entries
      // (256 B) apart; B rows are 32 bytes (32B swizzle), 8-row groups 256 B apart.
      constexpr uint32_t a8_lbo = (256u >> 4) << 16;
      const uint64_t b8_desc_hi = ((uint64_t)(256u >> 4) | (1ull << 14) | (6ull << 29)) << 32;       // SBO | version | SW32
      auto a8_desc = [&](uint32_t addr) { return a_desc_hi | a8_lbo | ((addr >> 4) & 0x3FFFu); };
      auto b8_desc = [&](uint32_t addr) { return b8_desc_hi | b_lbo | ((addr >> 4) & 0x3FFFu); };
      uint32_t slot = 0, bphase = 0, es = 0, cs = 0;
      uint32_t b_addr = base + p.off_b;
      const uint32_t b_end = b_addr + p.nbstages * stage_bytes;
      for (long long unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const KRange kr = k_range(p, (int)((unit % per_tile) / p.Gp) / p.npx, (int)(unit / per_tile));
        idesc = ptx::make_idesc_f16(kTileM * CG, tile_cols(p, (int)(unit / per_tile)));
        for (int c = 0; c < p.C; ++c, ++cs) {
          const int buf = cs & 1;
          ptx::mbar_wait(bar_accempty(buf), ((cs >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kTileN;
          uint32_t accumulate = 0;
          for (int sg = 0; sg < kr.nseg; ++sg, ++es) {
            const Seg s = seg_geometry(p, kr, sg);
            const int ebuf = es & 1;
            ptx::mbar_wait(bar_efull(ebuf), (es >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t e_hi = base + p.off_e + ebuf * e_buf_bytes;
            // E row 0 of this segment is template row u_first: K step ks starts 256*ks - 16*u_first*Pe bytes in
            uint32_t a_hi = e_hi + 256u * SPS * s.st0 - 16u * s.u_first * Pe;
            for (int st = s.st0; st < s.st1; ++st, a_hi += 256u * SPS) {
              ptx::mbar_wait(bar_full(slot), bphase);
              ptx::tc_fence_after();
              // K16 steps of this stage with a non-zero A operand (only the first and last stage of a unit are partial)
              const int k_lo = max(kr.ks_lo - SPS * st, 0), k_hi = min(kr.ks_hi - SPS * st, SPS);
#pragma unroll
              for (int kk = 0; kk < SPS; ++kk) {
                if (kk >= k_lo && kk < k_hi) {
                  const uint64_t da_hi = a_desc(a_hi + 256u * kk);
                  const uint64_t db_hi = b_desc(b_addr + 32u * kk);
                  mma(tmem_d, da_hi, db_hi, accumulate);
                  accumulate = 1;
                  if constexpr (SPS == 2) {
                    if (p.passes == 3) {
                      mma(tmem_d, a_desc(a_hi + p.e_half_bytes + 256u * kk), db_hi, 1);
                      mma(tmem_d, da_hi, b_desc(b_addr + kBHalfCta + 32u * kk), 1);
                    }
                  }
                }
              }
              if constexpr (SPS == 2) {
                if (p.passes == 2 && k_hi > k_lo) {
                  // corrections over the whole 32-tap stage: (A_lo*4)(B_hi/4) and (A_hi/4)(B_lo*4), both e4m3
                  mma8(tmem_d, a8_desc(a_hi + p.e_half_bytes), b8_desc(b_addr + kBHalfCta));
                  mma8(tmem_d, a8_desc(a_hi + 2 * p.e_half_bytes), b8_desc(b_addr + kBHalfCta + kBHalfCta / 2));
                }
              }
              commit(bar_empty(slot));
              b_addr += stage_bytes;
              if (++slot == (uint32_t)p.nbstages) {
                slot = 0;
                bphase ^= 1;
                b_addr = b_end - p.nbstages * stage_bytes;
              }
            }
            commit(bar_eempty(ebuf));
          }
          commit(bar_accfull(buf));
        }
      }
    }
  } else if (warp < 4) {
    // ================================================================== generators (E buffers)
    // Per (unit, channel, segment): TMA stages the window of the packed channel the segment can
    // touch (out-of-map cells arrive as zeros), issued one segment ahead into the other staging
    // buffer; the 64 threads then expand it into the shifted-entry array E.
    const int tg = threadIdx.x - 64;
    // staged cells per row: TMA needs a 16-byte aligned start column, so the window starts at the
    // multiple of 8 at or below (8*px - b) and is 8 cells wider than the entries need
    const int SP = Pe + 16;
    const uint32_t gs_bytes_half = 2u * p.gs_half_elems;
    const int SP8 = Pe + 32;  // fp8c: staged BYTES per row of the 1-byte windows (start column aligned to 16)
    // staging buffer: [fp16 hi window][fp16 lo window | fp8 lo window][fp8 hi/4 window]
    const uint32_t gs_buf_bytes = gs_bytes_half + (p.passes == 3 ? gs_bytes_half : 0) + (p.passes == 2 ? 2 * p.gs8_bytes : 0);
    const uint32_t gs_tx = 2u * SP * p.gs_rows * (p.passes == 3 ? 2 : 1) + (p.passes == 2 ? 2u * SP8 * p.gs_rows : 0);

    struct Cursor {
      long long unit;
      int c, sg, g, py, px;
      KRange kr;
    };
    auto decode = [&](Cursor& cu) {
      const long long rem = cu.unit % per_tile;
      const int pidx = (int)(rem / p.Gp);
      cu.g = min((int)(rem % p.Gp), p.G - 1);
      cu.py = pidx / p.npx;
      cu.px = pidx % p.npx;
      cu.kr = k_range(p, cu.py, (int)(cu.unit / per_tile));
    };
    auto advance = [&](Cursor& cu) {
      if (++cu.sg < cu.kr.nseg) return;
      cu.sg = 0;
      if (++cu.c < p.C) return;
      cu.c = 0;
      cu.unit += gridDim.x;
      if (cu.unit < p.nunits) decode(cu);
    };
    auto issue_stage = [&](const Cursor& cu, int sbuf) {  // one thread
      const Seg sgm = seg_geometry(p, cu.kr, cu.sg);
      const uint32_t dst = base + p.off_gs + sbuf * gs_buf_bytes;
      ptx::fence_proxy_async_smem();  // earlier generic reads of this buffer precede the async write
      ptx::mbar_arrive_expect_tx(bar_gsfull(sbuf), gs_tx);
      const int x0 = (8 * cu.px - b) & ~7, y0 = 16 * cu.py + sgm.u_first - a, z0 = cu.g * p.C + cu.c;
      ptx::tma_load_3d(dst, &tm_ghi, bar_gsfull(sbuf), x0, y0, z0);
      if (p.passes == 3) ptx::tma_load_3d(dst + gs_bytes_half, &tm_glo, bar_gsfull(sbuf), x0, y0, z0);
      if (p.passes == 2) {
        const int x8 = (8 * cu.px - b) & ~15;
        ptx::tma_load_3d(dst + gs_bytes_half, &tm_glo, bar_gsfull(sbuf), x8, y0, z0);
        ptx::tma_load_3d(dst + gs_bytes_half + p.gs8_bytes, &tm_gx, bar_gsfull(sbuf), x8, y0, z0);
      }
    };

    Cursor cur{};
    cur.unit = blockIdx.x;
    if (cur.unit < p.nunits) {
      decode(cur);
      Cursor nxt = cur;
      advance(nxt);
      const bool ahead = p.gs_bufs == 2;
      if (tg == 0 && ahead) issue_stage(cur, 0);
      uint32_t es = 0;
      while (cur.unit < p.nunits) {
        const int ebuf = es & 1, sbuf = ahead ? (es & 1) : 0;
        if (tg == 0) {
          if (!ahead) issue_stage(cur, 0);
          else if (nxt.unit < p.nunits) issue_stage(nxt, sbuf ^ 1);
        }
        const Seg s = seg_geometry(p, cur.kr, cur.sg);
        ptx::mbar_wait(bar_eempty(ebuf), ((es >> 1) & 1) ^ 1);
        ptx::mbar_wait(bar_gsfull(sbuf), ahead ? ((es >> 1) & 1) : (es & 1));
        const __half* gs_hi = reinterpret_cast<const __half*>(base_ptr + p.off_gs + sbuf * gs_buf_bytes);
        const __half* gs_lo = gs_hi + p.gs_half_elems;
        // shifted entries E[r][i] = gs[r][i .. i+7]; one thread emits an even/odd pair from five
        // aligned 32-bit words (the odd entry is the even one funnel-shifted by one cell)
        uint8_t* e_hi = base_ptr + p.off_e + ebuf * e_buf_bytes;
        uint8_t* e_lo = e_hi + p.e_half_bytes;
        const int pairs_per_row = Pe / 2;
        const int delta = (8 * cur.px - b) & 7;  // first needed cell inside the aligned window
        const bool odd = delta & 1;
        for (int pr = tg; pr < s.rows * pairs_per_row; pr += kGenThreads) {
          const int r = pr / pairs_per_row, ip = pr - r * pairs_per_row;
          const int wofs = (delta >> 1) + ip;  // 32-bit word holding the pair's first (even) cell
          {
            const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(gs_hi + r * SP) + wofs;
            const uint32_t w0 = wsrc[0], w1 = wsrc[1], w2 = wsrc[2], w3 = wsrc[3], w4 = wsrc[4];
            const uint4 al = odd ? make_uint4(w1, w2, w3, w4) : make_uint4(w0, w1, w2, w3);
            const uint4 sh = make_uint4(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16),
                                        __funnelshift_r(w3, w4, 16));
            uint4* dst = reinterpret_cast<uint4*>(e_hi + 16u * (r * Pe + 2 * ip));
            dst[0] = odd ? sh : al;
            dst[1] = odd ? al : sh;
          }
          if (p.passes == 3) {
            const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(gs_lo + r * SP) + wofs;
            const uint32_t w0 = wsrc[0], w1 = wsrc[1], w2 = wsrc[2], w3 = wsrc[3], w4 = wsrc[4];
            const uint4 al = odd ? make_uint4(w1, w2, w3, w4) : make_uint4(w0, w1, w2, w3);
            const uint4 sh = make_uint4(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16),
                                        __funnelshift_r(w3, w4, 16));
            uint4* dst = reinterpret_cast<uint4*>(e_lo + 16u * (r * Pe + 2 * ip));
            dst[0] = odd ? sh : al;
            dst[1] = odd ? al : sh;
          }
        }
        if (p.passes == 2) {
          // fp8 entries: E8[r][i] = 16 one-byte cells gs8[r][d8 + i .. +15], assembled from five aligned
          // 32-bit words and a byte-granular funnel shift; arrays: [1] = lo*4, [2] = hi/4
          const int d8 = (8 * cur.px - b) & 15;
#pragma unroll 1
          for (int arr = 0; arr < 2; ++arr) {
            const uint8_t* g8 = reinterpret_cast<const uint8_t*>(gs_hi) + gs_bytes_half + arr * p.gs8_bytes;
            uint8_t* e8 = e_hi + (1 + arr) * p.e_half_bytes;
            for (int e = tg; e < s.rows * Pe; e += kGenThreads) {
              const int r = e / Pe, i = e - r * Pe;
              const int o = d8 + i;
              const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(g8 + r * SP8) + (o >> 2);
              const uint32_t sh = (o & 3) * 8;
              const uint32_t w0 = wsrc[0], w1 = wsrc[1], w2 = wsrc[2], w3 = wsrc[3], w4 = wsrc[4];
              *reinterpret_cast<uint4*>(e8 + 16u * e) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                                                  __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
            }
          }
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async reads
        if constexpr (CG == 2) ptx::mbar_arrive_leader(bar_efull(ebuf));
        else ptx::mbar_arrive(bar_efull(ebuf));
        ptx::named_bar_sync(1, kGenThreads);  // everyone is done reading staging[sbuf]
        cur = nxt;
        advance(nxt);
        ++es;
      }
    }
  }
  } else {
    // ================================================================== epilogue (8 warps)
    ptx::setmaxnreg_inc<208>();
    const int ew = warp - 4;
    const int q4 = warp & 3;        // TMEM lane quarter this warp may touch
    const int half = ew >> 2;       // which 128 of the 256 columns
    const int m = q4 * 32 + lane;   // tile row = TMEM lane
    const int ml = m & 7, mh = m >> 3;
    float* colmax = reinterpret_cast<float*>(base_ptr + p.off_cm);  // [4][256]
    uint32_t cs = 0;
    for (long long unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const int nt = (int)(unit / per_tile);
      const long long rem = unit % per_tile;
      const int pidx = (int)(rem / p.Gp), g = min((int)(rem % p.Gp), p.G - 1);
      const int py = pidx / p.npx, px = pidx % p.npx;
      const int y = 16 * py + mh, x = 8 * px + ml;
      const bool valid = (y < p.Hp) && (x < p.Wp);
      const int ncol_half = tile_cols(p, nt) - half * 128;  // columns of this warp's half that exist
      // window-norm rows of this thread's position, one per kNormChunkCols-column chunk of its column half: with
      // multi-shape tiles the chunks of a tile may belong to templates of different true shapes (the table pointers are
      // warp-uniform loads, re-read per channel rather than held in 2 x 16 registers)
      constexpr int kChunks = 128 / kNormChunkCols;
      const size_t roff = (size_t)g * p.C * M + (valid ? y * p.Wp + x : 0);
      const float* const* tabp = p.rnorm_tab ? p.rnorm_tab + ((size_t)nt * (kTileN / kNormChunkCols) + half * kChunks) : nullptr;
      auto load_norms = [&](float (&r)[kChunks], int c) {
        if (tabp) {
#pragma unroll
          for (int j = 0; j < kChunks; ++j)
            r[j] = (valid && c < p.C && j * kNormChunkCols < ncol_half) ? __ldg(tabp[j] + roff + (size_t)c * M) : 0.0f;
        } else {
          const float v = (valid && c < p.C) ? __ldg(p.rnorm + roff + (size_t)c * M) : 0.0f;
#pragma unroll
          for (int j = 0; j < kChunks; ++j) r[j] = v;
        }
      };

      float total[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) total[j] = 0.0f;
      float r_cur[kChunks], r_next[kChunks];
      load_norms(r_cur, 0);
      for (int c = 0; c < p.C; ++c, ++cs) {
        load_norms(r_next, c + 1);
        const int buf = cs & 1;
        ptx::mbar_wait(bar_accfull(buf), (cs >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * kTileN + half * 128;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (j4 * 32 >= ncol_half) break;  // warp-uniform
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + j4 * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            total[j4 * 32 + j] = fmaf(r_cur[(32 / kNormChunkCols) * j4 + j / kNormChunkCols], __uint_as_float(v[j]), total[j4 * 32 + j]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) ptx::mbar_arrive_leader(bar_accempty(buf));
          else ptx::mbar_arrive(bar_accempty(buf));
        }
#pragma unroll
        for (int j = 0; j < kChunks; ++j) r_cur[j] = r_next[j];
      }
      // max over the tile's valid positions, then over the 4 lane quarters, then into scores
      if (p.rec == nullptr) {
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const float v = warp_max(valid ? total[j] : -INFINITY);
          if (lane == 0) colmax[q4 * kTileN + half * 128 + j] = v;
        }
        ptx::named_bar_sync(2, kEpiWarps * 32);
        {
          const int col = ew * 32 + lane;
          const float v = fmaxf(fmaxf(colmax[col], colmax[kTileN + col]), fmaxf(colmax[2 * kTileN + col], colmax[3 * kTileN + col]));
          const int n = nt * kTileN + col;
          if (n < p.ncols) atomic_max_nonneg(&p.scores[(size_t)p.col2probe[n] * p.score_ld + p.g0 + g], v * p.out_scale);
        }
      } else {
        // Screening: the fp16 surface is only good to ~2e-4 relative, so next to the maximum the unit reports WHICH
        // rows could hold the true maximum: all rows within tau of the tile maximum.  f(m) = m - tau(m) is increasing,
        // so the rows within tau of a lane quarter's own maximum are a superset of that quarter's share.
        uint32_t* colmask = reinterpret_cast<uint32_t*>(base_ptr + p.off_mask);  // [4][256]
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const float t = valid ? total[j] : -INFINITY;
          const float m = warp_max(t);
          const uint32_t bits = __ballot_sync(0xffffffffu, t >= m - (p.tau_rel * fabsf(m) + p.tau_abs));
          if (lane == 0) {
            colmax[q4 * kTileN + half * 128 + j] = m;
            colmask[q4 * kTileN + half * 128 + j] = bits;
          }
        }
        ptx::named_bar_sync(2, kEpiWarps * 32);
        {
          const int col = ew * 32 + lane;
          const int n = nt * kTileN + col;
          if (n < p.ncols) {
            float mq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) mq[q] = colmax[q * kTileN + col];
            const float v = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
            const float thr = v - (p.tau_rel * fabsf(v) + p.tau_abs);
            // info: up to three candidate rows (8 bits each) + their total count, saturated at 255; more than
            // three makes the refinement evaluate every position of the patch (rare)
            uint32_t info = 0, cnt = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (mq[q] >= thr) {
                uint32_t bits = colmask[q * kTileN + col];
                while (bits) {
                  const uint32_t r = __ffs(bits) - 1;
                  bits &= bits - 1;
                  if (cnt < 3) info |= (uint32_t)(q * 32 + r) << (8 * cnt);
                  ++cnt;
                }
              }
            }
            info |= min(cnt, 255u) << 24;
            p.rec[((size_t)n * p.G + g) * NP + pidx] = make_uint2(__float_as_uint(v * p.out_scale), info);
            atomic_max_nonneg(&p.scores[(size_t)p.col2probe[n] * p.score_ld + p.g0 + g], v * p.out_scale);
          }
        }
      }
      ptx::named_bar_sync(2, kEpiWarps * 32);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx::cluster_sync();  // nobody frees TMEM or exits while the peer still uses this CTA's memory
  if (warp == 1) {
    if constexpr (CG == 2) ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
    else ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_template_map(CUtensorMap* tm, const uint16_t* ptr, int Kpad, int ncols_alloc, int C, int box_cols, int stage_k = 32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SIR_E_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)ncols_alloc, (cuuint64_t)C};
  cuuint64_t strides[2] = {(cuuint64_t)Kpad * 2, (cuuint64_t)Kpad * 2 * (cuuint64_t)ncols_alloc};
  // a 64-tap box may reach past Kpad (a multiple of 32): the cells outside the tensor arrive as zeros
  cuuint32_t box[3] = {(cuuint32_t)stage_k, (cuuint32_t)box_cols, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, stage_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (Kpad=%d ncols=%d C=%d)", (int)r, Kpad, ncols_alloc, C);
    return SIR_E_CUDA;
  }
  return SIR_OK;
}
// Packed gallery channels [planes][Hp][pitch] f16; box = one staging window (cols x rows), no
// swizzle, out-of-bounds cells (the "same"-mode zero padding, negative coordinates included) read 0.
int make_gallery_map(CUtensorMap* tm, const uint16_t* ptr, int planes, int Hp, int Wp, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SIR_E_CUDA;
  }
  const int pitch = gal_pitch(Wp);
  // the logical width is Wp (pad cells are zero anyway, but keeping them out of bounds is free)
  cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * (cuuint64_t)Hp};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(gallery) failed with CUresult %d (planes=%d Hp=%d Wp=%d box=%dx%d)", (int)r, planes, Hp,
              Wp, box_cols, box_rows);
    return SIR_E_CUDA;
  }
  return SIR_OK;
}
// e4m3 template operands [C][ncols][Kpad] bytes: box 32 bytes x box_cols, 32-byte swizzle
int make_template_map8(CUtensorMap* tm, const uint8_t* ptr, int Kpad, int ncols_alloc, int C, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SIR_E_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)ncols_alloc, (cuuint64_t)C};
  cuuint64_t strides[2] = {(cuuint64_t)Kpad, (cuuint64_t)Kpad * (cuuint64_t)ncols_alloc};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_cols, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(e4m3 templates) failed with CUresult %d (Kpad=%d ncols=%d C=%d)", (int)r, Kpad, ncols_alloc, C);
    return SIR_E_CUDA;
  }
  return SIR_OK;
}
// e4m3 gallery channels [planes][Hp][pitch8] bytes; box = one staging window
int make_gallery_map8(CUtensorMap* tm, const uint8_t* ptr, int planes, int Hp, int Wp, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SIR_E_CUDA;
  }
  const int pitch = gal_pitch8(Wp);
  cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)Hp};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(e4m3 gallery) failed with CUresult %d (planes=%d Hp=%d Wp=%d box=%dx%d)", (int)r, planes, Hp, Wp,
              box_cols, box_rows);
    return SIR_E_CUDA;
  }
  return SIR_OK;
}
}  // namespace

// passes 3 / 1: d_glo, d_tlo are the fp16 lo operands, the *8* pointers are unused.
// passes 2 (fp8c): d_g8l/d_g8a and d_t8l/d_t8b are the e4m3 companions (lo*4, hi/4), d_glo/d_tlo unused.
int launch_ncc_tc(const uint16_t* d_ghi, const uint16_t* d_glo, const uint8_t* d_g8a, const uint8_t* d_g8l, const float* d_rnorm, int G,
                  int C, int Hp, int Wp, const uint16_t* d_thi, const uint16_t* d_tlo, const uint8_t* d_t8b, const uint8_t* d_t8l,
                  int ncols, int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                  int passes, cudaStream_t st, double* cost_out, const float* const* d_rnorm_tab, uint2* d_rec, float tau_rel,
                  float tau_abs, const int32_t* d_tile_rows) {
  // cost_out != NULL: dry run -- plan only, report the estimated SM cycles per (gallery, 256-column tile,
  // channel) and return without touching any pointer (sir_ncc_cost)
  SIR_CHECK_ARG(d_ghi && d_thi, "sir_ncc_scores(tcgen05): needs packed fp16 operands");
  if (passes == 2) SIR_CHECK_ARG(d_g8a && d_g8l && d_t8b && d_t8l, "sir_ncc_scores_fp8c: needs the e4m3 companion operands");
  else if (passes == 3) SIR_CHECK_ARG(d_glo && d_tlo, "sir_ncc_scores(tcgen05): needs the fp16 lo operands");
  SIR_CHECK_ARG((reinterpret_cast<uintptr_t>(d_thi) & 15) == 0, "sir_ncc_scores: template operands must be 16-byte aligned");
  TcParams p{};
  p.ghi = (const __half*)d_ghi;
  p.glo = (const __half*)d_glo;
  p.rnorm = d_rnorm;
  p.rnorm_tab = d_rnorm_tab;
  p.tile_rows = reinterpret_cast<const int2*>(d_tile_rows);
  p.col2probe = d_col2probe;
  p.scores = d_scores;
  p.score_ld = score_ld;
  p.g0 = g0;
  p.G = G; p.C = C; p.Hp = Hp; p.Wp = Wp; p.Hm = Hm; p.Wm = Wm;
  p.ncols = ncols;
  const int row_align = passes == 2 ? 16 : 8;
  p.nkc = tpl_row_taps(Wm, row_align) / 8;
  const int Kpad = tpl_kpad_aligned(Hm, Wm, row_align);
  p.nsteps = ceil_div(Hm * p.nkc, 2);
  // one-pass fp16 (screening, fp16x1): 64-tap stages so that the issuing thread has four MMAs per barrier round trip;
  // SIR_STAGE_TAPS=32 forces the short stages
  p.sps = passes == 1 ? 4 : 2;
  if (const char* env = getenv("SIR_STAGE_TAPS")) p.sps = (passes == 1 && atoi(env) != 32) ? 4 : 2;
  const int stage_k = 16 * p.sps;
  p.nkstages = ceil_div(Kpad, stage_k);
  p.npy = ceil_div(Hp, 16);
  p.npx = ceil_div(Wp, 8);
  p.ntiles_n = ceil_div(ncols, kTileN);
  // CTA pairs (cta_group::2) whenever there are at least two galleries; SIR_CTA_GROUP=1 forces single CTAs
  int cg = G >= 2 ? 2 : 1;
  if (const char* env = getenv("SIR_CTA_GROUP")) cg = atoi(env) == 1 ? 1 : cg;
  p.Gp = round_up(G, cg);
  p.nunits = (long long)p.ntiles_n * p.Gp * p.npy * p.npx;
  p.passes = passes;
  p.out_scale = 1.0f / ((float)C * (float)(1 << kTemplateScaleLog2));
  p.rec = d_rec;
  p.tau_rel = tau_rel;
  p.tau_abs = tau_abs / p.out_scale;  // the epilogue compares unscaled accumulator sums

  // shared memory plan: B ring, 2 E buffers (x halves), row staging, column maxima, barriers
  const int halves = passes == 3 ? 2 : passes == 2 ? 3 : 1;          // operand arrays per E buffer
  const int gs16 = passes == 3 ? 2 : 1;                             // staged fp16 windows
  const uint32_t stage_bytes = b_half_bytes(p.sps) / cg * (passes == 1 ? 1 : 2);  // per CTA
  const int Pe = 8 * p.nkc;
  const size_t limit = 227 * 1024 - 1024;  // alignment slack
  // Plan: long E segments first (every segment rebuilds 16 + rows-in-segment E rows, so short segments
  // make the generators the bottleneck), then as deep a B ring as the remaining shared memory allows
  // (at least 3 stages; 2 as a last resort), staging double-buffered unless that is what does not fit.
  bool ok = false;
  const int nb_max = passes == 1 ? (p.sps == 4 ? 6 : 12) : (cg == 2 ? 6 : 4);
  for (int nb_min = 3; nb_min >= 2 && !ok; --nb_min) {
    for (int gsb = 2; gsb >= 1 && !ok; --gsb) {
      for (int seg = p.nkstages; seg >= 1; --seg) {
        // rows touched by a segment of `seg` stages: worst case over alignments
        const int max_rows = 16 + (2 * p.sps * seg - 1) / p.nkc + 1;
        const size_t e_half = (size_t)(max_rows + 1) * Pe * 16;
        const size_t gs_half = (size_t)(max_rows + 1) * (Pe + 16);  // cells
        const size_t gs8 = passes == 2 ? (size_t)round_up((max_rows + 1) * (Pe + 32), 128) : 0;  // bytes of one 1-byte window
        const size_t gs_buf = gs16 * (size_t)round_up((int)gs_half, 64) * 2 + 2 * gs8;
        const size_t fixed = 2 * halves * (e_half + 128) + (size_t)gsb * gs_buf + (d_rec ? 2 : 1) * 4 * kTileN * 4 + 512;
        if (fixed + (size_t)nb_min * stage_bytes <= limit) {
          p.nbstages = (int)std::min<size_t>(nb_max, (limit - fixed) / stage_bytes);
          p.seg_stages = seg;
          p.e_half_bytes = (uint32_t)round_up((int)e_half, 128);
          p.gs_half_elems = (uint32_t)round_up((int)gs_half, 64);  // 128-byte aligned TMA destinations
          p.gs_rows = max_rows + 1;
          p.gs_bufs = gsb;
          p.gs8_bytes = (uint32_t)gs8;
          p.nhalf = halves;
          ok = true;
          break;
        }
        if (seg > 64) seg -= seg / 8;  // coarse search for very long K
      }
    }
  }
  SIR_CHECK_ARG(ok, "sir_ncc_scores: template %dx%d does not fit the shared-memory plan", Hm, Wm);
  p.nseg = ceil_div(p.nkstages, p.seg_stages);
  if (cost_out) {
    // MMA cycles of the non-skipped stages vs generator cycles (calibrated on B200 with 64 generator
    // threads: ~1.3 cycles per fp16 entry, built in pairs, ~2.5 per fp8 entry); whichever is larger paces
    // a (unit, channel)
    const double cyc_stage = passes == 3 ? 768.0 : passes == 2 ? 512.0 : 128.0 * p.sps;
    const int a = Hm / 2;
    double total = 0.0;
    for (int py = 0; py < p.npy; ++py) {
      const int u_lo = std::max(0, a - 16 * py - 15), u_hi = std::min(Hm, Hp + a - 16 * py);
      const int ks_lo = (u_lo * p.nkc) / 2, ks_hi = std::min(p.nsteps, (u_hi * p.nkc + 1) / 2);
      const int stages = (ks_hi + p.sps - 1) / p.sps - ks_lo / p.sps;
      const int nseg_u = ceil_div(stages, p.seg_stages);
      const int rows_seg = 16 + (2 * p.sps * std::min(stages, p.seg_stages) - 1) / p.nkc + 2;
      const double mma = stages * cyc_stage;
      const double per_entry = passes == 2 ? 1.3 + 2 * 2.5 : 1.3 * p.nhalf;
      const double gen = per_entry * nseg_u * (double)rows_seg * Pe;
      total += p.npx * std::max(mma, gen);
    }
    *cost_out = total;
    return SIR_OK;
  }
  p.off_b = 0;
  p.off_e = p.off_b + p.nbstages * stage_bytes;
  p.off_gs = p.off_e + 2 * p.nhalf * p.e_half_bytes;
  p.off_cm = (uint32_t)round_up((int)(p.off_gs + p.gs_bufs * (gs16 * p.gs_half_elems * 2 + 2 * p.gs8_bytes)), 16);
  p.off_mask = p.off_cm + 4 * kTileN * 4;
  p.off_bar = p.off_mask + (d_rec ? 4 * kTileN * 4 : 0);
  const size_t smem = 1024 + p.off_bar + 512;
  SIR_CHECK_ARG(smem <= 227 * 1024, "sir_ncc_scores: shared-memory plan overflow (%zu bytes)", smem);

  SIR_CHECK_ARG(Pe + (passes == 2 ? 32 : 16) <= 256 && p.gs_rows <= 256, "sir_ncc_scores: template %dx%d exceeds the staging TMA box", Hm, Wm);
  SIR_CHECK_ARG((reinterpret_cast<uintptr_t>(d_ghi) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_glo) & 15) == 0,
                "sir_ncc_scores: gallery operands must be 16-byte aligned");
  CUtensorMap tm_hi, tm_lo, tm_x, tm_ghi, tm_glo, tm_gx;
  int rc = make_template_map(&tm_hi, d_thi, Kpad, ncols_alloc, C, kTileN / cg, stage_k);
  if (rc) return rc;
  rc = make_gallery_map(&tm_ghi, d_ghi, G * C, Hp, Wp, Pe + 16, p.gs_rows);
  if (rc) return rc;
  if (passes == 2) {
    rc = make_template_map8(&tm_lo, d_t8l, Kpad, ncols_alloc, C, kTileN / cg);
    if (rc) return rc;
    rc = make_template_map8(&tm_x, d_t8b, Kpad, ncols_alloc, C, kTileN / cg);
    if (rc) return rc;
    rc = make_gallery_map8(&tm_glo, d_g8l, G * C, Hp, Wp, Pe + 32, p.gs_rows);
    if (rc) return rc;
    rc = make_gallery_map8(&tm_gx, d_g8a, G * C, Hp, Wp, Pe + 32, p.gs_rows);
    if (rc) return rc;
  } else {
    // one-pass modes never touch the lo maps: alias them to the hi operands when the caller has none
    rc = make_template_map(&tm_lo, d_tlo ? d_tlo : d_thi, Kpad, ncols_alloc, C, kTileN / cg, stage_k);
    if (rc) return rc;
    rc = make_gallery_map(&tm_glo, d_glo ? d_glo : d_ghi, G * C, Hp, Wp, Pe + 16, p.gs_rows);
    if (rc) return rc;
    tm_x = tm_lo;
    tm_gx = tm_glo;
  }

  auto kernel = cg == 2 ? (p.sps == 4 ? ncc_tc_kernel<2, 4> : ncc_tc_kernel<2, 2>) : (p.sps == 4 ? ncc_tc_kernel<1, 4> : ncc_tc_kernel<1, 2>);
  SIR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device, cheap: no cache
  int dev = 0, sms = 0;
  SIR_CUDA(cudaGetDevice(&dev));
  SIR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  unsigned grid = (unsigned)std::min<long long>(p.nunits, sms);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  if (cg == 2) {
    grid &= ~1u;  // whole pairs only (nunits is even)
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  SIR_CUDA(cudaLaunchKernelEx(&cfg, kernel, tm_hi, tm_lo, tm_x, tm_ghi, tm_glo, tm_gx, p));
  SIR_LAUNCH_CHECK("ncc_tc_kernel");
  return SIR_OK;
}

}  // namespace sir

extern "C" int sir_ncc_cost(int precision, int G, int Hp, int Wp, int Hm, int Wm, double* h_cost) {
  SIR_CHECK_ARG(h_cost && G > 0 && Hp > 0 && Wp > 0 && Hm > 0 && Wm > 0, "sir_ncc_cost: bad argument");
  const int passes = precision == SIR_PREC_FP16X3 ? 3 : precision == SIR_PREC_FP16_FP8C ? 2
                     : (precision == SIR_PREC_FP16X1 || precision == SIR_PREC_FP16_REFINE) ? 1 : 0;
  SIR_CHECK_ARG(passes != 0, "sir_ncc_cost: precision %d has no tensor-core plan", precision);
  alignas(16) static const uint16_t dummy16[8] = {0};
  alignas(16) static const uint8_t dummy8[16] = {0};
  return sir::launch_ncc_tc(dummy16, dummy16, dummy8, dummy8, nullptr, G, 1, Hp, Wp, dummy16, dummy16, dummy8, dummy8, 256, 256, Hm, Wm,
                            nullptr, nullptr, 0, 0, passes, nullptr, h_cost, nullptr, nullptr, 0.0f, 0.0f, nullptr);
}
