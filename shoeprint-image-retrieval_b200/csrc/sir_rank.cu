// K8 / K9: rank of the true match without a sort, and warp-select top-k.
//
// Replaces _get_rank (similarity.py:378-386): flip(argsort(row)) + where == true is
// 1 + #{g : s[g] > s[true]} up to the order of exact ties, so one streaming pass over the score
// row is enough.  HBM bound: 4*Q*G bytes read once (16-byte vector loads), Q*(8k+8) written.
#include "sir_common.cuh"

#include <algorithm>
#include <cfloat>

namespace sir {

constexpr int kRankThreads = 256;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kMaxTopK = 128;

__global__ void true_scores_kernel(const float* __restrict__ scores, int Q, int G, int ld,
                                   const int32_t* __restrict__ true_idx, int g0, float* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int t = true_idx[q] - g0;
  out[q] = (t >= 0 && t < G) ? scores[(size_t)q * ld + t] : -INFINITY;
}

// (value desc, index asc) strict order
__device__ __forceinline__ bool beats(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

// Warp-select: each warp keeps an UNSORTED list of its k best (value, index) pairs in shared memory
// together with the current worst entry (value `thr`, slot `wslot`).  A candidate that beats the worst
// replaces it and the worst is recomputed by a warp reduction -- O(k/32) per insertion, no shifting.
// Ordering (value desc, index asc) is only established once, by the final merge.
struct Worst {
  float v;
  int idx;
  int slot;
};
__device__ __forceinline__ Worst warp_find_worst(const float* lv, const int* li, int k, int lane) {
  Worst w{INFINITY, -1, 0};
  for (int j = lane; j < k; j += 32) {
    const float v = lv[j];
    const int id = li[j];
    if (v < w.v || (v == w.v && id > w.idx)) w = Worst{v, id, j};
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, w.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, w.idx, o);
    const int os = __shfl_xor_sync(0xffffffffu, w.slot, o);
    if (ov < w.v || (ov == w.v && oi > w.idx)) w = Worst{ov, oi, os};
  }
  return w;
}

__global__ void __launch_bounds__(kRankThreads) rank_topk_kernel(const float* __restrict__ scores, int Q, int G, int ld,
                                                                 const float* __restrict__ true_score, int g0, int k,
                                                                 int32_t* __restrict__ count_gt, int32_t* __restrict__ count_ge,
                                                                 float* __restrict__ topk_val, int32_t* __restrict__ topk_idx) {
  __shared__ float lv[kRankWarps][kMaxTopK];
  __shared__ int li[kRankWarps][kMaxTopK];
  __shared__ int cnt[2][kRankWarps];
  const int q = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* row = scores + (size_t)q * ld;
  const float ts = true_score[q];

  for (int j = lane; j < k; j += 32) { lv[wid][j] = -INFINITY; li[wid][j] = INT_MAX; }
  __syncwarp();
  Worst worst{-INFINITY, INT_MAX, 0};  // every slot is empty: any finite value beats it
  int gt = 0, ge = 0;

  auto consider = [&](float v, int g, bool valid) {
    if (valid) { gt += v > ts; ge += v >= ts; }
    if (k > 0) {
      // indices arrive in increasing order, so an equal value never displaces an earlier one
      unsigned m = __ballot_sync(0xffffffffu, valid && v > worst.v);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float cv = __shfl_sync(0xffffffffu, v, src);
        const int ci = __shfl_sync(0xffffffffu, g, src);
        if (cv > worst.v) {
          if (lane == 0) { lv[wid][worst.slot] = cv; li[wid][worst.slot] = ci; }
          __syncwarp();
          worst = warp_find_worst(lv[wid], li[wid], k, lane);
        }
      }
    }
  };

  // each warp owns a contiguous slice; 16-byte loads, four in flight per lane, when the row is aligned
  const int per = ceil_div(G, kRankWarps);
  const int beg = wid * per, end = min(G, beg + per);
  const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  int g = beg;
  if (aligned) {
    const int head = min(end, round_up(beg, 4));
    for (int base = g; base < head; base += 32) { const int i = base + lane; consider(i < head ? row[i] : 0.f, g0 + i, i < head); }
    g = head;
    const int nvec = (end - g) / 4;
    const float4* rv = reinterpret_cast<const float4*>(row + g);
    for (int base = 0; base < nvec; base += 128) {
      float4 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 32 + lane;
        ok[u] = i < nvec;
        v[u] = ok[u] ? __ldg(rv + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gi = g0 + g + 4 * (base + u * 32 + lane);
        consider(v[u].x, gi, ok[u]); consider(v[u].y, gi + 1, ok[u]); consider(v[u].z, gi + 2, ok[u]); consider(v[u].w, gi + 3, ok[u]);
      }
    }
    g += nvec * 4;
  }
  for (int base = g; base < end; base += 32) { const int i = base + lane; consider(i < end ? row[i] : 0.f, g0 + i, i < end); }

  gt = warp_sum(gt);
  ge = warp_sum(ge);
  if (lane == 0) { cnt[0][wid] = gt; cnt[1][wid] = ge; }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {  // slots left empty when G < k
    topk_val[(size_t)q * k + j] = -INFINITY;
    topk_idx[(size_t)q * k + j] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < kRankWarps; ++w) { a += cnt[0][w]; b += cnt[1][w]; }
    count_gt[q] = a;
    count_ge[q] = b;
  }
  // merge the per-warp lists: every candidate counts how many candidates beat it
  const int ncand = kRankWarps * k;
  for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
    const float v = lv[i / k][i % k];
    const int id = li[i / k][i % k];
    int r = 0;
    for (int w = 0; w < kRankWarps; ++w)
      for (int j = 0; j < k; ++j) r += beats(lv[w][j], li[w][j], v, id) ? 1 : 0;
    if (r < k) {
      topk_val[(size_t)q * k + r] = v;
      topk_idx[(size_t)q * k + r] = (id == INT_MAX) ? -1 : id;
    }
  }
}

// K9 final step: [P][Q][k] gathered lists -> global [Q][k].  Empty slots carry idx -1 / -inf.
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int P,
                                                         int Q, int k, float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int q = blockIdx.x, ncand = P * k;
  auto val_at = [&](int j) { return vals[((size_t)(j / k) * Q + q) * k + (j % k)]; };
  auto idx_at = [&](int j) { const int v = idx[((size_t)(j / k) * Q + q) * k + (j % k)]; return v < 0 ? INT_MAX : v; };
  for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
    const float v = val_at(i);
    const int id = idx_at(i);
    int r = 0;
    for (int j = 0; j < ncand; ++j) r += beats(val_at(j), idx_at(j), v, id) ? 1 : 0;
    if (r < k) {
      out_val[(size_t)q * k + r] = v;
      out_idx[(size_t)q * k + r] = (id == INT_MAX) ? -1 : id;
    }
  }
}

// out[q][order[j]] = in[q][j]: puts score columns that were computed shape group by shape group back into the
// caller's gallery order (reads coalesced, one 4-byte scatter per score).
__global__ void __launch_bounds__(256) scatter_columns_kernel(const float* __restrict__ in, int Q, int G, int ld_in,
                                                              const int32_t* __restrict__ order, float* __restrict__ out, int ld_out) {
  const size_t total = (size_t)Q * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t q = i / G;
    const int j = (int)(i - q * G);
    out[q * ld_out + __ldg(order + j)] = in[q * ld_in + j];
  }
}

}  // namespace sir

using namespace sir;

extern "C" int sir_scatter_columns(const float* d_in, int Q, int G, int ld_in, const int32_t* d_order, float* d_out, int ld_out,
                                   void* stream) {
  SIR_CHECK_ARG(d_in && d_order && d_out, "sir_scatter_columns: null pointer");
  SIR_CHECK_ARG(Q > 0 && G > 0 && ld_in >= G && ld_out >= G, "sir_scatter_columns: bad shape Q=%d G=%d ld=%d/%d", Q, G, ld_in, ld_out);
  const size_t total = (size_t)Q * G;
  const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  scatter_columns_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, Q, G, ld_in, d_order, d_out, ld_out);
  SIR_LAUNCH_CHECK("scatter_columns_kernel");
  return SIR_OK;
}

extern "C" int sir_true_scores(const float* d_scores, int Q, int G, int score_ld, const int32_t* d_true_idx, int g0,
                               float* d_true_score, void* stream) {
  SIR_CHECK_ARG(d_scores && d_true_idx && d_true_score, "sir_true_scores: null pointer");
  SIR_CHECK_ARG(Q > 0 && G > 0 && score_ld >= G, "sir_true_scores: bad shape Q=%d G=%d ld=%d", Q, G, score_ld);
  true_scores_kernel<<<ceil_div(Q, 128), 128, 0, (cudaStream_t)stream>>>(d_scores, Q, G, score_ld, d_true_idx, g0,
                                                                         d_true_score);
  SIR_LAUNCH_CHECK("true_scores_kernel");
  return SIR_OK;
}

extern "C" int sir_rank_topk(const float* d_scores, int Q, int G, int score_ld, const float* d_true_score, int g0, int k,
                             int32_t* d_count_gt, int32_t* d_count_ge, float* d_topk_val, int32_t* d_topk_idx,
                             void* stream) {
  SIR_CHECK_ARG(d_scores && d_true_score && d_count_gt && d_count_ge, "sir_rank_topk: null pointer");
  SIR_CHECK_ARG(Q > 0 && G > 0 && score_ld >= G, "sir_rank_topk: bad shape Q=%d G=%d ld=%d", Q, G, score_ld);
  SIR_CHECK_ARG(k >= 0 && k <= kMaxTopK, "sir_rank_topk: k=%d outside [0,%d]", k, kMaxTopK);
  SIR_CHECK_ARG(k == 0 || (d_topk_val && d_topk_idx), "sir_rank_topk: k>0 needs output lists");
  rank_topk_kernel<<<Q, kRankThreads, 0, (cudaStream_t)stream>>>(d_scores, Q, G, score_ld, d_true_score, g0, k,
                                                                 d_count_gt, d_count_ge, d_topk_val, d_topk_idx);
  SIR_LAUNCH_CHECK("rank_topk_kernel");
  return SIR_OK;
}

extern "C" int sir_merge_topk(const float* d_vals, const int32_t* d_idx, int P, int Q, int k, float* d_out_val,
                              int32_t* d_out_idx, void* stream) {
  SIR_CHECK_ARG(d_vals && d_idx && d_out_val && d_out_idx, "sir_merge_topk: null pointer");
  SIR_CHECK_ARG(P > 0 && Q > 0 && k > 0 && k <= kMaxTopK, "sir_merge_topk: bad shape P=%d Q=%d k=%d", P, Q, k);
  merge_topk_kernel<<<Q, 256, 0, (cudaStream_t)stream>>>(d_vals, d_idx, P, Q, k, d_out_val, d_out_idx);
  SIR_LAUNCH_CHECK("merge_topk_kernel");
  return SIR_OK;
}
