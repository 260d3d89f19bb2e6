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
constexpr int kMaxTopK = 128;  // <= kSelCap - kSelChunk: the k best always fit next to one chunk of new keys

__global__ void true_scores_kernel(const float* __restrict__ scores, int Q, int G, int ld,
                                   const int32_t* __restrict__ true_idx, int g0, float* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int t = true_idx[q] - g0;
  out[q] = (t >= 0 && t < G) ? scores[(size_t)q * ld + t] : -INFINITY;
}

// (value desc, index asc) strict order
__device__ __forceinline__ bool beats(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

// Top-k by threshold + compaction.  One CTA streams one score row (16-byte loads).  Every value is turned into a
// 64-bit key whose unsigned order is (value descending, index ascending); keys above the CTA's current threshold are
// appended to a shared-memory buffer (one shared-memory atomicAdd per candidate; candidates are rare after the first cut).  When the buffer could
// overflow during the next chunk it is sorted (bitonic, in place), cut to the k best, and the k-th key becomes the new
// threshold -- after the first cut only ~k/2048 of the values pass, so a 100,000-column row is cut two or three times
// and the kernel stays a single streaming pass over HBM.
constexpr int kSelCap = 2048;                  // keys the buffer holds
constexpr int kSelChunk = kRankThreads * 4;    // values per iteration (one float4 per thread)

__device__ __forceinline__ unsigned long long select_key(float v, int idx) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // unsigned order == float order
  return ((unsigned long long)u << 32) | (unsigned)(0x7fffffff - idx);
}
__device__ __forceinline__ float key_value(unsigned long long key) {
  unsigned u = (unsigned)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// Cuts buf[0..n) (n > k) to its k largest keys, unordered, and sets *thr to the smallest of them.  The k-th largest key
// is found by 8-bit radix passes from the top (a 256-bin shared-memory histogram of the keys that still match the
// prefix, a block scan from the top bin down); the walk stops as soon as the bin that holds the k-th key is needed
// whole -- usually after two to four passes, never more than eight because keys are unique (they carry the index).
// ~500 instructions per thread against ~2,600 for a bitonic sort of 2,048 keys.
struct SelectScratch {
  int hist[256];
  int warp_tot[kRankWarps];
  int bin, above, bin_cnt, kept_cnt;
  unsigned long long kept_min;
  unsigned long long tmp[kMaxTopK];
};

__device__ void select_cut(unsigned long long* buf, int n, int k, unsigned long long* thr, SelectScratch* sc) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned long long prefix = 0ull;
  int need = k, shift = 56;
  for (;; shift -= 8) {
    sc->hist[tid] = 0;  // kRankThreads == 256 bins
    __syncthreads();
    for (int i = tid; i < n; i += kRankThreads) {
      const unsigned long long key = buf[i];
      if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&sc->hist[(int)((key >> shift) & 255ull)], 1);
    }
    __syncthreads();
    // thread t owns bin 255 - t: inclusive prefix over t = number of matching keys in bins >= 255 - t
    const int h = sc->hist[255 - tid];
    int inc = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) sc->warp_tot[wid] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; ++w) base += sc->warp_tot[w];
    inc += base;
    if (inc >= need && inc - h < need) {  // exactly one thread: the bin that holds the need-th key from the top
      sc->bin = 255 - tid;
      sc->above = inc - h;
      sc->bin_cnt = h;
    }
    __syncthreads();
    prefix |= (unsigned long long)sc->bin << shift;
    need -= sc->above;
    const bool whole = sc->bin_cnt == need;  // every key of that bin is among the k best: the prefix is the boundary
    __syncthreads();
    if (whole || shift == 0) break;
  }
  // keep the keys >= prefix (exactly k of them)
  if (tid == 0) {
    sc->kept_cnt = 0;
    sc->kept_min = ~0ull;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kRankThreads) {
    const unsigned long long key = buf[i];
    if (key >= prefix) {
      const int pos = atomicAdd(&sc->kept_cnt, 1);
      if (pos < kMaxTopK) sc->tmp[pos] = key;
      atomicMin(&sc->kept_min, key);
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kRankThreads) buf[i] = sc->tmp[i];
  if (tid == 0) *thr = sc->kept_min;
  __syncthreads();
}

// sorts buf[0..n) descending (n <= kSelCap), keeps the first min(n, k) keys; returns the new count
__device__ int select_compact(unsigned long long* buf, int n, int k, unsigned long long* thr) {
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) buf[i] = 0ull;  // below every real key
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = buf[lo], b = buf[hi];
        if ((a < b) == desc) {
          buf[lo] = b;
          buf[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  const int kept = min(n, k);
  if (threadIdx.x == 0 && n >= k && k > 0) *thr = buf[k - 1];
  __syncthreads();
  return kept;
}

// Streaming loop of the top-k kernel.  Until the first cut every value is a candidate, so the row is walked one chunk
// (kSelChunk values) per barrier.  Afterwards only ~k/2048 of the values pass the threshold and the loop takes kSelSpan
// chunks per barrier with all of their 16-byte loads in flight at once; should more candidates arrive than the buffer
// holds (a row that keeps improving, e.g. sorted ascending) the span is discarded and redone chunk by chunk.
constexpr int kSelSpan = 8;

__global__ void __launch_bounds__(kRankThreads) rank_topk_kernel(const float* __restrict__ scores, int Q, int G, int ld,
                                                                 const float* __restrict__ true_score, int g0, int k,
                                                                 int32_t* __restrict__ count_gt, int32_t* __restrict__ count_ge,
                                                                 float* __restrict__ topk_val, int32_t* __restrict__ topk_idx) {
  __shared__ unsigned long long buf[kSelCap];
  __shared__ SelectScratch scratch;
  __shared__ unsigned long long thr_s;
  __shared__ int cnt_s, overflow_s;
  __shared__ int red[2][kRankWarps];
  const int q = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* row = scores + (size_t)q * ld;
  const float ts = true_score[q];
  if (threadIdx.x == 0) {
    thr_s = 0ull;
    cnt_s = 0;
    overflow_s = 0;
  }
  __syncthreads();
  const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  int gt = 0, ge = 0, cnt = 0;
  auto load4 = [&](int i0, float (&v)[4]) {
    if (aligned && i0 + 3 < G) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(row + i0));
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (i0 + j < G) ? __ldg(row + i0 + j) : 0.0f;
    }
  };
  // candidates of one float4 against the threshold the span started with; appends beyond the buffer set the overflow flag
  auto consider4 = [&](const float (&v)[4], int i0, unsigned long long thr, float thr_v, int& gts, int& ges) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool valid = i0 + j < G;
      gts += valid && v[j] > ts;
      ges += valid && v[j] >= ts;
      if (valid && (thr == 0ull || v[j] >= thr_v)) {
        const unsigned long long key = select_key(v[j], g0 + i0 + j);
        if (key > thr) {
          const int pos = atomicAdd(&cnt_s, 1);
          if (pos < kSelCap) buf[pos] = key;
          else overflow_s = 1;
        }
      }
    }
  };
  int slow_left = 0;  // chunks still to be walked one per barrier after an overflow
  for (int base = 0; base < G;) {
    const unsigned long long thr = thr_s;
    const float thr_v = key_value(thr);
    const bool fast = thr != 0ull && slow_left == 0 && k > 0;
    const int span = fast ? kSelSpan : 1;
    int gts = 0, ges = 0;
    if (fast) {
      float v[kSelSpan][4];
#pragma unroll
      for (int c = 0; c < kSelSpan; ++c) load4(base + c * kSelChunk + 4 * threadIdx.x, v[c]);
#pragma unroll
      for (int c = 0; c < kSelSpan; ++c) consider4(v[c], base + c * kSelChunk + 4 * threadIdx.x, thr, thr_v, gts, ges);
    } else {
      float v[4];
      load4(base + 4 * threadIdx.x, v);
      if (k > 0) {
        consider4(v, base + 4 * threadIdx.x, thr, thr_v, gts, ges);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool valid = base + 4 * threadIdx.x + j < G;
          gts += valid && v[j] > ts;
          ges += valid && v[j] >= ts;
        }
      }
    }
    __syncthreads();
    if (overflow_s) {  // block-uniform: forget this span and walk it again one chunk per barrier
      __syncthreads();
      if (threadIdx.x == 0) {
        cnt_s = cnt;
        overflow_s = 0;
      }
      slow_left = kSelSpan;
      __syncthreads();
      continue;
    }
    gt += gts;
    ge += ges;
    base += span * kSelChunk;
    if (slow_left > 0) --slow_left;
    cnt = min(cnt_s, kSelCap);
    if (k > 0 && cnt > kSelCap - kSelChunk) {  // the next chunk could overflow: cut to the k best, raise the threshold
      select_cut(buf, cnt, k, &thr_s, &scratch);
      cnt = k;
      if (threadIdx.x == 0) cnt_s = cnt;
      __syncthreads();
    }
  }
  gt = warp_sum(gt);
  ge = warp_sum(ge);
  if (lane == 0) {
    red[0][wid] = gt;
    red[1][wid] = ge;
  }
  if (k > 0) {
    cnt = cnt_s;
    if (cnt > k) {
      select_cut(buf, cnt, k, &thr_s, &scratch);
      cnt = k;
    }
    cnt = select_compact(buf, cnt, k, &thr_s);  // order the (at most k) survivors; also the barrier for red[]
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < kRankWarps; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    count_gt[q] = a;
    count_ge[q] = b;
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool have = j < cnt;
    topk_val[(size_t)q * k + j] = have ? key_value(buf[j]) : -INFINITY;
    topk_idx[(size_t)q * k + j] = have ? (int)(0x7fffffff - (unsigned)(buf[j] & 0xffffffffu)) : -1;
  }
}

// k == 0: counts only -- no candidate buffer, so many rows are resident per SM and four 16-byte loads are in flight per thread.
__global__ void __launch_bounds__(kRankThreads) rank_count_kernel(const float* __restrict__ scores, int Q, int G, int ld,
                                                                  const float* __restrict__ true_score, int32_t* __restrict__ count_gt,
                                                                  int32_t* __restrict__ count_ge) {
  __shared__ int red[2][kRankWarps];
  const int q = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* row = scores + (size_t)q * ld;
  const float ts = true_score[q];
  int gt = 0, ge = 0, g = 0;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    const int nvec = G / 4;
    const float4* rv = reinterpret_cast<const float4*>(row);
    for (int base = 0; base < nvec; base += 4 * kRankThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * kRankThreads + threadIdx.x;
        v[u] = i < nvec ? __ldg(rv + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        gt += (v[u].x > ts) + (v[u].y > ts) + (v[u].z > ts) + (v[u].w > ts);
        ge += (v[u].x >= ts) + (v[u].y >= ts) + (v[u].z >= ts) + (v[u].w >= ts);
      }
    }
    g = nvec * 4;
  }
  for (int i = g + threadIdx.x; i < G; i += kRankThreads) {
    const float v = __ldg(row + i);
    gt += v > ts;
    ge += v >= ts;
  }
  gt = warp_sum(gt);
  ge = warp_sum(ge);
  if (lane == 0) {
    red[0][wid] = gt;
    red[1][wid] = ge;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < kRankWarps; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    count_gt[q] = a;
    count_ge[q] = b;
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
  if (k == 0) {
    rank_count_kernel<<<Q, kRankThreads, 0, (cudaStream_t)stream>>>(d_scores, Q, G, score_ld, d_true_score, d_count_gt, d_count_ge);
  } else {
    rank_topk_kernel<<<Q, kRankThreads, 0, (cudaStream_t)stream>>>(d_scores, Q, G, score_ld, d_true_score, g0, k,
                                                                   d_count_gt, d_count_ge, d_topk_val, d_topk_idx);
  }
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
