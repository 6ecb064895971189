// Retrieval metrics: segmented recall@k / AP / MRR / nDCG / precision@k reductions.
//
// (1) retrieval_metrics_kernel -- metrics of ranked top-K lists, restating the per-query
//     Python loops of /root/reference/notebooks/retrieval.ipynb:
//        calculate_mrr_at_k / calculate_precision_at_k   :310-324
//        AP  = (sum over hits of hits_so_far / rank) / |positives|   :411-420
//        nDCG (binary gains, log2)                                    :430-437
//        Recall@k = |top-k ∩ P| / |P|                                 :439-443
//     evaluated on the first K entries of the ranking (K = full gallery reproduces the
//     notebook exactly; K < N gives the @K variants).  |P| counts ALL ground-truth positives
//     of the query, also those absent from the gallery (as the notebook does).
// (2) ap_full_kernel -- average precision over a FULL score row per query, in both of the
//     reference's conventions:
//        sklearn.average_precision_score (ties grouped)   src/train.py:3285, src/auxiliary.py:200-224
//        ranking order with index tie-break               notebooks/retrieval.ipynb:411-420
//     by rank counting instead of sorting:  AP = (1/|P|) sum_p  tp(s_p) / #{j : s_j >= s_p}.
// (3) column means in a fixed order (deterministic, fp64).
//
// All per-query arithmetic is fp64, like the Python floats of the reference loops.  One warp per
// query; positives arrive as CSR (offsets, items).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int MT_WARPS = 4;
constexpr int MAX_KS = 8;

struct KsArg {
  int n;
  int k[MAX_KS];
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// per_query row layout: [mrr, ap, ndcg, (mrr@k, precision@k, recall@k) for each k]
__global__ void __launch_bounds__(MT_WARPS * 32)
retrieval_metrics_kernel(const int64_t* __restrict__ ranked, int64_t Q, int K, const int64_t* __restrict__ pos_off,
                         const int64_t* __restrict__ pos_items, const int32_t* __restrict__ n_pos_total, KsArg ks,
                         double* __restrict__ per_query) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * MT_WARPS + warp;
  if (q >= Q) return;
  const int64_t p0 = pos_off[q], p1 = pos_off[q + 1];
  const int n_in = (int)(p1 - p0);
  const int n_pos = n_pos_total != nullptr ? n_pos_total[q] : n_in;
  const int ncol = 3 + 3 * ks.n;
  double* out = per_query + q * ncol;

  int n_valid = 0;            // length of the ranking actually present (entries >= 0)
  int carry = 0;              // hits before the current block of 32 ranks
  int first_hit = 0x7fffffff; // rank (1-based) of the first hit
  double ap = 0.0, dcg = 0.0;
  int hits_at[MAX_KS];
#pragma unroll
  for (int i = 0; i < MAX_KS; ++i) hits_at[i] = 0;

  for (int base = 0; base < K; base += 32) {
    const int pos = base + lane;
    int64_t id = -1;
    if (pos < K) id = ranked[q * K + pos];
    int hit = 0;
    if (id >= 0) {
      for (int64_t t = p0; t < p1; ++t) hit |= (pos_items[t] == id);
    }
    n_valid += __popc(__ballot_sync(0xffffffffu, id >= 0));
    const int incl = warp_incl_scan(hit, lane) + carry;
    const int rank = pos + 1;
    if (hit) {
      ap += (double)incl / (double)rank;
      dcg += 1.0 / log2((double)rank + 1.0);
      first_hit = min(first_hit, rank);
#pragma unroll
      for (int i = 0; i < MAX_KS; ++i)
        if (i < ks.n && rank <= ks.k[i]) hits_at[i] += 1;
    }
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  ap = warp_sum(ap);
  dcg = warp_sum(dcg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first_hit = min(first_hit, __shfl_xor_sync(0xffffffffu, first_hit, o));
#pragma unroll
  for (int i = 0; i < MAX_KS; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hits_at[i] += __shfl_xor_sync(0xffffffffu, hits_at[i], o);
  }
  if (lane == 0) {
    double idcg = 0.0;
    for (int j = 0; j < n_pos; ++j) idcg += 1.0 / log2((double)j + 2.0);
    out[0] = first_hit != 0x7fffffff ? 1.0 / (double)first_hit : 0.0;
    out[1] = n_pos > 0 ? ap / (double)n_pos : 0.0;
    out[2] = idcg > 0.0 ? dcg / idcg : 0.0;
    for (int i = 0; i < ks.n; ++i) {
      const int k = ks.k[i];
      out[3 + 3 * i + 0] = (first_hit <= k) ? 1.0 / (double)first_hit : 0.0;
      out[3 + 3 * i + 1] = (k <= n_valid) ? (double)hits_at[i] / (double)k : 0.0;
      out[3 + 3 * i + 2] = n_pos > 0 ? (double)hits_at[i] / (double)n_pos : 0.0;
    }
  }
}

// AP over full score rows (higher score = better, i.e. pass -distance).  out[q] = NaN-free AP,
// valid[q] = 0 for rows that the reference skips (no in-range positive, NaN/inf score).
__global__ void __launch_bounds__(128)
ap_full_kernel(const float* __restrict__ scores, int64_t Q, int64_t N, const int64_t* __restrict__ pos_off,
               const int64_t* __restrict__ pos_items, int grouped_ties, double* __restrict__ ap_out,
               int32_t* __restrict__ valid_out) {
  const int64_t q = blockIdx.x;
  const float* row = scores + q * N;
  const int64_t p0 = pos_off[q], p1 = pos_off[q + 1];
  __shared__ double red[128];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  // rows with a NaN / inf score are skipped (src/train.py:3262)
  int mybad = 0;
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) mybad |= !isfinite(row[j]);
  if (mybad) bad = 1;
  __syncthreads();
  int n_pos = 0;
  for (int64_t t = p0; t < p1; ++t) n_pos += (pos_items[t] >= 0 && pos_items[t] < N);
  if (bad || n_pos == 0) {
    if (threadIdx.x == 0) { ap_out[q] = 0.0; valid_out[q] = 0; }
    return;
  }
  double acc = 0.0;
  for (int64_t t = p0; t < p1; ++t) {
    const int64_t pi = pos_items[t];
    if (pi < 0 || pi >= N) continue;            // block-uniform
    const float sp = row[pi];
    // #items ranked at or above this positive
    int cnt = 0;
    for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
      const float sj = row[j];
      cnt += grouped_ties ? (sj >= sp) : (sj > sp || (sj == sp && j <= pi));
    }
    red[threadIdx.x] = (double)cnt;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double total = red[0];
    __syncthreads();
    if (threadIdx.x == 0) {
      int tp = 0;                               // positives ranked at or above it
      for (int64_t u = p0; u < p1; ++u) {
        const int64_t pu = pos_items[u];
        if (pu < 0 || pu >= N) continue;
        const float su = row[pu];
        tp += grouped_ties ? (su >= sp) : (su > sp || (su == sp && pu <= pi));
      }
      acc += (double)tp / total;
    }
  }
  if (threadIdx.x == 0) {
    ap_out[q] = acc / (double)n_pos;
    valid_out[q] = 1;
  }
}

// means[c] = mean over rows (optionally only rows with valid != 0) of per_query[:, c]; fixed order.
__global__ void __launch_bounds__(256)
column_means_kernel(const double* __restrict__ per_query, const int32_t* __restrict__ valid, int64_t Q, int ncol,
                    double* __restrict__ means) {
  const int c = blockIdx.x;
  __shared__ double red[256];
  __shared__ int cnt[256];
  double s = 0.0;
  int n = 0;
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x) {
    if (valid == nullptr || valid[q]) { s += per_query[q * ncol + c]; n += 1; }
  }
  red[threadIdx.x] = s;
  cnt[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; cnt[threadIdx.x] += cnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) means[c] = cnt[0] > 0 ? red[0] / (double)cnt[0] : 0.0;
}

}  // namespace

int hypret_launch_retrieval_metrics(const int64_t* ranked, int64_t Q, int K, const int64_t* pos_off,
                                    const int64_t* pos_items, const int32_t* n_pos_total, const int32_t* ks_host,
                                    int n_ks, double* per_query, double* means, cudaStream_t stream) {
  if (n_ks < 0 || n_ks > MAX_KS) return HYPRET_EINVAL;
  KsArg ks;
  ks.n = n_ks;
  for (int i = 0; i < MAX_KS; ++i) ks.k[i] = i < n_ks ? ks_host[i] : 0;
  if (Q > 0) {
    const int64_t grid = (Q + MT_WARPS - 1) / MT_WARPS;
    retrieval_metrics_kernel<<<(unsigned)grid, MT_WARPS * 32, 0, stream>>>(ranked, Q, K, pos_off, pos_items,
                                                                          n_pos_total, ks, per_query);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (means != nullptr) {
    const int ncol = 3 + 3 * n_ks;
    column_means_kernel<<<ncol, 256, 0, stream>>>(per_query, nullptr, Q, ncol, means);
  }
  return (int)cudaGetLastError();
}

int hypret_launch_ap_full(const float* scores, int64_t Q, int64_t N, const int64_t* pos_off, const int64_t* pos_items,
                          int grouped_ties, double* ap, int32_t* valid, double* mean_ap, cudaStream_t stream) {
  if (Q > 0) {
    ap_full_kernel<<<(unsigned)Q, 128, 0, stream>>>(scores, Q, N, pos_off, pos_items, grouped_ties, ap, valid);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (mean_ap != nullptr) column_means_kernel<<<1, 256, 0, stream>>>(ap, valid, Q, 1, mean_ap);
  return (int)cudaGetLastError();
}
