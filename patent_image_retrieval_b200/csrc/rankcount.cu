// Exact full-ranking AP without the [Q,N] score matrix, and shardable (SURVEY.md 8e "collective 2").
//
// The reference computes AP over the FULL ranking of every query:
//   sklearn average_precision_score(target, -dist), ties grouped      /root/reference/src/train.py:3259-3293
//   argsort of the cosine row, AP = sum(hits_so_far / rank) / |P|     /root/reference/notebooks/retrieval.ipynb:383,411-420
// Both need, per (query, positive) pair, only RANK COUNTS: how many gallery rows score strictly
// better than the positive and how many tie with it.  Three kernels:
//   pair_keys_kernel     key of every (query, positive) pair whose gallery row lives on this shard
//   rank_count_kernel    one sweep over the shard: exact 64x64 distance tiles (the arithmetic of
//                        pairdist.cu, bit for bit) whose epilogue compares each entry with the keys of its
//                        row's positives and accumulates {#better, #tied with a lower index, #tied}
//   ap_from_counts_kernel  AP per query from the (all-reduced) counts, in either tie convention
// "key" = Poincare distance, or minus cosine similarity: smaller is better in both metrics.
// Multi-GPU: keys are summed over shards (exactly one shard owns a positive), counts are summed over
// shards; nothing else crosses NVLink.  FP32-FMA bound like pairdist.cu.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int RC_TILE = 64;
constexpr int RC_K = 16;
constexpr int RC_PCAP = 1024;   // positives of one 64-query tile handled per sweep (more -> another sweep)

// key of one pair from the three fp32 sums; identical expression to pairdist.cu's epilogue
__device__ __forceinline__ float key_from_sums(float acc, float na, float nb, float dot, float c, int metric) {
  if (metric == HYPRET_METRIC_HYPERBOLIC) {
    const double cc = (double)c, rs = 1.0 / sqrt(cc);
    const double al = 1.0 - cc * (double)na, be = 1.0 - cc * (double)nb;
    const double t = 2.0 * cc * (double)acc / (al * be);
    return (float)(log1p(t + sqrt(t * (t + 2.0))) * rs);
  }
  const double nx = sqrt((double)na), ny = sqrt((double)nb);
  return (float)(-((double)dot / ((nx == 0.0 ? 1.0 : nx) * (ny == 0.0 ? 1.0 : ny))));
}

// One thread per (query, positive) pair, k ascending, the same fmaf chain a tile thread runs.
__global__ void __launch_bounds__(128)
pair_keys_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int64_t Q, int64_t n_local, int d, float c,
                 int metric, const int64_t* __restrict__ pos_off, const int64_t* __restrict__ pos_items,
                 int64_t idx_offset, float* __restrict__ keys) {
  const int64_t q = blockIdx.x;
  if (q >= Q) return;
  for (int64_t t = pos_off[q] + threadIdx.x; t < pos_off[q + 1]; t += blockDim.x) {
    const int64_t j = pos_items[t] - idx_offset;
    if (j < 0 || j >= n_local) { keys[t] = 0.f; continue; }     // another shard owns it (or invalid id)
    const float* a = q32 + q * d;
    const float* b = g32 + j * d;
    float acc = 0.f, na = 0.f, nb = 0.f, dot = 0.f;
    for (int k = 0; k < d; ++k) {
      const float av = a[k], bv = b[k];
      na = fmaf(av, av, na);
      const float e = av - bv;
      acc = fmaf(e, e, acc);
      dot = fmaf(av, bv, dot);
      nb = fmaf(bv, bv, nb);
    }
    keys[t] = key_from_sums(acc, na, nb, dot, c, metric);
  }
}

__global__ void __launch_bounds__(256)
rank_count_kernel(const float* __restrict__ a, const float* __restrict__ p, int64_t n, int64_t m, int d, float c,
                  int metric, const int64_t* __restrict__ pos_off, const int64_t* __restrict__ pos_items,
                  const float* __restrict__ pos_keys, int64_t idx_offset, int64_t cols_per_cta,
                  unsigned long long* __restrict__ counts, int32_t* __restrict__ bad) {
  __shared__ float As[RC_K][RC_TILE + 4];
  __shared__ float Ps[RC_K][RC_TILE + 4];
  __shared__ float s_key[RC_PCAP];
  __shared__ int64_t s_gid[RC_PCAP];
  __shared__ int s_cnt[RC_PCAP][3];
  __shared__ int s_rowptr[RC_TILE + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * RC_TILE;
  const int64_t jb = (int64_t)blockIdx.x * cols_per_cta;
  const int64_t je = jb + cols_per_cta < m ? jb + cols_per_cta : m;
  const int rows = (int)(n - i0 < RC_TILE ? n - i0 : RC_TILE);
  const int64_t pbase = pos_off[i0], pend = pos_off[i0 + rows];
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;

  for (int64_t pc = pbase; pc < pend; pc += RC_PCAP) {     // one sweep of the CTA's columns per chunk of positives
    const int np = (int)(pend - pc < RC_PCAP ? pend - pc : RC_PCAP);
    __syncthreads();
    for (int t = threadIdx.x; t < np; t += 256) {
      s_key[t] = pos_keys[pc + t];
      s_gid[t] = pos_items[pc + t];
      s_cnt[t][0] = s_cnt[t][1] = s_cnt[t][2] = 0;
    }
    for (int r = threadIdx.x; r <= RC_TILE; r += 256) {      // positives of row r inside this chunk: [rowptr[r], rowptr[r+1])
      int64_t o = (r <= rows ? pos_off[i0 + r] : pend) - pc;
      o = o < 0 ? 0 : (o > np ? np : o);
      s_rowptr[r] = (int)o;
    }
    __syncthreads();
    for (int64_t j0 = jb; j0 < je; j0 += RC_TILE) {
      float acc[4][4], dot[4][4], na[4], nb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        na[r] = 0.f;
        nb[r] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc[r][s] = 0.f; dot[r][s] = 0.f; }
      }
      for (int k0 = 0; k0 < d; k0 += RC_K) {
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vp = va;
        if (i0 + lr < n && k0 + lk < d) va = *reinterpret_cast<const float4*>(a + (i0 + lr) * d + k0 + lk);
        if (j0 + lr < je && k0 + lk < d) vp = *reinterpret_cast<const float4*>(p + (j0 + lr) * d + k0 + lk);
        __syncthreads();
        As[lk + 0][lr] = va.x; As[lk + 1][lr] = va.y; As[lk + 2][lr] = va.z; As[lk + 3][lr] = va.w;
        Ps[lk + 0][lr] = vp.x; Ps[lk + 1][lr] = vp.y; Ps[lk + 2][lr] = vp.z; Ps[lk + 3][lr] = vp.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RC_K; ++k) {
          float av[4], pv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) av[r] = As[k][ty * 4 + r];
#pragma unroll
          for (int s = 0; s < 4; ++s) pv[s] = Ps[k][tx * 4 + s];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            na[r] = fmaf(av[r], av[r], na[r]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              if (metric == HYPRET_METRIC_HYPERBOLIC) {
                const float e = av[r] - pv[s];
                acc[r][s] = fmaf(e, e, acc[r][s]);
              } else {
                dot[r][s] = fmaf(av[r], pv[s], dot[r][s]);
              }
            }
          }
#pragma unroll
          for (int s = 0; s < 4; ++s) nb[s] = fmaf(pv[s], pv[s], nb[s]);
        }
      }
      // ---- counting epilogue -------------------------------------------------------------------------
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int lrow = ty * 4 + r;
        if (lrow >= rows) continue;
        float key[4];
        bool ok[4];
        int n_bad = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          ok[s] = j0 + tx * 4 + s < je;
          key[s] = key_from_sums(acc[r][s], na[r], nb[s], dot[r][s], c, metric);
          n_bad += (ok[s] && !isfinite(key[s])) ? 1 : 0;
        }
        if (n_bad > 0 && pc == pbase) atomicAdd(&bad[i0 + lrow], n_bad);
        for (int t = s_rowptr[lrow]; t < s_rowptr[lrow + 1]; ++t) {
          const float kp = s_key[t];
          const int64_t gp = s_gid[t];
          int lt = 0, eq = 0, eq_lo = 0;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const int64_t gj = idx_offset + j0 + tx * 4 + s;
            lt += (ok[s] && key[s] < kp) ? 1 : 0;
            eq += (ok[s] && key[s] == kp) ? 1 : 0;
            eq_lo += (ok[s] && key[s] == kp && gj < gp) ? 1 : 0;
          }
          if (lt) atomicAdd(&s_cnt[t][0], lt);
          if (eq_lo) atomicAdd(&s_cnt[t][1], eq_lo);
          if (eq) atomicAdd(&s_cnt[t][2], eq);
        }
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * 3; t += 256) {
      const int v = s_cnt[t / 3][t % 3];
      if (v) atomicAdd(&counts[(pc + t / 3) * 3 + t % 3], (unsigned long long)v);
    }
  }
  // rows without any positive still need their non-finite scores counted (they decide `valid`)
  if (pbase == pend) {
    // nothing to rank for this tile: AP is undefined for these queries, bad[] is irrelevant
  }
}

// AP per query from global rank counts.  grouped_ties != 0: sklearn semantics (src/train.py:3285);
// == 0: ranking order with lower-index tie-break (retrieval.ipynb:411-420).  Mirrors ap_full_kernel.
__global__ void __launch_bounds__(128)
ap_from_counts_kernel(const int64_t* __restrict__ pos_off, const int64_t* __restrict__ pos_items,
                      const float* __restrict__ pos_keys, const unsigned long long* __restrict__ counts,
                      const int32_t* __restrict__ bad, int64_t Q, int64_t n_total, int grouped_ties,
                      double* __restrict__ ap_out, int32_t* __restrict__ valid_out) {
  const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int64_t p0 = pos_off[q], p1 = pos_off[q + 1];
  int n_pos = 0;
  for (int64_t t = p0 + lane; t < p1; t += 32) n_pos += (pos_items[t] >= 0 && pos_items[t] < n_total);
  n_pos = (int)warp_sum((float)n_pos);
  if (n_pos == 0 || (bad != nullptr && bad[q] != 0)) {
    if (lane == 0) { ap_out[q] = 0.0; valid_out[q] = 0; }
    return;
  }
  double acc = 0.0;
  for (int64_t t = p0 + lane; t < p1; t += 32) {
    const int64_t pi = pos_items[t];
    if (pi < 0 || pi >= n_total) continue;
    const float kp = pos_keys[t];
    const double total = grouped_ties ? (double)(counts[t * 3 + 0] + counts[t * 3 + 2])
                                      : (double)(counts[t * 3 + 0] + counts[t * 3 + 1] + 1ull);
    int tp = 0;     // positives ranked at or above this one
    for (int64_t u = p0; u < p1; ++u) {
      const int64_t pu = pos_items[u];
      if (pu < 0 || pu >= n_total) continue;
      const float ku = pos_keys[u];
      tp += grouped_ties ? (ku <= kp) : (ku < kp || (ku == kp && pu <= pi));
    }
    acc += (double)tp / total;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    ap_out[q] = acc / (double)n_pos;
    valid_out[q] = 1;
  }
}

__global__ void __launch_bounds__(256)
masked_mean_kernel(const double* __restrict__ v, const int32_t* __restrict__ valid, int64_t Q, double* __restrict__ out) {
  __shared__ double red[256];
  __shared__ int cnt[256];
  double s = 0.0;
  int n = 0;
  for (int64_t q = threadIdx.x; q < Q; q += blockDim.x)
    if (valid[q]) { s += v[q]; n += 1; }
  red[threadIdx.x] = s;
  cnt[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; cnt[threadIdx.x] += cnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = cnt[0] > 0 ? red[0] / (double)cnt[0] : 0.0;
}

}  // namespace

int hypret_launch_pair_keys(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                            const int64_t* pos_off, const int64_t* pos_items, int64_t idx_offset, float* keys,
                            cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  pair_keys_kernel<<<(unsigned)Q, 128, 0, stream>>>(q32, g32, Q, n_local, d, c, metric, pos_off, pos_items, idx_offset,
                                                   keys);
  return (int)cudaGetLastError();
}

int hypret_launch_rank_count(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                             const int64_t* pos_off, const int64_t* pos_items, const float* pos_keys,
                             int64_t idx_offset, unsigned long long* counts, int32_t* bad, cudaStream_t stream) {
  if (Q == 0 || n_local == 0) return HYPRET_OK;
  const int64_t row_tiles = (Q + RC_TILE - 1) / RC_TILE;
  if (row_tiles > 65535) return HYPRET_EUNSUPPORTED;
  // column splits: enough CTAs for ~4 waves of 148 SMs x 2 resident CTAs, each a whole number of 64-column tiles
  int64_t col_tiles = (n_local + RC_TILE - 1) / RC_TILE;
  int64_t splits = (148 * 8 + row_tiles - 1) / row_tiles;
  if (splits > col_tiles) splits = col_tiles;
  if (splits < 1) splits = 1;
  const int64_t cols_per_cta = ((col_tiles + splits - 1) / splits) * RC_TILE;
  splits = (n_local + cols_per_cta - 1) / cols_per_cta;
  dim3 grid((unsigned)splits, (unsigned)row_tiles);
  rank_count_kernel<<<grid, 256, 0, stream>>>(q32, g32, Q, n_local, d, c, metric, pos_off, pos_items, pos_keys,
                                              idx_offset, cols_per_cta, counts, bad);
  return (int)cudaGetLastError();
}

int hypret_launch_ap_from_counts(const int64_t* pos_off, const int64_t* pos_items, const float* pos_keys,
                                 const unsigned long long* counts, const int32_t* bad, int64_t Q, int64_t n_total,
                                 int grouped_ties, double* ap, int32_t* valid, double* mean_ap, cudaStream_t stream) {
  if (Q > 0) {
    ap_from_counts_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, stream>>>(pos_off, pos_items, pos_keys, counts, bad, Q,
                                                                     n_total, grouped_ties, ap, valid);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (mean_ap != nullptr) masked_mean_kernel<<<1, 256, 0, stream>>>(ap, valid, Q, mean_ap);
  return (int)cudaGetLastError();
}
