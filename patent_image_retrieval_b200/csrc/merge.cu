// k-way merge of per-shard top-k lists (multi-GPU exchange step).
//
// After every rank has searched its gallery row-shard, the ranks all-gather their [Q,k]
// (score, global index) lists; this kernel merges the W lists of each query into the global
// top-k.  The reference has no multi-GPU path (single process); the merge
// implements the same ordering as its single-device ranking (np.argsort / torch.topk,
// notebooks/retrieval.ipynb:383, src/auxiliary.py:374): ascending distance or descending
// similarity, ties -> lower gallery index.  One warp per query, k rounds of warp arg-min.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int MG_WARPS = 4;

__global__ void __launch_bounds__(MG_WARPS * 32)
merge_topk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int W, int64_t Q, int k,
                  int descending, float* __restrict__ out_score, int64_t* __restrict__ out_idx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * MG_WARPS + warp;
  if (q >= Q) return;
  const int n = W * k;                       // <= 8 * 32 candidates: up to 8 per lane, kept in registers
  float s[8];
  int64_t id[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int c = lane + 32 * t;
    if (c < n) {
      const int w = c / k, j = c - w * k;
      const float v = scores[((int64_t)w * Q + q) * k + j];
      s[t] = descending ? -v : v;
      id[t] = idx[((int64_t)w * Q + q) * k + j];
    } else {
      s[t] = INFINITY;
      id[t] = -1;
    }
  }
  for (int r = 0; r < k; ++r) {
    float bs = INFINITY;
    int64_t bi = -1;
    int bt = -1;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (id[t] >= 0 && (bt < 0 || s[t] < bs || (s[t] == bs && id[t] < bi))) { bs = s[t]; bi = id[t]; bt = t; }
    }
    int owner = lane;
    float ws = bs;
    int64_t wi = bi;
    int valid = bt >= 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      const int ov = __shfl_xor_sync(0xffffffffu, valid, o);
      const bool take = ov && (!valid || os < ws || (os == ws && (oi < wi || (oi == wi && oo < owner))));
      if (take) { ws = os; wi = oi; owner = oo; valid = 1; }
    }
    if (lane == 0) {
      out_score[q * k + r] = valid ? (descending ? -ws : ws) : (descending ? -INFINITY : INFINITY);
      out_idx[q * k + r] = valid ? wi : (int64_t)-1;
    }
    if (valid && owner == lane) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t == bt) id[t] = -1;
    }
  }
}

// out[q] = the kth smallest (1-based) of the n_parts * m values of query q, vals [n_parts, Q, m]; +inf when
// fewer than kth finite values exist.  Rank counting in shared memory, one warp per query (n <= 2048).
__global__ void __launch_bounds__(MG_WARPS * 32)
kth_smallest_kernel(const float* __restrict__ vals, int n_parts, int64_t Q, int m, int kth, float* __restrict__ out,
                    const PeerRoute route, int64_t out_off) {
  extern __shared__ float sm_vals[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * MG_WARPS + warp;
  if (q >= Q) return;
  const int n = n_parts * m;
  float* v = sm_vals + (size_t)warp * n;
  for (int t = lane; t < n; t += 32) {
    const int w = t / m, j = t - w * m;
    const float x = vals[((int64_t)w * Q + q) * m + j];
    v[t] = (x == x) ? x : INFINITY;            // NaN counts as +inf
  }
  __syncwarp();
  // routed: the all_gather of the thresholds -- the value also lands at [me*Q + q] of every rank's region
  if (kth > n) {
    if (lane == 0) {
      out[q] = INFINITY;
      for (int r = 0; r < route.n; ++r)
        reinterpret_cast<float*>(route.base[r] + out_off)[(int64_t)route.me * Q + q] = INFINITY;
    }
    return;
  }
  for (int t = lane; t < n; t += 32) {
    const float x = v[t];
    int rank = 0;
    for (int u = 0; u < n; ++u) {
      const float y = v[u];
      rank += (y < x) || (y == x && u < t);
    }
    if (rank == kth - 1) {                     // exactly one t has this rank
      out[q] = x;
      for (int r = 0; r < route.n; ++r)
        reinterpret_cast<float*>(route.base[r] + out_off)[(int64_t)route.me * Q + q] = x;
    }
  }
}

}  // namespace

int hypret_launch_kth_smallest(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out,
                               const hypret_peer_route* route, int64_t out_off, cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  const size_t smem = (size_t)MG_WARPS * n_parts * m * sizeof(float);
  kth_smallest_kernel<<<(unsigned)((Q + MG_WARPS - 1) / MG_WARPS), MG_WARPS * 32, smem, stream>>>(
      vals, n_parts, Q, m, kth, out, make_route(route), out_off);
  return (int)cudaGetLastError();
}

int hypret_launch_merge_topk(const float* scores, const int64_t* idx, int W, int64_t Q, int k, int descending,
                             float* out_score, int64_t* out_idx, cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  const int64_t grid = (Q + MG_WARPS - 1) / MG_WARPS;
  merge_topk_kernel<<<(unsigned)grid, MG_WARPS * 32, 0, stream>>>(scores, idx, W, Q, k, descending, out_score,
                                                                  out_idx);
  return (int)cudaGetLastError();
}
