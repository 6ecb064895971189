// Exact top-k by full scan: the guarantee behind the tensor-core filter.
//
// hypret_rerank_cert certifies, per query, that no gallery row outside the fp16-filtered candidate set can precede the
// k-th result (margin > rounding-error bound E, csrc/rerank.cu).  Queries it cannot certify -- near-duplicate
// galleries, where more than k' - k rows sit inside the fp16 rounding band of the k-th best -- are listed on the device,
// and this kernel recomputes their top-k from ALL gallery rows with the arithmetic of the rerank kernel (explicit
// differences, fp64 accumulation in the same order, arccosh closed form == pmath.dist,
// /root/reference/src/train.py:3259; cosine of notebooks/retrieval.ipynb:368), i.e. exactly what the reference's
// per-query loop + torch.topk (src/auxiliary.py:374) does, and overwrites the query's result rows.  No host
// round trip: the list length is read on the device, the grid is fixed, and with an empty list every CTA exits at once.
//
// Work decomposition: grid = (row chunks, slots).  CTA (c, s) serves GROUPS of up to 4 list entries (a gallery row is
// read once per group); for each group it scans gallery rows [c * chunk, (c+1) * chunk): 8 warps take groups of 32 rows
// (4 rows per pass with all their 128-bit loads issued before the first use; an fp32 lower bound decides whether a row
// needs the fp64 arithmetic at all), each warp keeps per query its best k as packed 64-bit keys (ordered fp32 score << 32 | row id: one unsigned compare = the (score, id) order of
// the rerank), one per lane, sorted.  The CTA's list is merged into the query's result rows under a per-query spin
// lock (the holder never waits on anything, so the lock cannot deadlock whatever part of the grid is resident); the
// first merger of a query discards the filtered result that is there.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int EX_WARPS = 8;
constexpr unsigned long long EX_NONE = ~0ull;

__device__ __forceinline__ unsigned ex_ordered_key(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ex_ordered_val(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// sorted insert of x into the warp's list (lane r = r-th best, ascending); x is warp-uniform
__device__ __forceinline__ void ex_insert(unsigned long long& mine, unsigned long long x, int k, int lane) {
  const unsigned long long worst = __shfl_sync(0xffffffffu, mine, k - 1);
  if (x >= worst) return;                                    // warp-uniform
  const int pos = __popc(__ballot_sync(0xffffffffu, mine < x));
  const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
  mine = lane < pos ? mine : (lane == pos ? x : up);
}

// the k smallest of two ascending lists (one element per lane each): bitonic merge
__device__ __forceinline__ unsigned long long ex_merge(unsigned long long a, unsigned long long b, int lane) {
  const unsigned long long rev = __shfl_sync(0xffffffffu, b, 31 - lane);
  unsigned long long v = a < rev ? a : rev;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
    const bool take_min = (lane & j) == 0;
    v = take_min ? (v < o ? v : o) : (v < o ? o : v);
  }
  return v;
}

// QB listed queries share every gallery row a warp reads, and a row reaches the fp64 arithmetic only if an fp32
// lower bound of its key can still enter the query's list: per (row, query) the warp forms ||x - y||^2 (or <x,y>) in fp32
// -- 21 roundings of non-negative terms, relative error < 2e-6 -- and compares it with the list's k-th best turned back
// into a squared-distance threshold (d = arccosh(1 + 2 c s / (alpha beta)) / sqrt(c) is increasing in s).  The scan is
// then HBM-bound (N * d * 4 bytes per pass over the shard, whatever QB is); before, the float -> double conversions of
// every element (a quarter-rate pipe) made it 0.6 ms per query and 5M rows.
template <int NV, int QB>
__global__ void __launch_bounds__(EX_WARPS * 32, NV <= 4 ? 2 : 1)
exact_topk_kernel(const float* __restrict__ q32, const float* __restrict__ g32, const double* __restrict__ g_sq64,
                  int64_t N, int d, float c, int metric, int k, int64_t idx_offset,
                  const int32_t* __restrict__ q_list, const int32_t* __restrict__ q_count, int32_t* state,
                  float* out_score, int64_t* out_idx, int64_t chunk, const unsigned long long* __restrict__ after,
                  const float* __restrict__ init_bound) {
  __shared__ unsigned long long lists[EX_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = *q_count;
  const int64_t r_lo = (int64_t)blockIdx.x * chunk;
  const int64_t r_hi = r_lo + chunk < N ? r_lo + chunk : N;
  const int nvec = d >> 2;
  const double cc = (double)c;
  const bool hyp = metric == HYPRET_METRIC_HYPERBOLIC;
  for (int it0 = blockIdx.y * QB; it0 < count; it0 += gridDim.y * QB) {
    const int nq = count - it0 < QB ? count - it0 : QB;      // queries of this group (block-uniform)
    float4 qv[QB][NV];
    double xsq[QB];
    unsigned long long aft[QB];                              // paging: only rows whose key is ABOVE this one are taken
#pragma unroll
    for (int u = 0; u < QB; ++u) {
      const int64_t qid = q_list[it0 + (u < nq ? u : 0)];    // pad the group with its first query (never inserted)
      aft[u] = (after != nullptr && u < nq) ? after[qid] : 0ull;      // indexed by QUERY id, like every per-query array
      HYPRET_CHECK(qid >= 0 && it0 + u < count + QB);
      const float4* qrow = reinterpret_cast<const float4*>(q32 + qid * d);
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int j = i * 32 + lane;
        qv[u][i] = (j < nvec) ? __ldg(qrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc += (double)qv[u][i].x * qv[u][i].x + (double)qv[u][i].y * qv[u][i].y + (double)qv[u][i].z * qv[u][i].z +
               (double)qv[u][i].w * qv[u][i].w;
      }
      xsq[u] = warp_sum(acc);
    }
    unsigned long long mine[QB];                             // per query: this warp's best k, lane r = r-th best
    // a row can enter query u's list only if  s32 * (1 - 2e-6) <= sthr[u] * beta_row  (hyperbolic: s = ||x-y||^2,
    // sthr = (cosh(sqrt(c) d_k) - 1) alpha / (2c));  dot32 + 2e-6 |x||y| >= cthr[u] |y|  (cosine).  +inf / -inf: list not full
    float thr[QB];
#pragma unroll
    for (int u = 0; u < QB; ++u) {
      mine[u] = EX_NONE;
      thr[u] = hyp ? INFINITY : -INFINITY;
      // init_bound[q]: a score no better than the query's true k-th best (the k-th entry of the filtered result, whose
      // rows are real gallery rows with exact scores).  The scan then starts warm: only rows at or inside that bound
      // are contenders, CTAs that see none skip the merge and its lock, and the union of the CTA lists still holds
      // every row of the true top-k.
      if (init_bound != nullptr && u < nq) {
        const double b = (double)init_bound[q_list[it0 + u]];
        if (hyp && b >= 0.0 && b < (double)INFINITY)
          thr[u] = (float)((cosh(sqrt(cc) * b * (1.0 + 1e-5)) - 1.0) * (1.0 - cc * xsq[u]) / (2.0 * cc) * (1.0 + 1e-5));
        else if (!hyp && b > -(double)INFINITY && b == b)
          thr[u] = (float)(b * sqrt(xsq[u]) - 1e-5 * fabs(b * sqrt(xsq[u])) - 1e-30);
      }
    }
    constexpr int PASS = NV <= 8 ? 2 : 1;                   // rows in flight per warp: their registers bound it
    constexpr int CH = NV < 4 ? NV : 4;
    for (int64_t g0 = r_lo + (int64_t)warp * PASS; g0 < r_hi; g0 += EX_WARPS * PASS) {
      const float4* g[PASS];
      bool val[PASS];
#pragma unroll
      for (int t = 0; t < PASS; ++t) {
        val[t] = g0 + t < r_hi;
        HYPRET_CHECK(!val[t] || (g0 + t >= 0 && g0 + t < N));
        g[t] = reinterpret_cast<const float4*>(g32 + (val[t] ? g0 + t : r_lo) * d);
      }
      double ysq[PASS];                                      // requested with the rows, used after them
#pragma unroll
      for (int t = 0; t < PASS; ++t) ysq[t] = val[t] ? g_sq64[g0 + t] : 0.0;
      float s32[QB][PASS];
#pragma unroll
      for (int u = 0; u < QB; ++u)
#pragma unroll
        for (int t = 0; t < PASS; ++t) s32[u][t] = 0.f;
      float4 b[PASS][NV];                                    // the rows stay in registers for the fp64 pass
#pragma unroll
      for (int i0 = 0; i0 < NV; i0 += CH) {
#pragma unroll
        for (int t = 0; t < PASS; ++t)
#pragma unroll
          for (int ii = 0; ii < CH; ++ii) {
            const int j = (i0 + ii) * 32 + lane;
            if (i0 + ii < NV) b[t][i0 + ii] = (val[t] && j < nvec) ? __ldg(g[t] + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int ii = 0; ii < CH; ++ii) {
          if (i0 + ii >= NV) continue;
          const int i = i0 + ii;
#pragma unroll
          for (int t = 0; t < PASS; ++t) {
            const float4 bb = b[t][i];
#pragma unroll
            for (int u = 0; u < QB; ++u) {
              if (hyp) {
                const float e0 = qv[u][i].x - bb.x, e1 = qv[u][i].y - bb.y, e2 = qv[u][i].z - bb.z, e3 = qv[u][i].w - bb.w;
                s32[u][t] = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, s32[u][t]))));
              } else {
                s32[u][t] = fmaf(qv[u][i].x, bb.x, fmaf(qv[u][i].y, bb.y, fmaf(qv[u][i].z, bb.z, fmaf(qv[u][i].w, bb.w, s32[u][t]))));
              }
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < PASS; ++t) {
        if (!val[t]) continue;                               // warp-uniform
        const double y0 = ysq[t];
        const float beta = (float)(1.0 - cc * y0), ynorm = sqrtf((float)y0);
#pragma unroll
        for (int u = 0; u < QB; ++u) {
          if (u >= nq) continue;                             // block-uniform
          const float s = warp_sum(s32[u][t]);
          const bool cand = hyp ? s * (1.0f - 2e-6f) <= thr[u] * beta
                                : s + 2e-6f * (float)sqrt(xsq[u]) * ynorm >= thr[u] * ynorm;
          if (!cand) continue;                               // warp-uniform: s is the same in every lane
          // exact key of this row, with the arithmetic (and summation order) of the rerank kernel
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float4 bb = b[t][i];
            if (hyp) {
              const float e0 = qv[u][i].x - bb.x, e1 = qv[u][i].y - bb.y, e2 = qv[u][i].z - bb.z, e3 = qv[u][i].w - bb.w;
              acc += (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2 + (double)e3 * e3;
            } else {
              acc += (double)qv[u][i].x * bb.x + (double)qv[u][i].y * bb.y + (double)qv[u][i].z * bb.z +
                     (double)qv[u][i].w * bb.w;
            }
          }
          acc = warp_sum(acc);
          double kd;
          if (hyp) {
            const double t0 = 2.0 * cc * acc / ((1.0 - cc * xsq[u]) * (1.0 - cc * y0));
            kd = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
          } else {
            const double nx = sqrt(xsq[u]);
            kd = -(acc / ((nx == 0.0 ? 1.0 : nx) * (y0 == 0.0 ? 1.0 : sqrt(y0))));
          }
          const unsigned long long key = ((unsigned long long)ex_ordered_key((float)kd) << 32) | (unsigned)(g0 + t);
          if (after != nullptr && key <= aft[u]) continue;   // warp-uniform: already returned by an earlier page
          ex_insert(mine[u], key, k, lane);
          // refresh the prefilter threshold from the list's k-th best (fp32 value, widened by one part in 1e5)
          const unsigned long long worst = __shfl_sync(0xffffffffu, mine[u], k - 1);
          if (worst != EX_NONE) {
            const double dk = (double)ex_ordered_val((unsigned)(worst >> 32));
            if (hyp) thr[u] = (float)((cosh(sqrt(cc) * dk * (1.0 + 1e-5)) - 1.0) * (1.0 - cc * xsq[u]) / (2.0 * cc) * (1.0 + 1e-5));
            else thr[u] = (float)(-dk * sqrt(xsq[u]) - 1e-5 * fabs(dk * sqrt(xsq[u])) - 1e-30);
          }
        }
      }
    }
    // ---- per query: the CTA's best k (warp 0 folds the eight lists), merged into the result rows under its lock -------
#pragma unroll 1
    for (int u = 0; u < nq; ++u) {
      unsigned long long mu = EX_NONE;
#pragma unroll
      for (int uu = 0; uu < QB; ++uu) mu = uu == u ? mine[uu] : mu;
      lists[warp][lane] = lane < k ? mu : EX_NONE;
      __syncthreads();
      if (warp == 0) {
        const int64_t q = q_list[it0 + u];
        unsigned long long best = lists[0][lane];
#pragma unroll
        for (int w = 1; w < EX_WARPS; ++w) best = ex_merge(best, lists[w][lane], lane);
        int32_t* lock = state + 2 * q;
        if (__any_sync(0xffffffffu, best != EX_NONE)) {      // nothing to contribute: no lock, no traffic
        if (lane == 0) {
          while (atomicCAS(lock, 0, 1) != 0) __nanosleep(100);
        }
        __syncwarp();
        __threadfence();
        const int initialised = *reinterpret_cast<volatile int32_t*>(lock + 1);
        unsigned long long cur = EX_NONE;
        if (initialised && lane < k) {
          const float sc = __ldcg(out_score + q * k + lane);
          const int64_t id = __ldcg(out_idx + q * k + lane);
          if (id >= 0) {
            const float kf = hyp ? sc : -sc;
            cur = ((unsigned long long)ex_ordered_key(kf) << 32) | (unsigned)(id - idx_offset);
          }
        }
        best = ex_merge(best, cur, lane);
        if (lane < k) {
          const bool valid = best != EX_NONE;
          const float kf = ex_ordered_val((unsigned)(best >> 32));
          const float sc = hyp ? kf : -kf;
          out_score[q * k + lane] = valid ? sc : (hyp ? INFINITY : -INFINITY);
          out_idx[q * k + lane] = valid ? (int64_t)(unsigned)(best & 0xffffffffull) + idx_offset : (int64_t)-1;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          *reinterpret_cast<volatile int32_t*>(lock + 1) = 1;
          __threadfence();
          atomicExch(lock, 0);
        }
        }
      }
      __syncthreads();                                       // lists[] is rewritten by the next query
    }
  }
}

// ---- the certificate of a MERGED result (row-sharded gallery, queries owned by one rank) ---------------------------
// The owner of query q holds its merged exact top-k (from every shard's pruned rerank) and the global k'-th best filter
// score thr[q] (hypret_kth_smallest over the shards' lists): every row of every shard outside the global candidate set
// has a filter score >= thr[q], and an exact surrogate within E of it (E from the query's own rounding residual and
// the MAXIMA over all shards of the gallery statistics).  thr - S_kth > E proves the merged list exact, as in
// rerank_kernel; S_kth is recovered from the emitted distance: S = (cosh(sqrt(c) d) - 1) (1 - c |x|^2) / 2.
// flags[q] = 1 when the proof FAILS (the query then goes through the exact scan of every shard).
__global__ void __launch_bounds__(128)
cert_merged_kernel(const float* __restrict__ q32, int64_t Q, int d, float c, int metric, const float* __restrict__ score,
                   const int64_t* __restrict__ idx, int k, const float* __restrict__ thr,
                   const float* __restrict__ q_err, const float* __restrict__ g_stats, float slack,
                   int32_t* __restrict__ flags, float* __restrict__ out_margin) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= Q) return;
  double xsq = 0.0;
  for (int j = lane; j < (d >> 2); j += 32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(q32 + q * d) + j);
    xsq += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  xsq = warp_sum(xsq);
  if (lane != 0) return;
  const double cc = (double)c;
  const bool hyp = metric == HYPRET_METRIC_HYPERBOLIC;
  const double dk = (double)score[q * k + k - 1];
  const bool open_set = idx[q * k + k - 1] < 0 || !(thr[q] < INFINITY);   // fewer than k rows / k' candidates in all
  double sur;
  if (hyp) {
    const double a = sqrt(cc) * dk;
    sur = (cosh(a) - 1.0) * (1.0 - cc * xsq) * 0.5 * (1.0 + 4e-7);      // d is the fp32 rounding of the fp64 distance
  } else {
    sur = -dk + 1e-7 * fabs(dk);
  }
  const double qn = hyp ? sqrt(cc * xsq) : 1.0;
  const double zmax = g_stats[0], dzmax = g_stats[1], rbmax = g_stats[2], bmax = g_stats[3];
  const double E = (double)q_err[q] * zmax + qn * dzmax + (double)slack * (qn * zmax + qn * qn * rbmax + bmax);
  const double margin = (double)thr[q] - sur;
  if (out_margin != nullptr) out_margin[q] = open_set ? INFINITY : (float)margin;
  flags[q] = (open_set || margin > E) ? 0 : 1;
}

// flags [n] (non-zero = listed) -> list of the flagged positions + count, and the lock words of those queries reset
// for exact_topk_kernel.  One CTA; the order of the list does not matter.
__global__ void __launch_bounds__(256)
flag_compact_kernel(const int32_t* __restrict__ flags, int64_t n, int32_t* __restrict__ list,
                    int32_t* __restrict__ count, int32_t* __restrict__ state) {
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    if (flags[i] != 0) {
      const int at = atomicAdd(&total, 1);
      HYPRET_CHECK(at >= 0 && at < n);
      list[at] = (int32_t)i;
      state[2 * i] = 0;
      state[2 * i + 1] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *count = total;
}

}  // namespace

int hypret_launch_exact_topk(const float* q32, const float* g32, const double* g_sq64, int64_t Q, int64_t N, int d,
                             float c, int metric, int k, int64_t idx_offset, const int32_t* q_list,
                             const int32_t* q_count, int32_t* state, float* out_score, int64_t* out_idx,
                             const unsigned long long* after, const float* init_bound, cudaStream_t stream) {
  if (Q == 0 || N == 0) return HYPRET_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // ~2 chunks per SM so that a single listed query already spreads over the whole GPU; a chunk is a multiple of the
  // 256 rows one pass of the 8 warps covers
  int64_t chunk = (N + 2 * sms - 1) / (2 * sms);
  chunk = (chunk + 31) / 32 * 32;
  const int64_t n_chunks = (N + chunk - 1) / chunk;
  int64_t slots = (2 * (int64_t)sms + n_chunks - 1) / n_chunks;      // query groups in flight: ~2 CTAs per SM
  if (slots > (Q + 3) / 4) slots = (Q + 3) / 4;
  if (slots < 1) slots = 1;
  if (slots > 65535) slots = 65535;
  const dim3 grid((unsigned)n_chunks, (unsigned)slots);
  const int need = (d + 127) / 128;
#define HYPRET_EXACT_LAUNCH(NV)                                                                                     \
  do {                                                                                                              \
    exact_topk_kernel<NV, (NV <= 4 ? 4 : 2)><<<grid, EX_WARPS * 32, 0, stream>>>(                                   \
        q32, g32, g_sq64, N, d, c, metric, k, idx_offset, q_list, q_count, state, out_score, out_idx, chunk, after,  \
        init_bound);                                                                                                \
    return (int)cudaGetLastError();                                                                                 \
  } while (0)
  if (need <= 1) HYPRET_EXACT_LAUNCH(1);
  if (need <= 2) HYPRET_EXACT_LAUNCH(2);
  if (need <= 4) HYPRET_EXACT_LAUNCH(4);
  if (need <= 6) HYPRET_EXACT_LAUNCH(6);
  if (need <= 8) HYPRET_EXACT_LAUNCH(8);
  if (need <= 16) HYPRET_EXACT_LAUNCH(16);
#undef HYPRET_EXACT_LAUNCH
  return HYPRET_EUNSUPPORTED;
}

int hypret_launch_cert_merged(const float* q32, int64_t Q, int d, float c, int metric, const float* score,
                              const int64_t* idx, int k, const float* thr, const float* q_err, const float* g_stats,
                              float slack, int32_t* flags, float* out_margin, cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  cert_merged_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, stream>>>(q32, Q, d, c, metric, score, idx, k, thr, q_err,
                                                                  g_stats, slack, flags, out_margin);
  return (int)cudaGetLastError();
}

int hypret_launch_flag_compact(const int32_t* flags, int64_t n, int32_t* list, int32_t* count, int32_t* state,
                               cudaStream_t stream) {
  flag_compact_kernel<<<1, 256, 0, stream>>>(flags, n, list, count, state);
  return (int)cudaGetLastError();
}
