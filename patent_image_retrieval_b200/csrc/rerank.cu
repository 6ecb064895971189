// Candidate merge + exact rerank.
//
// Input: per query, n_splits lists of k' (surrogate score, gallery index) candidates from
// score_topk.cu.  One warp per query:
//   1. keep the k' best candidates by surrogate score (they are comparable across splits, so
//      this equals a single global approximate top-k');
//   2. recompute each survivor exactly from the fp32 rows: ||x-y||^2 with the difference
//      formed explicitly (no ||x||^2+||y||^2-2<x,y> cancellation), fp64 accumulation, then
//         d = arccosh(1 + 2c||x-y||^2 / ((1-c||x||^2)(1-c||y||^2))) / sqrt(c)
//      which is analytically pmath.dist (/root/reference/src/train.py:3259), or the cosine
//      similarity of notebooks/retrieval.ipynb:368;
//   3. sort (ascending distance / descending similarity, ties -> lower gallery index) with a
//      warp bitonic network and write the first k.
// Gather-bound: k' * D * 4 bytes of gallery rows per query, read as coalesced 128-bit loads.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int RR_WARPS = 4;

__device__ __forceinline__ bool key_less(double a, int ia, double b, int ib) {
  // total order: value ascending, then index ascending; invalid (index < 0) last
  if (ia < 0) return false;
  if (ib < 0) return true;
  return (a < b) || (a == b && ia < ib);
}

template <int NV>   // float4 chunks of a row per lane: D <= NV * 128
__global__ void __launch_bounds__(RR_WARPS * 32)
rerank_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int64_t Q, int64_t N, int d, float c,
              int metric, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx, int n_cand,
              int kprime, int k, int64_t idx_offset, float* __restrict__ out_score, int64_t* __restrict__ out_idx,
              float* __restrict__ out_margin) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * RR_WARPS + warp;
  if (q >= Q) return;
  float* cs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * n_cand;
  int* ci = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)RR_WARPS * n_cand) +
            (size_t)warp * n_cand;

  // ---- 1. approximate merge: k' rounds of warp arg-min over the candidates ---------------
  for (int t = lane; t < n_cand; t += 32) {
    cs[t] = cand_score[q * n_cand + t];
    ci[t] = cand_idx[q * n_cand + t];
  }
  __syncwarp();
  int my_idx = -1;          // lane r holds the r-th selected candidate
  float worst_approx = -INFINITY;
  for (int r = 0; r < kprime; ++r) {
    float bs = INFINITY;
    int bi = 0x7fffffff, bpos = -1;
    for (int t = lane; t < n_cand; t += 32) {
      const float s = cs[t];
      const int i = ci[t];
      if (i >= 0 && (s < bs || (s == bs && i < bi))) { bs = s; bi = i; bpos = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      if (op >= 0 && (bpos < 0 || os < bs || (os == bs && oi < bi))) { bs = os; bi = oi; bpos = op; }
    }
    if (bpos < 0) break;      // fewer than k' valid candidates (warp-uniform)
    if (lane == r) my_idx = bi;
    worst_approx = bs;
    if ((bpos & 31) == lane) ci[bpos] = -1;
    __syncwarp();
  }

  // ---- 2. exact scores of the survivors ---------------------------------------------------
  // the query row stays in registers; four gallery rows are streamed per pass (independent
  // 128-bit loads in flight), fp64 accumulation of explicitly formed differences
  const int nvec = d >> 2;
  const float4* qrow = reinterpret_cast<const float4*>(q32 + q * d);
  float4 qv[NV];
  double xsq = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = i * 32 + lane;
    qv[i] = (j < nvec) ? __ldg(qrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    xsq += (double)qv[i].x * qv[i].x + (double)qv[i].y * qv[i].y + (double)qv[i].z * qv[i].z +
           (double)qv[i].w * qv[i].w;
  }
  xsq = warp_sum(xsq);

  double my_key = INFINITY;     // sort key (distance, or minus similarity)
  double my_sur = INFINITY;     // exact surrogate, comparable with the approximate scores
  constexpr int PASS = 4;
  for (int r0 = 0; r0 < kprime; r0 += PASS) {
    int idx[PASS];
    bool val[PASS];
    const float4* g[PASS];
    bool any = false;
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      idx[t] = __shfl_sync(0xffffffffu, my_idx, (r0 + t) & 31);
      val[t] = (r0 + t < kprime) && idx[t] >= 0;
      any |= val[t];
      g[t] = reinterpret_cast<const float4*>(g32 + (int64_t)(val[t] ? idx[t] : 0) * d);
    }
    if (!any) continue;
    double sacc[PASS], yacc[PASS];
#pragma unroll
    for (int t = 0; t < PASS; ++t) { sacc[t] = 0.0; yacc[t] = 0.0; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = i * 32 + lane;
      if (j < nvec) {
        float4 b[PASS];
#pragma unroll
        for (int t = 0; t < PASS; ++t) b[t] = __ldg(g[t] + j);
#pragma unroll
        for (int t = 0; t < PASS; ++t) {
          if (metric == HYPRET_METRIC_HYPERBOLIC) {
            const float e0 = qv[i].x - b[t].x, e1 = qv[i].y - b[t].y, e2 = qv[i].z - b[t].z, e3 = qv[i].w - b[t].w;
            sacc[t] += (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2 + (double)e3 * e3;
          } else {
            sacc[t] += (double)qv[i].x * b[t].x + (double)qv[i].y * b[t].y + (double)qv[i].z * b[t].z +
                       (double)qv[i].w * b[t].w;
          }
          yacc[t] += (double)b[t].x * b[t].x + (double)b[t].y * b[t].y + (double)b[t].z * b[t].z +
                     (double)b[t].w * b[t].w;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      const double s0 = warp_sum(sacc[t]), y0 = warp_sum(yacc[t]);
      double key0, sur0;
      if (metric == HYPRET_METRIC_HYPERBOLIC) {
        const double cc = (double)c;
        const double al = 1.0 - cc * xsq;
        const double t0 = 2.0 * cc * s0 / (al * (1.0 - cc * y0));
        key0 = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
        sur0 = s0 / (1.0 - cc * y0);
      } else {
        const double nx = sqrt(xsq);
        const double d0 = (nx == 0.0 ? 1.0 : nx) * (y0 == 0.0 ? 1.0 : sqrt(y0));
        key0 = -(s0 / d0);
        sur0 = key0;
      }
      if (val[t] && lane == r0 + t) { my_key = key0; my_sur = sur0; }
    }
  }

  // ---- 3. warp bitonic sort by (key, index) -----------------------------------------------
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, my_key, stride);
      const double os = __shfl_xor_sync(0xffffffffu, my_sur, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, my_idx, stride);
      const bool lower = (lane & stride) == 0;
      const bool asc = (lane & size) == 0;
      const bool mine_first = key_less(my_key, my_idx, ok, oi);
      const bool keep = (lower == asc) ? mine_first : !mine_first;
      // equal elements cannot occur twice (indices are distinct) unless both invalid
      if (!keep && !(my_idx < 0 && oi < 0)) { my_key = ok; my_sur = os; my_idx = oi; }
    }
  }
  if (lane < k) {
    const bool valid = my_idx >= 0;
    const double val = (metric == HYPRET_METRIC_HYPERBOLIC) ? my_key : -my_key;
    out_score[q * k + lane] = valid ? (float)val : ((metric == HYPRET_METRIC_HYPERBOLIC) ? INFINITY : -INFINITY);
    out_idx[q * k + lane] = valid ? (int64_t)my_idx + idx_offset : (int64_t)-1;
  }
  if (out_margin != nullptr) {
    const double kth = __shfl_sync(0xffffffffu, my_sur, k - 1);
    const int kth_idx = __shfl_sync(0xffffffffu, my_idx, k - 1);
    // +inf: the candidate set was not truncated (fewer than k' valid candidates survive)
    const int n_valid = __popc(__ballot_sync(0xffffffffu, my_idx >= 0));
    if (lane == 0)
      out_margin[q] = (n_valid < kprime || kth_idx < 0) ? INFINITY : (float)((double)worst_approx - kth);
  }
}

}  // namespace

int hypret_launch_rerank(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, int n_cand, int kprime, int k,
                         int64_t idx_offset, float* out_score, int64_t* out_idx, float* out_margin,
                         cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  const size_t smem = (size_t)RR_WARPS * n_cand * 8;
  if (smem > 200 * 1024) return HYPRET_EUNSUPPORTED;
  const int64_t grid = (Q + RR_WARPS - 1) / RR_WARPS;
  const int need = (d + 127) / 128;
#define HYPRET_RERANK_LAUNCH(NV)                                                                                    \
  do {                                                                                                              \
    if (smem > 48 * 1024) {                                                                                         \
      cudaError_t e =                                                                                               \
          cudaFuncSetAttribute(rerank_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      if (e != cudaSuccess) return (int)e;                                                                          \
    }                                                                                                               \
    rerank_kernel<NV><<<(unsigned)grid, RR_WARPS * 32, smem, stream>>>(q32, g32, Q, N, d, c, metric, cand_score,    \
                                                                      cand_idx, n_cand, kprime, k, idx_offset,      \
                                                                      out_score, out_idx, out_margin);              \
    return (int)cudaGetLastError();                                                                                 \
  } while (0)
  if (need <= 1) HYPRET_RERANK_LAUNCH(1);
  if (need <= 2) HYPRET_RERANK_LAUNCH(2);
  if (need <= 4) HYPRET_RERANK_LAUNCH(4);
  if (need <= 6) HYPRET_RERANK_LAUNCH(6);
  if (need <= 8) HYPRET_RERANK_LAUNCH(8);
  if (need <= 16) HYPRET_RERANK_LAUNCH(16);
#undef HYPRET_RERANK_LAUNCH
  return HYPRET_EUNSUPPORTED;
}
