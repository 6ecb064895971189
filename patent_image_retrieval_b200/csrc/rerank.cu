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
  float my_approx = INFINITY;
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
    if (lane == r) { my_idx = bi; my_approx = bs; }
    worst_approx = bs;
    if ((bpos & 31) == lane) ci[bpos] = -1;
    __syncwarp();
  }

  // ---- 2. exact scores of the survivors ---------------------------------------------------
  const int nvec = d >> 2;
  const float4* qrow = reinterpret_cast<const float4*>(q32 + q * d);
  double xsq = 0.0;
  for (int j = lane; j < nvec; j += 32) {
    const float4 a = __ldg(qrow + j);
    xsq += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
  }
  xsq = warp_sum(xsq);

  double my_key = INFINITY;     // sort key (distance, or minus similarity)
  double my_sur = INFINITY;     // exact surrogate, comparable with the approximate scores
  for (int r0 = 0; r0 < kprime; r0 += 2) {
    // two candidates per pass: independent row streams in flight
    const int i0 = __shfl_sync(0xffffffffu, my_idx, r0);
    const int i1 = (r0 + 1 < 32) ? __shfl_sync(0xffffffffu, my_idx, (r0 + 1) & 31) : -1;
    const bool v0 = i0 >= 0, v1 = (r0 + 1 < kprime) && i1 >= 0;
    if (!v0 && !v1) continue;
    const float4* g0 = reinterpret_cast<const float4*>(g32 + (int64_t)(v0 ? i0 : 0) * d);
    const float4* g1 = reinterpret_cast<const float4*>(g32 + (int64_t)(v1 ? i1 : 0) * d);
    double s0 = 0.0, y0 = 0.0, s1 = 0.0, y1 = 0.0;   // hyperbolic: sum (x-y)^2 ; cosine: sum x*y
    for (int j = lane; j < nvec; j += 32) {
      const float4 a = __ldg(qrow + j);
      const float4 b0 = __ldg(g0 + j);
      const float4 b1 = __ldg(g1 + j);
      if (metric == HYPRET_METRIC_HYPERBOLIC) {
        const float e0 = a.x - b0.x, e1 = a.y - b0.y, e2 = a.z - b0.z, e3 = a.w - b0.w;
        const float f0 = a.x - b1.x, f1 = a.y - b1.y, f2 = a.z - b1.z, f3 = a.w - b1.w;
        s0 += (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2 + (double)e3 * e3;
        s1 += (double)f0 * f0 + (double)f1 * f1 + (double)f2 * f2 + (double)f3 * f3;
      } else {
        s0 += (double)a.x * b0.x + (double)a.y * b0.y + (double)a.z * b0.z + (double)a.w * b0.w;
        s1 += (double)a.x * b1.x + (double)a.y * b1.y + (double)a.z * b1.z + (double)a.w * b1.w;
      }
      y0 += (double)b0.x * b0.x + (double)b0.y * b0.y + (double)b0.z * b0.z + (double)b0.w * b0.w;
      y1 += (double)b1.x * b1.x + (double)b1.y * b1.y + (double)b1.z * b1.z + (double)b1.w * b1.w;
    }
    s0 = warp_sum(s0); y0 = warp_sum(y0);
    s1 = warp_sum(s1); y1 = warp_sum(y1);
    double key0, sur0, key1, sur1;
    if (metric == HYPRET_METRIC_HYPERBOLIC) {
      const double cc = (double)c;
      const double al = 1.0 - cc * xsq;
      const double t0 = 2.0 * cc * s0 / (al * (1.0 - cc * y0));
      const double t1 = 2.0 * cc * s1 / (al * (1.0 - cc * y1));
      key0 = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
      key1 = log1p(t1 + sqrt(t1 * (t1 + 2.0))) / sqrt(cc);
      sur0 = s0 / (1.0 - cc * y0);
      sur1 = s1 / (1.0 - cc * y1);
    } else {
      const double nx = sqrt(xsq);
      const double d0 = (nx == 0.0 ? 1.0 : nx) * (y0 == 0.0 ? 1.0 : sqrt(y0));
      const double d1 = (nx == 0.0 ? 1.0 : nx) * (y1 == 0.0 ? 1.0 : sqrt(y1));
      key0 = -(s0 / d0);
      key1 = -(s1 / d1);
      sur0 = key0;
      sur1 = key1;
    }
    if (v0 && lane == r0) { my_key = key0; my_sur = sur0; }
    if (v1 && lane == r0 + 1) { my_key = key1; my_sur = sur1; }
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
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t grid = (Q + RR_WARPS - 1) / RR_WARPS;
  rerank_kernel<<<(unsigned)grid, RR_WARPS * 32, smem, stream>>>(q32, g32, Q, N, d, c, metric, cand_score, cand_idx,
                                                                n_cand, kprime, k, idx_offset, out_score, out_idx,
                                                                out_margin);
  return (int)cudaGetLastError();
}
