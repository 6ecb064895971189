// Exact top-k by full scan: the guarantee behind the tensor-core filter.
//
// hypret_rerank_cert certifies, per query, that no gallery row outside the bf16-filtered candidate set can precede the
// k-th result (margin > rounding-error bound E, csrc/rerank.cu).  Queries it cannot certify -- near-duplicate
// galleries, where more than k' - k rows sit inside the bf16 error band of the k-th best -- are listed on the device,
// and this kernel recomputes their top-k from ALL gallery rows with the arithmetic of the rerank kernel (explicit
// differences, fp64 accumulation in the same order, arccosh closed form == pmath.dist,
// /root/reference/src/train.py:3259; cosine of notebooks/retrieval.ipynb:368), i.e. exactly what the reference's
// per-query loop + torch.topk (src/auxiliary.py:374) does, and overwrites the query's result rows.  No host
// round trip: the list length is read on the device, the grid is fixed, and with an empty list every CTA exits at once.
//
// Work decomposition: grid = (row chunks, slots).  CTA (c, s) serves list entries s, s + slots, ...; for each it scans
// gallery rows [c * chunk, (c+1) * chunk): 8 warps take groups of 32 rows (4 rows per pass with all their 128-bit loads
// issued before the first use; lane t keeps the sums of row t of the group and evaluates its own key), each warp keeps
// its best k as packed 64-bit keys (ordered fp32 score << 32 | row id: one unsigned compare = the (score, id) order of
// the rerank), one per lane, sorted.  The CTA's list is merged into the query's result rows under a per-query spin
// lock (the holder never waits on anything, so the lock cannot deadlock whatever part of the grid is resident); the
// first merger of a query discards the filtered result that is there.  FP64-pipe bound (16 F2F + 16 DFMA per row and
// lane): ~40 cycles per row and SM, 300k rows x 1 query = 45 us of the whole GPU.
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

template <int NV>
__global__ void __launch_bounds__(EX_WARPS * 32)
exact_topk_kernel(const float* __restrict__ q32, const float* __restrict__ g32, const double* __restrict__ g_sq64,
                  int64_t N, int d, float c, int metric, int k, int64_t idx_offset,
                  const int32_t* __restrict__ q_list, const int32_t* __restrict__ q_count, int32_t* state,
                  float* out_score, int64_t* out_idx, int64_t chunk) {
  __shared__ unsigned long long lists[EX_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = *q_count;
  const int64_t r_lo = (int64_t)blockIdx.x * chunk;
  const int64_t r_hi = r_lo + chunk < N ? r_lo + chunk : N;
  const int nvec = d >> 2;
  const double cc = (double)c;
  for (int it = blockIdx.y; it < count; it += gridDim.y) {
    const int64_t q = q_list[it];
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
    unsigned long long mine = EX_NONE;                       // this warp's best k, lane r = r-th best
    for (int64_t g0 = r_lo + (int64_t)warp * 32; g0 < r_hi; g0 += EX_WARPS * 32) {
      double my_s = 0.0;
      constexpr int PASS = 4;
      constexpr int CH = NV < 4 ? NV : 4;
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += PASS) {
        if (g0 + r0 >= r_hi) break;                          // warp-uniform
        const float4* g[PASS];
        bool val[PASS];
#pragma unroll
        for (int t = 0; t < PASS; ++t) {
          val[t] = g0 + r0 + t < r_hi;
          g[t] = reinterpret_cast<const float4*>(g32 + (val[t] ? g0 + r0 + t : r_lo) * d);
        }
        double sacc[PASS];
#pragma unroll
        for (int t = 0; t < PASS; ++t) sacc[t] = 0.0;
#pragma unroll
        for (int i0 = 0; i0 < NV; i0 += CH) {
          float4 b[PASS][CH];
#pragma unroll
          for (int t = 0; t < PASS; ++t)
#pragma unroll
            for (int ii = 0; ii < CH; ++ii) {
              const int j = (i0 + ii) * 32 + lane;
              b[t][ii] = (i0 + ii < NV && val[t] && j < nvec) ? __ldg(g[t] + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int ii = 0; ii < CH; ++ii) {
            const int i = i0 + ii < NV ? i0 + ii : NV - 1;
            if (i0 + ii >= NV) continue;
#pragma unroll
            for (int t = 0; t < PASS; ++t) {
              const float4 bb = b[t][ii];
              if (metric == HYPRET_METRIC_HYPERBOLIC) {
                const float e0 = qv[i].x - bb.x, e1 = qv[i].y - bb.y, e2 = qv[i].z - bb.z, e3 = qv[i].w - bb.w;
                sacc[t] += (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2 + (double)e3 * e3;
              } else {
                sacc[t] += (double)qv[i].x * bb.x + (double)qv[i].y * bb.y + (double)qv[i].z * bb.z +
                           (double)qv[i].w * bb.w;
              }
            }
          }
        }
#pragma unroll
        for (int t = 0; t < PASS; ++t) {
          const double s0 = warp_sum(sacc[t]);
          if (lane == r0 + t) my_s = s0;
        }
      }
      // lane t evaluates row g0 + t: the fp32 value the rerank kernel emits, as an ordered key
      const int64_t row = g0 + lane;
      unsigned long long key = EX_NONE;
      if (row < r_hi) {
        const double y0 = g_sq64[row];
        double kd;
        if (metric == HYPRET_METRIC_HYPERBOLIC) {
          const double t0 = 2.0 * cc * my_s / ((1.0 - cc * xsq) * (1.0 - cc * y0));
          kd = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
        } else {
          const double nx = sqrt(xsq);
          kd = -(my_s / ((nx == 0.0 ? 1.0 : nx) * (y0 == 0.0 ? 1.0 : sqrt(y0))));
        }
        key = ((unsigned long long)ex_ordered_key((float)kd) << 32) | (unsigned)row;
      }
      const unsigned long long worst = __shfl_sync(0xffffffffu, mine, k - 1);
      unsigned hits = __ballot_sync(0xffffffffu, key < worst);
      while (hits != 0u) {                                   // warp-uniform; rare once the list is warm
        const int src = __ffs((int)hits) - 1;
        hits &= hits - 1u;
        ex_insert(mine, __shfl_sync(0xffffffffu, key, src), k, lane);
      }
    }
    // ---- the CTA's best k: warp 0 folds the eight lists ---------------------------------------------------------
    lists[warp][lane] = lane < k ? mine : EX_NONE;
    __syncthreads();
    if (warp == 0) {
      unsigned long long best = lists[0][lane];
#pragma unroll
      for (int w = 1; w < EX_WARPS; ++w) best = ex_merge(best, lists[w][lane], lane);
      // ---- merge into the query's result rows under its lock -----------------------------------------------------
      int32_t* lock = state + 2 * q;
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
          const float kf = metric == HYPRET_METRIC_HYPERBOLIC ? sc : -sc;
          cur = ((unsigned long long)ex_ordered_key(kf) << 32) | (unsigned)(id - idx_offset);
        }
      }
      best = ex_merge(best, cur, lane);
      if (lane < k) {
        const bool valid = best != EX_NONE;
        const float kf = ex_ordered_val((unsigned)(best >> 32));
        const float sc = metric == HYPRET_METRIC_HYPERBOLIC ? kf : -kf;
        out_score[q * k + lane] = valid ? sc : (metric == HYPRET_METRIC_HYPERBOLIC ? INFINITY : -INFINITY);
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
    __syncthreads();                                         // lists[] is rewritten by the next entry
  }
}

}  // namespace

int hypret_launch_exact_topk(const float* q32, const float* g32, const double* g_sq64, int64_t Q, int64_t N, int d,
                             float c, int metric, int k, int64_t idx_offset, const int32_t* q_list,
                             const int32_t* q_count, int32_t* state, float* out_score, int64_t* out_idx,
                             cudaStream_t stream) {
  if (Q == 0 || N == 0) return HYPRET_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // ~2 chunks per SM so that a single listed query already spreads over the whole GPU; a chunk is a multiple of the
  // 256 rows one pass of the 8 warps covers
  int64_t chunk = (N + 2 * sms - 1) / (2 * sms);
  chunk = (chunk + 255) / 256 * 256;
  const int64_t n_chunks = (N + chunk - 1) / chunk;
  int64_t slots = (4 * (int64_t)sms + n_chunks - 1) / n_chunks;      // ~4 CTAs per SM in flight
  if (slots > Q) slots = Q;
  if (slots < 1) slots = 1;
  if (slots > 65535) slots = 65535;
  const dim3 grid((unsigned)n_chunks, (unsigned)slots);
  const int need = (d + 127) / 128;
#define HYPRET_EXACT_LAUNCH(NV)                                                                                     \
  do {                                                                                                              \
    exact_topk_kernel<NV><<<grid, EX_WARPS * 32, 0, stream>>>(q32, g32, g_sq64, N, d, c, metric, k, idx_offset,     \
                                                              q_list, q_count, state, out_score, out_idx, chunk);   \
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
