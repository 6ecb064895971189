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

__device__ __forceinline__ unsigned ordered_key(float x) {   // monotone float -> uint map
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_val(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Approximate merge of one query's candidate lists (warp-cooperative): selects, in ascending (surrogate, index)
// order, up to `kprime` valid candidates from cs/ci[0..n_cand) (shared memory, consumed: selected slots get
// index -1), stopping early at the first candidate whose surrogate exceeds `prune`.  Lane r returns the r-th
// selected candidate (index -1 beyond the count); *worst = surrogate of the last one selected.
// Each lane caches the best of its own slots (t = lane, lane+32, ...); a round is two hardware warp
// reductions (redux.sync.min on the ordered key, then on the index among the lanes that tie) and a rescan of
// the winning lane's slots only.
// Few candidates (<= 96: a query of the headline configuration has 2-4 lists of 16): sort instead of selecting.
// One (ordered score, index) pair per lane packed into a 64-bit key -- a single unsigned compare is the
// lexicographic (surrogate, index) order, invalid slots are all-ones -- each chunk of 32 goes through a warp bitonic
// network (15 shuffle stages) and is merged into the running best 32 (reverse, min, 5 stages).  ~170 instructions
// per chunk against ~35 per selected candidate plus the rescans of the arg-min loop below.
__device__ __forceinline__ unsigned long long warp_minmax(unsigned long long v, int j, bool take_min) {
  const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
  return take_min ? (v < o ? v : o) : (v < o ? o : v);
}

__device__ __forceinline__ int select_candidates_sorted(const float* cs, const int* ci, int n_cand, int kprime,
                                                        float prune, int lane, float* my_score, float* worst) {
  constexpr unsigned long long NONE = ~0ull;
  unsigned long long best = NONE;
  for (int c0 = 0; c0 < n_cand; c0 += 32) {
    const int t = c0 + lane;
    unsigned long long key = NONE;
    if (t < n_cand) {
      const int i = ci[t];
      if (i >= 0) key = ((unsigned long long)ordered_key(cs[t]) << 32) | (unsigned)i;
    }
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) key = warp_minmax(key, j, ((lane & k) == 0) == ((lane & j) == 0));
    if (c0 == 0) {
      best = key;
    } else {
      const unsigned long long rev = __shfl_sync(0xffffffffu, key, 31 - lane);
      best = best < rev ? best : rev;              // the 32 smallest of both, as a bitonic sequence
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) best = warp_minmax(best, j, (lane & j) == 0);
    }
  }
  const float sc = ordered_val((unsigned)(best >> 32));
  const bool sel = best != NONE && lane < kprime && !(sc > prune);   // ascending: the selected lanes are a prefix
  const unsigned m = __ballot_sync(0xffffffffu, sel);
  const int n_sel = __popc(m);
  *my_score = sel ? sc : INFINITY;
  const float last = __shfl_sync(0xffffffffu, sc, n_sel > 0 ? n_sel - 1 : 0);
  *worst = n_sel > 0 ? last : -INFINITY;
  return sel ? (int)(unsigned)(best & 0xffffffffull) : -1;
}

__device__ __forceinline__ int select_candidates(float* cs, int* ci, int n_cand, int kprime, float prune, int lane,
                                                 float* my_score, float* worst) {
  if (n_cand <= 96) return select_candidates_sorted(cs, ci, n_cand, kprime, prune, lane, my_score, worst);
  unsigned bk = 0xffffffffu;   // ordered key of this lane's best slot (0xffffffff: none)
  int bi = 0x7fffffff, bpos = -1;
  for (int t = lane; t < n_cand; t += 32) {
    const int i = ci[t];
    const unsigned kk = ordered_key(cs[t]);
    if (i >= 0 && (bpos < 0 || kk < bk || (kk == bk && i < bi))) { bk = kk; bi = i; bpos = t; }
  }
  if (bpos < 0) bk = 0xffffffffu;
  int my_idx = -1;
  *my_score = INFINITY;
  *worst = -INFINITY;
  for (int r = 0; r < kprime; ++r) {
    const unsigned mk = __reduce_min_sync(0xffffffffu, bpos >= 0 ? bk : 0xffffffffu);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (bpos >= 0 && bk == mk) ? (unsigned)bi : 0xffffffffu);
    if (mi == 0xffffffffu) break;                       // no valid candidate left (warp-uniform)
    const float ms = ordered_val(mk);
    if (ms > prune) break;                              // everything left is worse than the global k'-th best
    if (lane == r) { my_idx = (int)mi; *my_score = ms; }
    *worst = ms;
    if (bpos >= 0 && bk == mk && bi == (int)mi) {       // the winning lane retires the slot and rescans its own
      ci[bpos] = -1;
      bk = 0xffffffffu; bi = 0x7fffffff; bpos = -1;
      for (int t = lane; t < n_cand; t += 32) {
        const int i = ci[t];
        const unsigned kk = ordered_key(cs[t]);
        if (i >= 0 && (bpos < 0 || kk < bk || (kk == bk && i < bi))) { bk = kk; bi = i; bpos = t; }
      }
    }
  }
  return my_idx;
}

// PRUNED (sharded serving: only the ~k'/W candidates at or below the global threshold are rescored): two rows per
// pass instead of four -- half the row registers, so 6 CTAs fit an SM instead of 4; the kernel is a chain of
// dependent latencies per query (list -> select -> rows -> sort), and only more resident warps hide it.
template <int NV, bool PRUNED>   // float4 chunks of a row per lane: D <= NV * 128
__global__ void __launch_bounds__(RR_WARPS * 32, NV <= 4 ? (PRUNED ? 6 : 4) : 1)
rerank_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int64_t Q, int64_t N, int d, float c,
              int metric, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
              const int32_t* __restrict__ list_count, int n_cand, int kprime, int k, int64_t idx_offset,
              const float* __restrict__ prune_thr,
              float* __restrict__ out_score, int64_t* __restrict__ out_idx, float* __restrict__ out_margin,
              const PeerRoute route, int64_t score_off, int64_t idx_off, const CertArgs cert) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * RR_WARPS + warp;
  if (q >= Q) return;
  int64_t out_row = q;
  if (route.n > 0) {        // the [Q,k] list goes straight into the query owner's receive region (NVLink stores)
    int owner;
    out_row = route_row(route, q, &owner);
    out_score = reinterpret_cast<float*>(route.base[owner] + score_off);
    out_idx = reinterpret_cast<int64_t*>(route.base[owner] + idx_off);
  }
  float* cs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * n_cand;
  int* ci = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)RR_WARPS * n_cand) +
            (size_t)warp * n_cand;

  // the query row is requested first so that its latency hides behind the candidate merge
  const int nvec = d >> 2;
  const float4* qrow = reinterpret_cast<const float4*>(q32 + q * d);
  float4 qv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = i * 32 + lane;
    qv[i] = (j < nvec) ? __ldg(qrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // ---- 1. approximate merge: the k' best candidates by surrogate score --------------------
  const int n_mine = list_count != nullptr ? min(list_count[q] * kprime, n_cand) : n_cand;   // compact list slots
  HYPRET_CHECK(n_mine >= 0 && n_mine <= n_cand && (list_count == nullptr || list_count[q] * kprime <= n_cand));
  for (int t = lane; t < n_mine; t += 32) {
    cs[t] = cand_score[q * n_cand + t];
    ci[t] = cand_idx[q * n_cand + t];
  }
  __syncwarp();
  // ksel > k' (register-list scoring: 16-slot lists that share the bound of the ksel-th best score): the ksel best of
  // the union are rescored; a FULL list may have dropped rows for lack of slots, and those are bounded only by the
  // list's own worst entry -- the certificate below takes the smaller of the two limits
  const int ksel = (cert.ksel > kprime && cert.ksel <= 32) ? cert.ksel : kprime;
  float w_trunc = INFINITY;
  if (ksel > kprime) {
    for (int t = lane; t < n_mine / kprime; t += 32) {
      bool full = true;
      float w = -INFINITY;
      for (int e = 0; e < kprime; ++e) {
        full = full && ci[t * kprime + e] >= 0;
        w = fmaxf(w, cs[t * kprime + e]);
      }
      if (full) w_trunc = fminf(w_trunc, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w_trunc = fminf(w_trunc, __shfl_xor_sync(0xffffffffu, w_trunc, o));
  }
  float my_approx, worst_approx;
  int my_idx = select_candidates(cs, ci, n_mine, ksel, prune_thr != nullptr ? prune_thr[q] : INFINITY, lane,
                                 &my_approx, &worst_approx);
  const int n_sel = __popc(__ballot_sync(0xffffffffu, my_idx >= 0));

  // ---- 2. exact scores of the survivors ---------------------------------------------------
  // four gallery rows per pass, all their 128-bit loads issued before the first use (memory-level
  // parallelism: the kernel is a latency-bound gather otherwise); fp64 accumulation of explicitly
  // formed differences; lane r keeps the sums of survivor r and evaluates its own distance afterwards
  double xsq = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    xsq += (double)qv[i].x * qv[i].x + (double)qv[i].y * qv[i].y + (double)qv[i].z * qv[i].z +
           (double)qv[i].w * qv[i].w;
  xsq = warp_sum(xsq);

  double my_s = 0.0, my_y = 0.0;
  constexpr int PASS = PRUNED ? 2 : 4;
  constexpr int CH = NV < 4 ? NV : 4;            // float4 chunks per lane and row in flight at once
  for (int r0 = 0; r0 < n_sel; r0 += PASS) {
    const float4* g[PASS];
    bool val[PASS];
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      const int id = __shfl_sync(0xffffffffu, my_idx, (r0 + t) & 31);
      val[t] = (r0 + t < n_sel);
      HYPRET_CHECK(!val[t] || (id >= 0 && id < N));
      g[t] = reinterpret_cast<const float4*>(g32 + (int64_t)(val[t] ? id : 0) * d);
    }
    double sacc[PASS], yacc[PASS];
#pragma unroll
    for (int t = 0; t < PASS; ++t) { sacc[t] = 0.0; yacc[t] = 0.0; }
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
        const int i = i0 + ii < NV ? i0 + ii : NV - 1;      // NV = 6: the last chunk is half empty (b == 0 there,
        if (i0 + ii >= NV) continue;                        // and the compile-time guard keeps qv[] in range)
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
          yacc[t] += (double)bb.x * bb.x + (double)bb.y * bb.y + (double)bb.z * bb.z + (double)bb.w * bb.w;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      const double s0 = warp_sum(sacc[t]), y0 = warp_sum(yacc[t]);
      if (lane == r0 + t) { my_s = s0; my_y = y0; }
    }
  }
  double my_key = INFINITY;     // sort key (distance, or minus similarity)
  double my_sur = INFINITY;     // exact surrogate, comparable with the approximate scores
  if (my_idx >= 0) {
    if (metric == HYPRET_METRIC_HYPERBOLIC) {
      const double cc = (double)c;
      const double al = 1.0 - cc * xsq;
      const double t0 = 2.0 * cc * my_s / (al * (1.0 - cc * my_y));
      my_key = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
      my_sur = cc * my_s / (1.0 - cc * my_y);          // unit-ball coordinates, like the filter's surrogate
    } else {
      const double nx = sqrt(xsq);
      const double d0 = (nx == 0.0 ? 1.0 : nx) * (my_y == 0.0 ? 1.0 : sqrt(my_y));
      my_key = -(my_s / d0);
      my_sur = my_key;
    }
    // order by the fp32 value that is written out (ties -> lower index): the result is then a function of the
    // emitted (score, index) pairs alone, so per-shard lists merge to exactly the single-GPU list
    my_key = (double)(float)my_key;
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
    out_score[out_row * k + lane] = valid ? (float)val : ((metric == HYPRET_METRIC_HYPERBOLIC) ? INFINITY : -INFINITY);
    out_idx[out_row * k + lane] = valid ? (int64_t)my_idx + idx_offset : (int64_t)-1;
  }
  if (out_margin != nullptr || cert.q_err != nullptr) {
    const double kth = __shfl_sync(0xffffffffu, my_sur, k - 1);
    const float kth_val = __shfl_sync(0xffffffffu, (float)((metric == HYPRET_METRIC_HYPERBOLIC) ? my_key : -my_key), k - 1);
    const int kth_idx = __shfl_sync(0xffffffffu, my_idx, k - 1);
    // +inf: the candidate set was not truncated (fewer than k' valid candidates survive)
    const int n_valid = __popc(__ballot_sync(0xffffffffu, my_idx >= 0));
    if (lane == 0) {
      // limit = a lower bound of the filter score of every row OUTSIDE the rescored set (+inf: nothing was left out)
      const float limit = fminf(n_valid >= ksel ? worst_approx : INFINITY, w_trunc);
      const bool open_set = !(limit < INFINITY);
      const double margin = (double)limit - kth;
      if (out_margin != nullptr) out_margin[q] = open_set ? INFINITY : (float)margin;
      if (cert.q_err != nullptr) {
        // Every row outside the candidate set has a tensor-core surrogate >= worst_approx and an exact surrogate within
        // E of it, so margin > E proves that none of them precedes the k-th result.  qn = norm of the
        // query operand's main columns (x on the ball; a unit vector for cosine).
        const double qn = metric == HYPRET_METRIC_HYPERBOLIC ? sqrt((double)c * xsq) : 1.0;
        const double zmax = cert.g_stats[0], dzmax = cert.g_stats[1], rbmax = cert.g_stats[2], bmax = cert.g_stats[3];
        const double E = (double)cert.q_err[q] * zmax + qn * dzmax +
                         (double)cert.slack * (qn * zmax + qn * qn * rbmax + bmax);
        const bool ok = open_set || (kth_idx >= 0 && margin > E);
        if (cert.flags != nullptr) cert.flags[q] = ok ? 1 : 0;
        if (!ok) {
          if (cert.bound != nullptr) cert.bound[q] = kth_idx >= 0 ? kth_val : (metric == HYPRET_METRIC_HYPERBOLIC ? INFINITY : -INFINITY);
          cert.state[2 * q] = 0;
          cert.state[2 * q + 1] = 0;
          const int at = atomicAdd(cert.count, 1);
          HYPRET_CHECK(at >= 0 && at < Q);
          cert.list[at] = (int)q;
        }
      }
    }
  }
}

// Approximate merge only: the k' best candidates of every query by surrogate score, ascending (surrogate, index),
// padded with (+inf, -1).  Multi-GPU pruning step: the surrogates of different gallery shards are comparable, so
// the owner of a query can find its GLOBAL k'-th best surrogate from W such lists before any exact rescoring.
__global__ void __launch_bounds__(RR_WARPS * 32)
cand_select_kernel(const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                   const int32_t* __restrict__ list_count, int64_t Q, int n_cand, int kprime,
                   float* __restrict__ sel_score, int32_t* __restrict__ sel_idx, const PeerRoute route,
                   int64_t recv_off) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * RR_WARPS + warp;
  if (q >= Q) return;
  float* cs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * n_cand;
  int* ci = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)RR_WARPS * n_cand) +
            (size_t)warp * n_cand;
  const int n_mine = list_count != nullptr ? min(list_count[q] * kprime, n_cand) : n_cand;
  for (int t = lane; t < n_mine; t += 32) {
    cs[t] = cand_score[q * n_cand + t];
    ci[t] = cand_idx[q * n_cand + t];
  }
  __syncwarp();
  float my_score, worst;
  const int my_idx = select_candidates(cs, ci, n_mine, kprime, INFINITY, lane, &my_score, &worst);
  if (lane < kprime) {
    sel_score[q * kprime + lane] = my_score;
    sel_idx[q * kprime + lane] = my_idx;
    if (route.n > 0) {      // the all_to_all of the surrogates: the owner of query q receives this shard's list
      int owner;
      const int64_t row = route_row(route, q, &owner);
      reinterpret_cast<float*>(route.base[owner] + recv_off)[row * kprime + lane] = my_score;
    }
  }
}

// ----------------------------------------------------------------------------- wide top-k (k up to 128)
// One CTA (128 threads) per query.  The candidate lists (n_lists x kprime, built WITHOUT threshold
// sharing so that each list is exactly its strip's top-kprime) are sorted by surrogate score with a
// block bitonic network in shared memory; the best 256 survive, are rescored exactly (one warp per
// 64 survivors, same fp64 arithmetic as above), sorted again by (score, index), and the first k are
// written.  out_margin = (smallest surrogate a non-candidate can have) - (exact surrogate of the
// k-th result): a non-candidate is either cut by the 256-survivor truncation or hidden behind a FULL
// list's worst entry.
constexpr int RW_THREADS = 128;
constexpr int RW_SURV = 256;   // survivors rescored exactly: 2x the largest k, the fp16 filter error is comparable
                               // with the score gap between rank k and rank 1.3k on concentrated data

__device__ __forceinline__ bool pair_less(float a, int ia, float b, int ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  return (a < b) || (a == b && ia < ib);
}

// Two launches share the queries by candidate count (size_class): the lists a query actually has (list_count) are
// far fewer than the slots the schedule reserves, so class 0 -- queries with at most RW_SMALL candidates -- runs with
// 8 KB of dynamic shared memory and 4 resident CTAs per SM (ncu of the single-launch version: 64 KB of candidate
// space per CTA and 168 registers held it at 3 CTAs = 12 warps per SM, DRAM at 17 % of peak: latency-bound), class 1
// takes the rest with the full-size buffer.  A CTA whose query belongs to the other class exits at once.
constexpr int RW_SMALL = 1024;

// YPRE: ||g||^2 of the gallery rows comes precomputed in fp64 (hypret_row_sqnorm64, once per index) instead of
// being re-accumulated for each of the 256 survivors of each query: ncu's source page had the float -> double
// conversions (F2F, a quarter-rate pipe: 13k warp-instructions per query, 7 % of all stall samples on one of them)
// and the DFMAs (20k) at the top; the row norm was half of both.
template <int NV, bool YPRE>
__global__ void __launch_bounds__(RW_THREADS, NV <= 6 ? 4 : 1)
rerank_wide_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int64_t Q, int64_t N, int d, float c,
                   int metric, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                   const int32_t* __restrict__ list_count, int n_lists_alloc, int kprime, int n_pad, int k,
                   int64_t idx_offset, float* __restrict__ out_score,
                   int64_t* __restrict__ out_idx, float* __restrict__ out_margin, int size_class,
                   const double* __restrict__ g_sq64) {
  extern __shared__ uint8_t smem_raw[];
  float* ks = reinterpret_cast<float*>(smem_raw);            // [n_pad] surrogate scores
  int* ki = reinterpret_cast<int*>(ks + n_pad);              // [n_pad] gallery ids
  __shared__ double ekey[RW_SURV];
  __shared__ double esur[RW_SURV];
  __shared__ int eidx[RW_SURV];
  __shared__ unsigned hidden_key;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q = blockIdx.x;
  const int n_lists = list_count != nullptr ? min(list_count[q], n_lists_alloc) : n_lists_alloc;   // compact slots
  const int n_cand = n_lists * kprime;
  const int64_t row_stride = (int64_t)n_lists_alloc * kprime;
  if (size_class >= 0 && (n_cand <= RW_SMALL) != (size_class == 0)) return;       // the other launch's query

  __shared__ int n_valid_s;
  if (tid == 0) { hidden_key = 0xffffffffu; n_valid_s = 0; }
  __syncthreads();
  // a FULL list may hide rows no better than its worst entry
  for (int l = tid; l < n_lists; l += RW_THREADS) {
    bool full = true;
    float worst = -INFINITY;
    for (int e = 0; e < kprime; ++e) {
      const int id = cand_idx[q * row_stride + l * kprime + e];
      full &= id >= 0;
      if (id >= 0) worst = fmaxf(worst, cand_score[q * row_stride + l * kprime + e]);
    }
    if (full) atomicMin(&hidden_key, ordered_key(worst));
  }
  // compact the valid candidates to the front (most list slots of a query are empty: slots belong to
  // strips of other query tiles); the order is fixed by the sort below
  for (int t = tid; t < n_cand; t += RW_THREADS) {
    const int id = cand_idx[q * row_stride + t];
    if (id >= 0) {
      const int pos = atomicAdd(&n_valid_s, 1);
      ks[pos] = cand_score[q * row_stride + t];
      ki[pos] = id;
    }
  }
  __syncthreads();
  const int n_valid = n_valid_s;
  int n_sort = RW_SURV;
  while (n_sort < n_valid) n_sort <<= 1;                 // block-uniform, <= n_pad
  for (int t = n_valid + tid; t < n_sort; t += RW_THREADS) { ks[t] = INFINITY; ki[t] = -1; }
  __syncthreads();
  // ---- block bitonic sort of the candidates by (surrogate, id) -------------------------------------------
  for (int size = 2; size <= n_sort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (n_sort >> 1); t += RW_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = ks[lo], b = ks[hi];
        const int ia = ki[lo], ib = ki[hi];
        const bool swap = asc ? pair_less(b, ib, a, ia) : pair_less(a, ia, b, ib);
        if (swap) { ks[lo] = b; ks[hi] = a; ki[lo] = ib; ki[hi] = ia; }
      }
      __syncthreads();
    }
  }
  const int n_surv = RW_SURV;
  const float cut = (n_sort > RW_SURV && ki[RW_SURV] >= 0) ? ks[RW_SURV] : INFINITY;   // first row cut by truncation

  // ---- exact rescoring: warp w takes survivors [64w, 64w+64) --------------------------------------------------
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
  constexpr int PASS = 4;
  constexpr int PER_WARP = RW_SURV / (RW_THREADS / 32);
  for (int r0 = warp * PER_WARP; r0 < (warp + 1) * PER_WARP && r0 < n_surv; r0 += PASS) {
    int idx[PASS];
    bool val[PASS];
    const float4* g[PASS];
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      idx[t] = (r0 + t < n_surv) ? ki[r0 + t] : -1;
      val[t] = idx[t] >= 0;
      g[t] = reinterpret_cast<const float4*>(g32 + (int64_t)(val[t] ? idx[t] : 0) * d);
    }
    double sacc[PASS], yacc[PASS];
#pragma unroll
    for (int t = 0; t < PASS; ++t) { sacc[t] = 0.0; yacc[t] = 0.0; }
    constexpr int CH = NV < 4 ? NV : 4;          // all loads of a chunk are issued before their first use
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
          if (!YPRE) yacc[t] += (double)bb.x * bb.x + (double)bb.y * bb.y + (double)bb.z * bb.z + (double)bb.w * bb.w;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < PASS; ++t) {
      const double s0 = warp_sum(sacc[t]);
      const double y0 = YPRE ? (val[t] ? g_sq64[idx[t]] : 0.0) : warp_sum(yacc[t]);
      double key0 = INFINITY, sur0 = INFINITY;
      if (val[t]) {
        if (metric == HYPRET_METRIC_HYPERBOLIC) {
          const double cc = (double)c;
          const double al = 1.0 - cc * xsq;
          const double t0 = 2.0 * cc * s0 / (al * (1.0 - cc * y0));
          key0 = log1p(t0 + sqrt(t0 * (t0 + 2.0))) / sqrt(cc);
          sur0 = cc * s0 / (1.0 - cc * y0);
        } else {
          const double nx = sqrt(xsq);
          const double d0 = (nx == 0.0 ? 1.0 : nx) * (y0 == 0.0 ? 1.0 : sqrt(y0));
          key0 = -(s0 / d0);
          sur0 = key0;
        }
        key0 = (double)(float)key0;      // order by the emitted fp32 value, ties -> lower index (see rerank_kernel)
      }
      if (lane == 0 && r0 + t < RW_SURV) { ekey[r0 + t] = key0; esur[r0 + t] = sur0; eidx[r0 + t] = val[t] ? idx[t] : -1; }
    }
  }
  __syncthreads();
  // ---- sort the exact scores by (key, id): one compare-exchange pair per thread -----------------------------------
  for (int size = 2; size <= RW_SURV; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < RW_SURV / 2) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const double a = ekey[lo], b = ekey[hi];
        const int ia = eidx[lo], ib = eidx[hi];
        const bool swap = asc ? key_less(b, ib, a, ia) : key_less(a, ia, b, ib);
        if (swap) {
          ekey[lo] = b; ekey[hi] = a; eidx[lo] = ib; eidx[hi] = ia;
          const double sa = esur[lo]; esur[lo] = esur[hi]; esur[hi] = sa;
        }
      }
      __syncthreads();
    }
  }
  if (tid < k) {
    const bool valid = eidx[tid] >= 0;
    const double val = (metric == HYPRET_METRIC_HYPERBOLIC) ? ekey[tid] : -ekey[tid];
    out_score[q * k + tid] = valid ? (float)val : ((metric == HYPRET_METRIC_HYPERBOLIC) ? INFINITY : -INFINITY);
    out_idx[q * k + tid] = valid ? (int64_t)eidx[tid] + idx_offset : (int64_t)-1;
  }
  if (out_margin != nullptr && tid == 0) {
    const float hb = fminf(hidden_key == 0xffffffffu ? INFINITY : ordered_val(hidden_key), cut);
    out_margin[q] = (eidx[k - 1] < 0 || hb == INFINITY) ? INFINITY : (float)((double)hb - esur[k - 1]);
  }
}

}  // namespace

int hypret_launch_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                              int n_cand, int kprime, float* sel_score, int32_t* sel_idx,
                              const hypret_peer_route* route, int64_t recv_off, cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  const size_t smem = (size_t)RR_WARPS * n_cand * 8;
  if (smem > 200 * 1024) return HYPRET_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(cand_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cand_select_kernel<<<(unsigned)((Q + RR_WARPS - 1) / RR_WARPS), RR_WARPS * 32, smem, stream>>>(
      cand_score, cand_idx, list_count, Q, n_cand, kprime, sel_score, sel_idx, make_route(route), recv_off);
  return (int)cudaGetLastError();
}

int hypret_launch_rerank(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_cand,
                         int kprime, int k, int64_t idx_offset, const float* prune_thr, float* out_score,
                         int64_t* out_idx, float* out_margin, const hypret_peer_route* route, int64_t score_off,
                         int64_t idx_off, const double* g_sq64, const CertArgs& cert, cudaStream_t stream) {
  if (Q == 0) return HYPRET_OK;
  if (kprime > 32 || k > 32 || k > kprime) {
    if (prune_thr != nullptr || (route != nullptr && route->n_ranks > 0)) return HYPRET_EUNSUPPORTED;
    if (cert.q_err != nullptr) return HYPRET_EUNSUPPORTED;     // the wide path has its own hidden-bound certificate
    const int n_lists = n_cand / kprime;
    int n_pad = RW_SURV;
    while (n_pad < n_cand) n_pad <<= 1;
    const size_t smem_w = (size_t)n_pad * 8;
    const int need_w = (d + 127) / 128;
#define HYPRET_RERANK_WIDE_Y(NV, YP)                                                                                \
  do {                                                                                                              \
    if (smem_w > 40 * 1024) {                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(rerank_wide_kernel<NV, YP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem_w);                                                            \
      if (e != cudaSuccess) return (int)e;                                                                          \
    }                                                                                                               \
    if (n_pad > RW_SMALL && list_count != nullptr) {                                                                \
      rerank_wide_kernel<NV, YP><<<(unsigned)Q, RW_THREADS, (size_t)RW_SMALL * 8, stream>>>(                        \
          q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_lists, kprime, RW_SMALL, k, idx_offset, \
          out_score, out_idx, out_margin, 0, g_sq64);                                                               \
      cudaError_t e0 = cudaGetLastError();                                                                          \
      if (e0 != cudaSuccess) return (int)e0;                                                                        \
    }                                                                                                               \
    rerank_wide_kernel<NV, YP><<<(unsigned)Q, RW_THREADS, smem_w, stream>>>(                                        \
        q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_lists, kprime, n_pad, k, idx_offset,      \
        out_score, out_idx, out_margin, (n_pad > RW_SMALL && list_count != nullptr) ? 1 : -1, g_sq64);              \
    return (int)cudaGetLastError();                                                                                 \
  } while (0)
#define HYPRET_RERANK_WIDE(NV)                                                                                      \
  do {                                                                                                              \
    if (g_sq64 != nullptr) HYPRET_RERANK_WIDE_Y(NV, true);                                                          \
    HYPRET_RERANK_WIDE_Y(NV, false);                                                                                \
  } while (0)
    if (need_w <= 1) HYPRET_RERANK_WIDE(1);
    if (need_w <= 2) HYPRET_RERANK_WIDE(2);
    if (need_w <= 4) HYPRET_RERANK_WIDE(4);
    if (need_w <= 6) HYPRET_RERANK_WIDE(6);
    if (need_w <= 8) HYPRET_RERANK_WIDE(8);
    if (need_w <= 16) HYPRET_RERANK_WIDE(16);
#undef HYPRET_RERANK_WIDE
#undef HYPRET_RERANK_WIDE_Y
    return HYPRET_EUNSUPPORTED;
  }
  const size_t smem = (size_t)RR_WARPS * n_cand * 8;
  if (smem > 200 * 1024) return HYPRET_EUNSUPPORTED;
  const int64_t grid = (Q + RR_WARPS - 1) / RR_WARPS;
  const int need = (d + 127) / 128;
#define HYPRET_RERANK_LAUNCH_P(NV, PR)                                                                              \
  do {                                                                                                              \
    if (smem > 48 * 1024) {                                                                                         \
      cudaError_t e =                                                                                               \
          cudaFuncSetAttribute(rerank_kernel<NV, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      if (e != cudaSuccess) return (int)e;                                                                          \
    }                                                                                                               \
    rerank_kernel<NV, PR><<<(unsigned)grid, RR_WARPS * 32, smem, stream>>>(                                         \
        q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_cand, kprime, k, idx_offset, prune_thr,   \
        out_score, out_idx, out_margin, make_route(route), score_off, idx_off, cert);                               \
    return (int)cudaGetLastError();                                                                                 \
  } while (0)
#define HYPRET_RERANK_LAUNCH(NV)                                                                                    \
  do {                                                                                                              \
    if (prune_thr != nullptr) HYPRET_RERANK_LAUNCH_P(NV, true);                                                     \
    HYPRET_RERANK_LAUNCH_P(NV, false);                                                                              \
  } while (0)
  if (need <= 1) HYPRET_RERANK_LAUNCH(1);
  if (need <= 2) HYPRET_RERANK_LAUNCH(2);
  if (need <= 4) HYPRET_RERANK_LAUNCH(4);
  if (need <= 6) HYPRET_RERANK_LAUNCH(6);
  if (need <= 8) HYPRET_RERANK_LAUNCH(8);
  if (need <= 16) HYPRET_RERANK_LAUNCH(16);
#undef HYPRET_RERANK_LAUNCH
#undef HYPRET_RERANK_LAUNCH_P
  return HYPRET_EUNSUPPORTED;
}

namespace {
// ||x_i||^2 in fp64, one warp per row, lanes striding the row's float4 chunks and the butterfly sum of the rerank
// kernels (same accumulation order as the in-kernel version it replaces).
__global__ void __launch_bounds__(256)
row_sqnorm64_kernel(const float* __restrict__ x, int64_t n, int d, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float4* row = reinterpret_cast<const float4*>(x + i * d);
  double acc = 0.0;
  for (int j = lane; j < (d >> 2); j += 32) {
    const float4 b = __ldg(row + j);
    acc += (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z + (double)b.w * b.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[i] = acc;
}
}  // namespace

int hypret_launch_row_sqnorm64(const float* x, int64_t n, int d, double* out, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  row_sqnorm64_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(x, n, d, out);
  return (int)cudaGetLastError();
}
