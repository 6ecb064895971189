// train_hyp (BASELINE config 5) as a flash-style step: the n x m Poincare distance matrix of the in-batch InfoNCE is
// NEVER written to memory, forward or backward, and every dense product runs on tcgen05.
//
// Replaces the reference's O(n^2) Python double loop of 1x1 pmath.dist calls + autograd through ~40 nodes per pair
// (/root/reference/src/train.py:1832-1846 rows-only CE; 2304-2334 symmetric CE).  Maths: SURVEY 7.4.
//
//   flash_prep        x [n,D] fp32 -> fp16 2-way split Gram operands in the row layout [hi|lo|hi] and the column layout
//                     [hi|hi|lo] (<R(a), C(p)> = hi.hi + lo.hi + hi.lo = <a,p> to 2^-22), transposed bf16 hi/mid planes
//                     [2, D, n] (B operand of the gradient product), and |x|^2
//   flash_lse         per row i: logsumexp_j(-d_ij / tau).  Gram tile (128 x 128) by tcgen05 into TMEM; the epilogue
//                     (lane = row) forms s = |a|^2 + |p|^2 - 2<a,p>, t = 2 c s / (alpha beta), the logit
//                     -lg2(1 + t + sqrt(t (t+2))) / (tau sqrt c) in log2 units and folds it into a running (max, sum).
//                     Only O(n) partials leave the SM.
//   flash_grad        per row block: the tile is RECOMPUTED, the softmax weights come from the stored log-sum-exps,
//                     w_ij = g_ij 4 sqrt(c) / (alpha_i beta_j sqrt(z^2 - 1)) is formed in registers, cut into two bf16
//                     planes (hi + mid) written to SHARED MEMORY in the UMMA K-major 128B-swizzle layout, and a second
//                     tcgen05.mma accumulates W Y into a TMEM accumulator [128 x D] that lives across all column tiles
//                     of the row block (three bf16 products: hi.hi + hi.mid + mid.hi).  dX = x rowsum - W Y.
//                     Called twice per step with the roles of anchors and positives swapped (W^T A is W' P' of the
//                     transposed problem), so each launch owns its output rows: no atomics, deterministic.
//
// Near pairs (s < (|a|^2+|p|^2)/4: the diagonal of a contrastive batch), where the Gram form cancels, are recomputed
// from the fp32 rows with explicit differences, as in gramdist.cu.  Only ABSOLUTE accuracy of the logits matters to the
// loss and to the softmax weights (lg2.approx: 2^-22 absolute; times 1/(tau sqrt c) ~ 14-20 => ~5e-6).
//
// Work decomposition: tiles in row-major order, CTA c owns the contiguous run [c T / P, (c+1) T / P): at most a few
// CTAs share a row block; each writes its partial (running max / sum, or the [128, D] partial product and row sums) into
// slot (c - first CTA of the row block) and a small finishing kernel combines the slots in a fixed order.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int FT_M = 128;                 // rows per tile (TMEM lanes)
constexpr int FT_N = 128;                 // columns per tile
constexpr int F_NWG = 2;                  // epilogue warpgroups: each takes FT_N / F_NWG columns of every tile
constexpr int F_CW = FT_N / F_NWG;        // columns per warpgroup and tile (a multiple of 32)
constexpr int F_EPI = 128 * F_NWG;
constexpr int F_THREADS = 64 + F_EPI;     // warp 0 TMA, warp 1 MMA, then the epilogue warpgroups
constexpr int F_STAGES_FWD = 4, F_STAGES_BWD = 2;
constexpr int FA_BLK = FT_M * HYPRET_KBLK * 2;      // 16 KB: one K-block of the row operand
constexpr int FB_BLK = FT_N * HYPRET_KBLK * 2;      // 16 KB: ... of the column operand
constexpr int F_STAGE = FA_BLK + FB_BLK;
constexpr int FW_PLANE = FT_M * FT_N * 2;           // 32 KB: one bf16 plane of W
constexpr float F_NEAR = 0.25f;
constexpr int F_MAX_D = 128;                        // gradient product: N = D columns of one TMEM accumulator
constexpr int F_COLS = 3;                           // per-column constants staged in shared memory

__host__ __device__ inline int flash_kpad(int d) { return (3 * d + HYPRET_KBLK - 1) / HYPRET_KBLK * HYPRET_KBLK; }

// ------------------------------------------------------------------------------------------------ operand preparation
__global__ void __launch_bounds__(256)
flash_prep_kernel(const float* __restrict__ x, int64_t n, int d, __half* __restrict__ row_op, __half* __restrict__ col_op,
                  float* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int kp = flash_kpad(d);
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = x[i * d + k];
    acc = fmaf(v, v, acc);
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    if (row_op != nullptr) { __half* o = row_op + i * kp; o[k] = hi; o[d + k] = lo; o[2 * d + k] = hi; }
    if (col_op != nullptr) { __half* o = col_op + i * kp; o[k] = hi; o[d + k] = hi; o[2 * d + k] = lo; }
  }
  for (int k = 3 * d + lane; k < kp; k += 32) {
    if (row_op != nullptr) row_op[i * kp + k] = __float2half_rn(0.f);
    if (col_op != nullptr) col_op[i * kp + k] = __float2half_rn(0.f);
  }
  acc = warp_sum(acc);
  if (sq != nullptr && lane == 0) sq[i] = acc;
}

// x [n,d] fp32 -> t [2, d, n_pad] bf16: plane 0 = hi, plane 1 = mid (hi + mid = x to 2^-17), transposed through a
// 32 x 33 shared-memory tile so that both the reads and the writes are coalesced
__global__ void __launch_bounds__(256)
flash_transpose_kernel(const float* __restrict__ x, int64_t n, int d, int64_t n_pad, __nv_bfloat16* __restrict__ t) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = (r0 + r < n && k0 + tx < d) ? x[(r0 + r) * d + k0 + tx] : 0.f;
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    if (k0 + k < d && r0 + tx < n_pad) {
      const float v = tile[tx][k];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      t[(int64_t)(k0 + k) * n_pad + r0 + tx] = hi;
      t[((int64_t)d + k0 + k) * n_pad + r0 + tx] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

// ------------------------------------------------------------------------------------------------ the tile kernels
struct FBarriers {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t yt_full;
  uint64_t w_full;
  uint64_t w_free;
  uint64_t acc_full;
  uint64_t acc_free;
  uint32_t tmem_ptr;
};

struct FParams {
  const float* x32; const float* y32;      // fp32 rows (exact recompute of near pairs)
  const float* xsq; const float* ysq;
  const float* x_lse; const float* y_lse;  // BWD: natural-log log-sum-exps of the rows / of the columns (or NULL)
  int64_t n, m;
  int d, kb;                               // feature dimension, Gram K blocks
  float c, kappa;                          // kappa = 1 / (tau sqrt c)
  float wx, wy;                            // BWD: weights of the row-wise / column-wise CE terms
  const float* grad_scale;                 // BWD: device scalar dL/dloss (NULL = 1)
  float coef;                              // BWD: 1 / (tau n_total)
  int64_t diag_offset;                     // the target of row i is column i + diag_offset
  int n_rt, n_ct, n_slots;
  float* part;                             // FWD: [n_rt, n_slots, F_NWG, 2, 128] (max, sum); BWD: [n_slots, n_rt*128, d]
  float* part_rs;                          // BWD: [n_slots, F_NWG, n_rt*128] partial row sums
  int64_t yt_cols;                         // BWD: padded column count of the transposed planes
  unsigned long long* stats;               // HYPRET_FLASH_STATS=1: [grid][8] wait-cycle counters of epilogue warp 2, or NULL
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float flash_exact_sq(const float* __restrict__ a, const float* __restrict__ p, int d) {
  float s = 0.f;
  for (int k = 0; k < d; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + k), y = *reinterpret_cast<const float4*>(p + k);
    const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
    s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
  }
  return s;
}


// ---- per-chunk epilogue arithmetic (32 columns of one row per thread; v[] = Gram entries on entry) --------------------
// s = |x|^2 + |y|^2 - 2 <x,y> in place; returns the chunk minimum (near-pair detection)
__device__ __forceinline__ float chunk_sqdist(float (&v)[32], const float4* __restrict__ cn, float na) {
  float smin = INFINITY;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 nb = cn[j4];
    v[4 * j4 + 0] = fmaf(-2.0f, v[4 * j4 + 0], na) + nb.x;
    v[4 * j4 + 1] = fmaf(-2.0f, v[4 * j4 + 1], na) + nb.y;
    v[4 * j4 + 2] = fmaf(-2.0f, v[4 * j4 + 2], na) + nb.z;
    v[4 * j4 + 3] = fmaf(-2.0f, v[4 * j4 + 3], na) + nb.w;
    smin = fminf(smin, fminf(fminf(v[4 * j4], v[4 * j4 + 1]), fminf(v[4 * j4 + 2], v[4 * j4 + 3])));
  }
  return smin;
}

// The arithmetic below is written STAGE-WISE over sub-chunks of F_SUB entries (F_SUB independent instructions per
// stage), so that no instruction waits on the one before it.  Measured (ncu, n = 8192, D = 128): 23 warp instructions
// per entry in the forward pass, issue slots 48 % busy, XU (rsqrt / lg2 / ex2) 53 %, FMA 25 %, ALU 23 %, tensor 25 %:
// the pass is latency-bound with 2.5 resident warps per scheduler, not bound by any one pipe; neither four epilogue
// warpgroups (96 registers, spills) nor other block sizes (4 / 16) moved it.
//
// forward: v[] <- L = lg2(z + sqrt(z^2 - 1)), z = 1 + s (2c/beta) / alpha; returns min L  (logit = -kappa L, log2 units)
constexpr int F_SUB = 8;                  // entries per stage block (see above); multiple of 4, divides 32

__device__ __forceinline__ float chunk_logits(float (&v)[32], const float4* __restrict__ ca, float rho) {
  float lmin = INFINITY;
#pragma unroll
  for (int h = 0; h < 32 / F_SUB; ++h) {
    float z[F_SUB], q[F_SUB];
#pragma unroll
    for (int j4 = 0; j4 < F_SUB / 4; ++j4) {
      const float4 a1 = ca[(F_SUB / 4) * h + j4];
      z[4 * j4 + 0] = v[F_SUB * h + 4 * j4 + 0] * (a1.x * rho);          // z - 1
      z[4 * j4 + 1] = v[F_SUB * h + 4 * j4 + 1] * (a1.y * rho);
      z[4 * j4 + 2] = v[F_SUB * h + 4 * j4 + 2] * (a1.z * rho);
      z[4 * j4 + 3] = v[F_SUB * h + 4 * j4 + 3] * (a1.w * rho);
    }
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) q[j] = fmaxf(fmaf(z[j], z[j], z[j] + z[j]), 1e-30f);      // z^2 - 1 without cancellation
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) q[j] *= rsqrt_approx(q[j]);          // sqrt(z^2 - 1)
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = (1.0f + z[j]) + q[j];
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = lg2_approx(z[j]);
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) { v[F_SUB * h + j] = z[j]; lmin = fminf(lmin, z[j]); }
  }
  return lmin;
}

// forward, software-pipelined by one chunk: the logits of THIS chunk (FMA-heavy, 2 MUFU per entry) together with the
// exponential sum of the PREVIOUS chunk (pv[], 1 MUFU + 2 FP32 per entry) in one branch-free block.  On their own the 32
// ex2 of a chunk's sum issue back to back -- the XU pipe (8 cycles per warp instruction) is the only unit working and
// the other warp of the scheduler, in the same phase, wants it too -- while the logits' FMA stages leave it idle: ncu
// put 34 % of all warp samples on MUFU instructions at an XU utilisation of 56 %.  Interleaved, the FMA work of one
// chunk hides behind the XU time of the other.  km = kappa * (running min), 0 while nothing has been seen.
__device__ __forceinline__ float chunk_logits_sum(float (&v)[32], const float4* __restrict__ ca, float rho,
                                                  const float (&pv)[32], float km, float kappa, float& run_s) {
  float lmin = INFINITY;
#pragma unroll
  for (int h = 0; h < 32 / F_SUB; ++h) {
    float z[F_SUB], q[F_SUB], e[F_SUB];
#pragma unroll
    for (int j4 = 0; j4 < F_SUB / 4; ++j4) {
      const float4 a1 = ca[(F_SUB / 4) * h + j4];
      z[4 * j4 + 0] = v[F_SUB * h + 4 * j4 + 0] * (a1.x * rho);          // z - 1
      z[4 * j4 + 1] = v[F_SUB * h + 4 * j4 + 1] * (a1.y * rho);
      z[4 * j4 + 2] = v[F_SUB * h + 4 * j4 + 2] * (a1.z * rho);
      z[4 * j4 + 3] = v[F_SUB * h + 4 * j4 + 3] * (a1.w * rho);
    }
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) e[j] = fmaf(-kappa, pv[F_SUB * h + j], km);               // previous chunk: exponents
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) q[j] = fmaxf(fmaf(z[j], z[j], z[j] + z[j]), 1e-30f);      // z^2 - 1 without cancellation
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) { q[j] *= rsqrt_approx(q[j]); e[j] = ex2_approx(e[j]); }  // sqrt(z^2 - 1) | 2^e
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = (1.0f + z[j]) + q[j];
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = lg2_approx(z[j]);
#pragma unroll
    for (int j = 0; j < F_SUB; j += 2) run_s += e[j] + e[j + 1];
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) { v[F_SUB * h + j] = z[j]; lmin = fminf(lmin, z[j]); }
  }
  return lmin;
}

// backward: the weights w_ij of the chunk, rounded to two bf16 planes (packed pairs hp / mp), and the row sum of the
// ROUNDED weights times (1 + c s / alpha): the two terms of dX = x rowsum - W Y then carry the same perturbation of w
// and their (large) common part still cancels.  gx / gy / gd already contain gs * 2 / sqrt(c).  Invalid columns have
// 2c/beta = 0 in shared memory and invalid rows gx = gy = gd = 0, so both come out as w = 0 without a test.
template <bool DIAG, bool COLS>
__device__ __forceinline__ void chunk_weights(const float (&v)[32], const float4* __restrict__ ca,
                                              const float4* __restrict__ cl, float rho, float crho, float kappa, float lx,
                                              float gx, float gy, float gd, int dj, float& rowsum, uint32_t (&hp)[16],
                                              uint32_t (&mp)[16]) {
#pragma unroll
  for (int h = 0; h < 32 / F_SUB; ++h) {
    float ar[F_SUB], z[F_SUB], q[F_SUB], g[F_SUB];
#pragma unroll
    for (int j4 = 0; j4 < F_SUB / 4; ++j4) {
      const float4 a1 = ca[(F_SUB / 4) * h + j4];
      ar[4 * j4 + 0] = a1.x * rho; ar[4 * j4 + 1] = a1.y * rho; ar[4 * j4 + 2] = a1.z * rho; ar[4 * j4 + 3] = a1.w * rho;
    }
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = v[F_SUB * h + j] * ar[j];          // z - 1
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) q[j] = fmaxf(fmaf(z[j], z[j], z[j] + z[j]), 1e-30f);      // z^2 - 1 without cancellation
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) g[j] = rsqrt_approx(q[j]);            // 1 / sqrt(z^2 - 1)
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) { z[j] = (1.0f + z[j]) + q[j] * g[j]; ar[j] *= g[j]; }    // z + sqrt(z^2-1);  ar / sqrt(z^2-1)
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) z[j] = lg2_approx(z[j]);
    // softmax probabilities are <= 1: the clamp changes nothing for real entries, and keeps the padding columns
    // (logit 0 against a possibly very negative log-sum-exp) from overflowing into inf * 0
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) q[j] = fminf(fmaf(-kappa, z[j], -lx), 0.f);
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) g[j] = gx * ex2_approx(q[j]);
    if (COLS) {
#pragma unroll
      for (int j4 = 0; j4 < F_SUB / 4; ++j4) {
        const float4 l4 = cl[(F_SUB / 4) * h + j4];
        q[4 * j4 + 0] = fminf(fmaf(-kappa, z[4 * j4 + 0], -l4.x), 0.f);
        q[4 * j4 + 1] = fminf(fmaf(-kappa, z[4 * j4 + 1], -l4.y), 0.f);
        q[4 * j4 + 2] = fminf(fmaf(-kappa, z[4 * j4 + 2], -l4.z), 0.f);
        q[4 * j4 + 3] = fminf(fmaf(-kappa, z[4 * j4 + 3], -l4.w), 0.f);
      }
#pragma unroll
      for (int j = 0; j < F_SUB; ++j) g[j] = fmaf(gy, ex2_approx(q[j]), g[j]);
    }
    if (DIAG) {
#pragma unroll
      for (int j = 0; j < F_SUB; ++j) g[j] += (F_SUB * h + j) == dj ? gd : 0.f;
    }
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) g[j] *= ar[j];                        // w
#pragma unroll
    for (int j = 0; j < F_SUB; j += 2) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(g[j], g[j + 1]);
      const float h0 = __low2float(hh), h1 = __high2float(hh);
      const __nv_bfloat162 md = __floats2bfloat162_rn(g[j] - h0, g[j + 1] - h1);
      hp[(F_SUB / 2) * h + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
      mp[(F_SUB / 2) * h + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&md);
      z[j] = h0 + __low2float(md);
      z[j + 1] = h1 + __high2float(md);
    }
#pragma unroll
    for (int j = 0; j < F_SUB; ++j) rowsum = fmaf(z[j], fmaf(crho, v[F_SUB * h + j], 1.0f), rowsum);
  }
}

// first CTA that touches row block rb / CTA that owns tile t, for runs [c T / P, (c+1) T / P)
__host__ __device__ inline int flash_cta_of_tile(int64_t t, int64_t T, int P) { return (int)(((t + 1) * P - 1) / T); }

template <bool BWD>
__global__ void __launch_bounds__(F_THREADS, 1)
flash_tile_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                  const __grid_constant__ CUtensorMap map_yt, const FParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGES = BWD ? F_STAGES_BWD : F_STAGES_FWD;
  uint8_t* ring = smem;
  uint8_t* w_sm = smem + STAGES * F_STAGE;                          // BWD: 2 planes [kblk][128 rows][64 cols] bf16
  uint8_t* yt_sm = w_sm + (BWD ? 2 * FW_PLANE : 0);                 // BWD: 2 planes [kblk][D rows][64 cols] bf16
  const int yt_plane = (FT_N / HYPRET_KBLK) * p.d * 128;            // bytes of one transposed plane tile
  float* cols = reinterpret_cast<float*>(yt_sm + (BWD ? 2 * yt_plane : 0));    // [2][F_COLS][128]
  FBarriers* bars = reinterpret_cast<FBarriers*>(cols + 2 * F_COLS * FT_N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t T = (int64_t)p.n_rt * p.n_ct;
  const int P = gridDim.x;
  // tile indices are 32-bit and (rt, ct) are tracked incrementally: 64-bit divisions by a run-time n_ct in every role's
  // tile loop cost several hundred instructions per tile (ncu: I2F / IMAD.WIDE chains among the hottest lines)
  const int t_begin = (int)((int64_t)blockIdx.x * T / P), t_end = (int)((int64_t)(blockIdx.x + 1) * T / P);
  const int rt_begin = t_begin / p.n_ct, ct_begin = t_begin - rt_begin * p.n_ct;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_y);
    if (BWD) tma_prefetch_desc(&map_yt);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->s_full[a], 1); mbar_init(&bars->s_empty[a], F_EPI); }
    mbar_init(&bars->yt_full, 1);
    mbar_init(&bars->w_full, F_EPI);
    mbar_init(&bars->w_free, 1);
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_free, F_EPI);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const uint32_t tmem_acc = tmem_base + 2 * FT_N;                  // BWD: the [128, D] product accumulator

  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t stage = 0, phase = 0, wfree_par = 0;
    int rt = rt_begin, ct = ct_begin;
    for (int t = t_begin; t < t_end; ++t, ct = ct + 1 == p.n_ct ? 0 : ct + 1, rt += ct == 0 ? 1 : 0) {
      for (int k = 0; k < p.kb; ++k) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = ring + stage * F_STAGE;
          mbar_arrive_expect_tx(&bars->full[stage], F_STAGE);
          tma_load_2d_hint(st, &map_x, &bars->full[stage], k * HYPRET_KBLK, rt * FT_M, TMA_EVICT_LAST);
          tma_load_2d_hint(st + FA_BLK, &map_y, &bars->full[stage], k * HYPRET_KBLK, ct * FT_N, TMA_EVICT_LAST);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (BWD) {
        // transposed planes of this column tile: the buffer is free once the previous tile's product has retired
        if (t > t_begin) { mbar_wait(&bars->w_free, wfree_par); wfree_par ^= 1; }
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->yt_full, 2 * yt_plane);
          for (int pl = 0; pl < 2; ++pl)
            for (int kk = 0; kk < FT_N / HYPRET_KBLK; ++kk)
              tma_load_2d_hint(yt_sm + pl * yt_plane + kk * p.d * 128, &map_yt, &bars->yt_full,
                               ct * FT_N + kk * HYPRET_KBLK, pl * p.d, TMA_EVICT_LAST);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc_g = umma_idesc_f16(FT_M, FT_N);
    const uint32_t idesc_p = umma_idesc_bf16(FT_M, p.d);
    constexpr uint64_t DESC_SW128 = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
                                    (UMMA_LAYOUT_SW128 << 61);
    const uint32_t ring_lo = smem_u32(ring) >> 4, w_lo = smem_u32(w_sm) >> 4, yt_lo = smem_u32(yt_sm) >> 4;
    uint32_t stage = 0, phase = 0, sempty_par[2] = {0, 0}, wfull_par = 0, ytfull_par = 0, accfree_par = 0;
    auto gram = [&](int t) {
      const uint32_t a = (uint32_t)((t - t_begin) & 1);
      mbar_wait(&bars->s_empty[a], sempty_par[a] ^ 1);
      sempty_par[a] ^= 1;
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + a * FT_N;
      for (int k = 0; k < p.kb; ++k) {
        mbar_wait(&bars->full[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ring_lo + stage * (F_STAGE >> 4), b_lo = a_lo + (FA_BLK >> 4);
#pragma unroll
          for (int kk = 0; kk < HYPRET_KBLK / 16; ++kk)
            umma_f16_ss(d_tmem, DESC_SW128 | (a_lo + 2 * kk), DESC_SW128 | (b_lo + 2 * kk), idesc_g,
                         (k | kk) != 0 ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (k == p.kb - 1) umma_commit(&bars->s_full[a]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    if (t_begin < t_end) gram(t_begin);
    int ct = ct_begin;
    for (int t = t_begin; t < t_end; ++t, ct = ct + 1 == p.n_ct ? 0 : ct + 1) {
      if (t + 1 < t_end) gram(t + 1);                    // the next Gram tile runs under this tile's epilogue
      if (BWD) {
        const bool seg_first = t == t_begin || ct == 0;
        const bool seg_last = t + 1 == t_end || ct + 1 == p.n_ct;
        mbar_wait(&bars->w_full, wfull_par); wfull_par ^= 1;
        mbar_wait(&bars->yt_full, ytfull_par); ytfull_par ^= 1;
        if (seg_first && t != t_begin) { mbar_wait(&bars->acc_free, accfree_par); accfree_par ^= 1; }
        tcgen05_fence_after();
        if (elect_one()) {
          // W Y = Wh Yh + Wh Ym + Wm Yh; operands: W plane [kblk][128][64] (A, K = column index), transposed Y plane
          // [kblk][D][64] (B): both K-major, 128B swizzle, 16 KB / D*128 B per 64-wide K block
          const uint32_t ypl = (uint32_t)(yt_plane >> 4), ykb = (uint32_t)(p.d * 128) >> 4;
          bool first = seg_first;
          for (int prod = 0; prod < 3; ++prod) {
            const uint32_t wa = w_lo + (prod == 2 ? (FW_PLANE >> 4) : 0);
            const uint32_t yb = yt_lo + (prod == 1 ? ypl : 0);
            for (int kk = 0; kk < FT_N / 16; ++kk) {
              const uint32_t blk = kk >> 2, in = (kk & 3) * 2;
              umma_f16_ss(tmem_acc, DESC_SW128 | (wa + blk * (FA_BLK >> 4) + in), DESC_SW128 | (yb + blk * ykb + in),
                           idesc_p, first ? 0u : 1u);
              first = false;
            }
          }
          umma_commit(&bars->w_free);
          if (seg_last) umma_commit(&bars->acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue
    // Instruction budget per matrix entry (ncu, first version: 39 warp instructions per entry in the forward pass at
    // 58 % issue utilisation -- the kernel is issue-bound, not tensor- or memory-bound): no 64-bit index arithmetic
    // and no bounds tests in the entry loops (full tiles take a path without them), near-pair detection as one
    // min-reduction per chunk, the logit never leaves log2 units.
    const int quad = warp & 3, row = quad * 32 + lane, et = threadIdx.x - 64;
    const int wg = (warp - 2) >> 2;                       // warpgroup: columns [F_CW wg, F_CW wg + F_CW) of every tile
    const float two_c = 2.0f * p.c;
    const float gs = BWD ? (p.grad_scale != nullptr ? *p.grad_scale : 1.0f) * p.coef * 2.0f * rsqrtf(p.c) : 0.f;
    uint32_t sfull_par[2] = {0, 0}, wfree_par = 0, accfull_par = 0;
    float run_m = INFINITY, run_s = 0.f;                  // FWD: running MIN of L = lg2(1+u) (max logit = -kappa min L), sum
    float pv[32];                                         // FWD: the previous chunk's L values, summed one chunk later
#pragma unroll
    for (int j = 0; j < 32; ++j) pv[j] = INFINITY;        // +inf contributes 2^-inf = 0
    float rowsum = 0.f;                                   // BWD: sum_j w_ij (1 + c s_ij / alpha_i)
    // per-column constants of a tile, double-buffered by tile parity and staged ONE TILE AHEAD (their global loads hide
    // behind the current tile's arithmetic):  [0] |y|^2   [1] 2c / beta (0 for columns >= m)   [2] BWD: column lse, log2
    // Global loads are issued one tile AHEAD and consumed one tile later (a dependent L2 round trip per tile cost ~1 us
    // of the ~3 us a tile takes): cols_load() puts the next tile's column norms / log-sum-exps into registers right
    // after the barrier, cols_store() turns them into the shared-memory constants at the END of the tile.
    float pre_nb = 0.f, pre_ll = 0.f;
    auto cols_load = [&](int ct_) {
      if (et < FT_N) {
        const int64_t j = (int64_t)ct_ * FT_N + et;
        const bool in = j < p.m;
        pre_nb = in ? p.ysq[j] : -1.0f;
        if (BWD) pre_ll = (p.y_lse != nullptr && in) ? p.y_lse[j] : 0.f;
      }
    };
    auto cols_store = [&](int tile) {
      if (et < FT_N) {
        float* cq_ = cols + ((tile - t_begin) & 1) * F_COLS * FT_N;
        const bool in = pre_nb >= 0.f;
        cq_[et] = in ? pre_nb : 0.f;
        cq_[FT_N + et] = in ? two_c / (1.0f - p.c * pre_nb) : 0.f;
        if (BWD) cq_[2 * FT_N + et] = pre_ll * 1.4426950408889634f;
      }
    };
    if (t_begin < t_end) { cols_load(ct_begin); cols_store(t_begin); }
    int cur_rt = -1;
    float na = 0.f, rho = 1.f, near_thr = -1.f, lx = 0.f;
    int rt = rt_begin, ct = ct_begin;
    for (int t = t_begin; t < t_end; ++t, ct = ct + 1 == p.n_ct ? 0 : ct + 1, rt += ct == 0 ? 1 : 0) {
      const uint32_t a = (uint32_t)((t - t_begin) & 1);
      const int64_t i = (int64_t)rt * FT_M + row;
      const int64_t j0 = (int64_t)ct * FT_N;
      const bool seg_last = t + 1 == t_end || ct + 1 == p.n_ct;
      const bool row_ok = i < p.n;
      const int n_cols = (int)(p.m - j0 < FT_N ? p.m - j0 : FT_N);     // valid columns of this tile
      const float* cq = cols + a * F_COLS * FT_N;
      // all threads are past tile t-1 here: buffer a^1 is free for tile t+1, buffer a is complete
      const long long c0 = p.stats ? clock64() : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI) : "memory");
      const long long c1 = p.stats ? clock64() : 0;
      if (t + 1 < t_end) cols_load(ct + 1 == p.n_ct ? 0 : ct + 1);
      if (rt != cur_rt) {                                               // per-row constants: once per row block
        cur_rt = rt;
        na = row_ok ? p.xsq[i] : 0.f;
        rho = 1.0f / (1.0f - p.c * na);
        near_thr = row_ok ? F_NEAR * na : -1.0f;                        // s below this: the Gram form has cancelled
        lx = (BWD && p.x_lse != nullptr && row_ok) ? p.x_lse[i] * 1.4426950408889634f : 0.f;
      }
      const int64_t jd64 = i + p.diag_offset - j0;                     // this row's target column inside the tile
      const int jd = (row_ok && jd64 >= 0 && jd64 < n_cols) ? (int)jd64 : -1;
      const bool tile_has_diag = __any_sync(0xffffffffu, jd >= 0);     // warp-uniform: 1 tile in n_ct
      const long long c2 = p.stats ? clock64() : 0;
      mbar_wait(&bars->s_full[a], sfull_par[a]);
      sfull_par[a] ^= 1;
      tcgen05_fence_after();
      const long long c3 = p.stats ? clock64() : 0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * FT_N + wg * F_CW;
      constexpr int HALVES = F_CW / 32;
      float vv[HALVES][32];
      __syncwarp();
#pragma unroll
      for (int h = 0; h < HALVES; ++h) tmem_ld_32x32(taddr + 32 * h, vv[h]);
#pragma unroll
      for (int h = 0; h < HALVES; ++h) tmem_ld_wait(vv[h]);
      tcgen05_fence_before();
      mbar_arrive(&bars->s_empty[a]);                     // the accumulator is in registers: the next Gram tile may land
      const long long c4 = p.stats ? clock64() : 0;
      if (p.stats && warp == 2 && lane == 0) {
        unsigned long long* st = p.stats + (size_t)blockIdx.x * 8;
        st[0] += (unsigned long long)(c1 - c0); st[1] += (unsigned long long)(c3 - c2); st[2] += (unsigned long long)(c4 - c3);
        st[3] += 1; if (st[4] == 0) st[4] = (unsigned long long)c0; st[5] = (unsigned long long)c4;
      }
#pragma unroll
      for (int half = 0; half < HALVES; ++half) {
        float (&v)[32] = vv[half];                        // Gram entries, overwritten in place
        const int cb = wg * F_CW + half * 32;             // first column of the chunk inside the tile
        const float4* cn = reinterpret_cast<const float4*>(cq + cb);
        const float4* ca = reinterpret_cast<const float4*>(cq + FT_N + cb);
        const float smin = chunk_sqdist(v, cn, na);
        if (smin < near_thr) {                            // rare: the diagonal of a contrastive batch, duplicates
#pragma unroll 1
          for (int j = 0; j < 32; ++j) {
            float sj = 0.f;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) sj = jj == j ? v[jj] : sj;
            if (sj < near_thr && cb + j < n_cols) {
              HYPRET_CHECK(i >= 0 && i < p.n && j0 + cb + j < p.m);
              const float ex = flash_exact_sq(p.x32 + i * p.d, p.y32 + (j0 + cb + j) * p.d, p.d);
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = jj == j ? ex : v[jj];     // keeps v[] in registers
            }
          }
        }
        if (!BWD) {
          // this chunk's logits + the previous chunk's exponential sum (against the running min that already covers it)
          float lmin = chunk_logits_sum(v, ca, rho, pv, run_m < INFINITY ? p.kappa * run_m : 0.f, p.kappa, run_s);
          if (cb + 32 > n_cols) {                         // warp-uniform, last column tile only: drop the padding columns
            lmin = INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (cb + j >= n_cols) v[j] = INFINITY;
              lmin = fminf(lmin, v[j]);
            }
          }
          // branch-free update of the running min: the sum is rescaled to the new reference (x 1 when it does not move)
          const float nm = fminf(run_m, lmin);
          run_s *= run_m < INFINITY ? ex2_approx(p.kappa * (nm - run_m)) : 0.f;
          run_m = nm;
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = v[j];
        } else {
          const float crho = p.c * rho;
          const float gx = row_ok ? -gs * p.wx : 0.f, gy = row_ok ? -gs * p.wy : 0.f, gd = gs * (p.wx + p.wy);
          const float4* cl = reinterpret_cast<const float4*>(cq + 2 * FT_N + cb);
          uint32_t hp[16], mp[16];
          if (tile_has_diag) {
            if (p.wy != 0.f) chunk_weights<true, true>(v, ca, cl, rho, crho, p.kappa, lx, gx, gy, gd, jd - cb, rowsum, hp, mp);
            else chunk_weights<true, false>(v, ca, cl, rho, crho, p.kappa, lx, gx, gy, gd, jd - cb, rowsum, hp, mp);
          } else {
            if (p.wy != 0.f) chunk_weights<false, true>(v, ca, cl, rho, crho, p.kappa, lx, gx, gy, gd, -1, rowsum, hp, mp);
            else chunk_weights<false, false>(v, ca, cl, rho, crho, p.kappa, lx, gx, gy, gd, -1, rowsum, hp, mp);
          }
          // the W buffer is rewritten now: the previous tile's product must have retired
          if (half == 0 && t > t_begin) { mbar_wait(&bars->w_free, wfree_par); wfree_par ^= 1; }
          // W planes, UMMA K-major 128B swizzle: [kblk = col / 64][row][64 cols]; 16-byte chunk index XOR (row & 7)
          const uint32_t base = smem_u32(w_sm) + (uint32_t)((cb >> 6) * FA_BLK) + (uint32_t)row * 128u;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint32_t off = ((((uint32_t)(((cb & 63) >> 3) + ch)) ^ (uint32_t)(row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(hp[4 * ch]), "r"(hp[4 * ch + 1]),
                         "r"(hp[4 * ch + 2]), "r"(hp[4 * ch + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + FW_PLANE + off), "r"(mp[4 * ch]),
                         "r"(mp[4 * ch + 1]), "r"(mp[4 * ch + 2]), "r"(mp[4 * ch + 3]) : "memory");
          }
        }
      }
      if (t + 1 < t_end) cols_store(t + 1);
      if (BWD) {
        fence_proxy_async_smem();                         // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(&bars->w_full);
      }
      if (seg_last) {
        const int slot = (int)blockIdx.x - flash_cta_of_tile((int64_t)rt * p.n_ct, T, P);
        HYPRET_CHECK(slot >= 0 && slot < p.n_slots && rt >= 0 && rt < p.n_rt);
        if (!BWD) {
          if (run_m < INFINITY) {                         // the last chunk of the segment is still pending
            const float km = p.kappa * run_m;
#pragma unroll
            for (int j = 0; j < 32; ++j) run_s += ex2_approx(fmaf(-p.kappa, pv[j], km));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = INFINITY;
          float* o = p.part + ((((int64_t)rt * p.n_slots + slot) * F_NWG + wg) * 2) * FT_M;
          o[row] = -p.kappa * run_m;                      // max logit (log2 units); -inf when no column was seen
          o[FT_M + row] = run_s;
          run_m = INFINITY;
          run_s = 0.f;
        } else {
          const int64_t rows_pad = (int64_t)p.n_rt * FT_M;
          p.part_rs[((int64_t)slot * F_NWG + wg) * rows_pad + (int64_t)rt * FT_M + row] = rowsum;
          rowsum = 0.f;
          mbar_wait(&bars->acc_full, accfull_par);
          accfull_par ^= 1;
          tcgen05_fence_after();
          // the [128, D] partial product: warpgroup g takes the 32-column chunks g, g+2, ...
          float* o = p.part + ((int64_t)slot * rows_pad + (int64_t)rt * FT_M + row) * p.d;
          const uint32_t tacc = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
          for (int cc = wg; cc * 32 < p.d; cc += F_NWG) {
            float v[32];
            __syncwarp();
            tmem_ld_32x32(tacc + cc * 32, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (cc * 32 + j < p.d)
                *reinterpret_cast<float4*>(o + cc * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          tcgen05_fence_before();
          mbar_arrive(&bars->acc_free);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// lse[i] = ln sum over the slots / warpgroups of 2^m s  (natural log), in a fixed order
__global__ void flash_lse_finish_kernel(const float* __restrict__ part, int64_t n, int n_rt, int n_ct, int n_slots, int P,
                                        float* __restrict__ lse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rt = (int)(i / FT_M), row = (int)(i - (int64_t)rt * FT_M);
  const int64_t T = (int64_t)n_rt * n_ct;
  const int c0 = flash_cta_of_tile((int64_t)rt * n_ct, T, P), c1 = flash_cta_of_tile((int64_t)(rt + 1) * n_ct - 1, T, P);
  float m = -INFINITY;
  for (int s = 0; s <= c1 - c0; ++s)
    for (int g = 0; g < F_NWG; ++g) m = fmaxf(m, part[((((int64_t)rt * n_slots + s) * F_NWG + g) * 2) * FT_M + row]);
  float acc = 0.f;
  for (int s = 0; s <= c1 - c0; ++s)
    for (int g = 0; g < F_NWG; ++g) {
      const float* o = part + ((((int64_t)rt * n_slots + s) * F_NWG + g) * 2) * FT_M;
      if (o[row] > -INFINITY) acc += o[FT_M + row] * exp2f(o[row] - m);
    }
  lse[i] = (m + log2f(acc)) * 0.6931471805599453f;
}

// dx[i,:] = x[i,:] * rowsum_i - sum over the slots of the partial products, in a fixed order
__global__ void flash_grad_finish_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                         const float* __restrict__ part_rs, int64_t n, int d, int n_rt, int n_ct,
                                         int P, float* __restrict__ dx) {
  const int64_t i = blockIdx.x;
  const int rt = (int)(i / FT_M);
  const int64_t T = (int64_t)n_rt * n_ct, rows_pad = (int64_t)n_rt * FT_M;
  const int c0 = flash_cta_of_tile((int64_t)rt * n_ct, T, P), c1 = flash_cta_of_tile((int64_t)(rt + 1) * n_ct - 1, T, P);
  float rs = 0.f;
  for (int s = 0; s <= c1 - c0; ++s)
    for (int g = 0; g < F_NWG; ++g) rs += part_rs[((int64_t)s * F_NWG + g) * rows_pad + i];
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s <= c1 - c0; ++s) acc += part[((int64_t)s * rows_pad + i) * d + k];
    dx[i * d + k] = x[i * d + k] * rs - acc;
  }
}

// out[j] = ln sum_w exp(parts[w, j]): combines per-rank partial log-sum-exps (sharded negatives), fixed order
__global__ void lse_combine_kernel(const float* __restrict__ parts, int w, int64_t n, float* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float m = -INFINITY;
  for (int r = 0; r < w; ++r) m = fmaxf(m, parts[(int64_t)r * n + j]);
  float acc = 0.f;
  for (int r = 0; r < w; ++r) acc += expf(parts[(int64_t)r * n + j] - m);
  out[j] = m > -INFINITY ? m + logf(acc) : -INFINITY;
}

// out[j] = sum_r parts[r, j], fixed order (the row / column sums of the W tiles of the non-flash backward)
__global__ void sum_parts_kernel(const float* __restrict__ parts, int w, int64_t n, float* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float acc = 0.f;
  for (int r = 0; r < w; ++r) acc += parts[(int64_t)r * n + j];
  out[j] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int flash_make_map(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int64_t rows, int64_t cols, int box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (enc == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return HYPRET_EDRIVER;
    enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)HYPRET_KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("HYPRET_FLASH_STATS") != nullptr)
    fprintf(stderr, "flash_make_map: CUresult %d base %p dt %d rows %lld cols %lld box_rows %d\n", (int)r, base, (int)dt,
            (long long)rows, (long long)cols, box_rows);
  return r == CUDA_SUCCESS ? HYPRET_OK : HYPRET_EINVAL;
}

int flash_grid(int64_t tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(tiles < sms ? tiles : sms);
}

int flash_slots(int n_rt, int n_ct, int P) {
  const int64_t T = (int64_t)n_rt * n_ct;
  int mx = 1;
  for (int rt = 0; rt < n_rt; ++rt) {
    const int s = flash_cta_of_tile((int64_t)(rt + 1) * n_ct - 1, T, P) - flash_cta_of_tile((int64_t)rt * n_ct, T, P) + 1;
    if (s > mx) mx = s;
  }
  return mx;
}

}  // namespace

int64_t hypret_flash_kpad_impl(int d) { return flash_kpad(d); }

int hypret_launch_lse_combine(const float* parts, int w, int64_t n, float* out, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  lse_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(parts, w, n, out);
  return (int)cudaGetLastError();
}

int hypret_launch_sum_parts(const float* parts, int w, int64_t n, float* out, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  sum_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(parts, w, n, out);
  return (int)cudaGetLastError();
}

int hypret_launch_flash_prep(const float* x, int64_t n, int d, void* row_op, void* col_op, void* t_planes, int64_t n_pad,
                             float* sq, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  if (row_op != nullptr || col_op != nullptr || sq != nullptr) {
    flash_prep_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(x, n, d, static_cast<__half*>(row_op),
                                                                  static_cast<__half*>(col_op), sq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (t_planes != nullptr) {
    const dim3 grid((unsigned)((n_pad + 31) / 32), (unsigned)((d + 31) / 32));
    flash_transpose_kernel<<<grid, 256, 0, stream>>>(x, n, d, n_pad, static_cast<__nv_bfloat16*>(t_planes));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return HYPRET_OK;
}

// workspace (floats): forward n_rt * slots * F_NWG * 2 * 128; backward slots * n_rt*128 * (d + F_NWG)
int64_t hypret_flash_workspace_floats(int64_t n, int64_t m, int d) {
  const int n_rt = (int)((n + FT_M - 1) / FT_M), n_ct = (int)((m + FT_N - 1) / FT_N);
  const int P = flash_grid((int64_t)n_rt * n_ct);
  const int64_t slots = flash_slots(n_rt, n_ct, P);
  const int64_t fwd = (int64_t)n_rt * slots * F_NWG * 2 * FT_M, bwd = slots * (int64_t)n_rt * FT_M * (d + F_NWG);
  return fwd > bwd ? fwd : bwd;
}

int hypret_launch_flash(int bwd, const void* x_row_op, const void* y_col_op, const void* y_t_planes, int64_t yt_cols,
                        const float* x32, const float* y32, const float* xsq, const float* ysq, const float* x_lse,
                        const float* y_lse, int64_t n, int64_t m, int d, float c, float inv_tau, float wx, float wy,
                        const float* grad_scale, int64_t diag_offset, int64_t n_total, float* workspace, float* out,
                        cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  const int kp = flash_kpad(d);
  CUtensorMap map_x, map_y, map_yt;
  int rc;
  if ((rc = flash_make_map(&map_x, x_row_op, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, n, kp, FT_M))) return rc;
  if ((rc = flash_make_map(&map_y, y_col_op, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, m, kp, FT_N))) return rc;
  map_yt = map_y;
  if (bwd && (rc = flash_make_map(&map_yt, y_t_planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2 * (int64_t)d, yt_cols, d)))
    return rc;
  FParams p;
  p.x32 = x32; p.y32 = y32; p.xsq = xsq; p.ysq = ysq; p.x_lse = x_lse; p.y_lse = y_lse;
  p.n = n; p.m = m; p.d = d; p.kb = kp / HYPRET_KBLK;
  p.c = c; p.kappa = inv_tau / sqrtf(c);
  p.wx = wx; p.wy = wy; p.grad_scale = grad_scale; p.coef = inv_tau / (float)n_total;
  p.diag_offset = diag_offset;
  p.n_rt = (int)((n + FT_M - 1) / FT_M); p.n_ct = (int)((m + FT_N - 1) / FT_N);
  const int P = flash_grid((int64_t)p.n_rt * p.n_ct);
  p.n_slots = flash_slots(p.n_rt, p.n_ct, P);
  p.yt_cols = yt_cols;
  p.stats = nullptr;
  const char* st_env = getenv("HYPRET_FLASH_STATS");      // instrumentation only: synchronises and prints to stderr
  if (st_env != nullptr && st_env[0] == '1') {
    if (cudaMalloc(&p.stats, (size_t)P * 64) == cudaSuccess) cudaMemsetAsync(p.stats, 0, (size_t)P * 64, stream);
    else p.stats = nullptr;
  }
  const int64_t rows_pad = (int64_t)p.n_rt * FT_M;
  p.part = workspace;
  p.part_rs = workspace + (int64_t)p.n_slots * rows_pad * d;
  if (bwd) {
    const int smem = 1024 + F_STAGES_BWD * F_STAGE + 2 * FW_PLANE + 2 * (FT_N / HYPRET_KBLK) * d * 128 +
                     2 * F_COLS * FT_N * 4 + 256;
    cudaError_t e = cudaFuncSetAttribute(flash_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    flash_tile_kernel<true><<<P, F_THREADS, smem, stream>>>(map_x, map_y, map_yt, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    flash_grad_finish_kernel<<<(unsigned)n, 128, 0, stream>>>(x32, p.part, p.part_rs, n, d, p.n_rt, p.n_ct, P, out);
  } else {
    const int smem = 1024 + F_STAGES_FWD * F_STAGE + 2 * F_COLS * FT_N * 4 + 256;
    cudaError_t e = cudaFuncSetAttribute(flash_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    flash_tile_kernel<false><<<P, F_THREADS, smem, stream>>>(map_x, map_y, map_yt, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    flash_lse_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p.part, n, p.n_rt, p.n_ct, p.n_slots, P, out);
  }
  if (p.stats != nullptr) {
    static unsigned long long h[148 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.stats, (size_t)P * 64, cudaMemcpyDeviceToHost);
    cudaFree(p.stats);
    double s[6] = {0};
    for (int c = 0; c < P; ++c) {
      for (int i = 0; i < 4; ++i) s[i] += (double)h[c * 8 + i] / P;
      s[4] += (double)(h[c * 8 + 5] - h[c * 8 + 4]) / P;
    }
    fprintf(stderr, "flash stats (%s): per CTA, epilogue warp 2: tiles %.1f, cycles: barrier %.0f, wait S %.0f, tmem ld %.0f of %.0f\n",
            bwd ? "bwd" : "fwd", s[3], s[0], s[1], s[2], s[4]);
  }
  return (int)cudaGetLastError();
}
