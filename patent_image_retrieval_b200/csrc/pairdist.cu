// Exact pairwise Poincare distances  D[i,j] = dist(a_i, p_j)  as a dense [n,m] fp32 matrix.
//
// Replaces the reference's Python loops of 1x1 / 1xN pmath.dist calls where the caller wants the
// matrix itself (moderate sizes):
//   n x n in-batch matrix           /root/reference/src/train.py:1832-1840, 2304-2320
//   one-vs-all rows for sklearn AP  /root/reference/src/train.py:3259
//   B x B broadcast                 /root/reference/src/train.py:1033
// Unlike the tensor-core filter (score_topk.cu) this kernel is exact: ||a-p||^2 is accumulated
// from explicitly formed differences in fp32 (no ||a||^2 + ||p||^2 - 2<a,p> cancellation, which
// matters for the near-duplicate diagonal of the contrastive loss), and the transcendental tail
//   d = log1p(t + sqrt(t (t + 2))) / sqrt(c),  t = 2c s / ((1 - c|a|^2)(1 - c|p|^2))
// (== 2 artanh(sqrt(c) |(-a) (+) p|) / sqrt(c), geoopt's dist) is evaluated in fp64.
// 64x64 output tile per CTA of 256 threads, 4x4 micro-tile per thread, K staged through
// shared memory in 16-wide slabs (transposed, padded: conflict-free).  FP32-FMA bound.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int PD_TILE = 64;
constexpr int PD_K = 16;

// FAST_TAIL: arccosh tail in fp32 (training path: ~3e-7 relative, next to the ~1e-6 of the fp32 accumulation
// itself); otherwise fp64 (evaluation path; hypret_rank_count reproduces that arithmetic bit for bit).
template <bool FAST_TAIL>
__global__ void __launch_bounds__(256)
pairdist_kernel(const float* __restrict__ a, const float* __restrict__ p, int64_t n, int64_t m, int d, float c,
                float* __restrict__ out) {
  __shared__ float As[PD_K][PD_TILE + 4];
  __shared__ float Ps[PD_K][PD_TILE + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * PD_TILE, j0 = (int64_t)blockIdx.x * PD_TILE;
  float acc[4][4];
  float na[4], np_[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    na[r] = 0.f;
    np_[r] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[r][s] = 0.f;
  }
  // loader mapping: 256 threads load a 64 x 16 slab of each operand (one float4 per thread)
  const int lr = threadIdx.x >> 2;          // row in tile 0..63
  const int lk = (threadIdx.x & 3) * 4;     // k offset 0,4,8,12
  for (int k0 = 0; k0 < d; k0 += PD_K) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vp = va;
    if (i0 + lr < n && k0 + lk < d) va = *reinterpret_cast<const float4*>(a + (i0 + lr) * d + k0 + lk);
    if (j0 + lr < m && k0 + lk < d) vp = *reinterpret_cast<const float4*>(p + (j0 + lr) * d + k0 + lk);
    __syncthreads();
    As[lk + 0][lr] = va.x; As[lk + 1][lr] = va.y; As[lk + 2][lr] = va.z; As[lk + 3][lr] = va.w;
    Ps[lk + 0][lr] = vp.x; Ps[lk + 1][lr] = vp.y; Ps[lk + 2][lr] = vp.z; Ps[lk + 3][lr] = vp.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PD_K; ++k) {
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
          const float e = av[r] - pv[s];
          acc[r][s] = fmaf(e, e, acc[r][s]);
        }
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) np_[s] = fmaf(pv[s], pv[s], np_[s]);
    }
  }
  if (FAST_TAIL) {
    const float rs = 1.0f / sqrtf(c);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t i = i0 + ty * 4 + r;
      if (i >= n) continue;
      const float al = 1.0f - c * na[r];
      float o[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float be = 1.0f - c * np_[s];
        const float t = 2.0f * c * acc[r][s] / (al * be);
        o[s] = log1pf(t + sqrtf(t * (t + 2.0f))) * rs;
      }
      const int64_t j = j0 + tx * 4;
      if (j + 3 < m && (m & 3) == 0) {
        *reinterpret_cast<float4*>(out + i * m + j) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (j + s < m) out[i * m + j + s] = o[s];
      }
    }
    return;
  }
  const double cc = (double)c, rs = 1.0 / sqrt(cc);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + ty * 4 + r;
    if (i >= n) continue;
    const double al = 1.0 - cc * (double)na[r];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int64_t j = j0 + tx * 4 + s;
      if (j >= m) continue;
      const double be = 1.0 - cc * (double)np_[s];
      const double t = 2.0 * cc * (double)acc[r][s] / (al * be);
      out[i * m + j] = (float)(log1p(t + sqrt(t * (t + 2.0))) * rs);
    }
  }
}

// Backward of the distance matrix (reference: autograd through ~40 elementwise ops per pair,
// /root/reference/src/train.py:1846).  With g = dL/dD and, for the pair (i,j),
//   alpha = 1 - c|a_i|^2, beta = 1 - c|p_j|^2, x = sqrt(c) d, h = sinh(x/2), s = h^2 alpha beta / c (= |a_i - p_j|^2),
//   sinh x = 2 h sqrt(1 + h^2),  w = g * 4 sqrt(c) / (alpha beta sinh x)
// the gradients are  dA_i = a_i * sum_j w (1 + c s / alpha) - (W P)_i  and
//                    dP_j = p_j * sum_i w (1 + c s / beta)  - (W^T A)_j   (SURVEY.md 7.4).
// ONE pass over the matrix (fp32 math: one sinhf and one sqrtf per element; HBM-bound): a CTA owns BW_ROWS
// rows and walks all columns in blocks of 256 -- a thread owns one column of the block (coalesced rows), keeps
// the BW_ROWS row accumulators in registers and the column accumulator of the block in one register.  Row sums
// are reduced in a fixed order at the end (deterministic); column sums leave as one partial row per CTA.
// CE: the upstream gradient is not read from memory but formed on the fly from the row / column log-sum-exps of
// sim = -D / tau (the in-batch InfoNCE losses of src/train.py:1832-1844 and 2304-2334):
//   g_ij = -(gs / tau / n) * [ wr (exp(sim_ij - lse_r[i]) - [i==j]) + wc (exp(sim_ij - lse_c[j]) - [i==j]) ].
// The two dense products W P and W^T A are left to the caller (plain GEMMs).
// MUFU approximations with flush-to-zero (1-2 ulp): the IEEE-rounded __frcp_rn and the denormal handling of __expf
// cost ~10 instructions and a branch each in a kernel that is bound by its instruction count
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rsq(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int BW_ROWS = 16;     // = HYPRET_BWD_ROWS: with 32 the row accumulators spilled under the 128-register cap
static_assert(BW_ROWS == HYPRET_BWD_ROWS, "include/hypret.h documents the row-block size");
constexpr int BW_BATCH = 16;   // rows loaded (all loads issued) before the first use

// SPLIT: W leaves as three bf16 planes [3][n][m] (hi, mid, lo with hi + mid + lo = W to fp32 accuracy) so that the
// two dense products of the backward run as bf16 tensor-core GEMMs (six cross products each) instead of SGEMMs.
template <bool CE, bool SPLIT>
__global__ void __launch_bounds__(256, 2)
pairdist_bwd_fused_kernel(const float* __restrict__ g, const float* __restrict__ dmat, const float* __restrict__ asq,
                          const float* __restrict__ psq, int64_t n, int64_t m, float c, const float* __restrict__ row_lse,
                          const float* __restrict__ col_lse, float inv_tau, float wr, float wc,
                          const float* __restrict__ grad_scale, void* __restrict__ w_out_raw,
                          float* __restrict__ row_partial, float* __restrict__ col_partial, int64_t diag_offset,
                          int64_t n_total) {
  // grid (row blocks of BW_ROWS, column chunks): CTA (x, y) owns the column blocks [y * cb_per, (y+1) * cb_per)
  // of its rows, writes row y of row_partial (its rows) and row x of col_partial (its columns) -- one writer per
  // element, no atomics; the chunks make ~7 waves of CTAs out of the 512 row blocks of an 8192-row batch
  __shared__ float s_al[BW_ROWS], s_ial[BW_ROWS], s_lse[BW_ROWS];
  __shared__ float s_red[BW_ROWS][8];
  const int64_t i0 = (int64_t)blockIdx.x * BW_ROWS;
  const int rows = (int)(n - i0 < BW_ROWS ? n - i0 : BW_ROWS);
  if (threadIdx.x < BW_ROWS) {
    const bool ok = threadIdx.x < rows;
    s_al[threadIdx.x] = ok ? 1.0f - c * asq[i0 + threadIdx.x] : 1.0f;
    s_ial[threadIdx.x] = 1.0f / s_al[threadIdx.x];
    s_lse[threadIdx.x] = (CE && ok) ? row_lse[i0 + threadIdx.x] : 0.f;
  }
  __syncthreads();
  const float sc = sqrtf(c), four_sc = 4.0f * sc;
  const float gsc = CE ? -(grad_scale != nullptr ? grad_scale[0] : 1.0f) * inv_tau / (float)n_total : 0.f;
  float racc[BW_ROWS];
#pragma unroll
  for (int r = 0; r < BW_ROWS; ++r) racc[r] = 0.f;
  // CTAs start at different column blocks (and wrap around): with a power-of-two row pitch the 32 rows of a
  // block land on the same memory channels, and CTAs marching through the columns in lock-step would all camp
  // on them at once
  const int64_t n_cb_all = (m + 255) / 256;
  const int64_t cb_per = (n_cb_all + gridDim.y - 1) / gridDim.y;
  const int64_t cb_lo = (int64_t)blockIdx.y * cb_per;
  const int64_t n_cb = cb_lo >= n_cb_all ? 0 : (n_cb_all - cb_lo < cb_per ? n_cb_all - cb_lo : cb_per);
  const int64_t cb0 = n_cb > 0 ? ((int64_t)blockIdx.x * 5) % n_cb : 0;
  for (int64_t cbi = 0; cbi < n_cb; ++cbi) {
    const int64_t cb = cb_lo + (cb0 + cbi < n_cb ? cb0 + cbi : cb0 + cbi - n_cb);
    const int64_t j = cb * 256 + threadIdx.x;
    const bool jok = j < m;
    const float be = jok ? 1.0f - c * psq[j] : 1.0f;
    const float ibe = 1.0f / be;
    const float lse_c = (CE && jok && wc != 0.f) ? col_lse[j] : INFINITY;    // +inf: exp(sim - inf) = 0, no branch
    float cacc = 0.f;
    // all 32 loads of the block are issued before the first use (the per-element version, with its branches, did
    // one dependent HBM round trip per element: 1.34 ms for 0.5 GB of traffic); out-of-range rows / columns read
    // a clamped address and are masked at the stores and sums
    // addresses: one per-thread pointer per stream, walked down the rows by the (CTA-uniform) pitch -- the
    // per-element 64-bit index arithmetic was 43 of the 156 instructions per element (ncu source page)
    const int64_t jc = jok ? j : m - 1;
    const float* dcur = dmat + i0 * m + jc;
    const float* gcur = CE ? nullptr : g + i0 * m + jc;
    char* wcur = static_cast<char*>(w_out_raw) + (i0 * m + j) * (SPLIT ? 2 : 4);
    const int64_t w_pitch = m * (SPLIT ? 2 : 4), plane_bytes = n * m * 2;
#pragma unroll
    for (int rb = 0; rb < BW_ROWS; rb += BW_BATCH) {
    float dv[BW_BATCH], gv[BW_BATCH];
#pragma unroll
    for (int q = 0; q < BW_BATCH; ++q) {
      const int r = rb + q;
      dv[q] = *dcur;
      gv[q] = CE ? 0.f : *gcur;
      const int64_t step = (r + 1 < rows) ? m : 0;      // rows past the end re-read the last valid row (masked below)
      dcur += step;
      if (!CE) gcur += step;
    }
#pragma unroll
    for (int q = 0; q < BW_BATCH; ++q) {
      const int r = rb + q;
      const bool ok = r < rows && jok;
      const float dd = dv[q];
      float gg;
      if (CE) {
        const float sim = -dd * inv_tau;
        const float diag = (i0 + r + diag_offset == j) ? 1.0f : 0.0f;   // target column of row i
        gg = gsc * (wr * (fast_exp(sim - s_lse[r]) - diag) + wc * (fast_exp(sim - lse_c) - diag));
      } else {
        gg = gv[q];
      }
      // t = cosh(x) - 1 = 2 sinh^2(x/2), x = sqrt(c) d, without library calls or IEEE-division slow paths:
      // (e^x + e^-x)/2 - 1 from one fast exp2 and one approximate reciprocal; below x = 0.35 that form
      // cancels, so the even series x^2/2 (1 + x^2/12 (1 + x^2/30 (1 + x^2/56))) takes over (its first
      // neglected term is < 2e-9 relative there).  sinh x = sqrt(t (t + 2)).
      const float xx = sc * dd, x2 = xx * xx;
      const float e = fast_exp(xx);
      const float t_big = 0.5f * (e + fast_rcp(e)) - 1.0f;
      const float t_small = 0.5f * x2 * (1.0f + x2 * (1.0f / 12.0f) * (1.0f + x2 * (1.0f / 30.0f) * (1.0f + x2 * (1.0f / 56.0f))));
      const float t = xx < 0.35f ? t_small : t_big;
      const float t2 = fmaxf(t * (t + 2.0f), 1e-30f);
      const float sh = t2 * fast_rsq(t2);       // sinh x
      const float ab = s_al[r] * be;
      const float w = ok ? gg * four_sc * fast_rcp(ab * sh) : 0.f;
      if (ok) {
        if (SPLIT) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(w);
          const float r1 = w - __bfloat162float(hi);
          const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
          *reinterpret_cast<__nv_bfloat16*>(wcur) = hi;
          *reinterpret_cast<__nv_bfloat16*>(wcur + plane_bytes) = mid;
          *reinterpret_cast<__nv_bfloat16*>(wcur + 2 * plane_bytes) = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        } else {
          *reinterpret_cast<float*>(wcur) = w;
        }
      }
      wcur += w_pitch;
      const float cs = 0.5f * t * ab;           // c * s = h^2 alpha beta, h^2 = t / 2
      racc[r] += w * (1.0f + cs * s_ial[r]);
      cacc += w * (1.0f + cs * ibe);
    }
    }
    if (jok) col_partial[(int64_t)blockIdx.x * m + j] = cacc;
  }
  // row sums: warp tree, then the 8 warps of the CTA in a fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < BW_ROWS; ++r) {
    const float v = warp_sum(racc[r]);
    if (lane == 0) s_red[r][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < rows) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += s_red[threadIdx.x][w8];
    row_partial[(int64_t)blockIdx.y * n + i0 + threadIdx.x] = v;
  }
}

// Log-sum-exp of sim = -D / tau along rows (one warp per row) and along columns (thread per column over a block of
// rows -> partial (max, sum) pairs, combined by a second tiny kernel).  Online max/sum in fp32.
__global__ void __launch_bounds__(256)
lse_rows_kernel(const float* __restrict__ dmat, int64_t n, int64_t m, float inv_tau, float* __restrict__ row_lse) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  float mx = -INFINITY, sm = 0.f;
  for (int64_t j = lane; j < m; j += 32) {
    const float z = -dmat[i * m + j] * inv_tau;
    if (z > mx) { sm = sm * expf(mx - z) + 1.0f; mx = z; }
    else sm += expf(z - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), osm = __shfl_xor_sync(0xffffffffu, sm, o);
    const float nm = fmaxf(mx, omx);
    sm = (mx == -INFINITY ? 0.f : sm * expf(mx - nm)) + (omx == -INFINITY ? 0.f : osm * expf(omx - nm));
    mx = nm;
  }
  if (lane == 0) row_lse[i] = mx + logf(sm);
}

__global__ void __launch_bounds__(256)
lse_cols_partial_kernel(const float* __restrict__ dmat, int64_t n, int64_t m, float inv_tau, int64_t rows_per_block,
                        float* __restrict__ part_max, float* __restrict__ part_sum) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= m) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  float mx = -INFINITY, sm = 0.f;
  for (int64_t i = r0; i < r1; ++i) {
    const float z = -dmat[i * m + j] * inv_tau;
    if (z > mx) { sm = sm * expf(mx - z) + 1.0f; mx = z; }
    else sm += expf(z - mx);
  }
  part_max[(int64_t)blockIdx.y * m + j] = mx;
  part_sum[(int64_t)blockIdx.y * m + j] = sm;
}

__global__ void __launch_bounds__(256)
lse_cols_combine_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, int64_t m, int n_part,
                        float* __restrict__ col_lse) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= m) return;
  float mx = -INFINITY;
  for (int b = 0; b < n_part; ++b) mx = fmaxf(mx, part_max[(int64_t)b * m + j]);
  float sm = 0.f;
  for (int b = 0; b < n_part; ++b) {
    const float pm = part_max[(int64_t)b * m + j];
    if (pm > -INFINITY) sm += part_sum[(int64_t)b * m + j] * expf(pm - mx);
  }
  col_lse[j] = mx + logf(sm);
}

// fp32 [count] -> three bf16 planes [3][count] (hi, mid, lo; hi + mid + lo = x to fp32 accuracy), 16-byte loads and
// 8-byte stores.  The backward pass emits W fastest as plain fp32 (one 128-byte line per warp store; 0.23 ms at
// 8192^2) -- writing the planes from inside it costs three 64-byte stores per warp and element and the registers
// spill (0.44 ms) -- so the planes are cut in this streaming pass instead (0.67 GB of traffic).
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, int64_t count, __nv_bfloat16* __restrict__ out) {
  const int64_t n4 = count >> 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(x) + v);
    const float in[4] = {a.x, a.y, a.z, a.w};
    __nv_bfloat16 h[4], mid[4], l[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      h[t] = __float2bfloat16_rn(in[t]);
      const float r1 = in[t] - __bfloat162float(h[t]);
      mid[t] = __float2bfloat16_rn(r1);
      l[t] = __float2bfloat16_rn(r1 - __bfloat162float(mid[t]));
    }
    *reinterpret_cast<uint2*>(out + 4 * v) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(out + count + 4 * v) = *reinterpret_cast<const uint2*>(mid);
    *reinterpret_cast<uint2*>(out + 2 * count + 4 * v) = *reinterpret_cast<const uint2*>(l);
  }
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {          // tail
    const int64_t i = (n4 << 2) + threadIdx.x;
    const __nv_bfloat16 h = __float2bfloat16_rn(x[i]);
    const float r1 = x[i] - __bfloat162float(h);
    const __nv_bfloat16 md = __float2bfloat16_rn(r1);
    out[i] = h;
    out[count + i] = md;
    out[2 * count + i] = __float2bfloat16_rn(r1 - __bfloat162float(md));
  }
}

}  // namespace

int hypret_launch_split3(const float* x, int64_t count, void* out_bf16, cudaStream_t stream) {
  if (count == 0) return HYPRET_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (count / 4 + 255) / 256;
  const unsigned grid = (unsigned)(want < (int64_t)sms * 8 ? (want < 1 ? 1 : want) : (int64_t)sms * 8);
  split3_kernel<<<grid, 256, 0, stream>>>(x, count, static_cast<__nv_bfloat16*>(out_bf16));
  return (int)cudaGetLastError();
}

int hypret_launch_pairdist_bwd(const float* g, const float* dmat, const float* asq, const float* psq, int64_t n,
                               int64_t m, float c, void* w_out, int w_format, float* row_partial, int n_row_partial,
                               float* col_partial, int n_partial, cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  if ((int64_t)n_partial * BW_ROWS < n) return HYPRET_EINVAL;     // one partial row of column sums per CTA of BW_ROWS rows
  const dim3 grid((unsigned)((n + BW_ROWS - 1) / BW_ROWS), (unsigned)n_row_partial);
  if (w_format == 1)
    pairdist_bwd_fused_kernel<false, true><<<grid, 256, 0, stream>>>(g, dmat, asq, psq, n, m, c, nullptr, nullptr, 0.f,
                                                                    0.f, 0.f, nullptr, w_out, row_partial, col_partial,
                                                                    0, n);
  else
    pairdist_bwd_fused_kernel<false, false><<<grid, 256, 0, stream>>>(g, dmat, asq, psq, n, m, c, nullptr, nullptr, 0.f,
                                                                     0.f, 0.f, nullptr, w_out, row_partial,
                                                                     col_partial, 0, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if ((int64_t)n_partial > (int64_t)grid.x)    // partial rows no CTA writes must read as zero
    e = cudaMemsetAsync(col_partial + (int64_t)grid.x * m, 0, ((int64_t)n_partial - grid.x) * m * sizeof(float),
                        stream);
  return (int)e;
}

int hypret_launch_pairdist_ce_fwd(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float inv_tau,
                                  int want_cols, float* dmat, float* row_lse, float* col_lse, float* scratch,
                                  int n_part, cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  dim3 grid((unsigned)((m + PD_TILE - 1) / PD_TILE), (unsigned)((n + PD_TILE - 1) / PD_TILE));
  if (grid.y > 65535) return HYPRET_EUNSUPPORTED;
  pairdist_kernel<true><<<grid, 256, 0, stream>>>(a, p, n, m, d, c, dmat);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return hypret_launch_neg_lse(dmat, n, m, inv_tau, want_cols, row_lse, col_lse, scratch, n_part, stream);
}

int hypret_launch_neg_lse(const float* dmat, int64_t n, int64_t m, float inv_tau, int want_cols, float* row_lse,
                          float* col_lse, float* scratch, int n_part, cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  cudaError_t e;
  lse_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(dmat, n, m, inv_tau, row_lse);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (want_cols) {
    const int64_t rpb = (n + n_part - 1) / n_part;
    dim3 g2((unsigned)((m + 255) / 256), (unsigned)n_part);
    lse_cols_partial_kernel<<<g2, 256, 0, stream>>>(dmat, n, m, inv_tau, rpb, scratch, scratch + (int64_t)n_part * m);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    lse_cols_combine_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(scratch, scratch + (int64_t)n_part * m, m,
                                                                            n_part, col_lse);
  }
  return (int)cudaGetLastError();
}

int hypret_launch_pairdist_ce_bwd(const float* dmat, const float* asq, const float* psq, int64_t n, int64_t m, float c,
                                  const float* row_lse, const float* col_lse, float inv_tau, float wr, float wc,
                                  const float* grad_scale, void* w_out, int w_format, float* row_partial,
                                  int n_row_partial, float* col_partial, int64_t diag_offset, int64_t n_total,
                                  cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  const dim3 grid((unsigned)((n + BW_ROWS - 1) / BW_ROWS), (unsigned)n_row_partial);
  if (w_format == 1)
    pairdist_bwd_fused_kernel<true, true><<<grid, 256, 0, stream>>>(nullptr, dmat, asq, psq, n, m, c, row_lse, col_lse,
                                                                   inv_tau, wr, wc, grad_scale, w_out, row_partial,
                                                                   col_partial, diag_offset, n_total);
  else
    pairdist_bwd_fused_kernel<true, false><<<grid, 256, 0, stream>>>(nullptr, dmat, asq, psq, n, m, c, row_lse, col_lse,
                                                                    inv_tau, wr, wc, grad_scale, w_out, row_partial,
                                                                    col_partial, diag_offset, n_total);
  return (int)cudaGetLastError();
}

int hypret_launch_pairdist(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float* out,
                           cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  dim3 grid((unsigned)((m + PD_TILE - 1) / PD_TILE), (unsigned)((n + PD_TILE - 1) / PD_TILE));
  if (grid.y > 65535) return HYPRET_EUNSUPPORTED;
  pairdist_kernel<false><<<grid, 256, 0, stream>>>(a, p, n, m, d, c, out);
  return (int)cudaGetLastError();
}
