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
//   alpha = 1 - c|a_i|^2, beta = 1 - c|p_j|^2, z = cosh(sqrt(c) d), s = (z - 1) alpha beta / (2c),
//   w = g * 4 sqrt(c) / (alpha beta sinh(sqrt(c) d))
// the gradients are  dA_i = a_i * sum_j w (1 + c s / alpha) - (W P)_i  and
//                    dP_j = p_j * sum_i w (1 + c s / beta)  - (W^T A)_j   (SURVEY.md 7.4).
// Kernel A (one warp per row, coalesced along j) writes W and the row sums; kernel B (one
// thread per column, coalesced across the warp) re-reads W, D for the column sums.  The two
// dense products W P and W^T A are left to the caller (plain GEMMs).
__global__ void __launch_bounds__(256)
pairdist_bwd_rows_kernel(const float* __restrict__ g, const float* __restrict__ dmat, const float* __restrict__ asq,
                         const float* __restrict__ psq, int64_t n, int64_t m, float c, float* __restrict__ w_out,
                         float* __restrict__ row_sum) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const double cc = (double)c, sc = sqrt(cc);
  const double al = 1.0 - cc * (double)asq[i];
  double acc = 0.0;
  for (int64_t j = lane; j < m; j += 32) {
    const double be = 1.0 - cc * (double)psq[j];
    const double x = sc * (double)dmat[i * m + j];
    const double sh = fmax(sinh(x), 1e-15);
    const double hh = sinh(0.5 * x);                       // cosh(x) - 1 = 2 sinh^2(x/2), no cancellation
    const double s = hh * hh * al * be / cc;
    const double w = (double)g[i * m + j] * 4.0 * sc / (al * be * sh);
    w_out[i * m + j] = (float)w;
    acc += w * (1.0 + cc * s / al);
  }
  acc = warp_sum(acc);
  if (lane == 0) row_sum[i] = (float)acc;
}

__global__ void __launch_bounds__(256)
pairdist_bwd_cols_kernel(const float* __restrict__ w, const float* __restrict__ dmat, const float* __restrict__ asq,
                         const float* __restrict__ psq, int64_t n, int64_t m, float c, int64_t rows_per_block,
                         float* __restrict__ col_partial) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int64_t i0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t i1 = i0 + rows_per_block < n ? i0 + rows_per_block : n;
  const double cc = (double)c, sc = sqrt(cc);
  const double be = 1.0 - cc * (double)psq[j];
  double acc = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const double al = 1.0 - cc * (double)asq[i];
    const double x = sc * (double)dmat[i * m + j];
    const double hh = sinh(0.5 * x);
    const double s = hh * hh * al * be / cc;
    acc += (double)w[i * m + j] * (1.0 + cc * s / be);
  }
  col_partial[(int64_t)blockIdx.y * m + j] = (float)acc;
}

}  // namespace

int hypret_launch_pairdist_bwd(const float* g, const float* dmat, const float* asq, const float* psq, int64_t n,
                               int64_t m, float c, float* w_out, float* row_sum, float* col_partial, int n_partial,
                               cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  pairdist_bwd_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(g, dmat, asq, psq, n, m, c, w_out, row_sum);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int64_t rpb = (n + n_partial - 1) / n_partial;
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)n_partial);
  pairdist_bwd_cols_kernel<<<grid, 256, 0, stream>>>(w_out, dmat, asq, psq, n, m, c, rpb, col_partial);
  return (int)cudaGetLastError();
}

int hypret_launch_pairdist(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float* out,
                           cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  dim3 grid((unsigned)((m + PD_TILE - 1) / PD_TILE), (unsigned)((n + PD_TILE - 1) / PD_TILE));
  if (grid.y > 65535) return HYPRET_EUNSUPPORTED;
  pairdist_kernel<<<grid, 256, 0, stream>>>(a, p, n, m, d, c, out);
  return (int)cudaGetLastError();
}
