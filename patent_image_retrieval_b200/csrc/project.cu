// Fused row projection: Euclidean feature row -> Poincare ball point + GEMM operand.
//
// Replaces, for the no-weight / identity head used by the retrieval benchmark, the chain
//   pmath.expmap0 -> pmath.project            (/root/reference/src/models.py:310,317)
// and, for rows that are already on the ball (outputs of encode_figures, models.py:537-548),
// just the final pmath.project (models.py:504).  For the cosine path it replaces the row
// normalisation inside sklearn's cosine_similarity (notebooks/retrieval.ipynb:368).
//
// One warp owns one row: 128-bit coalesced loads, the whole row stays in registers, two
// warp-shuffle reductions (||u||^2, then ||y||^2 of the rounded result, as torch does),
// and up to three coalesced stores: the fp32 point (rerank operand), the fp16 GEMM operand
// row and ||y||^2.  HBM-bound: (4 + 4 + 2) * D bytes per row when all outputs are taken.
//
// fp16 operand row layout (Kpad = roundup(D,64) + 16 columns), in UNIT-BALL coordinates x^ = sqrt(c) x:
//   query   (side 0): [ x^_0..x^_{D-1} | 0.. | x1 x1 x2 x1 x2 x3  1  1  1  0 0 0 0 0 0 0 ]
//   gallery (side 1): [ -2*rb*y^_0..   | 0.. | r1 r2 r1 r3 r2 r1  b1 b2 b3 0 0 0 0 0 0 0 ]
// with x1+x2+x3 = ||x^||^2, r1+r2+r3 = rb = 1/(1-||y^||^2), b1+b2+b3 = rb*||y^||^2 (3-way fp16
// splits), so that the tensor-core inner product of a query row and a gallery row is the
// ranking surrogate  rb_j * ||x^_i - y^_j||^2 = c * rb_j * ||x_i - y_j||^2  (monotone in the Poincare distance for
// fixed i) with only the main-column products carrying rounding error.  fp16, not bf16: both run at the same
// tensor rate (kind::f16), but fp16 keeps 11 significant bits against 8, which makes the rounding bound of the
// exact-top-k certificate 8x tighter (with bf16 only 73 % of C2's queries could be certified at k' = 16, with fp16
// > 99.9 %).  The unit-ball scaling bounds every entry whatever c is (|x^_i| < 1, |2 rb y^_i| < 250, rb ||y^||^2 < 125
// with geoopt's eps = 4e-3 clip), far inside fp16's range.
// Cosine: query row = unit vector, gallery row = minus the unit vector, extension zero, so the
// inner product is  -cos(x, y)  (smaller = better, same as the hyperbolic surrogate).
//
// Certificate inputs (hypret_project_rows_cert; used by the exact-top-k guarantee of hypret_rerank_cert): with m the
// fp32 main-column row that is rounded to fp16 (x^ for a query, z = -2 rb y^ for a gallery row), the kernel can also emit
//   op_err[i]  = || fp16(m_i) - m_i ||_2          the rounding residual of the row, and
//   stats[0..3] (atomicMax over the rows: gallery side) = max ||fp16(m)||, max op_err, max rb, max rb ||y^||^2
// so that |<fp16 x, fp16 z> - <x, z>| = |<dx, z~> + <x, dz>| <= op_err_i * stats[0] + ||x_i|| * stats[1] bounds the error of
// the tensor-core surrogate of EVERY gallery row for query i (Cauchy-Schwarz; no distributional assumption).
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ void split3(float v, __half& a, __half& b, __half& c) {
  a = __float2half_rn(v);
  float r = v - __half2float(a);
  b = __float2half_rn(r);
  r -= __half2float(b);
  c = __float2half_rn(r);
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 p = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// Destinations of the fp16 operand row.  n == 1: the usual local buffer.  n > 1 (multi-GPU serving,
// hypret_project_rows_peers): the same row is stored into the exchange buffer of every rank of the box -- peer
// memory mapped over NVLink, plain posted stores -- so the projection IS the all-gather of the query operands.
struct OpDsts {
  __half* p[HYPRET_MAX_PEERS];
  int n;
};

template <int NV>  // float4 chunks per lane; supports D <= NV * 128
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
project_rows_kernel(const float* __restrict__ u, int64_t n, int d, float c, int mode, int side,
                    float* __restrict__ y32, const OpDsts ops, float* __restrict__ sqnorm,
                    float* __restrict__ op_err, unsigned* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const int dpad = hypret_dpad(d);
  const int kpad = dpad + HYPRET_KEXT;
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  const float sc = sqrtf(c);
  const float maxnorm = (1.0f - 4e-3f) / sc;   // geoopt project(): eps = 4e-3 for float32

  for (int64_t row = warp0; row < n; row += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(u + row * d);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = i * 32 + lane;
      v[i] = (j < nvec) ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    ss = warp_sum(ss);

    float ysq;
    if (mode == HYPRET_MODE_COSINE) {
      // sklearn normalize(): zero rows are left as they are
      const float nrm = sqrtf(ss);
      const float dv = (nrm == 0.f) ? 1.f : nrm;
#pragma unroll
      for (int i = 0; i < NV; ++i) { v[i].x /= dv; v[i].y /= dv; v[i].z /= dv; v[i].w /= dv; }
      ysq = (nrm == 0.f) ? 0.f : 1.f;
    } else {
      if (mode == HYPRET_MODE_EXPMAP0) {
        // expmap0: tan_k(||u||) * (u / ||u||), tanh argument clamped to +-15 (geoopt.utils.tanh)
        const float un = fmaxf(sqrtf(ss), 1e-15f);
        const float t = tanhf(fminf(fmaxf(un * sc, -15.f), 15.f)) / sc;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i].x = t * (v[i].x / un); v[i].y = t * (v[i].y / un);
          v[i].z = t * (v[i].z / un); v[i].w = t * (v[i].w / un);
        }
        ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        ss = warp_sum(ss);
      }
      // project: rows whose norm exceeds (1-eps)/sqrt(c) are rescaled onto that sphere
      const float nrm = fmaxf(sqrtf(ss), 1e-15f);
      if (nrm > maxnorm) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i].x = v[i].x / nrm * maxnorm; v[i].y = v[i].y / nrm * maxnorm;
          v[i].z = v[i].z / nrm * maxnorm; v[i].w = v[i].w / nrm * maxnorm;
        }
        ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        ss = warp_sum(ss);
      }
      ysq = ss;
    }

    if (y32 != nullptr) {
      float4* dst = reinterpret_cast<float4*>(y32 + row * d);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int j = i * 32 + lane;
        if (j < nvec) dst[j] = v[i];
      }
    }
    if (sqnorm != nullptr && lane == 0) sqnorm[row] = ysq;

    if (ops.n > 0) {
      const bool hyp = (mode != HYPRET_MODE_COSINE);
      const float ysq_u = hyp ? c * ysq : ysq;                 // squared norm in unit-ball coordinates
      const float rb = hyp ? 1.0f / (1.0f - ysq_u) : 1.0f;
      const float mul = (side == HYPRET_SIDE_QUERY) ? (hyp ? sc : 1.0f) : (hyp ? -2.0f * rb * sc : -1.0f);
      uint2 packed[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i)
        packed[i] = make_uint2(pack_f16(v[i].x * mul, v[i].y * mul), pack_f16(v[i].z * mul, v[i].w * mul));
      if (op_err != nullptr || stats != nullptr) {
        // rounding residual and norm of the row as the tensor core sees it (lanes past the row hold zeros)
        float e2 = 0.f, r2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float m4[4] = {v[i].x * mul, v[i].y * mul, v[i].z * mul, v[i].w * mul};
          const uint32_t w2[2] = {packed[i].x, packed[i].y};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float r = __half2float(__ushort_as_half((unsigned short)((t & 1) ? (w2[t >> 1] >> 16) : (w2[t >> 1] & 0xffffu))));
            const float e = r - m4[t];
            e2 = fmaf(e, e, e2);
            r2 = fmaf(r, r, r2);
          }
        }
        e2 = warp_sum(e2);
        r2 = warp_sum(r2);
        // round the norms UP (they are upper bounds): one ulp covers the fp32 summation error of <= 2048 terms
        // of one sign only approximately, so scale by (1 + 2^-10)
        const float en = sqrtf(e2) * 1.001f, rn = sqrtf(r2) * 1.001f;
        if (lane == 0) {
          if (op_err != nullptr) op_err[row] = en;
          if (stats != nullptr) {      // non-negative floats order like their bit patterns
            atomicMax(stats + 0, __float_as_uint(rn));
            atomicMax(stats + 1, __float_as_uint(en));
            atomicMax(stats + 2, __float_as_uint(hyp ? rb : 0.f));
            atomicMax(stats + 3, __float_as_uint(hyp ? rb * ysq_u : 0.f));
          }
        }
      }
      for (int t = 0; t < ops.n; ++t) {
        uint2* dst = reinterpret_cast<uint2*>(ops.p[t] + row * kpad);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int j = i * 32 + lane;
          if (j < nvec) dst[j] = packed[i];
        }
        // zero the K padding between D and Dpad
        for (int j = nvec + lane; j < (dpad >> 2); j += 32) dst[j] = make_uint2(0u, 0u);
      }
      if (lane == 0) {
        __align__(16) __half e[HYPRET_KEXT];
        const __half zero = __float2half_rn(0.f);
#pragma unroll
        for (int i = 0; i < HYPRET_KEXT; ++i) e[i] = zero;
        if (hyp) {
          if (side == HYPRET_SIDE_QUERY) {
            __half x1, x2, x3;
            split3(ysq_u, x1, x2, x3);
            const __half one = __float2half_rn(1.f);
            e[0] = x1; e[1] = x1; e[2] = x2; e[3] = x1; e[4] = x2; e[5] = x3;
            e[6] = one; e[7] = one; e[8] = one;
          } else {
            __half r1, r2, r3, b1, b2, b3;
            split3(rb, r1, r2, r3);
            split3(rb * ysq_u, b1, b2, b3);
            e[0] = r1; e[1] = r2; e[2] = r1; e[3] = r3; e[4] = r2; e[5] = r1;
            e[6] = b1; e[7] = b2; e[8] = b3;
          }
        }
        const uint4* es = reinterpret_cast<const uint4*>(e);
        for (int t = 0; t < ops.n; ++t) {
          uint4* ext = reinterpret_cast<uint4*>(ops.p[t] + row * kpad + dpad);
          ext[0] = es[0];
          ext[1] = es[1];
        }
      }
    }
  }
}

template <int NV>
int launch(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, const OpDsts& ops, float* sqnorm,
           float* op_err, unsigned* stats, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t blocks_needed = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  // grid: a multiple of the SM count; 8 resident CTAs of 8 warps per SM saturates HBM
  int64_t grid = (int64_t)sms * 8;
  if (blocks_needed < grid) grid = blocks_needed;
  if (grid < 1) grid = 1;
  project_rows_kernel<NV><<<(unsigned)grid, WARPS_PER_BLOCK * 32, 0, stream>>>(
      u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats);
  return (int)cudaGetLastError();
}

int dispatch(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, const OpDsts& ops,
             float* sqnorm, float* op_err, unsigned* stats, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  const int need = (d + 127) / 128;
  if (need <= 1) return launch<1>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  if (need <= 2) return launch<2>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  if (need <= 4) return launch<4>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  if (need <= 6) return launch<6>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  if (need <= 8) return launch<8>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  if (need <= 16) return launch<16>(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, stats, stream);
  return HYPRET_EINVAL;
}

}  // namespace

int hypret_launch_project_rows(const float* u, int64_t n, int d, float c, int mode, int side, float* y32,
                               void* op_f16, float* sqnorm, float* op_err, float* stats, cudaStream_t stream) {
  OpDsts ops;
  ops.n = op_f16 != nullptr ? 1 : 0;
  ops.p[0] = reinterpret_cast<__half*>(op_f16);
  return dispatch(u, n, d, c, mode, side, y32, ops, sqnorm, op_err, reinterpret_cast<unsigned*>(stats), stream);
}

int hypret_launch_project_rows_peers(const float* u, int64_t n, int d, float c, int mode, float* y32,
                                     void* const* op_dsts_host, int n_dst, float* op_err, cudaStream_t stream) {
  if (n_dst < 1 || n_dst > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  OpDsts ops;
  ops.n = n_dst;
  for (int t = 0; t < n_dst; ++t) ops.p[t] = reinterpret_cast<__half*>(op_dsts_host[t]);
  return dispatch(u, n, d, c, mode, HYPRET_SIDE_QUERY, y32, ops, nullptr, op_err, nullptr, stream);
}
