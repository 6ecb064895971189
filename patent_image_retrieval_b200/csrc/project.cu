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
// and up to three coalesced stores: the fp32 point (rerank operand), the bf16 GEMM operand
// row and ||y||^2.  HBM-bound: (4 + 4 + 2) * D bytes per row when all outputs are taken.
//
// bf16 operand row layout (Kpad = roundup(D,64) + 16 columns):
//   query   (side 0): [ x_0..x_{D-1} | 0.. | x1 x1 x2 x1 x2 x3  1  1  1  0 0 0 0 0 0 0 ]
//   gallery (side 1): [ -2*rb*y_0..  | 0.. | r1 r2 r1 r3 r2 r1  b1 b2 b3 0 0 0 0 0 0 0 ]
// with x1+x2+x3 = ||x||^2, r1+r2+r3 = rb = 1/(1-c||y||^2), b1+b2+b3 = rb*||y||^2 (3-way bf16
// splits), so that the tensor-core inner product of a query row and a gallery row is the
// ranking surrogate  rb_j * ||x_i - y_j||^2  (monotone in the Poincare distance for fixed i)
// with only the main-column products carrying bf16 rounding error.
// Cosine: query row = unit vector, gallery row = minus the unit vector, extension zero, so the
// inner product is  -cos(x, y)  (smaller = better, same as the hyperbolic surrogate).
#include "common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ void split3(float v, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(a);
  b = __float2bfloat16_rn(r);
  r -= __bfloat162float(b);
  c = __float2bfloat16_rn(r);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// Destinations of the bf16 operand row.  n == 1: the usual local buffer.  n > 1 (multi-GPU serving,
// hypret_project_rows_peers): the same row is stored into the exchange buffer of every rank of the box -- peer
// memory mapped over NVLink, plain posted stores -- so the projection IS the all-gather of the query operands.
struct OpDsts {
  __nv_bfloat16* p[HYPRET_MAX_PEERS];
  int n;
};

template <int NV>  // float4 chunks per lane; supports D <= NV * 128
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
project_rows_kernel(const float* __restrict__ u, int64_t n, int d, float c, int mode, int side,
                    float* __restrict__ y32, const OpDsts ops, float* __restrict__ sqnorm) {
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
      const float rb = hyp ? 1.0f / (1.0f - c * ysq) : 1.0f;
      const float mul = (side == HYPRET_SIDE_QUERY) ? 1.0f : (hyp ? -2.0f * rb : -1.0f);
      uint2 packed[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i)
        packed[i] = make_uint2(pack_bf16(v[i].x * mul, v[i].y * mul), pack_bf16(v[i].z * mul, v[i].w * mul));
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
        __align__(16) __nv_bfloat16 e[HYPRET_KEXT];
        const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int i = 0; i < HYPRET_KEXT; ++i) e[i] = zero;
        if (hyp) {
          if (side == HYPRET_SIDE_QUERY) {
            __nv_bfloat16 x1, x2, x3;
            split3(ysq, x1, x2, x3);
            const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
            e[0] = x1; e[1] = x1; e[2] = x2; e[3] = x1; e[4] = x2; e[5] = x3;
            e[6] = one; e[7] = one; e[8] = one;
          } else {
            __nv_bfloat16 r1, r2, r3, b1, b2, b3;
            split3(rb, r1, r2, r3);
            split3(rb * ysq, b1, b2, b3);
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
           cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t blocks_needed = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  // grid: a multiple of the SM count; 8 resident CTAs of 8 warps per SM saturates HBM
  int64_t grid = (int64_t)sms * 8;
  if (blocks_needed < grid) grid = blocks_needed;
  if (grid < 1) grid = 1;
  project_rows_kernel<NV><<<(unsigned)grid, WARPS_PER_BLOCK * 32, 0, stream>>>(
      u, n, d, c, mode, side, y32, ops, sqnorm);
  return (int)cudaGetLastError();
}

int dispatch(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, const OpDsts& ops,
             float* sqnorm, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  const int need = (d + 127) / 128;
  if (need <= 1) return launch<1>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  if (need <= 2) return launch<2>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  if (need <= 4) return launch<4>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  if (need <= 6) return launch<6>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  if (need <= 8) return launch<8>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  if (need <= 16) return launch<16>(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
  return HYPRET_EINVAL;
}

}  // namespace

int hypret_launch_project_rows(const float* u, int64_t n, int d, float c, int mode, int side, float* y32,
                               void* op_bf16, float* sqnorm, cudaStream_t stream) {
  OpDsts ops;
  ops.n = op_bf16 != nullptr ? 1 : 0;
  ops.p[0] = reinterpret_cast<__nv_bfloat16*>(op_bf16);
  return dispatch(u, n, d, c, mode, side, y32, ops, sqnorm, stream);
}

int hypret_launch_project_rows_peers(const float* u, int64_t n, int d, float c, int mode, float* y32,
                                     void* const* op_dsts_host, int n_dst, cudaStream_t stream) {
  if (n_dst < 1 || n_dst > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  OpDsts ops;
  ops.n = n_dst;
  for (int t = 0; t < n_dst; ++t) ops.p[t] = reinterpret_cast<__nv_bfloat16*>(op_dsts_host[t]);
  return dispatch(u, n, d, c, mode, HYPRET_SIDE_QUERY, y32, ops, nullptr, stream);
}
