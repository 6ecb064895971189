// Fused epilogue of a MobiusLinear layer (the learned projection head).
//
// The reference evaluates a layer as one GEMM followed by ~25-30 elementwise ATen launches, each
// a full pass over the [B, Dout] tensor (/root/reference/src/models.py:291-318), and the encoder
// adds ~20 more for the tanh-in-tangent-space activation and the projections (:481-505):
//   first layer   F.linear -> pmath.expmap0 -> pmath.mobius_add(bias) -> pmath.project       (309-317)
//   activation    pmath.mobius_fn_apply(tanh) = expmap0(tanh(logmap0(.)))                    (491)
//   final layer   pmath.mobius_matvec -> pmath.mobius_add(bias) -> pmath.project -> project  (307-317, 504)
// Here everything after the GEMM is ONE pass: a warp owns a row of mx = x W^T (registers), all
// norms / inner products are warp-shuffle reductions, and the row is written once together with
// ||y||^2 (which the next layer's mobius_matvec needs).  HBM-bound: 8*Dout bytes per row.
// Arithmetic order follows the oracle restatement of geoopt (oracle/pmath.py) in fp32.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int HD_WARPS = 8;

__device__ __forceinline__ float tanh_c(float x) { return tanhf(fminf(fmaxf(x, -15.f), 15.f)); }
__device__ __forceinline__ float artanh_c(float x) {
  x = fminf(fmaxf(x, -1.f + 1e-7f), 1.f - 1e-7f);
  return 0.5f * (logf(1.f + x) - logf(1.f - x));
}

template <int NV>
__device__ __forceinline__ float row_sumsq(const float4 (&v)[NV]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  return warp_sum(s);
}
template <int NV>
__device__ __forceinline__ void row_scale(float4 (&v)[NV], float a) {
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i].x *= a; v[i].y *= a; v[i].z *= a; v[i].w *= a; }
}

template <int NV>
__global__ void __launch_bounds__(HD_WARPS * 32)
mobius_epilogue_kernel(const float* __restrict__ mx, int64_t n, int d, const float* __restrict__ xsq_in,
                       const float* __restrict__ bias, float c, int hyperbolic_input, int post_tanh, int n_project,
                       float* __restrict__ y_out, float* __restrict__ sqnorm_out) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const int64_t warp0 = (int64_t)blockIdx.x * HD_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * HD_WARPS;
  const float sc = sqrtf(c), k = -c;
  const float maxnorm = (1.0f - 4e-3f) / sc;
  float4 b[NV];
  float b2 = 0.f;
  if (bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = i * 32 + lane;
      b[i] = (j < nvec) ? __ldg(reinterpret_cast<const float4*>(bias) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    b2 = row_sumsq<NV>(b);
  }
  for (int64_t row = warp0; row < n; row += nwarps) {
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = i * 32 + lane;
      v[i] = (j < nvec) ? __ldg(reinterpret_cast<const float4*>(mx + row * d) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- expmap0 (Euclidean input) or the mobius_matvec rescale (hyperbolic input) ----------------
    {
      const float ss = row_sumsq<NV>(v);
      const float mn = fmaxf(sqrtf(ss), 1e-15f);
      float t;
      if (hyperbolic_input) {
        const float xn = fmaxf(sqrtf(xsq_in[row]), 1e-15f);
        t = tanh_c(sc * (mn / xn * (artanh_c(sc * xn) / sc))) / sc;
        if (ss == 0.f) t = 0.f;                      // mx == 0 -> zero row (geoopt's cond)
      } else {
        t = tanh_c(sc * mn) / sc;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x = t * (v[i].x / mn); v[i].y = t * (v[i].y / mn); v[i].z = t * (v[i].z / mn); v[i].w = t * (v[i].w / mn);
      }
    }
    // ---- mobius_add(out, bias) ------------------------------------------------------------------------
    if (bias != nullptr) {
      const float x2 = row_sumsq<NV>(v);
      float xy = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) xy += v[i].x * b[i].x + v[i].y * b[i].y + v[i].z * b[i].z + v[i].w * b[i].w;
      xy = warp_sum(xy);
      const float ca = 1.f - 2.f * k * xy - k * b2, cb = 1.f + k * x2;
      const float den = fmaxf(1.f - 2.f * k * xy + k * k * x2 * b2, 1e-15f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x = (ca * v[i].x + cb * b[i].x) / den; v[i].y = (ca * v[i].y + cb * b[i].y) / den;
        v[i].z = (ca * v[i].z + cb * b[i].z) / den; v[i].w = (ca * v[i].w + cb * b[i].w) / den;
      }
    }
    // ---- project (once per layer; the encoder's trailing project makes it twice on the last layer) -----
    float ysq = row_sumsq<NV>(v);
    for (int pj = 0; pj < n_project; ++pj) {
      const float nrm = fmaxf(sqrtf(ysq), 1e-15f);
      if (nrm > maxnorm) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i].x = v[i].x / nrm * maxnorm; v[i].y = v[i].y / nrm * maxnorm;
          v[i].z = v[i].z / nrm * maxnorm; v[i].w = v[i].w / nrm * maxnorm;
        }
        ysq = row_sumsq<NV>(v);
      }
    }
    // ---- mobius_fn_apply(tanh): expmap0(tanh(logmap0(y))) --------------------------------------------------
    if (post_tanh) {
      const float yn = fmaxf(sqrtf(ysq), 1e-15f);
      const float l = artanh_c(sc * yn) / sc;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x = tanhf(v[i].x / yn * l); v[i].y = tanhf(v[i].y / yn * l);
        v[i].z = tanhf(v[i].z / yn * l); v[i].w = tanhf(v[i].w / yn * l);
      }
      const float tn = fmaxf(sqrtf(row_sumsq<NV>(v)), 1e-15f);
      const float t = tanh_c(sc * tn) / sc;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x = t * (v[i].x / tn); v[i].y = t * (v[i].y / tn); v[i].z = t * (v[i].z / tn); v[i].w = t * (v[i].w / tn);
      }
      ysq = row_sumsq<NV>(v);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = i * 32 + lane;
      if (j < nvec) reinterpret_cast<float4*>(y_out + row * d)[j] = v[i];
    }
    if (sqnorm_out != nullptr && lane == 0) sqnorm_out[row] = ysq;
  }
}

template <int NV>
int launch(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c, int hyp_in, int post_tanh,
           int n_project, float* y, float* sq, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (n + HD_WARPS - 1) / HD_WARPS;
  if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
  mobius_epilogue_kernel<NV><<<(unsigned)grid, HD_WARPS * 32, 0, stream>>>(mx, n, d, xsq, bias, c, hyp_in, post_tanh,
                                                                          n_project, y, sq);
  return (int)cudaGetLastError();
}

}  // namespace

int hypret_launch_mobius_epilogue(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                                  int hyperbolic_input, int post_tanh, int n_project, float* y, float* sqnorm,
                                  cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  const int need = (d + 127) / 128;
  if (need <= 1) return launch<1>(mx, n, d, xsq, bias, c, hyperbolic_input, post_tanh, n_project, y, sqnorm, stream);
  if (need <= 2) return launch<2>(mx, n, d, xsq, bias, c, hyperbolic_input, post_tanh, n_project, y, sqnorm, stream);
  if (need <= 4) return launch<4>(mx, n, d, xsq, bias, c, hyperbolic_input, post_tanh, n_project, y, sqnorm, stream);
  if (need <= 8) return launch<8>(mx, n, d, xsq, bias, c, hyperbolic_input, post_tanh, n_project, y, sqnorm, stream);
  if (need <= 16) return launch<16>(mx, n, d, xsq, bias, c, hyperbolic_input, post_tanh, n_project, y, sqnorm, stream);
  return HYPRET_EUNSUPPORTED;
}
