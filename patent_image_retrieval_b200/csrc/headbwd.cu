// Backward of the MobiusLinear epilogue + the small dense products of the head's backward pass.
//
// The reference differentiates the projection head (/root/reference/src/models.py:291-318, 481-505) through ~75
// elementwise autograd nodes per batch.  Here the whole epilogue of a layer -- expmap0 | mobius_matvec rescale,
// mobius_add(bias), project (x n), tanh in the tangent space -- is differentiated in closed form by ONE kernel, a warp
// per row: the row's forward scalars are recomputed from the saved raw product mx, then the chain is walked backwards.
// Every stage is either RADIAL, out = phi(|in|, .) in  =>  g_in = phi g + dphi/d|in| <g, in> in/|in|, the Moebius
// addition (a rational function of u, b, <u,b>, |u|^2, |b|^2) or an elementwise tanh.
//   mobius_epilogue_bwd   gy [n,N] -> gmx [n,N], gbias [N] (+=, atomics), gxn [n] = dL/d|x_in| / |x_in| (hyperbolic input)
//   sgemm_strided         C[M,N] = sum_k A(m,k) B(k,n) with arbitrary strides: gW = gmx^T X, gX = gmx W -- batch-sized
//                         products (B = 128 in the reference's training loop): FP32 FMA tiles, not worth a tensor map
#include <math.h>

#include "common.cuh"

namespace {

constexpr int HB_WARPS = 4;

__device__ __forceinline__ float hb_artanh_c(float x, bool* clamped) {
  const float lim = 1.f - 1e-7f;
  *clamped = x > lim || x < -lim;
  x = fminf(fmaxf(x, -lim), lim);
  return 0.5f * (logf(1.f + x) - logf(1.f - x));
}

template <int NV>
__device__ __forceinline__ float hb_dot(const float4 (&a)[NV], const float4 (&b)[NV]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += a[i].x * b[i].x + a[i].y * b[i].y + a[i].z * b[i].z + a[i].w * b[i].w;
  return warp_sum(s);
}
template <int NV>
__device__ __forceinline__ void hb_axpby(float4 (&o)[NV], float a, const float4 (&x)[NV], float b, const float4 (&y)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    o[i].x = a * x[i].x + b * y[i].x; o[i].y = a * x[i].y + b * y[i].y;
    o[i].z = a * x[i].z + b * y[i].z; o[i].w = a * x[i].w + b * y[i].w;
  }
}

template <int NV>
__global__ void __launch_bounds__(HB_WARPS * 32)
mobius_epilogue_bwd_kernel(const float* __restrict__ mx, int64_t n, int d, const float* __restrict__ xsq_in,
                           const float* __restrict__ bias, float c, int post_tanh, int n_project,
                           const float* __restrict__ gy, float* __restrict__ gmx, float* __restrict__ gbias,
                           float* __restrict__ gxn) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const float sc = sqrtf(c), k = -c;
  const float maxnorm = (1.0f - 4e-3f) / sc;
  float4 b[NV], gb_acc[NV];                    // bias row; this warp's sum of dL/dbias over the rows it serves
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = i * 32 + lane;
    b[i] = (j < nvec && bias != nullptr) ? __ldg(reinterpret_cast<const float4*>(bias) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    gb_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // rows are taken grid-stride, so that the bias gradient costs one atomic per column and CTA, not per column and row
  for (int64_t row = (int64_t)blockIdx.x * HB_WARPS + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * HB_WARPS) {
  float4 m[NV], g[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = i * 32 + lane;
    const bool in = j < nvec;
    m[i] = in ? __ldg(reinterpret_cast<const float4*>(mx + row * d) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[i] = in ? __ldg(reinterpret_cast<const float4*>(gy + row * d) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // ------------------------------------------------------------------ forward scalars (as csrc/headgemm.cu)
  const float mm = hb_dot<NV>(m, m), mb = hb_dot<NV>(m, b), b2 = hb_dot<NV>(b, b);
  const float mn = fmaxf(sqrtf(mm), 1e-15f);
  float tt, dphi_dmn, dphi_dxn = 0.f;          // u = phi m, phi = tt / mn
  if (xsq_in != nullptr) {
    const float xn = fmaxf(sqrtf(xsq_in[row]), 1e-15f);
    bool cl;
    const float A = hb_artanh_c(sc * xn, &cl);
    const float r = mn * A / xn;                // sc * (mn / xn * (artanh / sc))
    const bool tcl = r > 15.f;
    const float th = tanhf(fminf(r, 15.f)), sech2 = tcl ? 0.f : 1.f - th * th;
    tt = th / sc;
    dphi_dmn = (sech2 * (A / xn) * mn - th) / (sc * mn * mn);
    const float dA = cl ? 0.f : sc / (1.f - c * xn * xn);
    dphi_dxn = sech2 * (dA * xn - A) / (xn * xn) / sc;
  } else {
    const float a = sc * mn;
    const bool tcl = a > 15.f;
    const float th = tanhf(fminf(a, 15.f)), sech2 = tcl ? 0.f : 1.f - th * th;
    tt = th / sc;
    dphi_dmn = (sech2 * sc * mn - th) / (sc * mn * mn);          // d/dmn [tanh(sc mn) / (sc mn)]
  }
  const bool zero_row = mm == 0.f;
  const float phi = zero_row ? 0.f : tt / mn;
  float4 u[NV], v[NV];
  hb_axpby<NV>(u, phi, m, 0.f, m);
  const float x2 = phi * phi * mm, xy = phi * mb;
  float ca = 1.f, cb = 0.f, den = 1.f;
  if (bias != nullptr) {
    ca = 1.f - 2.f * k * xy - k * b2;
    cb = 1.f + k * x2;
    den = fmaxf(1.f - 2.f * k * xy + k * k * x2 * b2, 1e-15f);
  }
  hb_axpby<NV>(v, ca / den, u, cb / den, b);                      // v = mobius_add(u, b)  (or u)
  // project (each application: scale s_p when the norm exceeds maxnorm)
  float ps[2] = {1.f, 1.f};
  float vsq = hb_dot<NV>(v, v);
  float4 y1[NV];
  hb_axpby<NV>(y1, 1.f, v, 0.f, v);
  for (int pj = 0; pj < n_project && pj < 2; ++pj) {
    const float nrm = fmaxf(sqrtf(vsq), 1e-15f);
    if (nrm > maxnorm) {
      ps[pj] = maxnorm / nrm;
      hb_axpby<NV>(y1, ps[pj], y1, 0.f, y1);
      vsq = hb_dot<NV>(y1, y1);
    }
  }
  // ------------------------------------------------------------------ backward
  if (post_tanh) {
    // out = e w, w = tanh(l y1) elementwise, l = artanh(sc yn) / (sc yn), e = tanh(sc tn) / (sc tn)
    const float yn = fmaxf(sqrtf(vsq), 1e-15f);
    bool cl;
    const float at = hb_artanh_c(sc * yn, &cl);
    const float l = at / (sc * yn);
    float4 w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      w[i].x = tanhf(y1[i].x * l); w[i].y = tanhf(y1[i].y * l); w[i].z = tanhf(y1[i].z * l); w[i].w = tanhf(y1[i].w * l);
    }
    const float wsq = hb_dot<NV>(w, w);
    const float tn = fmaxf(sqrtf(wsq), 1e-15f);
    const float a = sc * tn;
    const bool tcl = a > 15.f;
    const float th = tanhf(fminf(a, 15.f)), sech2 = tcl ? 0.f : 1.f - th * th;
    const float e = th / a;
    const float de = (sech2 - th / a) / tn;                        // de/dtn
    const float gw_dot = hb_dot<NV>(g, w);
    float4 ga[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {                                  // g_w = e g + de (g.w) w / tn ; g_a = g_w (1 - w^2)
      const float f = de * gw_dot / tn;
      ga[i].x = (e * g[i].x + f * w[i].x) * (1.f - w[i].x * w[i].x);
      ga[i].y = (e * g[i].y + f * w[i].y) * (1.f - w[i].y * w[i].y);
      ga[i].z = (e * g[i].z + f * w[i].z) * (1.f - w[i].z * w[i].z);
      ga[i].w = (e * g[i].w + f * w[i].w) * (1.f - w[i].w * w[i].w);
    }
    const float dl = cl ? -l / yn : (1.f / (1.f - c * yn * yn) - l) / yn;     // dl/dyn
    const float gy_dot = hb_dot<NV>(ga, y1);
    hb_axpby<NV>(g, l, ga, dl * gy_dot / yn, y1);                   // g_y1
  }
  // project^T, last application first: y = s v  =>  g_v = s (g - v^ (v^ . g)), v^ = y / |y|
  for (int pj = (n_project < 2 ? n_project : 2) - 1; pj >= 0; --pj) {
    if (ps[pj] != 1.f) {
      // the projected vector (after applications 0..pj) has norm maxnorm; its direction is that of y1
      const float yn2 = fmaxf(hb_dot<NV>(y1, y1), 1e-30f);
      const float gd = hb_dot<NV>(g, y1) / yn2;
      hb_axpby<NV>(g, ps[pj], g, -ps[pj] * gd, y1);
    }
  }
  // mobius_add^T
  float4 gu[NV];
  float4 gb[NV];
  if (bias != nullptr) {
    const float g_v_dot_v = hb_dot<NV>(g, v);
    const float inv_den = 1.f / den;
    const float g_den = -g_v_dot_v * inv_den;
    const float gN_u = hb_dot<NV>(g, u) * inv_den, gN_b = hb_dot<NV>(g, b) * inv_den;      // g_ca, g_cb
    const float g_xy = -2.f * k * gN_u - 2.f * k * g_den;
    const float g_x2 = k * gN_b + k * k * b2 * g_den;
    const float g_b2 = -k * gN_u + k * k * x2 * g_den;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      gu[i].x = ca * inv_den * g[i].x + g_xy * b[i].x + 2.f * g_x2 * u[i].x;
      gu[i].y = ca * inv_den * g[i].y + g_xy * b[i].y + 2.f * g_x2 * u[i].y;
      gu[i].z = ca * inv_den * g[i].z + g_xy * b[i].z + 2.f * g_x2 * u[i].z;
      gu[i].w = ca * inv_den * g[i].w + g_xy * b[i].w + 2.f * g_x2 * u[i].w;
      gb[i].x = cb * inv_den * g[i].x + g_xy * u[i].x + 2.f * g_b2 * b[i].x;
      gb[i].y = cb * inv_den * g[i].y + g_xy * u[i].y + 2.f * g_b2 * b[i].y;
      gb[i].z = cb * inv_den * g[i].z + g_xy * u[i].z + 2.f * g_b2 * b[i].z;
      gb[i].w = cb * inv_den * g[i].w + g_xy * u[i].w + 2.f * g_b2 * b[i].w;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      gb_acc[i].x += gb[i].x; gb_acc[i].y += gb[i].y; gb_acc[i].z += gb[i].z; gb_acc[i].w += gb[i].w;
    }
  } else {
    hb_axpby<NV>(gu, 1.f, g, 0.f, g);
  }
  // u = phi(mn, xn) m  =>  g_m = phi g_u + dphi/dmn (g_u . m) m / mn ;  dL/dxn = dphi/dxn (g_u . m)
  const float gum = hb_dot<NV>(gu, m);
  float4 gm[NV];
  if (zero_row) {
    hb_axpby<NV>(gm, 0.f, gu, 0.f, gu);
  } else {
    hb_axpby<NV>(gm, phi, gu, dphi_dmn * gum / mn, m);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = i * 32 + lane;
    if (j < nvec) reinterpret_cast<float4*>(gmx + row * d)[j] = gm[i];
  }
  if (gxn != nullptr && lane == 0)
    gxn[row] = zero_row ? 0.f : dphi_dxn * gum / fmaxf(sqrtf(xsq_in[row]), 1e-15f);   // so that dL/dx_in += gxn x_in
  }   // rows
  if (gbias != nullptr && bias != nullptr) {
    __shared__ float4 red[HB_WARPS][NV * 32];
#pragma unroll
    for (int i = 0; i < NV; ++i) red[threadIdx.x >> 5][i * 32 + lane] = gb_acc[i];
    __syncthreads();
    for (int j = threadIdx.x; j < nvec; j += HB_WARPS * 32) {
      float4 t = red[0][j];
#pragma unroll
      for (int w = 1; w < HB_WARPS; ++w) { t.x += red[w][j].x; t.y += red[w][j].y; t.z += red[w][j].z; t.w += red[w][j].w; }
      atomicAdd(gbias + 4 * j + 0, t.x); atomicAdd(gbias + 4 * j + 1, t.y);
      atomicAdd(gbias + 4 * j + 2, t.z); atomicAdd(gbias + 4 * j + 3, t.w);
    }
  }
}

// C[m,n] = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ row_scale[m] * addend[m,n]); 64 x 64 tile, 16-deep steps, 4 x 4 per thread
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
                     int64_t sbn, int M, int N, int K, const float* __restrict__ row_scale,
                     const float* __restrict__ addend, float* __restrict__ C) {
  __shared__ float As[16][64 + 4], Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  // split-K (gridDim.z > 1): this CTA takes K range [k_lo, k_hi) and adds its partial tile into C (zeroed by the
  // launcher) with atomics -- dW = gmx^T x has K = batch rows and only (N_out/64) x (D_in/64) output tiles
  const int k_per = (((K + (int)gridDim.z - 1) / (int)gridDim.z) + 15) / 16 * 16;
  const int k_lo = (int)blockIdx.z * k_per, k_hi = k_lo + k_per < K ? k_lo + k_per : K;
  float acc[4][4] = {};
  for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
    for (int t = threadIdx.x; t < 16 * 64; t += 256) {
      const int kk = t >> 6, mmi = t & 63;
      As[kk][mmi] = (m0 + mmi < M && k0 + kk < k_hi) ? A[(int64_t)(m0 + mmi) * sam + (int64_t)(k0 + kk) * sak] : 0.f;
      Bs[kk][mmi] = (n0 + mmi < N && k0 + kk < k_hi) ? B[(int64_t)(k0 + kk) * sbk + (int64_t)(n0 + mmi) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (m0 + ty * 4 + i < M && n0 + tx * 4 + j < N) {
        const int64_t at = (int64_t)(m0 + ty * 4 + i) * N + n0 + tx * 4 + j;
        const float v = acc[i][j] + ((row_scale != nullptr && blockIdx.z == 0) ? row_scale[m0 + ty * 4 + i] * addend[at] : 0.f);
        if (gridDim.z > 1) atomicAdd(C + at, v);
        else C[at] = v;
      }
}

}  // namespace

int hypret_launch_mobius_epilogue_bwd(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                                      int post_tanh, int n_project, const float* gy, float* gmx, float* gbias,
                                      float* gxn, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + HB_WARPS - 1) / HB_WARPS;
  const unsigned grid = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);      // grid-stride beyond
  const int need = (d + 127) / 128;
  if (need <= 1) mobius_epilogue_bwd_kernel<1><<<grid, HB_WARPS * 32, 0, stream>>>(mx, n, d, xsq, bias, c, post_tanh, n_project, gy, gmx, gbias, gxn);
  else if (need <= 2) mobius_epilogue_bwd_kernel<2><<<grid, HB_WARPS * 32, 0, stream>>>(mx, n, d, xsq, bias, c, post_tanh, n_project, gy, gmx, gbias, gxn);
  else if (need <= 4) mobius_epilogue_bwd_kernel<4><<<grid, HB_WARPS * 32, 0, stream>>>(mx, n, d, xsq, bias, c, post_tanh, n_project, gy, gmx, gbias, gxn);
  else return HYPRET_EUNSUPPORTED;
  return (int)cudaGetLastError();
}

int hypret_launch_sgemm_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                                int M, int N, int K, const float* row_scale, const float* addend, float* C,
                                cudaStream_t stream) {
  if (M == 0 || N == 0) return HYPRET_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  int splits = 1;                                             // split K when the output tiles cannot fill the GPU
  if (tiles < 2 * sms && K >= 1024) {
    splits = (2 * sms + tiles - 1) / tiles;
    if (splits > K / 256) splits = K / 256;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    cudaError_t e = cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), stream);
    if (e != cudaSuccess) return (int)e;
  }
  const dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)splits);
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, M, N, K, row_scale, addend, C);
  return (int)cudaGetLastError();
}
