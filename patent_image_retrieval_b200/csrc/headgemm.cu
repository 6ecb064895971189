// MobiusLinear as ONE kernel: tcgen05 GEMM + the whole hyperbolic epilogue on the accumulator.
//
// Replaces, per layer of the learned projection head (/root/reference/src/models.py:291-318, 481-505),
//   F.linear / the matrix product inside pmath.mobius_matvec   ->  a library SGEMM writing [B, Dout] fp32, then
//   expmap0 | matvec rescale, mobius_add(bias), project (x n), mobius_fn_apply(tanh)  ->  ~25-50 elementwise launches
// (round 1: one cuBLAS GEMM + one epilogue kernel per layer).  Here the product runs on tcgen05 from 2-way fp16 split
// operands (row layout [hi|lo|hi] x column layout [hi|hi|lo]: hi.hi + lo.hi + hi.lo = the fp32 product to 2^-22,
// operands from hypret_flash_prep), the accumulator tile [128 rows x Dout <= 256 columns] stays in TMEM, and the
// epilogue -- lane = row, so every norm / inner product of the Moebius arithmetic is a private running sum over the
// row's columns -- makes up to three passes over it: (1) ||mx||^2 and <mx, bias>, from which expmap0 / the matvec
// rescale, mobius_add and project collapse into two scalars (y = A mx + B bias) and ||y||^2 in closed form; (2) for the
// encoder's tanh-in-tangent-space activation, ||tanh(.)||^2; (3) the outputs.  It emits, as requested: the fp32 row
// (the layer output / the activations autograd keeps), ||y||^2 (the next layer's mobius_matvec needs ||x||), and the
// NEXT layer's GEMM operand row [hi|lo|hi] in fp16 -- the hidden activations go to the second GEMM without an fp32
// [B, 256] round trip.  Arithmetic follows the oracle's restatement of geoopt (oracle/pmath.py) in fp32.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int HG_M = 128;
constexpr int HG_STAGES = 4;
constexpr int HG_THREADS = 192;        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (lane = row)
constexpr int HG_EPI = 128;
constexpr int HG_A_BLK = HG_M * HYPRET_KBLK * 2;

struct HGBarriers {
  uint64_t full[HG_STAGES];
  uint64_t empty[HG_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_ptr;
};

struct HGParams {
  int64_t n;                 // rows
  int n_out, kb;             // output columns (multiple of 16, <= 256), K blocks of the split operands
  const float* bias;         // [n_out] on-ball bias or NULL
  const float* xsq;          // [n] ||x||^2 of the layer input (hyperbolic input) or NULL (Euclidean input: expmap0)
  float c;
  int post_tanh, n_project;
  float* mx_out;             // [n, n_out] raw product (kept for the backward pass) or NULL
  float* y_out;              // [n, n_out] fp32 output or NULL
  float* ysq_out;            // [n] or NULL
  __half* op_out;            // [n, op_kpad] next layer's row operand [hi|lo|hi] or NULL
  int op_kpad;
};

__device__ __forceinline__ float hg_tanh_c(float x) { return tanhf(fminf(fmaxf(x, -15.f), 15.f)); }
__device__ __forceinline__ float hg_artanh_c(float x) {
  x = fminf(fmaxf(x, -1.f + 1e-7f), 1.f - 1e-7f);
  return 0.5f * (logf(1.f + x) - logf(1.f - x));
}

__global__ void __launch_bounds__(HG_THREADS, 1)
mobius_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const HGParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_blk = p.n_out * HYPRET_KBLK * 2;
  const int stage_bytes = HG_A_BLK + b_blk;
  uint8_t* ring = smem;
  float* bias_s = reinterpret_cast<float*>(smem + HG_STAGES * stage_bytes);       // [n_out]
  HGBarriers* bars = reinterpret_cast<HGBarriers*>(bias_s + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (int)((p.n + HG_M - 1) / HG_M);
  const int acc_cols = p.n_out <= 128 ? 128 : 256;          // TMEM columns per accumulator (power of two)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HG_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->acc_full[a], 1); mbar_init(&bars->acc_empty[a], HG_EPI); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_ptr, 512);
  for (int j = threadIdx.x; j < 256; j += HG_THREADS) bias_s[j] = (p.bias != nullptr && j < p.n_out) ? p.bias[j] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int k = 0; k < p.kb; ++k) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = ring + stage * stage_bytes;
          mbar_arrive_expect_tx(&bars->full[stage], stage_bytes);
          tma_load_2d_hint(st, &map_x, &bars->full[stage], k * HYPRET_KBLK, t * HG_M, TMA_EVICT_FIRST);
          tma_load_2d_hint(st + HG_A_BLK, &map_w, &bars->full[stage], k * HYPRET_KBLK, 0, TMA_EVICT_LAST);
        }
        __syncwarp();
        if (++stage == HG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = umma_idesc_f16(HG_M, p.n_out);
    constexpr uint64_t DESC_SW128 = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
                                    (UMMA_LAYOUT_SW128 << 61);
    const uint32_t ring_lo = smem_u32(ring) >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * acc_cols;
      for (int k = 0; k < p.kb; ++k) {
        mbar_wait(&bars->full[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ring_lo + stage * (stage_bytes >> 4), b_lo = a_lo + (HG_A_BLK >> 4);
#pragma unroll
          for (int kk = 0; kk < HYPRET_KBLK / 16; ++kk)
            umma_f16_ss(d_tmem, DESC_SW128 | (a_lo + 2 * kk), DESC_SW128 | (b_lo + 2 * kk), idesc,
                         (k | kk) != 0 ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (k == p.kb - 1) umma_commit(&bars->acc_full[acc]);
        }
        __syncwarp();
        if (++stage == HG_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue: the Moebius arithmetic, lane = row
    const int quad = warp & 3, row = quad * 32 + lane;
    const float sc = sqrtf(p.c), k = -p.c;
    const float maxnorm = (1.0f - 4e-3f) / sc;                 // geoopt project(): eps = 4e-3 for float32
    float b2 = 0.f;
    for (int j = 0; j < p.n_out; ++j) b2 = fmaf(bias_s[j], bias_s[j], b2);
    const int n_chunks = p.n_out / 32 + ((p.n_out & 31) ? 1 : 0);
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t i = (int64_t)t * HG_M + row;
      const bool row_ok = i < p.n;
      mbar_wait(&bars->acc_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * acc_cols;
      float v[32];
      // ---- pass 1: ||mx||^2, <mx, bias>; the raw product leaves here when the backward pass will need it
      float mm = 0.f, mb = 0.f;
      for (int cc = 0; cc < n_chunks; ++cc) {
        __syncwarp();
        tmem_ld_32x32(taddr + cc * 32, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (cc * 32 + j < p.n_out) { mm = fmaf(v[j], v[j], mm); mb = fmaf(v[j], bias_s[cc * 32 + j], mb); }
        }
        if (p.mx_out != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (cc * 32 + j < p.n_out)
              *reinterpret_cast<float4*>(p.mx_out + i * p.n_out + cc * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      // ---- expmap0 (Euclidean input) or the mobius_matvec rescale, mobius_add(bias), project: y = A mx + B bias
      const float mn = fmaxf(sqrtf(mm), 1e-15f);
      float tt;
      if (p.xsq != nullptr) {
        const float xn = fmaxf(sqrtf(row_ok ? p.xsq[i] : 0.f), 1e-15f);
        tt = hg_tanh_c(sc * (mn / xn * (hg_artanh_c(sc * xn) / sc))) / sc;
        if (mm == 0.f) tt = 0.f;                               // mx == 0 -> zero row (geoopt's cond)
      } else {
        tt = hg_tanh_c(sc * mn) / sc;
      }
      float A = tt / mn, B = 0.f, ysq = tt * tt;               // u = A mx, ||u||^2 = tt^2
      if (mm == 0.f) ysq = 0.f;
      if (p.bias != nullptr) {
        const float x2 = ysq, xy = A * mb;
        const float ca = 1.f - 2.f * k * xy - k * b2, cb = 1.f + k * x2;
        const float den = fmaxf(1.f - 2.f * k * xy + k * k * x2 * b2, 1e-15f);
        A = ca * A / den;
        B = cb / den;
        ysq = fmaxf(A * A * mm + 2.f * A * B * mb + B * B * b2, 0.f);
      }
      for (int pj = 0; pj < p.n_project; ++pj) {
        const float nrm = fmaxf(sqrtf(ysq), 1e-15f);
        if (nrm > maxnorm) {
          const float s = maxnorm / nrm;
          A *= s; B *= s; ysq *= s * s;
        }
      }
      // ---- pass 2 (tanh in the tangent space): w = tanh(logmap0(y)) elementwise, ||w||^2
      float e = 1.f, l = 0.f;
      if (p.post_tanh) {
        const float yn = fmaxf(sqrtf(ysq), 1e-15f);
        l = (hg_artanh_c(sc * yn) / sc) / yn;
        float wsq = 0.f;
        for (int cc = 0; cc < n_chunks; ++cc) {
          __syncwarp();
          tmem_ld_32x32(taddr + cc * 32, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cc * 32 + j < p.n_out) {
              const float w = tanhf((A * v[j] + B * bias_s[cc * 32 + j]) * l);
              wsq = fmaf(w, w, wsq);
            }
          }
        }
        const float tn = fmaxf(sqrtf(wsq), 1e-15f);
        const float th = hg_tanh_c(sc * tn) / sc;
        e = th / tn;
        ysq = th * th;
        if (wsq == 0.f) ysq = 0.f;
      }
      // ---- pass 3: outputs (fp32 row, next layer's fp16 split operand row)
      for (int cc = 0; cc < n_chunks; ++cc) {
        __syncwarp();
        tmem_ld_32x32(taddr + cc * 32, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y = A * v[j] + B * bias_s[(cc * 32 + j) & 255];
          if (p.post_tanh) y = e * tanhf(y * l);
          v[j] = y;
        }
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (cc * 32 + j >= p.n_out) continue;
            HYPRET_CHECK(i >= 0 && i < p.n && cc * 32 + j + 8 <= p.n_out && 3 * p.n_out <= p.op_kpad);
            if (p.y_out != nullptr) {
              float* o = p.y_out + i * p.n_out + cc * 32 + j;
              *reinterpret_cast<float4*>(o) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
            }
            if (p.op_out != nullptr) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(v[j + 2 * q], v[j + 2 * q + 1]);
                const __half2 r = __floats2half2_rn(v[j + 2 * q] - __low2float(h), v[j + 2 * q + 1] - __high2float(h));
                hi[q] = *reinterpret_cast<const uint32_t*>(&h);
                lo[q] = *reinterpret_cast<const uint32_t*>(&r);
              }
              __half* o = p.op_out + i * p.op_kpad + cc * 32 + j;
              *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(o + p.n_out) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(o + 2 * p.n_out) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            }
          }
        }
      }
      if (row_ok) {
        if (p.ysq_out != nullptr) p.ysq_out[i] = ysq;
        if (p.op_out != nullptr)
          for (int j = 3 * p.n_out; j < p.op_kpad; ++j) p.op_out[i * p.op_kpad + j] = __float2half_rn(0.f);
      }
      tcgen05_fence_before();
      mbar_arrive(&bars->acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int hg_make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
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
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HYPRET_OK : HYPRET_EINVAL;
}

}  // namespace

int hypret_launch_mobius_gemm(const void* x_row_op, const void* w_col_op, int64_t n, int d_in, int n_out,
                              const float* xsq, const float* bias, float c, int post_tanh, int n_project, float* mx_out,
                              float* y_out, float* ysq_out, void* op_out, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  const int kp = (int)hypret_flash_kpad_impl(d_in);
  CUtensorMap map_x, map_w;
  int rc;
  if ((rc = hg_make_map(&map_x, x_row_op, n, kp, HG_M))) return rc;
  if ((rc = hg_make_map(&map_w, w_col_op, n_out, kp, n_out))) return rc;
  HGParams p;
  p.n = n; p.n_out = n_out; p.kb = kp / HYPRET_KBLK; p.bias = bias; p.xsq = xsq; p.c = c;
  p.post_tanh = post_tanh; p.n_project = n_project; p.mx_out = mx_out; p.y_out = y_out; p.ysq_out = ysq_out;
  p.op_out = static_cast<__half*>(op_out);
  p.op_kpad = (int)hypret_flash_kpad_impl(n_out);
  const int smem = 1024 + HG_STAGES * (HG_A_BLK + n_out * HYPRET_KBLK * 2) + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(mobius_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (n + HG_M - 1) / HG_M;
  mobius_gemm_kernel<<<(unsigned)(tiles < sms ? tiles : sms), HG_THREADS, smem, stream>>>(map_x, map_w, p);
  return (int)cudaGetLastError();
}
