// n x m Poincare distance matrix on the tensor cores (train_hyp, BASELINE config 5).
//
// Replaces the reference's O(n^2) Python double loop of 1x1 pmath.dist calls
// (/root/reference/src/train.py:1832-1840, 2304-2320) for the TRAINING path, where the matrix is dense and large
// (n = 8192: 67 M pairs).  pairdist.cu forms every difference explicitly on the CUDA cores (FP32-FMA bound,
// 2 n m D instructions); here the Gram matrix <a_i, p_j> comes from ONE tcgen05 GEMM whose operands carry a
// 3-way bf16 split of the fp32 rows along K,
//     A' = [ a_hi | a_mid | a_lo | a_hi | a_mid | a_hi ]        P' = [ p_hi | p_hi | p_hi | p_mid | p_mid | p_lo ]
// so that <A'_i, P'_j> = hi.hi + mid.hi + lo.hi + hi.mid + mid.mid + hi.lo = <a_i, p_j> to fp32 accuracy (the dropped
// cross terms are below 2^-24 relative), accumulated in fp32 in TMEM.  The epilogue turns it into
//     s = |a|^2 + |p|^2 - 2 <a,p>,   d = arccosh(1 + 2 c s / ((1 - c|a|^2)(1 - c|p|^2))) / sqrt(c)
// in registers.  s cancels for near pairs (the diagonal of the contrastive batch): whenever
// s < NEAR_FRAC (|a|^2 + |p|^2) the thread recomputes |a - p|^2 exactly from the fp32 rows (explicit differences);
// elsewhere the cancellation costs at most 1 / NEAR_FRAC ulps of fp32.
// Structure = the single-CTA skeleton of score_topk.cu: persistent CTAs, warp 0 TMA producer (A' block 128 x 64 and
// P' block 256 x 64 per stage, 128B swizzle, 4 stages), warp 1 MMA issuer (M=128, N=256, K=16 into one of two TMEM
// accumulators), warps 2-9 epilogue in two warpgroups (tcgen05.ld 32x32b: lane = row; odd / even column chunks).  Tiles are walked column-major so that the
// CTAs running together share the P' tile; both operands (25 MB at n = 8192, D = 128) live in L2.
#include <cuda.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int GT_M = 128, GT_N = 256, GT_ACC = 2, GT_STAGES = 3;
constexpr int GA_BLK = GT_M * HYPRET_KBLK * 2;   // 16 KB
constexpr int GB_BLK = GT_N * HYPRET_KBLK * 2;   // 32 KB
constexpr int G_STAGE = GA_BLK + GB_BLK;
constexpr int G_THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two warpgroups)
constexpr int G_EPI = G_THREADS - 64;
constexpr float NEAR_FRAC = 0.25f;
constexpr int G_PITCH = 36;       // floats per staged row: 16-byte aligned, 128-bit accesses conflict-free
constexpr int G_STAGING = 32 * G_PITCH * 4;                     // one 32 x 32 fp32 block per epilogue warp
constexpr int G_SMEM = 1024 + GT_STAGES * G_STAGE + GT_ACC * GT_N * 4 + 256 + (G_THREADS / 32 - 2) * G_STAGING;

__host__ __device__ inline int gram_kpad(int d) { return (6 * d + HYPRET_KBLK - 1) / HYPRET_KBLK * HYPRET_KBLK; }

// One warp per row: 3-way bf16 split laid out for side A (0) or side P (1), plus |x|^2.
__global__ void __launch_bounds__(256)
gram_split_kernel(const float* __restrict__ x, int64_t n, int d, int side, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int kp = gram_kpad(d);
  __nv_bfloat16* o = out + i * kp;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = x[i * d + k];
    acc = fmaf(v, v, acc);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    if (side == 0) {
      o[k] = hi; o[d + k] = mid; o[2 * d + k] = lo; o[3 * d + k] = hi; o[4 * d + k] = mid; o[5 * d + k] = hi;
    } else {
      o[k] = hi; o[d + k] = hi; o[2 * d + k] = hi; o[3 * d + k] = mid; o[4 * d + k] = mid; o[5 * d + k] = lo;
    }
  }
  for (int k = 6 * d + lane; k < kp; k += 32) o[k] = __float2bfloat16_rn(0.f);
  acc = warp_sum(acc);
  if (lane == 0) sq[i] = acc;
}

struct GBarriers {
  uint64_t full[GT_STAGES];
  uint64_t empty[GT_STAGES];
  uint64_t tmem_full[GT_ACC];
  uint64_t tmem_empty[GT_ACC];
  uint32_t tmem_ptr;
};

// d = arccosh(1 + t) / sqrt(c), t = 2 c s / (alpha beta): accurate log1pf, but no IEEE-division / IEEE-sqrt slow
// paths (approximate reciprocal and rsqrt, ~2 ulp each) -- the epilogue must keep up with the tensor pipe
__device__ __forceinline__ float dist_from_sq(float s, float al, float be, float two_c, float rs) {
  const float tt = __fdividef(two_c * s, al * be);
  const float t2 = fmaxf(tt * (tt + 2.0f), 1e-37f);
  return log1pf(tt + t2 * rsqrtf(t2)) * rs;
}

__device__ __forceinline__ float exact_sqdist(const float* __restrict__ a, const float* __restrict__ p, int d) {
  float s = 0.f;
  for (int k = 0; k < d; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + k), y = *reinterpret_cast<const float4*>(p + k);
    const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
    s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
  }
  return s;
}

__global__ void __launch_bounds__(G_THREADS, 1)
gram_dist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_p,
                 const float* __restrict__ a32, const float* __restrict__ p32, const float* __restrict__ asq,
                 const float* __restrict__ psq, int64_t n, int64_t m, int d, int kb, float c, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  float* psq_s = reinterpret_cast<float*>(smem + GT_STAGES * G_STAGE);          // [GT_ACC][GT_N]
  GBarriers* bars = reinterpret_cast<GBarriers*>(smem + GT_STAGES * G_STAGE + GT_ACC * GT_N * 4);
  float* staging = reinterpret_cast<float*>(smem + GT_STAGES * G_STAGE + GT_ACC * GT_N * 4 + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rt = (int)((n + GT_M - 1) / GT_M), n_ct = (int)((m + GT_N - 1) / GT_N);
  const int n_tiles = n_rt * n_ct;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_p);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < GT_ACC; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], G_EPI); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_ptr, GT_ACC * GT_N);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int ct = t / n_rt, rt = t - ct * n_rt;           // column-major walk
      for (int k = 0; k < kb; ++k) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = ring + stage * G_STAGE;
          mbar_arrive_expect_tx(&bars->full[stage], G_STAGE);
          tma_load_2d_hint(st, &map_a, &bars->full[stage], k * HYPRET_KBLK, rt * GT_M, TMA_EVICT_LAST);
          tma_load_2d_hint(st + GA_BLK, &map_p, &bars->full[stage], k * HYPRET_KBLK, ct * GT_N, TMA_EVICT_LAST);
        }
        __syncwarp();
        if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(GT_M, GT_N);
    constexpr uint64_t DESC_SW128 = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
                                    (UMMA_LAYOUT_SW128 << 61);
    const uint32_t ring_lo = smem_u32(ring) >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * GT_N;
      for (int k = 0; k < kb; ++k) {
        mbar_wait(&bars->full[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ring_lo + stage * (G_STAGE >> 4), b_lo = a_lo + (GA_BLK >> 4);
#pragma unroll
          for (int kk = 0; kk < HYPRET_KBLK / 16; ++kk)
            umma_f16_ss(d_tmem, DESC_SW128 | (a_lo + 2 * kk), DESC_SW128 | (b_lo + 2 * kk), idesc,
                         (k | kk) != 0 ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (k == kb - 1) umma_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == GT_ACC) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue: Gram entry -> distance
    // two warpgroups share every accumulator: warpgroup g takes the odd / even 32-column chunks.  With one
    // warpgroup each SM sub-partition held a single epilogue warp, whose dependent MUFU / FMA chains issue one
    // instruction every few cycles (ncu: 0.24 ms per 8192 x 8192 matrix, tensor pipe 7 % active, 9 % of the warp
    // slots in use); two warps per sub-partition hide each other's latencies
    const int quad = warp & 3, row = quad * 32 + lane, et = threadIdx.x - 64;     // et: 0..G_EPI-1
    const int wg = (warp - 2) >> 2;
    const float rs = 1.0f / sqrtf(c), two_c = 2.0f * c;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int ct = t / n_rt, rt = t - ct * n_rt;
      const int64_t i = (int64_t)rt * GT_M + row;
      const int64_t j0 = (int64_t)ct * GT_N;
      // this tile's column norms (the buffer of accumulator `acc` was last read two tiles ago)
      float* pq = psq_s + acc * GT_N;
      for (int u = et; u < GT_N; u += G_EPI) pq[u] = (j0 + u < m) ? psq[j0 + u] : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(G_EPI) : "memory");
      const float na = i < n ? asq[i] : 0.f;
      const float al = 1.0f - c * na;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * GT_N;
      float v[32];
#pragma unroll 1
      for (int cc = wg; cc < GT_N / 32; cc += 2) {
        __syncwarp();
        tmem_ld_32x32(taddr + cc * 32, v);
        tmem_ld_wait(v);
        const int64_t jb = j0 + cc * 32;
        if (jb + 32 <= m && (m & 3) == 0) {        // warp-uniform: a full chunk of an aligned matrix
          // Rows past n compute on zero-filled operands and are masked at the stores.  Main pass: no branches, the
          // 32 entries are independent (their MUFU chains overlap), log(1 + x) from the fast lg2 (absolute error
          // 2^-21.4: below 1e-6 relative once x > 0.5).  Two kinds of entries are only flagged here and redone below:
          // near pairs, whose |a|^2 + |p|^2 - 2<a,p> cancels (exact differences from the fp32 rows), and entries
          // with x <= 0.5, i.e. d sqrt(c) < 0.41 (accurate log1pf on the value already at hand).
          // lane = row, so a direct float4 store would touch 32 different lines per instruction (ncu: the warps
          // sat behind the store queue, 14 % of all stall samples on the first instruction that reuses a store's
          // data register): the 32 x 32 block goes through a per-warp staging buffer and leaves as 8 stores of 4
          // full 128-byte lines each.
          float* st = staging + (warp - 2) * (32 * G_PITCH);
          uint32_t near = 0, slow = 0;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float r[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 + jj;
              const float nb = pq[cc * 32 + j];
              const float s = fmaxf(na + nb - 2.0f * v[j], 0.f);
              near |= (s < NEAR_FRAC * (na + nb) ? 1u : 0u) << j;
              const float tt = __fdividef(two_c * s, al * (1.0f - c * nb));
              const float t2 = fmaxf(tt * (tt + 2.0f), 1e-37f);
              const float x = tt + t2 * rsqrtf(t2);
              slow |= (x <= 0.5f ? 1u : 0u) << j;
              r[jj] = __logf(1.0f + x) * rs;
            }
            *reinterpret_cast<float4*>(st + lane * G_PITCH + j4) = make_float4(r[0], r[1], r[2], r[3]);
          }
          if (i >= n) near = slow = 0u;
          slow &= ~near;
          if (slow != 0u) {                       // close pairs that do not cancel: same s, accurate logarithm
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((slow >> j) & 1u) {
                const float nb = pq[cc * 32 + j];
                st[lane * G_PITCH + j] = dist_from_sq(fmaxf(na + nb - 2.0f * v[j], 0.f), al, 1.0f - c * nb, two_c, rs);
              }
          }
          while (near != 0u) {                    // rare: the diagonal of a contrastive batch, duplicates
            const int j = __ffs((int)near) - 1;
            near &= near - 1u;
            const float s = exact_sqdist(a32 + i * d, p32 + (jb + j) * d, d);
            st[lane * G_PITCH + j] = dist_from_sq(s, al, 1.0f - c * pq[cc * 32 + j], two_c, rs);
          }
          __syncwarp();
          const int64_t r_base = (int64_t)rt * GT_M + quad * 32;
          const int sub = lane >> 3, col = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + sub;
            const float4 x = *reinterpret_cast<const float4*>(st + r * G_PITCH + col);
            if (r_base + r < n) *reinterpret_cast<float4*>(out + (r_base + r) * m + jb + col) = x;
          }
          __syncwarp();                           // the buffer is rewritten by the next chunk
        } else if (jb < m && i < n) {             // ragged edge / unaligned matrix: accurate path, direct stores
          uint32_t near = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float nb = pq[cc * 32 + j];
            const float s = fmaxf(na + nb - 2.0f * v[j], 0.f);
            near |= (s < NEAR_FRAC * (na + nb) ? 1u : 0u) << j;
            v[j] = dist_from_sq(s, al, 1.0f - c * nb, two_c, rs);
          }
          float* o = out + i * m + jb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (jb + j < m) o[j] = v[j];
          while (near != 0u) {
            const int j = __ffs((int)near) - 1;
            near &= near - 1u;
            if (jb + j < m) {
              const float s = exact_sqdist(a32 + i * d, p32 + (jb + j) * d, d);
              o[j] = dist_from_sq(s, al, 1.0f - c * pq[cc * 32 + j], two_c, rs);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&bars->tmem_empty[acc]);
      if (++acc == GT_ACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, GT_ACC * GT_N);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int gram_make_map(CUtensorMap* map, const void* base, int64_t rows, int kpad, int box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (enc == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return HYPRET_EDRIVER;
    enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)HYPRET_KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HYPRET_OK : HYPRET_EINVAL;
}

}  // namespace

int64_t hypret_gram_kpad_impl(int d) { return gram_kpad(d); }

int hypret_launch_gram_split(const float* x, int64_t n, int d, int side, void* out_bf16, float* sq, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  gram_split_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(x, n, d, side, static_cast<__nv_bfloat16*>(out_bf16), sq);
  return (int)cudaGetLastError();
}

int hypret_launch_gram_dist(const void* a_op, const void* p_op, const float* a32, const float* p32, const float* asq,
                            const float* psq, int64_t n, int64_t m, int d, float c, float* out, cudaStream_t stream) {
  if (n == 0 || m == 0) return HYPRET_OK;
  const int kp = gram_kpad(d);
  CUtensorMap map_a, map_p;
  int rc;
  if ((rc = gram_make_map(&map_a, a_op, n, kp, GT_M))) return rc;
  if ((rc = gram_make_map(&map_p, p_op, m, kp, GT_N))) return rc;
  cudaError_t e = cudaFuncSetAttribute(gram_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ((n + GT_M - 1) / GT_M) * ((m + GT_N - 1) / GT_N);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  gram_dist_kernel<<<grid, G_THREADS, G_SMEM, stream>>>(map_a, map_p, a32, p32, asq, psq, n, m, d, kp / HYPRET_KBLK, c,
                                                       out);
  return (int)cudaGetLastError();
}
