// Scoring GEMM + fused streaming top-k' for sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces the reference's one-vs-all scoring loops and the ranking that follows them
//   pmath.dist(q[1,D], G[P,D])            /root/reference/src/train.py:3259
//   cosine_similarity(Q, G)               /root/reference/notebooks/retrieval.ipynb:368
//   np.argsort / torch.topk               retrieval.ipynb:383,202 ; src/auxiliary.py:374
// as a candidate filter: S[i,j] = <q_op[i,:], g_op[j,:]> is the ranking surrogate built by
// project.cu (rb_j * ||x_i - y_j||^2 for the Poincare ball, -cos for cosine); per query the
// kernel keeps the k' smallest S per gallery split.  S lives only in TMEM and registers.
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   TMA producer   gallery K-blocks (256 rows x 64 bf16, 128B swizzle) into a
//                           shared-memory ring; the 128-row query tile is loaded ONCE per work
//                           item and stays resident in shared memory when D <= 512 (RESIDENT),
//                           otherwise it streams through the ring next to the gallery block
//   warp 1   MMA issuer     one thread issues tcgen05.mma (M=128, N=256, K=16, bf16 -> fp32)
//                           into one of two 256-column TMEM accumulators; tcgen05.commit
//                           releases ring stages and publishes finished accumulators
//   warps 2-5 epilogue      tcgen05.ld 32x32b: thread t owns query row t of the tile, so the
//                           running top-k' of a query is thread-private: register threshold,
//                           min-tree fast reject, rare insert into a per-thread list in
//                           shared memory (conflict-free [slot][thread] layout)
// Work item = (query tile, gallery split); items are ordered split-major and dealt
// round-robin so the CTAs that run concurrently stream the SAME gallery range through L2
// (the gallery is read from HBM about once per pass instead of once per query tile).
#include <cuda.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int TILE_N = 256;
constexpr int NUM_ACC = 2;
constexpr int TMEM_COLS = NUM_ACC * TILE_N;               // 512
constexpr int A_BLK_BYTES = TILE_M * HYPRET_KBLK * 2;     // 16 KB
constexpr int B_BLK_BYTES = TILE_N * HYPRET_KBLK * 2;     // 32 KB
constexpr int A_EXT_BYTES = TILE_M * HYPRET_KEXT * 2;     // 4 KB
constexpr int B_EXT_BYTES = TILE_N * HYPRET_KEXT * 2;     // 8 KB
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;                        // 227 KB opt-in maximum per CTA
constexpr int BAR_BYTES = 256;

struct Params {
  int64_t Q;
  int64_t N;
  int kb_main;          // number of 64-wide K blocks (Dpad / 64)
  int has_ext;          // 1: extension K block present (hyperbolic surrogate constants)
  int dpad;
  int n_qtiles, n_gtiles, n_splits, tiles_per_split, n_items;
  int kprime;
  int stages;
  int stage_bytes;
  int ring_off;         // byte offsets from the 1024-aligned shared-memory base
  int lists_off;
  int bar_off;
  float* cand_score;
  int32_t* cand_idx;
  float* debug_scores;
};

struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint64_t a_full;
  uint64_t a_empty;
  uint32_t tmem_ptr;
};
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block too large");

template <bool RESIDENT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap map_q_main, const __grid_constant__ CUtensorMap map_q_ext,
                  const __grid_constant__ CUtensorMap map_g_main, const __grid_constant__ CUtensorMap map_g_ext,
                  const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte aligned bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_res = smem;                                   // RESIDENT: kb_main blocks + ext block
  uint8_t* ring = smem + p.ring_off;
  float* list_s = reinterpret_cast<float*>(smem + p.lists_off);
  int* list_i = reinterpret_cast<int*>(list_s + p.kprime * TILE_M);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + p.bar_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.kb_main;
  const int KSTEPS = KB + (p.has_ext ? 1 : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q_main);
    tma_prefetch_desc(&map_q_ext);
    tma_prefetch_desc(&map_g_main);
    tma_prefetch_desc(&map_g_ext);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], NUM_EPI_THREADS);
    }
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_ptr, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_par = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int split = item / p.n_qtiles;
        const int qt = item - split * p.n_qtiles;
        const int gt0 = split * p.tiles_per_split;
        const int gt1 = min(p.n_gtiles, gt0 + p.tiles_per_split);
        if (RESIDENT) {
          if (it > 0) {  // the previous item's MMAs must be done reading the resident tile
            mbar_wait(&bars->a_empty, a_par);
            a_par ^= 1;
          }
          mbar_arrive_expect_tx(&bars->a_full, KB * A_BLK_BYTES + (p.has_ext ? A_EXT_BYTES : 0));
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d_hint(a_res + kb * A_BLK_BYTES, &map_q_main, &bars->a_full, kb * HYPRET_KBLK, qt * TILE_M,
                             TMA_EVICT_LAST);
          if (p.has_ext)
            tma_load_2d_hint(a_res + KB * A_BLK_BYTES, &map_q_ext, &bars->a_full, p.dpad, qt * TILE_M,
                             TMA_EVICT_LAST);
        }
        for (int gt = gt0; gt < gt1; ++gt) {
          for (int ks = 0; ks < KSTEPS; ++ks) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* st = ring + stage * p.stage_bytes;
            uint8_t* st_b = RESIDENT ? st : st + A_BLK_BYTES;
            const bool ext = (ks == KB);
            uint32_t bytes = ext ? B_EXT_BYTES : B_BLK_BYTES;
            if (!RESIDENT) bytes += ext ? A_EXT_BYTES : A_BLK_BYTES;
            mbar_arrive_expect_tx(&bars->full[stage], bytes);
            if (!ext) {
              tma_load_2d(st_b, &map_g_main, &bars->full[stage], ks * HYPRET_KBLK, gt * TILE_N);
              if (!RESIDENT)
                tma_load_2d_hint(st, &map_q_main, &bars->full[stage], ks * HYPRET_KBLK, qt * TILE_M, TMA_EVICT_LAST);
            } else {
              tma_load_2d(st_b, &map_g_ext, &bars->full[stage], p.dpad, gt * TILE_N);
              if (!RESIDENT)
                tma_load_2d_hint(st, &map_q_ext, &bars->full[stage], p.dpad, qt * TILE_M, TMA_EVICT_LAST);
            }
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, TILE_N);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_par = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int split = item / p.n_qtiles;
        const int gt0 = split * p.tiles_per_split;
        const int gt1 = min(p.n_gtiles, gt0 + p.tiles_per_split);
        if (RESIDENT) {
          mbar_wait(&bars->a_full, a_par);
          a_par ^= 1;
          tcgen05_fence_after();
        }
        for (int gt = gt0; gt < gt1; ++gt) {
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this accumulator
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TILE_N;
          for (int ks = 0; ks < KSTEPS; ++ks) {
            mbar_wait(&bars->full[stage], phase);
            tcgen05_fence_after();
            uint8_t* st = ring + stage * p.stage_bytes;
            const uint32_t b_addr = smem_u32(RESIDENT ? st : st + A_BLK_BYTES);
            if (ks < KB) {
              const uint32_t a_addr = smem_u32(RESIDENT ? a_res + ks * A_BLK_BYTES : st);
#pragma unroll
              for (int k = 0; k < HYPRET_KBLK / 16; ++k) {
                // +32 bytes per UMMA_K step inside the 128-byte swizzle span
                const uint64_t ad = umma_smem_desc(a_addr + k * 32, 1024, UMMA_LAYOUT_SW128);
                const uint64_t bd = umma_smem_desc(b_addr + k * 32, 1024, UMMA_LAYOUT_SW128);
                umma_bf16_ss(d_tmem, ad, bd, idesc, (ks | k) != 0 ? 1u : 0u);
              }
            } else {
              const uint32_t a_addr = smem_u32(RESIDENT ? a_res + KB * A_BLK_BYTES : st);
              const uint64_t ad = umma_smem_desc(a_addr, 256, UMMA_LAYOUT_SW32);
              const uint64_t bd = umma_smem_desc(b_addr, 256, UMMA_LAYOUT_SW32);
              umma_bf16_ss(d_tmem, ad, bd, idesc, 1u);
            }
            umma_commit(&bars->empty[stage]);                  // ring stage reusable once these MMAs retire
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&bars->tmem_full[acc]);                  // accumulator complete -> epilogue
          if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
        }
        if (RESIDENT) umma_commit(&bars->a_empty);
      }
    }
    __syncwarp();
  } else {
    // ======================================================================= epilogue / top-k'
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;          // query row inside the tile == TMEM lane
    float* ls = list_s + row;                  // [slot][thread] layout: stride TILE_M
    int* li = list_i + row;
    const int KP = p.kprime;
    const int64_t N = p.N;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int split = item / p.n_qtiles;
      const int qt = item - split * p.n_qtiles;
      const int gt0 = split * p.tiles_per_split;
      const int gt1 = min(p.n_gtiles, gt0 + p.tiles_per_split);
      const int64_t qrow = (int64_t)qt * TILE_M + row;
      float thr = INFINITY;
      int maxpos = 0;
      for (int s = 0; s < KP; ++s) {
        ls[s * TILE_M] = INFINITY;
        li[s * TILE_M] = -1;
      }
      for (int gt = gt0; gt < gt1; ++gt) {
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TILE_N;
        const int64_t col0 = (int64_t)gt * TILE_N;
#pragma unroll 1
        for (int cc = 0; cc < TILE_N / 32; ++cc) {
          float v[32];
          tmem_ld_32x32(taddr + cc * 32, v);
          tmem_ld_wait();
          const int64_t cbase = col0 + cc * 32;
          if (p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j < N) p.debug_scores[qrow * N + cbase + j] = v[j];
          }
          float m = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) m = fminf(m, v[j]);
          if (m < thr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (v[j] < thr && cbase + j < N) {
                ls[maxpos * TILE_M] = v[j];
                li[maxpos * TILE_M] = static_cast<int>(cbase + j);
                float mx = -INFINITY;
                int mp = 0;
                for (int s = 0; s < KP; ++s) {
                  const float x = ls[s * TILE_M];
                  if (x > mx) { mx = x; mp = s; }
                }
                thr = mx;
                maxpos = mp;
              }
            }
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&bars->tmem_empty[acc]);
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (qrow < p.Q) {
        const int64_t base = (qrow * p.n_splits + split) * KP;
        for (int s = 0; s < KP; ++s) {
          p.cand_score[base + s] = ls[s * TILE_M];
          p.cand_idx[base + s] = li[s * TILE_M];
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor [rows, kpad] row-major; box = box_cols x box_rows starting at (col, row).
int make_map(CUtensorMap* map, const void* base, int64_t rows, int kpad, int box_cols, int box_rows,
             CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return HYPRET_EDRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HYPRET_OK : HYPRET_EINVAL;
}

int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;  // B200; lets the plan be queried on a machine without a GPU
  }
  return sms;
}

}  // namespace

extern "C" int hypret_score_plan(int64_t Q, int64_t N, int d, int kprime, int n_splits_hint,
                                 hypret_score_plan_t* plan) {
  if (plan == nullptr || Q < 1 || N < 1 || d < 1 || kprime < 1 || kprime > 32) return HYPRET_EINVAL;
  if (N > 0x7fffffffll - TILE_N) return HYPRET_EUNSUPPORTED;   // int32 candidate indices per shard
  const int sms = device_sms();
  const int kb = hypret_dpad(d) / HYPRET_KBLK;
  const int64_t n_qtiles = (Q + TILE_M - 1) / TILE_M;
  const int64_t n_gtiles = (N + TILE_N - 1) / TILE_N;
  if (n_qtiles > (1 << 24)) return HYPRET_EUNSUPPORTED;

  // shared-memory carve-up
  const int lists = kprime * TILE_M * 8;
  const int a_res_bytes = kb * A_BLK_BYTES + A_EXT_BYTES;
  int resident = 0, stages = 0, stage_bytes = 0;
  {
    const int avail = SMEM_LIMIT - 1024 - BAR_BYTES - lists - a_res_bytes;
    if (kb <= 8 && avail >= 2 * B_BLK_BYTES) {
      resident = 1;
      stage_bytes = B_BLK_BYTES;
      stages = avail / B_BLK_BYTES;
    } else {
      stage_bytes = A_BLK_BYTES + B_BLK_BYTES;
      stages = (SMEM_LIMIT - 1024 - BAR_BYTES - lists) / stage_bytes;
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return HYPRET_EUNSUPPORTED;
  }

  // gallery splits: minimise (waves) x (tiles per item + fixed per-item cost)
  int64_t best_s = 1, best_tps = n_gtiles;
  if (n_splits_hint > 0) {
    int64_t s = n_splits_hint < n_gtiles ? n_splits_hint : n_gtiles;
    best_tps = (n_gtiles + s - 1) / s;
    best_s = (n_gtiles + best_tps - 1) / best_tps;
  } else {
    double best_cost = 1e300;
    const int64_t smax = n_gtiles < 64 ? n_gtiles : 64;
    for (int64_t s = 1; s <= smax; ++s) {
      const int64_t tps = (n_gtiles + s - 1) / s;
      const int64_t s2 = (n_gtiles + tps - 1) / tps;
      const int64_t items = n_qtiles * s2;
      const int64_t waves = (items + sms - 1) / sms;
      const double cost = (double)waves * ((double)tps + 3.0) + 0.02 * (double)s2;
      if (cost < best_cost) { best_cost = cost; best_s = s2; best_tps = tps; }
    }
  }
  const int64_t items = n_qtiles * best_s;
  if (items > 0x7fffffffll) return HYPRET_EUNSUPPORTED;
  plan->n_qtiles = (int32_t)n_qtiles;
  plan->n_gtiles = (int32_t)n_gtiles;
  plan->n_splits = (int32_t)best_s;
  plan->tiles_per_split = (int32_t)best_tps;
  plan->grid = (int32_t)(items < sms ? items : sms);
  plan->stages = stages;
  plan->resident = resident;
  plan->smem_bytes = 1024 + (resident ? a_res_bytes : 0) + stages * stage_bytes + lists + BAR_BYTES;
  return HYPRET_OK;
}

int hypret_launch_score_topk(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime,
                             int n_splits, float* cand_score, int32_t* cand_idx, float* debug_scores,
                             cudaStream_t stream) {
  hypret_score_plan_t plan;
  int rc = hypret_score_plan(Q, N, d, kprime, n_splits, &plan);
  if (rc != HYPRET_OK) return rc;
  if (plan.n_splits != n_splits) return HYPRET_EINVAL;   // caller sized cand_* for a different plan
  if ((reinterpret_cast<uintptr_t>(q_op) & 15) || (reinterpret_cast<uintptr_t>(g_op) & 15)) return HYPRET_EINVAL;

  const int dpad = hypret_dpad(d);
  const int kpad = dpad + HYPRET_KEXT;
  CUtensorMap mq_main, mq_ext, mg_main, mg_ext;
  if ((rc = make_map(&mq_main, q_op, Q, kpad, HYPRET_KBLK, TILE_M, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mq_ext, q_op, Q, kpad, HYPRET_KEXT, TILE_M, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map(&mg_main, g_op, N, kpad, HYPRET_KBLK, TILE_N, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mg_ext, g_op, N, kpad, HYPRET_KEXT, TILE_N, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;

  Params p;
  p.Q = Q;
  p.N = N;
  p.kb_main = dpad / HYPRET_KBLK;
  p.has_ext = 1;
  p.dpad = dpad;
  p.n_qtiles = plan.n_qtiles;
  p.n_gtiles = plan.n_gtiles;
  p.n_splits = plan.n_splits;
  p.tiles_per_split = plan.tiles_per_split;
  p.n_items = plan.n_qtiles * plan.n_splits;
  p.kprime = kprime;
  p.stages = plan.stages;
  p.stage_bytes = plan.resident ? B_BLK_BYTES : A_BLK_BYTES + B_BLK_BYTES;
  p.ring_off = plan.resident ? p.kb_main * A_BLK_BYTES + A_EXT_BYTES : 0;
  p.lists_off = p.ring_off + plan.stages * p.stage_bytes;
  p.bar_off = p.lists_off + kprime * TILE_M * 8;
  p.cand_score = cand_score;
  p.cand_idx = cand_idx;
  p.debug_scores = debug_scores;

  cudaError_t e;
  if (plan.resident) {
    e = cudaFuncSetAttribute(score_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    score_topk_kernel<true><<<plan.grid, NUM_THREADS, plan.smem_bytes, stream>>>(mq_main, mq_ext, mg_main, mg_ext, p);
  } else {
    e = cudaFuncSetAttribute(score_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    score_topk_kernel<false><<<plan.grid, NUM_THREADS, plan.smem_bytes, stream>>>(mq_main, mq_ext, mg_main, mg_ext, p);
  }
  return (int)cudaGetLastError();
}
