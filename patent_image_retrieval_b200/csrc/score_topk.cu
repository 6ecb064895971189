// Scoring GEMM + fused streaming top-k' for sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces the reference's one-vs-all scoring loops and the ranking that follows them
//   pmath.dist(q[1,D], G[P,D])            /root/reference/src/train.py:3259
//   cosine_similarity(Q, G)               /root/reference/notebooks/retrieval.ipynb:368
//   np.argsort / torch.topk               retrieval.ipynb:383,202 ; src/auxiliary.py:374
// as a candidate filter: S[i,j] = <q_op[i,:], g_op[j,:]> is the ranking surrogate built by
// project.cu (c * rb_j * ||x_i - y_j||^2 for the Poincare ball, -cos for cosine); per query the
// kernel keeps the k' smallest S of every gallery strip it visits.  S lives only in TMEM and
// registers.
//
// Structure (one persistent CTA per SM, warp-specialised; CTAs run as PAIRS -- tcgen05 cta_group::2 -- whenever there
// are two query tiles to pair up; 320 threads for k' <= 16, 192 otherwise):
//   warp 0   TMA producer   gallery K-blocks (each CTA of a pair its half: 128 rows x 64 fp16, 128B swizzle) into a
//                           shared-memory ring; the CTA's 128-row query tile is loaded ONCE per strip and stays
//                           resident in shared memory when D <= 512 (RESIDENT), otherwise it streams through the
//                           ring next to the gallery block
//   warp 1   MMA issuer     one thread of the leader CTA issues tcgen05.mma (M=256 across the pair, N=256, K=16,
//                           fp16 -> fp32) into one of two 256-column TMEM accumulators; tcgen05.commit releases ring
//                           stages and publishes finished accumulators (multicast to both CTAs)
//   warps 2+ epilogue       tcgen05.ld 32x32b: lane t of a warp owns query row t of its TMEM quadrant, so the
//                           running k'-th best score of a query is a private register.  Fast path: 3-input min tree
//                           + one compare per 32 columns.  Hits go to a per-lane pending queue in shared memory and
//                           the warp drains all queues together.
//                           k' <= 16 (the search path): TWO epilogue warpgroups, each on its half of every tile,
//                           the candidate list of a row in the owning thread's REGISTERS (sorted, branch-free
//                           insert); a cold strip's first 64 columns go through a sorting network instead
//                           (sort16_pairs / reg_merge16).  k' > 16: one warpgroup, lists in shared memory.
//
// Strip schedule.  Work unit = (query tile, contiguous range of gallery tiles) = "strip"; a
// strip owns one candidate list per query row, so every strip start is a cold threshold and a
// cold list.  With T query tiles, G gallery tiles and P CTAs:
//   * full waves   while >= P query tiles remain, every CTA takes one whole row (strip = G);
//   * phase 1      the T' < P remaining rows get a = P / T' strips each (the first b = P % T'
//                  rows one more), all of length L1 = ceil(G/(a+1)) -- exactly P strips;
//   * phase 2      rows with only `a` strips still miss [a*L1, G); that remainder is cut into
//                  m = P / (T'-b) pieces per row so that again (almost) every CTA is busy.
// CTAs with consecutive ids run the same gallery range at the same time (lock-step through
// L2: the gallery is read from HBM about once per phase instead of once per query tile), the
// load balance is within one small piece, and a query sees only a handful of cold starts.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

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
constexpr int MIN_STRIP_TILES = 4;                        // do not spread tiny problems over all SMs (C1, measured with
                                                          // the dense cold start: floor 4 / 3 / 2 -> scoring 83 / 77 / 76 us
                                                          // but rerank 74 / 78 / 79 us for the extra lists: a wash)
constexpr unsigned FULL = 0xffffffffu;

struct Sched {
  int T, G, P;
  int n_full, tail_rows, a, b, L1, rem_rows, rem_g0, m, L2, n_steps, n_lists;
  // wide top-k: base strips are cut into consecutive sub-strips with their own list slots -- `sub` pieces for
  // full-wave strips (one strip per row otherwise), `sub_tail` pieces for phase-1 / phase-2 strips
  int sub, sub_tail;
};

__host__ __device__ inline bool strip_at_base(const Sched& s, int cta, int step, int& qt, int& g0, int& g1, int& slot) {
  if (step < s.n_full) {
    qt = step * s.P + cta;
    g0 = 0;
    g1 = s.G;
    slot = 0;
    return true;
  }
  step -= s.n_full;
  if (s.tail_rows == 0) return false;
  const int row0 = s.n_full * s.P;
  if (step == 0) {
    int j, row;
    if (cta < s.a * s.tail_rows) {
      j = cta / s.tail_rows;
      row = cta - j * s.tail_rows;
    } else {
      j = s.a;
      row = cta - s.a * s.tail_rows;
      if (row >= s.b) return false;
    }
    qt = row0 + row;
    g0 = j * s.L1;
    g1 = g0 + s.L1 < s.G ? g0 + s.L1 : s.G;
    slot = j;
    return g0 < g1;
  }
  if (step == 1 && s.rem_rows > 0) {
    if (cta >= s.m * s.rem_rows) return false;
    const int e = cta / s.rem_rows;
    const int row = s.b + (cta - e * s.rem_rows);
    qt = row0 + row;
    g0 = s.rem_g0 + e * s.L2;
    g1 = g0 + s.L2 < s.G ? g0 + s.L2 : s.G;
    slot = s.a + e;
    return g0 < g1;
  }
  return false;
}

__host__ __device__ inline bool strip_at_base(const Sched& s, int cta, int step, int& qt, int& g0, int& g1, int& slot);

// Public schedule = base schedule with every strip cut into s.sub sub-strips.  s.n_steps / s.n_lists
// are the PUBLIC counts (base counts x sub).
__host__ __device__ inline bool strip_at(const Sched& s, int cta, int step, int& qt, int& g0, int& g1, int& slot) {
  if (s.sub <= 1 && s.sub_tail <= 1) return strip_at_base(s, cta, step, qt, g0, g1, slot);
  int base, j, sub;
  const int full_steps = s.n_full * s.sub;
  if (step < full_steps) {
    sub = s.sub;
    base = step / sub;
    j = step - base * sub;
  } else {
    sub = s.sub_tail;
    const int t = step - full_steps;
    base = s.n_full + t / sub;
    j = t - (t / sub) * sub;
  }
  if (!strip_at_base(s, cta, base, qt, g0, g1, slot)) return false;
  const int piece = (g1 - g0 + sub - 1) / sub;
  g0 += j * piece;
  g1 = g0 + piece < g1 ? g0 + piece : g1;
  slot = slot * sub + j;
  return g0 < g1;
}

Sched make_sched(int64_t T, int64_t G, int sms, int max_ctas) {
  Sched s{};
  int64_t P = sms;
  if (max_ctas > 0 && max_ctas < P) P = max_ctas;
  int64_t cap = (T * G) / MIN_STRIP_TILES;
  if (cap < 1) cap = 1;
  if (cap < P) P = cap;
  s.T = (int)T;
  s.G = (int)G;
  s.P = (int)P;
  s.n_full = (int)(T / P);
  s.tail_rows = (int)(T % P);
  s.n_lists = 1;
  s.n_steps = s.n_full;
  s.sub = 1;
  s.sub_tail = 1;
  if (s.tail_rows > 0) {
    s.a = s.P / s.tail_rows;
    s.b = s.P % s.tail_rows;
    if (s.b == 0) {
      s.L1 = (int)((G + s.a - 1) / s.a);
      s.n_lists = s.a;
    } else {
      s.L1 = (int)((G + s.a) / (s.a + 1));
      s.n_lists = s.a + 1;
      s.rem_g0 = s.a * s.L1;
      if (s.rem_g0 < G) {
        s.rem_rows = s.tail_rows - s.b;
        s.m = s.P / s.rem_rows;
        if (s.m < 1) s.m = 1;
        const int rem = (int)G - s.rem_g0;
        if (s.m > rem) s.m = rem;
        s.L2 = (rem + s.m - 1) / s.m;
        if (s.a + s.m > s.n_lists) s.n_lists = s.a + s.m;
      }
    }
    s.n_steps += 1 + (s.rem_rows > 0 ? 1 : 0);
  }
  return s;
}

struct Params {
  int64_t Q;
  int64_t N;
  int kb_main;          // number of 64-wide K blocks (Dpad / 64)
  int has_ext;          // 1: extension K block present (hyperbolic surrogate constants)
  int dpad;
  int kprime;
  int kbound;          // rank of the score bound the query's lists share (>= kprime; > kprime only with register lists)
  int stages;
  int stage_bytes;
  int ring_off;         // byte offsets from the 1024-aligned shared-memory base
  int lists_off;
  int bar_off;
  Sched sched;
  float* cand_score;
  int32_t* cand_idx;
  int32_t* list_count;    // [Q] or NULL.  Given: a strip takes the next free list slot of each query (atomicAdd) and
                          // empty lists take none, so a query's lists are its first list_count[q] slots
  float* debug_scores;
  uint32_t* shared_thr;   // [Q] ordered-uint keys of the best known k'-th score per query, or NULL
  unsigned long long* stats;   // DEBUG builds: [grid][8] wait-cycle counters per warp role, or NULL
  int wait_mode;          // experiments: 0 = try_wait + suspend hint, 1 = plain try_wait, 2 = test_wait spin
  int pf_tiles;           // producer: L2 prefetch distance in gallery tiles (0 = off)
  int wg_off;             // REGLIST: byte offset of the warpgroup-exchange words [2][128] x 8 B
};

// mbarrier wait with optional cycle accounting (DEBUG kernels only)
template <bool DEBUG>
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, unsigned long long& acc_cycles,
                                           int wait_mode) {
  if (DEBUG) {
    const long long t0 = clock64();
    if (wait_mode == 2) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      }
    } else if (wait_mode == 1) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      }
    } else {
      mbar_wait(bar, parity);
    }
    acc_cycles += (unsigned long long)(clock64() - t0);
  } else {
    mbar_wait(bar, parity);
  }
}

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

// monotone float <-> uint32 map so that redux.max on the keys is a float max
__device__ __forceinline__ uint32_t f2key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");   // L2 (coherent), not L1
  return v;
}
constexpr uint32_t KEY_INF = 0xff800000u;   // f2key(+inf)

// smallest float above a FINITE x (the libm next-after toward +inf without its special-case branches: the
// threshold exchange runs it twice per tile and row)
__device__ __forceinline__ float next_up(float x) {
  const int b = __float_as_int(x);
  return x >= 0.0f ? __int_as_float((b & 0x7fffffff) + 1) : __int_as_float(b - 1);
}

// Lists live in shared memory as [slot][row] (row = thread of the epilogue, 0..127): when all
// lanes of a warp touch the same slot the access is conflict-free.  Shared-state-space
// addresses (32-bit) + explicit ld/st.shared keep the non-inlined helper free of generic loads.
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t a, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
constexpr int LIST_SLOT_STRIDE = TILE_M * 4;   // bytes between consecutive slots of one row

// Thread-private insert: overwrite the row's worst slot, rescan the KPP slots (independent
// loads, compile-time unrolled) and return (new worst score, its slot).  One call site per
// column of a chunk, so it is deliberately NOT inlined.
template <int KPP>
__device__ __noinline__ uint2 list_insert(uint32_t s_addr, uint32_t i_addr, float x, int col, int maxpos) {
  sts_f32(s_addr + maxpos * LIST_SLOT_STRIDE, x);
  sts_b32(i_addr + maxpos * LIST_SLOT_STRIDE, col);
  float w[KPP];
#pragma unroll
  for (int s = 0; s < KPP; ++s) w[s] = lds_f32(s_addr + s * LIST_SLOT_STRIDE);
  float mx = w[0];
#pragma unroll
  for (int s = 1; s < KPP; ++s) mx = fmaxf(mx, w[s]);
  int pos = 0;
#pragma unroll
  for (int s = KPP - 1; s >= 0; --s) pos = (w[s] == mx) ? s : pos;
  return make_uint2(__float_as_uint(mx), (uint32_t)pos);
}

// Pending-hit queue.  A hit is usually confined to one or two lanes of the warp, so inserting it
// on the spot costs a whole SIMT pass (one list rescan) per hit -- ~2400 passes per strip, almost
// all of them in the cold phase after a strip start, 0.29 ms per launch.  Instead each lane
// appends its hits to a small private queue ([slot][row] in shared memory, two predicated
// stores) and the warp drains all queues together: one rescan pass then serves up to 32 lanes'
// inserts.  The threshold a lane filters with is refreshed at every drain, so it is at most a
// few hits stale -- that only lets a few extra candidates into the queue, never drops one.
constexpr int QCAP = 12;   // queue slots per lane (12 x 8 B x 128 rows = 12 KB)
constexpr int QCAP_R = 12; // ... of the register-list epilogue, per warpgroup.  (4 slots would buy a fifth ring stage:
                           // measured 2.09 vs 2.08 ms at C2 and 2.73 vs 2.51 ms on 37.5k-row shards -- the overflow
                           // path and the poorer drain batching cost more than the stage gains)

struct RowState {
  float thr;        // effective filter threshold = min(thr_list, thr_g)
  float thr_list;   // worst score currently kept (+inf until the list is full)
  float thr_g;      // bound shared by the query's other strips
  int maxpos;       // slot of the worst kept score
  int cnt;          // pending hits in the queue
  int n_ins;        // DEBUG statistics: list inserts done by this lane
  int n_drain;      // DEBUG statistics: drain-loop passes of the warp
  int n_now;        // DEBUG statistics: inserts that bypassed the (full) queue
};

template <int KPP>
__device__ __forceinline__ void insert_now(RowState& st, float x, int col, uint32_t s_addr, uint32_t i_addr) {
  const uint2 r = list_insert<KPP>(s_addr, i_addr, x, col, st.maxpos);
  st.n_ins += 1;
  st.thr_list = __uint_as_float(r.x);
  st.maxpos = (int)r.y;
  st.thr = fminf(st.thr_list, st.thr_g);
}

// warp-uniform: every lane pops its own queue; lanes that run dry idle until the longest is empty
template <int KPP>
__device__ __forceinline__ void drain_queue(RowState& st, uint32_t s_addr, uint32_t i_addr, uint32_t qs_addr,
                                            uint32_t qi_addr) {
  while (__any_sync(FULL, st.cnt > 0)) {
    st.n_drain += 1;
    if (st.cnt > 0) {
      st.cnt -= 1;
      const float x = lds_f32(qs_addr + st.cnt * LIST_SLOT_STRIDE);
      if (x < st.thr) {   // the threshold may have tightened since the hit was queued
        int col;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(col) : "r"(qi_addr + st.cnt * LIST_SLOT_STRIDE) : "memory");
        insert_now<KPP>(st, x, col, s_addr, i_addr);
      }
    }
  }
}

// One 32-column chunk of the accumulator.  Every lane owns one query row.  Fast path: 3-input min
// tree + one compare.  Slow path: plain SIMT divergence -- only lanes with a hit walk their group
// and queue the hit (or insert it at once when the queue is full, which only happens in the
// first tile of a cold strip).
template <int KPP>
__device__ __forceinline__ void process_chunk(const float (&v)[32], int cbase, RowState& st, uint32_t s_addr,
                                              uint32_t i_addr, uint32_t qs_addr, uint32_t qi_addr) {
  float mg[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float m = v[g * 8];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fminf(m, v[g * 8 + j]);
    mg[g] = m;
  }
  const float m = fminf(fminf(mg[0], mg[1]), fminf(mg[2], mg[3]));
  if (m < st.thr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (mg[g] < st.thr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[g * 8 + j];
          if (x < st.thr) {
            if (st.cnt < QCAP) {
              sts_f32(qs_addr + st.cnt * LIST_SLOT_STRIDE, x);
              sts_b32(qi_addr + st.cnt * LIST_SLOT_STRIDE, cbase + g * 8 + j);
              st.cnt += 1;
            } else {
              st.n_now += 1;
              insert_now<KPP>(st, x, cbase + g * 8 + j, s_addr, i_addr);
            }
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- k' <= 16: lists in registers
// For k' <= 16 (top-10 retrieval, the headline configuration) a row's candidate list lives in the owning
// thread's REGISTERS as 16 (score, index) pairs sorted ascending.  An insert is a branch-free
// compare-and-shift over the 16 slots (each slot's new value depends only on old values: 16-way ILP, no
// shared-memory round trip, no rescan for the new worst entry -- it is simply the last slot), and the
// 16 KB of shared memory the lists used to occupy pays for the pending-hit queue of a SECOND epilogue
// warpgroup: warpgroup g scans columns [128 g, 128 g + 128) of every gallery tile, so each SM sub-partition
// has two epilogue warps to hide each other's latencies and a tile leaves the accumulator in half the time.
// A row therefore owns two lists per strip (slot 2*s+g); the two halves exchange their k'-th best through
// the same L2 threshold word the strips of a query already share, and a tighter pair bound through shared memory.
constexpr int RL = 16;
constexpr int WG_X_BYTES = 2 * TILE_M * 8;   // warpgroup exchange: [2][128] x {strip tag, score bits}
constexpr int OVF = 64;      // per-thread overflow of the pending queue (local memory; cold phase only)

// The list is a struct of 32 NAMED scalars, not two arrays: with arrays the compiler front end left the scores
// in a local-memory depot (the L1 that would back it is almost entirely carved out as shared memory here).
#define RL_FOR_EACH(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15)
#define RL_FOR_EACH_DOWN(M)                                                                                        \
  M(15, 14) M(14, 13) M(13, 12) M(12, 11) M(11, 10) M(10, 9) M(9, 8) M(8, 7) M(7, 6) M(6, 5) M(5, 4) M(4, 3) M(3, 2) \
      M(2, 1) M(1, 0)
struct RegList {
#define RL_DECL(t) float s##t; int i##t;
  RL_FOR_EACH(RL_DECL)
#undef RL_DECL
};

__device__ __forceinline__ void reg_init(RegList& L) {
#define RL_INIT(t) L.s##t = INFINITY; L.i##t = -1;
  RL_FOR_EACH(RL_INIT)
#undef RL_INIT
}

__device__ __forceinline__ void reg_insert(RegList& L, float x, int col) {
#define RL_STEP(t, u)                                    \
  {                                                      \
    const bool up = x < L.s##u; /* better neighbour moves down one slot */ \
    const bool in = x < L.s##t; /* ... or x lands here */ \
    L.s##t = up ? L.s##u : (in ? x : L.s##t);            \
    L.i##t = up ? L.i##u : (in ? col : L.i##t);          \
  }
  RL_FOR_EACH_DOWN(RL_STEP)
#undef RL_STEP
  const bool in0 = x < L.s0;
  L.s0 = in0 ? x : L.s0;
  L.i0 = in0 ? col : L.i0;
}

__device__ __forceinline__ float reg_kth(const RegList& L, int kp) {   // kp-th best (1-based), kp <= RL
  float v = L.s15;
#define RL_KTH(t) v = (kp - 1 == t) ? L.s##t : v;
  RL_FOR_EACH(RL_KTH)
#undef RL_KTH
  return v;
}

__device__ __forceinline__ void reg_publish(const RegList& L, int kp, float* cs, int32_t* ci) {
#define RL_PUB(t)        \
  if (t < kp) {          \
    cs[t] = L.s##t;      \
    ci[t] = L.i##t;      \
  }
  RL_FOR_EACH(RL_PUB)
#undef RL_PUB
}

// Cold strip start, dense path.  A list that starts with no bound (nothing published for its query yet) lets every
// column through: streaming inserts cost one ~100-instruction drain pass per hit, the passes of a warp are as many as
// its unluckiest lane has hits, and the first 52 hits of a lane spill from the 12-slot queue into local memory --
// ~90 passes and ~11k warp-instructions for the first tile of a strip, most of the epilogue's time on short strips
// (C1: 4-tile strips; 37.5k-row shards).  The first DENSE_COLS columns of such a strip therefore skip the filter:
// 16 columns at a time are sorted by a 63-exchange odd-even merge network (Batcher) on (score, column) pairs and
// merged into the sorted register list by one half-cleaner + a 4-stage bitonic merge -- branch-free, the same
// ~520 instructions for every lane, no queue, no local memory.  After 64 columns the list's own threshold sits at
// the 25 % quantile and the streaming path (a pass per hit) is the cheaper one again.
constexpr int DENSE_COLS = 64;
static_assert(DENSE_COLS % 64 == 0 && DENSE_COLS <= TILE_N / 2, "dense prefix: whole chunk pairs of a warpgroup's half tile");

__device__ __forceinline__ void tmem_ld_32x16_sync(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// compare-exchange on (score, column) pairs: afterwards a <= b
__device__ __forceinline__ void pair_ce(float& as, int& ai, float& bs, int& bi) {
  const bool sw = bs < as;
  const float s0 = sw ? bs : as, s1 = sw ? as : bs;
  const int i0 = sw ? bi : ai, i1 = sw ? ai : bi;
  as = s0; ai = i0; bs = s1; bi = i1;
}

// Batcher's odd-even merge sort of 16 pairs, ascending (63 exchanges; the pairs are compile-time constants, so the
// arrays stay in registers)
__device__ __forceinline__ void sort16_pairs(float (&s)[16], int (&i)[16]) {
  constexpr int A[63] = {0, 2, 4, 6, 8, 10, 12, 14, 0, 1, 4, 5, 8, 9, 12, 13, 1, 5, 9, 13, 0,
                         1, 2, 3, 8, 9, 10, 11, 2, 3, 10, 11, 1, 3, 5, 9, 11, 13, 0, 1, 2, 3,
                         4, 5, 6, 7, 4, 5, 6, 7, 2, 3, 6, 7, 10, 11, 1, 3, 5, 7, 9, 11, 13};
  constexpr int B[63] = {1, 3, 5, 7, 9, 11, 13, 15, 2, 3, 6, 7, 10, 11, 14, 15, 2, 6, 10, 14, 4,
                         5, 6, 7, 12, 13, 14, 15, 4, 5, 12, 13, 2, 4, 6, 10, 12, 14, 8, 9, 10, 11,
                         12, 13, 14, 15, 8, 9, 10, 11, 4, 5, 8, 9, 12, 13, 2, 4, 6, 8, 10, 12, 14};
#pragma unroll
  for (int c = 0; c < 63; ++c) pair_ce(s[A[c]], i[A[c]], s[B[c]], i[B[c]]);
}

// L <- the 16 smallest of L (sorted) and h (sorted): min(L[t], h[15-t]) is those 16 as a bitonic sequence, four
// half-cleaner stages sort it.  Ties keep the list's entry, like reg_insert.
__device__ __forceinline__ void reg_merge16(RegList& L, const float (&hs)[16], const int (&hi)[16]) {
  float ls[16];
  int li[16];
#define RL_GET(t) ls[t] = L.s##t; li[t] = L.i##t;
  RL_FOR_EACH(RL_GET)
#undef RL_GET
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const bool tk = hs[15 - t] < ls[t];
    ls[t] = tk ? hs[15 - t] : ls[t];
    li[t] = tk ? hi[15 - t] : li[t];
  }
#pragma unroll
  for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      if ((a ^ j) > a) pair_ce(ls[a], li[a], ls[a ^ j], li[a ^ j]);
    }
  }
#define RL_PUT(t) L.s##t = ls[t]; L.i##t = li[t];
  RL_FOR_EACH(RL_PUT)
#undef RL_PUT
}

struct RowStateR {
  float thr;        // effective filter threshold = min(thr_list, thr_g)
  float thr_list;   // k'-th best score currently kept (+inf until k' scores are in)
  float thr_g;      // bound shared by the query's other lists
  int cnt;          // pending hits (queue + overflow)
  int n_ins, n_drain, n_now;   // DEBUG statistics
};

// warp-uniform and branch-free per pass: every lane pops one pending hit (+inf when it has none) and runs
// the register insert; lanes that run dry idle until the longest queue is empty
// KFULL: kp == RL (the default k' = 16) -- the list's threshold is simply its last slot, not a 16-way select
template <bool KFULL>
__device__ __forceinline__ void drain_queue_reg(RowStateR& st, RegList& L, int kp, uint32_t qs_addr, uint32_t qi_addr,
                                                const float* ovf_s, const int* ovf_i) {
  while (__any_sync(FULL, st.cnt > 0)) {
    st.n_drain += 1;
    float x = INFINITY;
    int col = -1;
    if (st.cnt > 0) {
      st.cnt -= 1;
      if (st.cnt >= QCAP_R) {
        x = ovf_s[st.cnt - QCAP_R];
        col = ovf_i[st.cnt - QCAP_R];
      } else {
        x = lds_f32(qs_addr + st.cnt * LIST_SLOT_STRIDE);
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(col) : "r"(qi_addr + st.cnt * LIST_SLOT_STRIDE) : "memory");
      }
      st.n_ins += (x < st.thr) ? 1 : 0;
    }
    x = (x < st.thr) ? x : INFINITY;          // the threshold may have tightened since the hit was queued
    reg_insert(L, x, col);
    st.thr_list = KFULL ? L.s15 : reg_kth(L, kp);
    st.thr = fminf(st.thr_list, st.thr_g);
  }
}

// One 32-column chunk, lists in registers: same fast path; hits go to the pending queue (shared memory) or,
// when that is full (first tiles of a cold list only), to the thread's overflow array.
__device__ __forceinline__ void process_chunk_reg(const float (&v)[32], int cbase, RowStateR& st, uint32_t qs_addr,
                                                  uint32_t qi_addr, float* ovf_s, int* ovf_i) {
  float mg[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float m = v[g * 8];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fminf(m, v[g * 8 + j]);
    mg[g] = m;
  }
  const float m = fminf(fminf(mg[0], mg[1]), fminf(mg[2], mg[3]));
  if (m < st.thr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (mg[g] < st.thr) {
        if (st.cnt <= QCAP_R - 8) {
          // room for the whole group: no capacity check per column -- a compare, two predicated stores and a
          // predicated address step per column (5 instructions against 13 with the overflow alternative; the
          // walk is ~30 % of the epilogue's stall samples on short strips, profiles/r2zb_score_c1_ncu.txt)
          uint32_t a = qs_addr + st.cnt * LIST_SLOT_STRIDE;
          const uint32_t i_off = qi_addr - qs_addr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = v[g * 8 + j];
            const bool h = x < st.thr;
            if (h) {
              sts_f32(a, x);
              sts_b32(a + i_off, cbase + g * 8 + j);
            }
            a += h ? LIST_SLOT_STRIDE : 0;
          }
          st.cnt = (int)((a - qs_addr) / LIST_SLOT_STRIDE);
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[g * 8 + j];
          if (x < st.thr) {
            if (st.cnt < QCAP_R) {
              sts_f32(qs_addr + st.cnt * LIST_SLOT_STRIDE, x);
              sts_b32(qi_addr + st.cnt * LIST_SLOT_STRIDE, cbase + g * 8 + j);
            } else {
              st.n_now += 1;
              HYPRET_CHECK(st.cnt - QCAP_R < OVF);
              ovf_s[st.cnt - QCAP_R] = x;
              ovf_i[st.cnt - QCAP_R] = cbase + g * 8 + j;
            }
            st.cnt += 1;
          }
        }
      }
    }
  }
}

// PAIR: two CTAs of a cluster (one TPC) share every gallery tile through tcgen05 cta_group::2 --
// each CTA keeps its OWN 128-row query tile and loads only HALF (128 rows) of the 256-row gallery
// tile; the leader CTA issues one M=256 MMA for both.  Halves the L2->shared-memory traffic per
// SM (the measured limiter of the single-CTA kernel at D=512) and leaves room for a resident
// query tile plus a 4-stage ring.
template <bool RESIDENT, int KPP, bool DEBUG, bool PAIR>
__global__ void __launch_bounds__(KPP == RL ? NUM_THREADS + NUM_EPI_THREADS : NUM_THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap map_q_main, const __grid_constant__ CUtensorMap map_q_ext,
                  const __grid_constant__ CUtensorMap map_g_main, const __grid_constant__ CUtensorMap map_g_ext,
                  const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte aligned bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_res = smem;                                   // RESIDENT: kb_main blocks + ext block
  uint8_t* ring = smem + p.ring_off;
  constexpr bool REGLIST = KPP == RL;                      // k' <= 16: lists in registers, two epilogue warpgroups
  float* list_s = reinterpret_cast<float*>(smem + p.lists_off);   // [KPP slots][128 rows]   (unused when REGLIST)
  int* list_i = reinterpret_cast<int*>(list_s + KPP * TILE_M);
  float* queue_s = REGLIST ? list_s : reinterpret_cast<float*>(list_i + KPP * TILE_M);   // [QCAP slots][128 rows]
  int* queue_i = reinterpret_cast<int*>(queue_s + (REGLIST ? QCAP_R : QCAP) * TILE_M);   // REGLIST: one block per warpgroup
  Barriers* bars = reinterpret_cast<Barriers*>(smem + p.bar_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.kb_main;
  const int KSTEPS = KB + (p.has_ext ? 1 : 0);
  const Sched& sc = p.sched;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs)
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // scheduling unit: CTA or CTA pair
  constexpr int B_ROWS = PAIR ? TILE_N / 2 : TILE_N;            // gallery rows this CTA stages per tile
  constexpr uint32_t B_BLK = B_ROWS * HYPRET_KBLK * 2, B_EXT = B_ROWS * HYPRET_KEXT * 2;

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
      // every epilogue thread that reads the accumulator arrives: one warpgroup per CTA, or both (k' <= 16)
      mbar_init(&bars->tmem_empty[a], (PAIR ? 2 : 1) * (KPP == RL ? 2 : 1) * NUM_EPI_THREADS);
    }
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2cta(&bars->tmem_ptr, TMEM_COLS);
    else tmem_alloc(&bars->tmem_ptr, TMEM_COLS);
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers must exist before any remote arrive / TMA credit
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;

  if (warp == 0) {
    // ======================================================================= TMA producer
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues the copies.
    uint32_t stage = 0, phase = 0, a_par = 0;
    bool first = true;
    unsigned long long w_empty = 0, t_begin = DEBUG ? clock64() : 0;
    const uint32_t a_tile_bytes = KB * A_BLK_BYTES + (p.has_ext ? A_EXT_BYTES : 0);
    for (int step = 0; step < sc.n_steps; ++step) {
      int qt, gt0, gt1, slot;
      if (!strip_at(sc, cta, step, qt, gt0, gt1, slot)) continue;
      const int q_row = (PAIR ? 2 * qt + (int)rank : qt) * TILE_M;   // this CTA's own query tile
      // In PAIR mode both CTAs issue copies into their own shared memory, all transaction bytes are
      // credited to the LEADER's barriers, and only the leader posts the expected byte count.
      constexpr uint32_t NCTA = PAIR ? 2u : 1u;
      if (RESIDENT) {
        if (!first) {  // the previous strip's MMAs must be done reading the resident tile
          mbar_wait(&bars->a_empty, a_par);
          a_par ^= 1;
        }
        first = false;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&bars->a_full, a_tile_bytes * NCTA);
          for (int kb = 0; kb < KB; ++kb) {
            if (PAIR) tma_load_2d_pair(a_res + kb * A_BLK_BYTES, &map_q_main, &bars->a_full, kb * HYPRET_KBLK, q_row,
                                       TMA_EVICT_LAST);
            else tma_load_2d_hint(a_res + kb * A_BLK_BYTES, &map_q_main, &bars->a_full, kb * HYPRET_KBLK, q_row,
                                  TMA_EVICT_LAST);
          }
          if (p.has_ext) {
            if (PAIR) tma_load_2d_pair(a_res + KB * A_BLK_BYTES, &map_q_ext, &bars->a_full, p.dpad, q_row,
                                       TMA_EVICT_LAST);
            else tma_load_2d_hint(a_res + KB * A_BLK_BYTES, &map_q_ext, &bars->a_full, p.dpad, q_row, TMA_EVICT_LAST);
          }
        }
        __syncwarp();
      }
      for (int gt = gt0; gt < gt1; ++gt) {
        const int g_row = gt * TILE_N + (int)rank * B_ROWS;          // this CTA's half of the gallery tile
        for (int ks = 0; ks < KSTEPS; ++ks) {
          timed_wait<DEBUG>(&bars->empty[stage], phase ^ 1, w_empty, p.wait_mode);
          if (elect_one()) {
            uint8_t* st = ring + stage * p.stage_bytes;
            uint8_t* st_b = RESIDENT ? st : st + A_BLK_BYTES;
            uint64_t* full = &bars->full[stage];
            // pull the same K-block of the tile PF_TILES ahead into L2: the ring only looks half a tile ahead
            // (4 x 16 KB), less than an HBM round trip when the whole lock-step group misses together
            if (p.pf_tiles > 0 && gt + p.pf_tiles < gt1) {
              const int pf_row = (gt + p.pf_tiles) * TILE_N + (int)rank * B_ROWS;
              if (ks < KB) tma_prefetch_2d(&map_g_main, ks * HYPRET_KBLK, pf_row);
              else tma_prefetch_2d(&map_g_ext, p.dpad, pf_row);
            }
            if (ks < KB) {
              if (rank == 0) mbar_arrive_expect_tx(full, (RESIDENT ? B_BLK : A_BLK_BYTES + B_BLK) * NCTA);
              if (PAIR) {
                tma_load_2d_pair(st_b, &map_g_main, full, ks * HYPRET_KBLK, g_row, TMA_EVICT_NORMAL);
                if (!RESIDENT) tma_load_2d_pair(st, &map_q_main, full, ks * HYPRET_KBLK, q_row, TMA_EVICT_LAST);
              } else {
                tma_load_2d(st_b, &map_g_main, full, ks * HYPRET_KBLK, g_row);
                if (!RESIDENT) tma_load_2d_hint(st, &map_q_main, full, ks * HYPRET_KBLK, q_row, TMA_EVICT_LAST);
              }
            } else {
              if (rank == 0) mbar_arrive_expect_tx(full, (RESIDENT ? B_EXT : A_EXT_BYTES + B_EXT) * NCTA);
              if (PAIR) {
                tma_load_2d_pair(st_b, &map_g_ext, full, p.dpad, g_row, TMA_EVICT_NORMAL);
                if (!RESIDENT) tma_load_2d_pair(st, &map_q_ext, full, p.dpad, q_row, TMA_EVICT_LAST);
              } else {
                tma_load_2d(st_b, &map_g_ext, full, p.dpad, g_row);
                if (!RESIDENT) tma_load_2d_hint(st, &map_q_ext, full, p.dpad, q_row, TMA_EVICT_LAST);
              }
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (DEBUG && p.stats != nullptr && lane == 0) {
      p.stats[blockIdx.x * 8 + 0] = w_empty;
      p.stats[blockIdx.x * 8 + 1] = (unsigned long long)clock64() - t_begin;
    }
  } else if (warp == 1 && rank == 0) {
    // ======================================================================= MMA issuer
    // Warp-uniform control flow; one elected lane issues tcgen05.mma / tcgen05.commit.  The
    // shared-memory descriptors are a constant high word plus (address >> 4): stepping K by 16
    // elements inside the 128-byte swizzle span is "+2" on the low word.
    constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * TILE_M : TILE_M, TILE_N);
    constexpr uint64_t DESC_SW128 = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
                                    (UMMA_LAYOUT_SW128 << 61);
    constexpr uint64_t DESC_SW32 = (uint64_t(1) << 16) | (uint64_t(256 >> 4) << 32) | (uint64_t(1) << 46) |
                                   (UMMA_LAYOUT_SW32 << 61);
    const uint32_t ring_lo = smem_u32(ring) >> 4, ares_lo = smem_u32(a_res) >> 4;
    const uint32_t stage_lo = (uint32_t)p.stage_bytes >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_par = 0;
    unsigned long long w_full = 0, w_tmem = 0, t_begin = DEBUG ? clock64() : 0;
    for (int step = 0; step < sc.n_steps; ++step) {
      int qt, gt0, gt1, slot;
      if (!strip_at(sc, cta, step, qt, gt0, gt1, slot)) continue;
      if (RESIDENT) {
        mbar_wait(&bars->a_full, a_par);
        a_par ^= 1;
      }
      for (int gt = gt0; gt < gt1; ++gt) {
        timed_wait<DEBUG>(&bars->tmem_empty[acc], acc_phase ^ 1, w_tmem, p.wait_mode);   // epilogue drained it
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        for (int ks = 0; ks < KSTEPS; ++ks) {
          timed_wait<DEBUG>(&bars->full[stage], phase, w_full, p.wait_mode);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t st_lo = ring_lo + stage * stage_lo;
            const uint32_t b_lo = RESIDENT ? st_lo : st_lo + (A_BLK_BYTES >> 4);
            if (ks < KB) {
              const uint32_t a_lo = RESIDENT ? ares_lo + ks * (A_BLK_BYTES >> 4) : st_lo;
#pragma unroll
              for (int k = 0; k < HYPRET_KBLK / 16; ++k) {
                if (PAIR) umma_f16_ss_pair(d_tmem, DESC_SW128 | (a_lo + 2 * k), DESC_SW128 | (b_lo + 2 * k), idesc,
                                            (ks | k) != 0 ? 1u : 0u);
                else umma_f16_ss(d_tmem, DESC_SW128 | (a_lo + 2 * k), DESC_SW128 | (b_lo + 2 * k), idesc,
                                  (ks | k) != 0 ? 1u : 0u);
              }
            } else {
              const uint32_t a_lo = RESIDENT ? ares_lo + KB * (A_BLK_BYTES >> 4) : st_lo;
              if (PAIR) umma_f16_ss_pair(d_tmem, DESC_SW32 | a_lo, DESC_SW32 | b_lo, idesc, 1u);
              else umma_f16_ss(d_tmem, DESC_SW32 | a_lo, DESC_SW32 | b_lo, idesc, 1u);
            }
            // ring stage reusable once these MMAs retire; last K step: accumulator complete -> epilogue
            // (PAIR: the arrivals are multicast to the same barriers of both CTAs)
            if (PAIR) {
              umma_commit_pair(&bars->empty[stage]);
              if (ks == KSTEPS - 1) umma_commit_pair(&bars->tmem_full[acc]);
            } else {
              umma_commit(&bars->empty[stage]);
              if (ks == KSTEPS - 1) umma_commit(&bars->tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (RESIDENT) {
        if (elect_one()) {
          if (PAIR) umma_commit_pair(&bars->a_empty);
          else umma_commit(&bars->a_empty);
        }
        __syncwarp();
      }
    }
    if (DEBUG && p.stats != nullptr && lane == 0) {
      p.stats[blockIdx.x * 8 + 2] = w_full;
      p.stats[blockIdx.x * 8 + 3] = w_tmem;
      p.stats[blockIdx.x * 8 + 4] = (unsigned long long)clock64() - t_begin;
    }
  } else if (REGLIST && warp >= 2) {
    // ======================================================================= epilogue / top-k', lists in registers
    const int wg = (warp - 2) >> 2;            // epilogue warpgroup 0 / 1 <-> TMEM accumulator 0 / 1
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t qs_addr = smem_u32(queue_s + wg * 2 * QCAP_R * TILE_M + row);
    const uint32_t qi_addr = smem_u32(queue_i + wg * 2 * QCAP_R * TILE_M + row);
    float ovf_s[OVF];
    int ovf_i[OVF];
    const int KP = p.kprime;
    const int N = (int)p.N;
    // Both warpgroups work on EVERY tile, warpgroup g on columns [128 g, 128 g + 128): an accumulator is needed
    // again one MMA tile time after it was filled, so it is the epilogue's LATENCY per tile that must stay below
    // one tile time -- splitting tiles between the warpgroups (g takes accumulator g) doubled the throughput but
    // left the latency at 5.1 k cycles against 4.2 k, and the MMA issuer stalled 17 % of the time at C2.
    uint32_t acc = 0, acc_phase = 0;
    unsigned long long w_acc = 0, t_begin = DEBUG ? clock64() : 0;
    int tot_ins = 0, tot_drain = 0, tot_now = 0;
    for (int step = 0; step < sc.n_steps; ++step) {
      int qt, gt0, gt1, slot;
      if (!strip_at(sc, cta, step, qt, gt0, gt1, slot)) continue;
      if (PAIR) qt = 2 * qt + (int)rank;         // this CTA's own query tile of the pair
      const int64_t qrow = (int64_t)qt * TILE_M + row;
      uint32_t* gthr = (p.shared_thr != nullptr && qrow < p.Q) ? p.shared_thr + qrow : nullptr;
      RegList L;
      reg_init(L);
      // exchange word of this row: {strip tag, bits of this warpgroup's ceil(k'/2)-th best}.  The two lists of
      // a row hold disjoint columns, so max(half-th best of one, half-th best of the other) bounds the k'-th
      // best of their union -- about half the hits of filtering with min(k'-th best, k'-th best).
      const uint32_t my_x = smem_u32(smem + p.wg_off) + (uint32_t)(wg * TILE_M + row) * 8u;
      const uint32_t peer_x = smem_u32(smem + p.wg_off) + (uint32_t)((wg ^ 1) * TILE_M + row) * 8u;
      const int half_k = (p.kbound + 1) >> 1;       // <= RL: kbound <= 2 RL
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(my_x), "r"(step), "r"(0x7f800000u) : "memory");
      RowStateR st;
      st.thr_list = INFINITY;
      st.cnt = 0;
      st.n_ins = st.n_drain = st.n_now = 0;
      st.thr_g = INFINITY;
      float published = INFINITY;
      uint32_t gk_inflight = 0xffffffffu;
      if (gthr != nullptr) {
        const uint32_t gk = ld_cg_u32(gthr);
        if (gk < KEY_INF) st.thr_g = next_up(key2f(gk));
      }
      st.thr = fminf(st.thr_list, st.thr_g);
      // a row of this warp without any bound: the strip's first columns take the dense path (warp-uniform)
      bool dense = __any_sync(FULL, st.thr == INFINITY);
      for (int gt = gt0; gt < gt1; ++gt) {
        timed_wait<DEBUG>(&bars->tmem_full[acc], acc_phase, w_acc, p.wait_mode);
        tcgen05_fence_after();
        constexpr int HALF = TILE_N / 2;          // columns per warpgroup
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TILE_N + wg * HALF;
        const int col0 = gt * TILE_N + wg * HALF;
        const bool ragged = col0 + HALF > N;                 // last gallery tile: TMA zero-filled rows
        float va[32], vb[32];
        __syncwarp();
        int cc0 = 0;
        if (dense) {
          dense = false;
#pragma unroll 1
          for (int h = 0; h < DENSE_COLS / 16; ++h) {
            float hs[16];
            int hi[16];
            const int cb = col0 + h * 16;
            tmem_ld_32x16_sync(taddr + h * 16, hs);
            if (DEBUG && p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cb + j < N) p.debug_scores[qrow * p.N + cb + j] = hs[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const bool ok = hs[j] < INFINITY && cb + j < N;   // NaN / +inf / zero-filled rows never enter a list
              hs[j] = ok ? hs[j] : INFINITY;
              hi[j] = ok ? cb + j : -1;
            }
            sort16_pairs(hs, hi);
            reg_merge16(L, hs, hi);
            __syncwarp();
          }
          st.n_ins += DENSE_COLS;
          st.thr_list = reg_kth(L, KP);
          st.thr = fminf(st.thr_list, st.thr_g);
          cc0 = DENSE_COLS / 32;
        }
        if (cc0 < HALF / 32) tmem_ld_32x32(taddr + cc0 * 32, va);
#pragma unroll 1
        for (int cc = cc0; cc < HALF / 32; cc += 2) {
          tmem_ld_wait(va);
          tmem_ld_32x32(taddr + (cc + 1) * 32, vb);           // next chunk in flight while this one is scanned
          if (ragged) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + cc * 32 + j >= N) va[j] = INFINITY;
          }
          if (DEBUG && p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + cc * 32 + j < N) p.debug_scores[qrow * p.N + col0 + cc * 32 + j] = va[j];
          }
          process_chunk_reg(va, col0 + cc * 32, st, qs_addr, qi_addr, ovf_s, ovf_i);
          __syncwarp();                                       // tcgen05.ld/wait are warp-collective
          tmem_ld_wait(vb);
          if (cc + 2 < HALF / 32) tmem_ld_32x32(taddr + (cc + 2) * 32, va);
          if (ragged) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + (cc + 1) * 32 + j >= N) vb[j] = INFINITY;
          }
          if (DEBUG && p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + (cc + 1) * 32 + j < N) p.debug_scores[qrow * p.N + col0 + (cc + 1) * 32 + j] = vb[j];
          }
          process_chunk_reg(vb, col0 + (cc + 1) * 32, st, qs_addr, qi_addr, ovf_s, ovf_i);
          __syncwarp();
          // drain early when a queue is half full, so that thresholds do not go stale in the cold phase
          if (__any_sync(FULL, st.cnt >= QCAP_R / 2)) {
            if (KP == RL) drain_queue_reg<true>(st, L, KP, qs_addr, qi_addr, ovf_s, ovf_i);
            else drain_queue_reg<false>(st, L, KP, qs_addr, qi_addr, ovf_s, ovf_i);
          }
        }
        tcgen05_fence_before();
        if (PAIR) mbar_arrive_cluster(&bars->tmem_empty[acc], 0);   // the leader waits for both CTAs' epilogues
        else mbar_arrive(&bars->tmem_empty[acc]);
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
        // the accumulator is released: fold the pending hits in (off the MMA's critical path)
        // (measured, no gain: draining only once a queue holds 2-3 hits; exchanging thresholds every 4th / 8th tile
        // once a strip is 8 tiles old -- the steady-state epilogue is not what bounds the kernel)
        if (KP == RL) drain_queue_reg<true>(st, L, KP, qs_addr, qi_addr, ovf_s, ovf_i);
        else drain_queue_reg<false>(st, L, KP, qs_addr, qi_addr, ovf_s, ovf_i);
        // exchange thresholds with the query's other lists (other warpgroup, other strips) through L2
        if (gthr != nullptr) {
          // this list's k'-th best bounds the query's k'-th best -- but not its kbound-th when kbound > k' (lists of
          // 16 sharing the bound of the 24th best: the certificate's spare candidates at the cost of 16-slot lists)
          if (p.kbound <= KP && st.thr_list < published) {
            atomicMin(gthr, f2key(st.thr_list));
            published = st.thr_list;
          }
          if (gk_inflight < KEY_INF) st.thr_g = fminf(st.thr_g, next_up(key2f(gk_inflight)));
          gk_inflight = ld_cg_u32(gthr);
          // ... and with the other warpgroup of this CTA through shared memory
          float mine;
          if (half_k == 12) mine = L.s11;            // kbound = 24, the search path's default (warp-uniform branches)
          else if (half_k == 8) mine = L.s7;         // kbound = k' = 16
          else mine = reg_kth(L, half_k);
          uint32_t tag, bits;
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(my_x), "r"(step), "r"(__float_as_uint(mine)) : "memory");
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(tag), "=r"(bits) : "r"(peer_x) : "memory");
          if (tag == (uint32_t)step) {                 // the peer is on the same strip (same query rows)
            const float both = fmaxf(mine, __uint_as_float(bits));
            if (both < INFINITY) st.thr_g = fminf(st.thr_g, next_up(both));
            if (both < published) {                    // also a bound for the query's other strips
              atomicMin(gthr, f2key(both));
              published = both;
            }
          }
          st.thr = fminf(st.thr_list, st.thr_g);
        }
      }
      tot_ins += st.n_ins; tot_drain += st.n_drain; tot_now += st.n_now;
      // publish this warpgroup's list of the strip: the first k' entries of the sorted register list
      if (qrow < p.Q) {
        int pos = slot * 2 + wg;
        if (p.list_count != nullptr) pos = L.i0 < 0 ? -1 : atomicAdd(&p.list_count[qrow], 1);
        if (pos >= 0) {
          const int64_t o = (qrow * sc.n_lists + pos) * KP;
          HYPRET_CHECK(pos < sc.n_lists && qrow >= 0 && qrow < p.Q && KP <= RL);
          reg_publish(L, KP, p.cand_score + o, p.cand_idx + o);
        }
      }
    }
    if (DEBUG && p.stats != nullptr && warp == 2 && lane == 0) {
      p.stats[blockIdx.x * 8 + 5] = w_acc;
      p.stats[blockIdx.x * 8 + 6] = (unsigned long long)clock64() - t_begin;
      p.stats[blockIdx.x * 8 + 7] = ((unsigned long long)tot_ins << 40) | ((unsigned long long)tot_drain << 20) |
                                    (unsigned long long)tot_now;
    }
  } else if (warp >= 2) {
    // ======================================================================= epilogue / top-k'
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;          // query row inside the tile == TMEM lane
    float* my_s = list_s + row;                // [slot][row] layout, slot stride TILE_M
    int* my_i = list_i + row;
    const uint32_t s_addr = smem_u32(my_s), i_addr = smem_u32(my_i);
    const uint32_t qs_addr = smem_u32(queue_s + row), qi_addr = smem_u32(queue_i + row);
    const int KP = p.kprime;
    const int N = (int)p.N;
    uint32_t acc = 0, acc_phase = 0;
    unsigned long long w_acc = 0, t_begin = DEBUG ? clock64() : 0;
    int tot_ins = 0, tot_drain = 0, tot_now = 0;
    for (int step = 0; step < sc.n_steps; ++step) {
      int qt, gt0, gt1, slot;
      if (!strip_at(sc, cta, step, qt, gt0, gt1, slot)) continue;
      if (PAIR) qt = 2 * qt + (int)rank;         // this CTA's own query tile of the pair
      const int64_t qrow = (int64_t)qt * TILE_M + quad * 32 + lane;
      uint32_t* gthr = (p.shared_thr != nullptr && qrow < p.Q) ? p.shared_thr + qrow : nullptr;
      // cold list: active slots +inf (the list threshold stays +inf until k' scores are in),
      // inactive slots -inf (never the maximum)
      RowState st;
      st.thr_list = INFINITY;
      st.maxpos = 0;
      st.cnt = 0;
      st.n_ins = st.n_drain = st.n_now = 0;
#pragma unroll
      for (int s = 0; s < KPP; ++s) {
        my_s[s * TILE_M] = (s < KP) ? INFINITY : -INFINITY;
        my_i[s * TILE_M] = -1;
      }
      // warm threshold: what other strips of this query have already established.  Any bound
      // published there is >= the query's final k'-th best score, so filtering with the next
      // float above it can only drop rows that are not in the final top-k'.
      st.thr_g = INFINITY;
      float published = INFINITY;
      uint32_t gk_inflight = 0xffffffffu;      // bound loaded during the previous tile, consumed one tile later
      if (gthr != nullptr) {
        const uint32_t gk = ld_cg_u32(gthr);
        if (gk < KEY_INF) st.thr_g = next_up(key2f(gk));
      }
      st.thr = fminf(st.thr_list, st.thr_g);
      __syncwarp();
      for (int gt = gt0; gt < gt1; ++gt) {
        timed_wait<DEBUG>(&bars->tmem_full[acc], acc_phase, w_acc, p.wait_mode);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TILE_N;
        const int col0 = gt * TILE_N;
        const bool ragged = col0 + TILE_N > N;               // last gallery tile: TMA zero-filled rows
        float va[32], vb[32];
        __syncwarp();
        tmem_ld_32x32(taddr, va);
#pragma unroll 1
        for (int cc = 0; cc < TILE_N / 32; cc += 2) {
          tmem_ld_wait(va);
          tmem_ld_32x32(taddr + (cc + 1) * 32, vb);           // next chunk in flight while this one is scanned
          if (ragged) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + cc * 32 + j >= N) va[j] = INFINITY;
          }
          if (DEBUG && p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + cc * 32 + j < N) p.debug_scores[qrow * p.N + col0 + cc * 32 + j] = va[j];
          }
          process_chunk<KPP>(va, col0 + cc * 32, st, s_addr, i_addr, qs_addr, qi_addr);
          __syncwarp();                                       // tcgen05.ld/wait are warp-collective
          tmem_ld_wait(vb);
          if (cc + 2 < TILE_N / 32) tmem_ld_32x32(taddr + (cc + 2) * 32, va);
          if (ragged) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + (cc + 1) * 32 + j >= N) vb[j] = INFINITY;
          }
          if (DEBUG && p.debug_scores != nullptr && qrow < p.Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + (cc + 1) * 32 + j < N) p.debug_scores[qrow * p.N + col0 + (cc + 1) * 32 + j] = vb[j];
          }
          process_chunk<KPP>(vb, col0 + (cc + 1) * 32, st, s_addr, i_addr, qs_addr, qi_addr);
          __syncwarp();
          // drain early when a queue is half full, so that thresholds do not go stale in the cold phase
          if (__any_sync(FULL, st.cnt >= QCAP / 2)) drain_queue<KPP>(st, s_addr, i_addr, qs_addr, qi_addr);
        }
        tcgen05_fence_before();
        if (PAIR) mbar_arrive_cluster(&bars->tmem_empty[acc], 0);   // the leader waits for both CTAs' epilogues
        else mbar_arrive(&bars->tmem_empty[acc]);
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
        // the accumulator is released: fold the pending hits in (off the MMA's critical path)
        drain_queue<KPP>(st, s_addr, i_addr, qs_addr, qi_addr);
        // exchange thresholds with the other strips of this query (L2 atomics)
        if (gthr != nullptr) {
          if (st.thr_list < published) {
            atomicMin(gthr, f2key(st.thr_list));
            published = st.thr_list;
          }
          // consume the load issued one tile ago (its latency is hidden behind a whole tile),
          // then put the next one in flight
          if (gk_inflight < KEY_INF) st.thr_g = fminf(st.thr_g, next_up(key2f(gk_inflight)));
          st.thr = fminf(st.thr_list, st.thr_g);
          gk_inflight = ld_cg_u32(gthr);
        }
      }
      tot_ins += st.n_ins; tot_drain += st.n_drain; tot_now += st.n_now;
      // publish this strip's lists: one coalesced row of k' entries per query
      __syncwarp();
      int mypos = slot;
      if (p.list_count != nullptr) {          // compact slots: the lane that owns the row reserves one, if not empty
        bool any = false;
        for (int e = 0; e < KP; ++e) any |= my_i[e * TILE_M] >= 0;
        mypos = (any && qrow < p.Q) ? atomicAdd(&p.list_count[qrow], 1) : -1;
      }
      for (int r = 0; r < 32; ++r) {
        const int64_t qr = (int64_t)qt * TILE_M + quad * 32 + r;
        const int pos = __shfl_sync(FULL, mypos, r);
        if (qr < p.Q && pos >= 0) {
          for (int e = lane; e < KP; e += 32) {
            const int64_t o = (qr * sc.n_lists + pos) * KP + e;
            HYPRET_CHECK(pos < sc.n_lists && e < KPP && o < p.Q * sc.n_lists * KP);
            p.cand_score[o] = list_s[e * TILE_M + quad * 32 + r];
            p.cand_idx[o] = list_i[e * TILE_M + quad * 32 + r];
          }
        }
      }
      __syncwarp();
    }
    if (DEBUG && p.stats != nullptr && warp == 2 && lane == 0) {
      p.stats[blockIdx.x * 8 + 5] = w_acc;
      p.stats[blockIdx.x * 8 + 6] = (unsigned long long)clock64() - t_begin;
      p.stats[blockIdx.x * 8 + 7] = ((unsigned long long)tot_ins << 40) | ((unsigned long long)tot_drain << 20) |
                                    (unsigned long long)tot_now;
    }
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // neither CTA may exit while its peer can still reach its barriers / smem
  else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
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

// 2-D fp16 tensor [rows, kpad] row-major; box = box_cols x box_rows starting at (col, row).
int make_map(CUtensorMap* map, const void* base, int64_t rows, int kpad, int box_cols, int box_rows,
             CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return HYPRET_EDRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
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

Sched sched_from_plan(const hypret_score_plan_t& pl) {
  Sched s{};
  s.T = (pl.n_qtiles + pl.pair - 1) / pl.pair; s.G = pl.n_gtiles; s.P = pl.grid / pl.pair;
  s.n_full = pl.n_full; s.tail_rows = pl.tail_rows; s.a = pl.a; s.b = pl.b; s.L1 = pl.l1;
  s.rem_rows = pl.rem_rows; s.rem_g0 = pl.rem_g0; s.m = pl.m; s.L2 = pl.l2; s.n_steps = pl.n_steps;
  s.n_lists = pl.n_lists;          // PUBLIC count (x epi_groups): the stride of a query's list slots
  s.sub = pl.sub;
  s.sub_tail = pl.sub_tail;
  return s;
}

int kpp_of(int kprime) { return kprime <= 16 ? 16 : (kprime <= 32 ? 32 : 64); }

template <bool RESIDENT, int KPP, bool DEBUG, bool PAIR>
int launch_one(const hypret_score_plan_t& plan, const CUtensorMap& mq_main, const CUtensorMap& mq_ext,
               const CUtensorMap& mg_main, const CUtensorMap& mg_ext, const Params& p, cudaStream_t stream) {
  auto kern = score_topk_kernel<RESIDENT, KPP, DEBUG, PAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3(KPP == RL ? NUM_THREADS + NUM_EPI_THREADS : NUM_THREADS);
  cfg.dynamicSmemBytes = (size_t)plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, mq_main, mq_ext, mg_main, mg_ext, p);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}
template <bool RESIDENT, int KPP>
int launch_variant(const hypret_score_plan_t& plan, const CUtensorMap& mq_main, const CUtensorMap& mq_ext,
                   const CUtensorMap& mg_main, const CUtensorMap& mg_ext, const Params& p, cudaStream_t stream) {
  const bool dbg = p.debug_scores != nullptr || p.stats != nullptr;
  if (plan.pair == 2)
    return dbg ? launch_one<RESIDENT, KPP, true, true>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
               : launch_one<RESIDENT, KPP, false, true>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream);
  return dbg ? launch_one<RESIDENT, KPP, true, false>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
             : launch_one<RESIDENT, KPP, false, false>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream);
}

}  // namespace

extern "C" int hypret_score_plan(int64_t Q, int64_t N, int d, int kprime, int max_ctas, int min_lists,
                                 hypret_score_plan_t* plan) {
  if (plan == nullptr || Q < 1 || N < 1 || d < 1 || kprime < 1 || kprime > 64 || max_ctas < 0 || min_lists < 0 ||
      min_lists > 64)
    return HYPRET_EINVAL;
  if (N > 0x7fffffffll - TILE_N) return HYPRET_EUNSUPPORTED;   // int32 candidate indices per shard
  const int sms = device_sms();
  const int kb = hypret_dpad(d) / HYPRET_KBLK;
  const int64_t n_qtiles = (Q + TILE_M - 1) / TILE_M;
  const int64_t n_gtiles = (N + TILE_N - 1) / TILE_N;
  if (n_qtiles > (1 << 24)) return HYPRET_EUNSUPPORTED;

  // CTA pairs (tcgen05 cta_group::2) whenever there are at least two query tiles to pair up
  int ctas = sms;
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  const bool pair = n_qtiles >= 2 && ctas >= 2;
  const int b_blk = pair ? B_BLK_BYTES / 2 : B_BLK_BYTES;

  // shared-memory carve-up
  // candidate lists + pending-hit queues; k' <= 16: lists in registers, one queue per epilogue warpgroup
  const int epi_groups = kpp_of(kprime) == RL ? 2 : 1;
  const int lists = epi_groups == 2 ? 2 * QCAP_R * TILE_M * 8 + WG_X_BYTES
                                    : kpp_of(kprime) * TILE_M * 8 + QCAP * TILE_M * 8;
  const int a_res_bytes = kb * A_BLK_BYTES + A_EXT_BYTES;
  int resident = 0, stages = 0, stage_bytes = 0;
  {
    const int avail = SMEM_LIMIT - 1024 - BAR_BYTES - lists - a_res_bytes;
    // Resident query tile only when the gallery ring is still >= 4 stages deep; with the two
    // stages that D=512 leaves in single-CTA mode the MMA pipe starves (measured 863 vs 1103 TFLOP/s).
    if (kb <= 8 && avail >= 4 * b_blk) {
      resident = 1;
      stage_bytes = b_blk;
      stages = avail / b_blk;
    } else {
      stage_bytes = A_BLK_BYTES + b_blk;
      stages = (SMEM_LIMIT - 1024 - BAR_BYTES - lists) / stage_bytes;
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return HYPRET_EUNSUPPORTED;
  }
  const int64_t rows = pair ? (n_qtiles + 1) / 2 : n_qtiles;     // scheduling rows: query tiles or tile pairs
  Sched s = make_sched(rows, n_gtiles, pair ? ctas / 2 : ctas, 0);
  {
    // wide top-k (k > kprime): every query needs at least min_lists independent candidate lists
    if (min_lists > 1) {
      const int tail_steps = s.n_steps - s.n_full;          // 0, 1 or 2 base phases after the full waves
      const int tail_lists = s.tail_rows > 0 ? s.n_lists : 0;
      s.sub = s.n_full > 0 ? min_lists : 1;
      s.sub_tail = (s.tail_rows > 0 && s.a < min_lists) ? (min_lists + s.a - 1) / s.a : 1;
      s.n_steps = s.n_full * s.sub + tail_steps * s.sub_tail;
      const int full_lists = s.n_full > 0 ? s.sub : 0;
      s.n_lists = full_lists > tail_lists * s.sub_tail ? full_lists : tail_lists * s.sub_tail;
    }
  }
  plan->n_qtiles = (int32_t)n_qtiles;
  plan->n_gtiles = (int32_t)n_gtiles;
  plan->n_lists = s.n_lists * epi_groups;
  plan->epi_groups = epi_groups;
  plan->grid = pair ? 2 * s.P : s.P;
  plan->stages = stages;
  plan->resident = resident;
  plan->smem_bytes = 1024 + (resident ? a_res_bytes : 0) + stages * stage_bytes + lists + BAR_BYTES;
  plan->n_full = s.n_full;
  plan->tail_rows = s.tail_rows;
  plan->a = s.a;
  plan->b = s.b;
  plan->l1 = s.L1;
  plan->rem_rows = s.rem_rows;
  plan->rem_g0 = s.rem_g0;
  plan->m = s.m;
  plan->l2 = s.L2;
  plan->n_steps = s.n_steps;
  plan->pair = pair ? 2 : 1;
  plan->sub = s.sub;
  plan->sub_tail = s.sub_tail;
  return HYPRET_OK;
}

extern "C" int hypret_score_strip(const hypret_score_plan_t* plan, int cta, int step, int32_t* out4) {
  if (plan == nullptr || out4 == nullptr || cta < 0 || cta >= plan->grid / plan->pair || step < 0 ||
      step >= plan->n_steps)
    return HYPRET_EINVAL;
  const Sched s = sched_from_plan(*plan);
  int qt = 0, g0 = 0, g1 = 0, slot = 0;
  const bool ok = strip_at(s, cta, step, qt, g0, g1, slot);
  out4[0] = qt * plan->pair; out4[1] = g0; out4[2] = g1; out4[3] = slot * plan->epi_groups;
  return ok ? 1 : 0;
}

int hypret_launch_score_topk(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int kbound,
                             int n_lists, int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx,
                             uint32_t* thr_ws, int32_t* list_count, float* debug_scores, cudaStream_t stream) {
  hypret_score_plan_t plan;
  int rc = hypret_score_plan(Q, N, d, kprime, max_ctas, min_lists, &plan);
  if (rc != HYPRET_OK) return rc;
  if (plan.n_lists != n_lists) return HYPRET_EINVAL;   // caller sized cand_* for a different plan
  if (kbound > kprime && (kpp_of(kprime) != RL || kbound > 2 * RL)) return HYPRET_EINVAL;
  if ((reinterpret_cast<uintptr_t>(q_op) & 15) || (reinterpret_cast<uintptr_t>(g_op) & 15)) return HYPRET_EINVAL;

  const int dpad = hypret_dpad(d);
  const int kpad = dpad + HYPRET_KEXT;
  CUtensorMap mq_main, mq_ext, mg_main, mg_ext;
  if ((rc = make_map(&mq_main, q_op, Q, kpad, HYPRET_KBLK, TILE_M, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mq_ext, q_op, Q, kpad, HYPRET_KEXT, TILE_M, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  const int g_box = TILE_N / plan.pair;   // CTA pairs: each CTA stages half of the gallery tile
  if ((rc = make_map(&mg_main, g_op, N, kpad, HYPRET_KBLK, g_box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mg_ext, g_op, N, kpad, HYPRET_KEXT, g_box, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;

  Params p;
  p.Q = Q;
  p.N = N;
  p.kb_main = dpad / HYPRET_KBLK;
  p.has_ext = 1;
  p.dpad = dpad;
  p.kprime = kprime;
  p.kbound = kbound > kprime ? kbound : kprime;
  p.stages = plan.stages;
  p.stage_bytes = plan.resident ? B_BLK_BYTES / plan.pair : A_BLK_BYTES + B_BLK_BYTES / plan.pair;
  p.ring_off = plan.resident ? p.kb_main * A_BLK_BYTES + A_EXT_BYTES : 0;
  p.lists_off = p.ring_off + plan.stages * p.stage_bytes;
  p.bar_off = p.lists_off + (plan.epi_groups == 2 ? 2 * QCAP_R * TILE_M * 8 + WG_X_BYTES
                                                  : kpp_of(kprime) * TILE_M * 8 + QCAP * TILE_M * 8);
  p.sched = sched_from_plan(plan);
  p.cand_score = cand_score;
  p.cand_idx = cand_idx;
  p.list_count = list_count;
  p.debug_scores = debug_scores;
  p.shared_thr = thr_ws;
  p.stats = nullptr;
  p.wait_mode = 0;
  p.pf_tiles = 0;      // measured at C2 and on 37.5k-row shards: no gain from 2 or 4 tiles of L2 lookahead
  p.wg_off = p.bar_off - WG_X_BYTES;
  // Experiments only: HYPRET_STATS=1 runs the instrumented kernel variant, synchronises and prints
  // per-role wait-cycle totals to stderr.
  const char* stats_env = getenv("HYPRET_STATS");
  const bool want_stats = stats_env != nullptr && stats_env[0] == '1' && debug_scores == nullptr;
  if (want_stats) {
    if (cudaMalloc(&p.stats, (size_t)plan.grid * 8 * sizeof(unsigned long long)) != cudaSuccess) p.stats = nullptr;
    if (p.stats != nullptr) cudaMemsetAsync(p.stats, 0, (size_t)plan.grid * 8 * sizeof(unsigned long long), stream);
  }

  // list slots that no strip writes (rows with fewer strips than n_lists) must read as empty -- or, with
  // list_count, are never read: only the [Q] counters are cleared
  cudaError_t e = list_count != nullptr
                      ? cudaMemsetAsync(list_count, 0, (size_t)Q * sizeof(int32_t), stream)
                      : cudaMemsetAsync(cand_idx, 0xFF, (size_t)Q * n_lists * kprime * sizeof(int32_t), stream);
  if (e != cudaSuccess) return (int)e;
  if (thr_ws != nullptr) {   // 0xFFFFFFFF > f2key(+inf): "no bound published yet"
    e = cudaMemsetAsync(thr_ws, 0xFF, (size_t)Q * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return (int)e;
  }

  const int kpp = kpp_of(kprime);
  if (plan.resident)
    rc = kpp == 16   ? launch_variant<true, 16>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
         : kpp == 32 ? launch_variant<true, 32>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
                     : launch_variant<true, 64>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream);
  else
    rc = kpp == 16   ? launch_variant<false, 16>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
         : kpp == 32 ? launch_variant<false, 32>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream)
                     : launch_variant<false, 64>(plan, mq_main, mq_ext, mg_main, mg_ext, p, stream);
  if (p.stats != nullptr) {
    std::vector<unsigned long long> h((size_t)plan.grid * 8);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), p.stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(p.stats);
    double s[8] = {0};
    double ins = 0, drains = 0, nows = 0;
    for (int c = 0; c < plan.grid; ++c) {
      for (int i = 0; i < 7; ++i) s[i] += (double)h[(size_t)c * 8 + i] / plan.grid;
      const unsigned long long pk = h[(size_t)c * 8 + 7];
      ins += (double)(pk >> 40) / plan.grid;
      drains += (double)((pk >> 20) & 0xfffff) / plan.grid;
      nows += (double)(pk & 0xfffff) / plan.grid;
    }
    fprintf(stderr, "hypret stats: per CTA (warp 2 lane 0): list inserts %.0f, drain passes %.0f, queue-bypass inserts %.0f\n",
            ins, drains, nows);
    fprintf(stderr,
            "hypret stats (mean cycles per CTA, wait_mode %d): producer wait-empty %.0f of %.0f | mma wait-full %.0f "
            "wait-tmem-empty %.0f of %.0f | epilogue(warp2) wait-tmem-full %.0f of %.0f\n",
            p.wait_mode, s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
  }
  return rc;
}
