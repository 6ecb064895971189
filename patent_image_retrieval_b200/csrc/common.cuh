// Shared device helpers for the hypret kernels (sm_100a only).
//
// Thin inline-PTX wrappers around the Blackwell primitives the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the UMMA shared-memory / instruction descriptors.  Nothing here is generic:
// the project targets -gencode arch=compute_100a,code=sm_100a exclusively.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hypret.h"

// Checked build (python -m patent_image_retrieval_b200.build --checked, -DHYPRET_CHECKED): every HYPRET_CHECK guards
// an index or a count right before it is used -- a violation traps the kernel (cudaErrorAssert at the next
// synchronisation) instead of reading or writing out of bounds.  The product build compiles them away.
#ifdef HYPRET_CHECKED
#include <assert.h>
#define HYPRET_CHECK(cond) assert(cond)
#else
#define HYPRET_CHECK(cond) ((void)0)
#endif

// ----------------------------------------------------------------------------- geometry
// fp16 GEMM operand row (scoring) = [ Dpad main columns | 16 extension columns ], Dpad = roundup(D, 64).
constexpr int HYPRET_KBLK = 64;   // K elements per 128B-swizzled block
constexpr int HYPRET_MAX_PEERS = 16;   // ranks of one NVLink box an exchange buffer can address
constexpr int HYPRET_KEXT = 16;   // extension block: one UMMA_K step (32B-swizzled)

__host__ __device__ inline int hypret_dpad(int d) { return (d + HYPRET_KBLK - 1) / HYPRET_KBLK * HYPRET_KBLK; }
__host__ __device__ inline int hypret_kpad(int d) { return hypret_dpad(d) + HYPRET_KEXT; }

#define HYPRET_CUDA_CHECK(expr)                                  \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return (int)_e;                       \
  } while (0)

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: up to 10 ms inside the HW wait
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> cudaErrorLaunchFailure surfaced through the
// C ABI) instead of hanging the GPU.  4 s is >1000x the longest legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      // no printf here: a call inside the wait forces every value that is live across it (the epilogue's
      // register-resident candidate lists) onto the stack
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on `bar` (complete_tx bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(policy)
      : "memory");
}
constexpr uint64_t TMA_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t TMA_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t TMA_EVICT_LAST = 0x14F0000000000000ull;

// ----------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 16-bit inputs (kind::f16: bf16 or fp16 as the instruction descriptor says), fp32
// accumulate; one thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously-issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but with the 32 destination registers of a preceding tmem_ld_32x32 tied to the
// statement as in/out operands: no consumer of v[] can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster (same TPC) run one UMMA of M=256 together: each holds its own 128 rows
// of A and one half (N/2 rows) of B in shared memory at IDENTICAL offsets, the leader (cluster
// rank 0) issues the instruction, each CTA's tensor core fills its own TMEM.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the LEADER's
// mbarrier at the same shared-memory offset (rank bit cleared).
// L2 prefetch of a 2-D tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0),
        "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once the pair's MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
      "}\n"
      :
      : "r"(smem_u32(bar)), "r"(rank)
      : "memory");
}

// ----------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, canonical swizzled layout
// (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout type [61,64).  For swizzled K-major layouts LBO is the
// constant 1 and SBO is the byte distance between consecutive 8-row groups.
constexpr uint64_t UMMA_LAYOUT_SW128 = 2;
constexpr uint64_t UMMA_LAYOUT_SW32 = 6;
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= layout << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32,
// both operands K-major: c_format=F32 [4,6) | a_format=BF16 [7,10) | b_format=BF16 [10,13) |
// N>>3 [17,23) | M>>4 [24,29).
// the same with fp16 operands (a_format = b_format = F16 = 0): the scoring kernel's operands (csrc/project.cu)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ----------------------------------------------------------------------------- misc
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

// ----------------------------------------------------------------------------- launchers (one per .cu)
int hypret_launch_project_rows_peers(const float* u, int64_t n, int d, float c, int mode, float* y32,
                                     void* const* op_dsts_host, int n_dst, float* op_err, cudaStream_t stream);
int hypret_launch_peer_signal(void* const* flags_host, int n, uint32_t value, cudaStream_t stream);
int hypret_launch_peer_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, cudaStream_t stream);
int hypret_launch_project_rows(const float* u, int64_t n, int d, float c, int mode, int side, float* y32,
                               void* op_f16, float* sqnorm, float* op_err, float* stats, cudaStream_t stream);
int hypret_launch_score_topk(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int kbound,
                             int n_lists, int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx,
                             uint32_t* thr_ws, int32_t* list_count, float* debug_scores, cudaStream_t stream);
// Kernel-side view of hypret_peer_route (passed by value); n == 0: no routing.
struct PeerRoute {
  char* base[HYPRET_MAX_PEERS];
  int n, me;
  int64_t ql;
};
inline PeerRoute make_route(const hypret_peer_route* r) {
  PeerRoute o;
  o.n = 0; o.me = 0; o.ql = 1;
  for (int i = 0; i < HYPRET_MAX_PEERS; ++i) o.base[i] = nullptr;
  if (r != nullptr && r->n_ranks > 0) {
    o.n = r->n_ranks; o.me = r->me; o.ql = r->ql;
    for (int i = 0; i < r->n_ranks && i < HYPRET_MAX_PEERS; ++i) o.base[i] = static_cast<char*>(r->base[i]);
  }
  return o;
}
// row q of a rank-major [n*ql, .] result -> (owner rank, row in the owner's receive region)
__device__ __forceinline__ int64_t route_row(const PeerRoute& r, int64_t q, int* owner) {
  const int64_t o = q / r.ql;
  *owner = (int)o;
  return (int64_t)r.me * r.ql + (q - o * r.ql);
}

// Exact-top-k certificate of the rerank kernels (hypret_rerank_cert): per-query bound E on |tensor-core surrogate - exact
// surrogate| from the projection kernel's rounding residuals; a query whose margin does not exceed E is appended to
// `list` (and its lock words in `state` cleared) for hypret_exact_topk.  q_err == nullptr: no certificate.
struct CertArgs {
  const float* q_err;     // [Q] rounding residual norm of each query operand row
  const float* g_stats;   // [4] gallery maxima (csrc/project.cu)
  float slack;            // relative allowance for fp32 accumulation / split truncation
  int32_t* state;         // [2Q] {lock, initialised} words of the fallback merge
  int32_t* count;         // [1] number of uncertified queries
  int32_t* list;          // [Q] their ids
  uint8_t* flags;         // [Q] or NULL: 1 = certified by the filter pass, 0 = sent to the exact scan
  float* bound;           // [Q] or NULL: k-th score of the filtered result of an uncertified query (warm start of the scan)
  int ksel;               // candidates rescored per query (0 = k'); > k': lists of k' slots sharing a ksel-th-best bound
};
inline CertArgs no_cert() {
  CertArgs a;
  a.q_err = nullptr; a.g_stats = nullptr; a.slack = 0.f; a.state = nullptr; a.count = nullptr; a.list = nullptr;
  a.flags = nullptr; a.ksel = 0; a.bound = nullptr;
  return a;
}
int hypret_launch_rerank(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_cand,
                         int kprime, int k, int64_t idx_offset, const float* prune_thr, float* out_score,
                         int64_t* out_idx, float* out_margin, const hypret_peer_route* route, int64_t score_off,
                         int64_t idx_off, const double* g_sq64, const CertArgs& cert, cudaStream_t stream);
int hypret_launch_exact_topk(const float* q32, const float* g32, const double* g_sq64, int64_t Q, int64_t N, int d,
                             float c, int metric, int k, int64_t idx_offset, const int32_t* q_list,
                             const int32_t* q_count, int32_t* state, float* out_score, int64_t* out_idx,
                             const unsigned long long* after, const float* init_bound, cudaStream_t stream);
int hypret_launch_row_sqnorm64(const float* x, int64_t n, int d, double* out, cudaStream_t stream);
int hypret_launch_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                              int n_cand, int kprime, float* sel_score, int32_t* sel_idx,
                              const hypret_peer_route* route, int64_t recv_off, cudaStream_t stream);
int hypret_launch_kth_smallest(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out,
                               const hypret_peer_route* route, int64_t out_off, cudaStream_t stream);
int hypret_launch_merge_topk(const float* scores, const int64_t* idx, int W, int64_t Q, int k, int descending,
                             float* out_score, int64_t* out_idx, cudaStream_t stream);
int hypret_launch_pairdist(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float* out,
                           cudaStream_t stream);
int hypret_launch_retrieval_metrics(const int64_t* ranked, int64_t Q, int K, const int64_t* pos_off,
                                    const int64_t* pos_items, const int32_t* n_pos_total, const int32_t* ks_host,
                                    int n_ks, double* per_query, double* means, cudaStream_t stream);
int hypret_launch_ap_full(const float* scores, int64_t Q, int64_t N, const int64_t* pos_off, const int64_t* pos_items,
                          int grouped_ties, double* ap, int32_t* valid, double* mean_ap, cudaStream_t stream);
int hypret_launch_pair_keys(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                            const int64_t* pos_off, const int64_t* pos_items, int64_t idx_offset, float* keys,
                            cudaStream_t stream);
int hypret_launch_rank_count(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                             const int64_t* pos_off, const int64_t* pos_items, const float* pos_keys,
                             int64_t idx_offset, unsigned long long* counts, int32_t* bad, cudaStream_t stream);
int hypret_launch_ap_from_counts(const int64_t* pos_off, const int64_t* pos_items, const float* pos_keys,
                                 const unsigned long long* counts, const int32_t* bad, int64_t Q, int64_t n_total,
                                 int grouped_ties, double* ap, int32_t* valid, double* mean_ap, cudaStream_t stream);
int hypret_launch_pairdist_bwd(const float* g, const float* dmat, const float* asq, const float* psq, int64_t n,
                               int64_t m, float c, void* w_out, int w_format, float* row_partial, int n_row_partial,
                               float* col_partial, int n_partial, cudaStream_t stream);
int hypret_launch_pairdist_ce_fwd(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float inv_tau,
                                  int want_cols, float* dmat, float* row_lse, float* col_lse, float* scratch,
                                  int n_part, cudaStream_t stream);
int hypret_launch_pairdist_ce_bwd(const float* dmat, const float* asq, const float* psq, int64_t n, int64_t m, float c,
                                  const float* row_lse, const float* col_lse, float inv_tau, float wr, float wc,
                                  const float* grad_scale, void* w_out, int w_format, float* row_partial,
                                  int n_row_partial, float* col_partial, int64_t diag_offset, int64_t n_total,
                                  cudaStream_t stream);
int hypret_launch_split3(const float* x, int64_t count, void* out_bf16, cudaStream_t stream);
int64_t hypret_gram_kpad_impl(int d);
int hypret_launch_gram_split(const float* x, int64_t n, int d, int side, void* out_bf16, float* sq, cudaStream_t stream);
int hypret_launch_gram_dist(const void* a_op, const void* p_op, const float* a32, const float* p32, const float* asq,
                            const float* psq, int64_t n, int64_t m, int d, float c, float* out, cudaStream_t stream);
int hypret_launch_neg_lse(const float* dmat, int64_t n, int64_t m, float inv_tau, int want_cols, float* row_lse,
                          float* col_lse, float* scratch, int n_part, cudaStream_t stream);
int hypret_launch_mobius_epilogue(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                                  int hyperbolic_input, int post_tanh, int n_project, float* y, float* sqnorm,
                                  cudaStream_t stream);

int hypret_launch_rowpair_dist(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs,
                               int d, float c, float* out, const float* grad_out, float* gx, float* gy,
                               cudaStream_t stream);
int hypret_launch_hmi_pairs(const float* emb, const int64_t* pairs, int64_t n_pairs, int d, float c, int mode,
                            float margin, float proj_eps, float* values, double* loss_sum, const float* grad_scale,
                            float* grad_emb, cudaStream_t stream);
int hypret_launch_dist0_reg(const float* x, int64_t n, int d, float c, float lo, float hi, double* loss_sum,
                            const float* grad_scale, float* grad_x, cudaStream_t stream);
int hypret_launch_radam_ball(float* x, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int d, float c,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                             cudaStream_t stream);

int64_t hypret_flash_kpad_impl(int d);
int64_t hypret_flash_workspace_floats(int64_t n, int64_t m, int d);
int hypret_launch_flash_prep(const float* x, int64_t n, int d, void* row_op, void* col_op, void* t_planes, int64_t n_pad,
                             float* sq, cudaStream_t stream);
int hypret_launch_flash(int bwd, const void* x_row_op, const void* y_col_op, const void* y_t_planes, int64_t yt_cols,
                        const float* x32, const float* y32, const float* xsq, const float* ysq, const float* x_lse,
                        const float* y_lse, int64_t n, int64_t m, int d, float c, float inv_tau, float wx, float wy,
                        const float* grad_scale, int64_t diag_offset, int64_t n_total, float* workspace, float* out,
                        cudaStream_t stream);
int hypret_launch_mobius_epilogue_bwd(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                                      int post_tanh, int n_project, const float* gy, float* gmx, float* gbias,
                                      float* gxn, cudaStream_t stream);
int hypret_launch_sgemm_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                                int M, int N, int K, const float* row_scale, const float* addend, float* C,
                                cudaStream_t stream);
int hypret_launch_cert_merged(const float* q32, int64_t Q, int d, float c, int metric, const float* score,
                              const int64_t* idx, int k, const float* thr, const float* q_err, const float* g_stats,
                              float slack, int32_t* flags, float* out_margin, cudaStream_t stream);
int hypret_launch_flag_compact(const int32_t* flags, int64_t n, int32_t* list, int32_t* count, int32_t* state,
                               cudaStream_t stream);
int hypret_launch_sum_parts(const float* parts, int w, int64_t n, float* out, cudaStream_t stream);
int hypret_launch_lse_combine(const float* parts, int w, int64_t n, float* out, cudaStream_t stream);
int hypret_launch_mobius_gemm(const void* x_row_op, const void* w_col_op, int64_t n, int d_in, int n_out,
                              const float* xsq, const float* bias, float c, int post_tanh, int n_project, float* mx_out,
                              float* y_out, float* ysq_out, void* op_out, cudaStream_t stream);
