/* hypret -- C ABI of the B200-native hyperbolic / cosine retrieval hot path.
 *
 * One shared library (libhypret.so), plain pointers and sizes, no torch or C++
 * types in any signature.  The reference (Alvarodelamaza/patent-image-retrieval)
 * has no FFI of its own: its hot path is plain Python calling geoopt / sklearn.
 * Each entry point below therefore names the reference *call site* it replaces
 * (file:line under /root/reference); INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every buffer (outputs and workspaces included) is allocated by the caller;
 *   - row-major, contiguous; rows of fp32 matrices 16-byte aligned (D % 4 == 0);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, negative = HYPRET_E* argument error, positive = cudaError_t;
 *   - no C++ exception crosses the boundary; there is no CPU fallback.
 */
#ifndef HYPRET_H_
#define HYPRET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYPRET_OK 0
#define HYPRET_EINVAL (-1)      /* bad argument (shape, alignment, enum)            */
#define HYPRET_EUNSUPPORTED (-2) /* valid request this build cannot serve (e.g. D)   */
#define HYPRET_EDRIVER (-3)     /* driver entry point (cuTensorMapEncodeTiled) missing */
#define HYPRET_ENOTSM100 (-4)   /* device is not compute capability 10.x             */

/* projection modes (hypret_project_rows) */
#define HYPRET_MODE_EXPMAP0 0 /* u Euclidean  -> y = project(expmap0(u))             */
#define HYPRET_MODE_ONBALL 1  /* u on the ball -> y = project(u)                      */
#define HYPRET_MODE_COSINE 2  /* u Euclidean  -> y = u / ||u||  (zero rows stay zero) */
#define HYPRET_SIDE_QUERY 0
#define HYPRET_SIDE_GALLERY 1

/* scoring metrics (hypret_rerank) */
#define HYPRET_METRIC_COSINE 0
#define HYPRET_METRIC_HYPERBOLIC 1

const char* hypret_strerror(int rc);
int hypret_version(void);

/* Row length (elements) of the fp16 GEMM operand for feature dimension d:
 * roundup(d, 64) + 16 extension columns. */
int64_t hypret_operand_kpad(int d);

/* Fused projection + operand build, one pass over the rows.
 * Replaces pmath.expmap0 -> pmath.project (src/models.py:310,317), the final
 * pmath.project of DeeperHyperbolicEncoder.forward (src/models.py:504), and the row
 * normalisation inside sklearn cosine_similarity (notebooks/retrieval.ipynb:368).
 *   u        [n,d] fp32 in
 *   y32      [n,d] fp32 out or NULL  (the point; exact-rerank operand)
 *   op_f16   [n,kpad] fp16 out or NULL (tensor-core operand in unit-ball coordinates, see csrc/project.cu)
 *   sqnorm   [n] fp32 out or NULL    (||y||^2)
 * d % 4 == 0, d <= 2048. */
int hypret_project_rows(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, void* op_f16,
                        float* sqnorm, void* stream);
/* The same, plus the inputs of the exact-top-k certificate (hypret_rerank_cert).  With m_i the fp32 row that is rounded
 * to the fp16 main columns of the operand (x^_i = sqrt(c) x_i for a query; -2 rb_j y^_j for a gallery row; the unit
 * vector for cosine):
 *   op_err [n] fp32 out or NULL   || fp16(m_i) - m_i ||_2, rounded up
 *   stats  [4] fp32 in/out or NULL  running maxima over the rows (caller zeroes once per gallery; atomicMax):
 *          max ||fp16(m)||, max op_err, max rb, max rb ||y^||^2   (rb = 1 / (1 - c ||y||^2); 0 for cosine)
 * |<fp16 x, fp16 z> - <x, z>| <= op_err(x) ||fp16 z|| + ||x|| op_err(z) then bounds the filter's error for every pair. */
int hypret_project_rows_cert(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, void* op_f16,
                             float* sqnorm, float* op_err, float* stats, void* stream);

/* ---- Peer-memory exchange (multi-GPU serving on one NVLink / NVSwitch box; csrc/peer.cu) ------------------
 * The reference is single-GPU; these replace the NCCL all_gather of the per-rank query batches in the
 * sharded-serving protocol.  One process per GPU; every device of the box visible to every process.
 *   hypret_peer_alloc   cudaMalloc + zero-fill `bytes`, and the 64-byte CUDA IPC handle to send to the peers
 *   hypret_peer_open    map a peer's buffer from its handle (peer access is enabled on demand)
 *   hypret_peer_close / hypret_peer_free   undo the two above
 *   hypret_peer_copy    asynchronous copy between any two mapped buffers (copy engines; no SM)
 *   hypret_project_rows_peers   hypret_project_rows for query rows, with the fp16 operand row stored into
 *                       n_dst destination buffers (op_dsts_host: HOST array of device pointers, each already
 *                       offset to the block of this rank) -- the projection is the all-gather; op_err [n] fp32 out or
 *                       NULL as in hypret_project_rows_cert
 *   hypret_peer_signal  behind everything queued on `stream` so far: *flags_host[i] = value (release, system scope)
 *   hypret_peer_wait    hold `stream` until flags[i] >= value for all i < n (20 s bound: then *err = 1 + i) */
#define HYPRET_IPC_HANDLE_BYTES 64
int hypret_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out_host);
int hypret_peer_free(void* dev_ptr);
int hypret_peer_open(const void* handle_host, void** peer_ptr);
int hypret_peer_close(void* peer_ptr);
int hypret_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
int hypret_project_rows_peers(const float* u, int64_t n, int d, float c, int mode, float* y32,
                              void* const* op_dsts_host, int n_dst, float* op_err, void* stream);
int hypret_peer_signal(void* const* flags_host, int n, uint32_t value, void* stream);

/* Output routing of the sharded-serving exchange: the all_to_all / all_gather that follows a kernel is done BY that
 * kernel with posted NVLink stores into the receivers' exchange buffers (then hypret_peer_signal / hypret_peer_wait).
 * Row q of a [n_ranks*ql, width] result belongs to rank q / ql and is stored at row (me*ql + q % ql) of the
 * [n_ranks*ql, width] receive region at byte offset *_off of that rank's buffer (the layout an all_to_all of
 * equal blocks produces).  n_ranks == 0 or a NULL route: local outputs only. */
typedef struct {
  void* base[16];    /* exchange-buffer base of every rank, own included, as mapped in THIS process */
  int32_t n_ranks;
  int32_t me;
  int64_t ql;        /* rows per rank */
} hypret_peer_route;
int hypret_peer_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, void* stream);

/* Work decomposition of hypret_score_topk for a problem size on the current device.
 * The gallery is swept in "strips" (query tile x contiguous gallery-tile range); a query
 * receives one candidate list per strip that visits it.  See csrc/score_topk.cu. */
typedef struct {
  int32_t n_qtiles;   /* T = ceil(Q / 128)                                               */
  int32_t n_gtiles;   /* G = ceil(N / 256)                                               */
  int32_t n_lists;    /* candidate lists (slots) per query in cand_score / cand_idx      */
  int32_t grid;       /* persistent CTAs launched (P = grid / pair scheduling units)     */
  int32_t stages;     /* shared-memory pipeline stages                                   */
  int32_t resident;   /* 1: query tile resident in shared memory for a whole strip       */
  int32_t smem_bytes; /* dynamic shared memory per CTA                                   */
  int32_t n_full;     /* full waves: every CTA takes one whole gallery row of tiles      */
  int32_t tail_rows;  /* T' = T mod P query tiles left for phase 1 / phase 2             */
  int32_t a, b;       /* phase 1: rows < b get a+1 strips, the others a                  */
  int32_t l1;         /* phase-1 strip length in gallery tiles                           */
  int32_t rem_rows;   /* rows whose range [rem_g0, G) is finished in phase 2             */
  int32_t rem_g0;
  int32_t m, l2;      /* phase 2: m pieces per remaining row, l2 tiles each              */
  int32_t n_steps;    /* strips a CTA walks through at most: n_full + phases             */
  int32_t sub;        /* wide top-k: full-wave strips are cut into `sub` sub-strips with their own list slots */
  int32_t sub_tail;   /* ... and phase-1 / phase-2 strips into `sub_tail`                                  */
  int32_t pair;       /* 2: CTA pairs (tcgen05 cta_group::2) -- a scheduling unit is a cluster of 2 CTAs and a
                         strip covers 2 consecutive query tiles; 1: single-CTA MMAs            */
  int32_t epi_groups; /* epilogue warpgroups per CTA.  2 (kprime <= 16): a strip owns TWO lists per query, slots
                         hypret_score_strip's slot and slot+1, one per warpgroup (the two 128-column halves
                         of every gallery tile of the strip); n_lists already counts both                              */
} hypret_score_plan_t;

/* max_ctas: 0 = one CTA per SM; >0 caps the grid (tests use it to force multi-wave schedules).
 * min_lists: 0/1 = no constraint; >1 = every query gets at least that many independent candidate lists
 * (used for top-k with k > kprime: the union of the lists must contain the top-k, see hypret_rerank). */
int hypret_score_plan(int64_t Q, int64_t N, int d, int kprime, int max_ctas, int min_lists,
                      hypret_score_plan_t* plan);

/* Strip of scheduling unit `cta` (0 <= cta < grid / pair) at step `step` of `plan`:
 * out4 = {first query tile (the strip covers `pair` consecutive tiles), first gallery tile, end gallery
 * tile, list slot}.  Returns 1 if the CTA has work at that step, 0 if idle, <0 on bad arguments.
 * Host-only (no GPU needed); the kernel evaluates the same closed form on the device. */
int hypret_score_strip(const hypret_score_plan_t* plan, int cta, int step, int32_t* out4);

/* Scoring GEMM + fused streaming top-k' (tcgen05 / TMEM / TMA).
 * Replaces the one-vs-all scoring loops  pmath.dist(q[1,D], G[P,D])  (src/train.py:3259)
 * and  cosine_similarity(Q, G)  (notebooks/retrieval.ipynb:368) together with the ranking
 * that follows them (np.argsort, retrieval.ipynb:383,202; torch.topk, src/auxiliary.py:374)
 * as a *candidate filter*: per query and strip it keeps the kprime smallest fp16-operand
 * surrogate scores.  The [Q,N] matrix is never written to memory.
 *   q_op [Q,kpad] fp16, g_op [N,kpad] fp16   operands from hypret_project_rows
 *   cand_score [Q, n_lists, kprime] fp32 out, cand_idx same shape int32 out (-1 = empty)
 *   n_lists   must equal plan.n_lists of hypret_score_plan(Q, N, d, kprime, max_ctas, min_lists)
 *   thr_workspace  [Q] uint32 scratch, or NULL.  When given, the strips of a query exchange
 *             their running k'-th best score through it (L2 atomics), so later / concurrent
 *             strips start warm: the UNION of a query's lists still contains its global
 *             top-kprime, but a single list is no longer the top-kprime of its own strip.
 *             With NULL every list is exactly its strip's top-kprime (slower; tests).
 *   list_count  [Q] int32 out, or NULL.  Given: list slots are handed out per query in arrival order (and empty
 *             lists take none), so the lists of query q are exactly its first list_count[q] slots and nothing
 *             else of cand_* is written or needs clearing; hypret_rerank / hypret_cand_select take the same array.
 *             NULL: slot = the strip's slot of hypret_score_strip, unwritten slots read as empty (idx -1).
 *   debug_scores  NULL, or [Q,N] fp32 that receives every surrogate score (tests only)
 * 1 <= kprime <= 64. */
int hypret_score_topk(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int n_lists,
                      int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx, uint32_t* thr_workspace,
                      int32_t* list_count, float* debug_scores, void* stream);
/* hypret_score_topk whose lists share the bound of the query's kbound-th best score instead of the kprime-th
 * (kprime <= 16 < kbound <= 32, thr_workspace required): the union of a query's 16-slot register lists then holds its
 * global top-kbound unless one half-strip alone holds more than 16 of them -- the spare candidates the exact-top-k
 * certificate needs (hypret_rerank_cert, ksel = kbound) at the speed of the 16-slot kernel.  kbound == kprime: as
 * hypret_score_topk. */
int hypret_score_topk_bound(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int kbound,
                            int n_lists, int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx,
                            uint32_t* thr_workspace, int32_t* list_count, float* debug_scores, void* stream);

/* Candidate merge + exact rerank.  For each query: keep the kprime best of its
 * n_lists*kprime candidates by surrogate score, recompute their distance exactly from the
 * fp32 rows (differences formed explicitly, fp64 accumulation, arccosh closed form ==
 * pmath.dist, src/train.py:3259; or the cosine similarity, retrieval.ipynb:368), sort
 * (ascending distance / descending similarity, ties -> lower index) and emit the first k.
 *   list_count [Q] int32 from hypret_score_topk, or NULL (all n_lists slots are scanned)
 *   q32 [Q,d], g32 [N,d] fp32   (hyperbolic: points on the ball; cosine: raw features)
 *   out_score [Q,k] fp32, out_idx [Q,k] int64 (+ idx_offset; -1 when fewer than k rows)
 *   out_margin [Q] fp32 or NULL: (smallest surrogate a NON-candidate can have) - (exact surrogate of the
 *              k-th result) in unit-ball coordinates (c * rb * ||x - y||^2); hypret_rerank_cert compares it with a
 *              bound on the fp16 filter's error
 * k <= kprime <= 32: one warp per query.  Wide top-k (kprime <= 64, k <= 128, n_lists*kprime <= 16384,
 * k <= n_lists*kprime): one CTA per query; requires lists built WITHOUT threshold sharing and
 * min_lists >= 3 so that the union of a query's lists contains its top-k (certified by out_margin).
 *   g_sqnorm64 [N] fp64 or NULL: ||g_j||^2 from hypret_row_sqnorm64 (once per index).  The wide path rescoring 256
 *              survivors per query then skips the per-survivor norm (half of its fp64 work); NULL: computed in
 *              the kernel.  The one-warp path ignores it. */
int hypret_rerank(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                  const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_lists, int kprime,
                  int k, int64_t idx_offset, float* out_score, int64_t* out_idx, float* out_margin,
                  const double* g_sqnorm64, void* stream);
/* hypret_rerank (k <= kprime <= 32) with the EXACT-TOP-K GUARANTEE.  The filter is an fp16 GEMM, so a gallery row outside
 * the candidate set could in principle precede the k-th result.  Per query the kernel bounds the filter's error,
 *   E = q_err[q] * stats[0] + ||x_q|| * stats[1] + slack * (||x_q|| stats[0] + ||x_q||^2 stats[2] + stats[3]),
 * (q_err / stats from hypret_project_rows_cert; slack covers the fp32 accumulation of the tensor core), and accepts the
 * result only if  (k'-th best filter score) - (exact surrogate of the k-th result) > E  -- then no non-candidate can
 * have an exact surrogate at or below the k-th result's.  Otherwise the query id is appended to fb_list (fb_count
 * counts them; zeroed by this call) and its two lock words in fb_state [2Q] are cleared: hypret_exact_topk, queued
 * behind this call, recomputes exactly those queries from all gallery rows.  certified [Q] uint8 out or NULL.
 *   fb_bound [Q] fp32 out or NULL: for an uncertified query, the k-th score of its filtered result (an exact score of
 *         a real row, so no better than the true k-th best) -- hypret_exact_topk's init_bound.
 *   ksel  0, or kprime < ksel <= 32 for lists built by hypret_score_topk_bound(kbound = ksel): the ksel best candidates
 *         of the union of the kprime-slot lists are rescored, and the margin is taken against the smaller of the
 *         ksel-th best filter score and the worst entry of any FULL list (rows a full list had no slot for). */
int hypret_rerank_cert(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                       const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_lists,
                       int kprime, int ksel, int k, int64_t idx_offset, float* out_score, int64_t* out_idx,
                       float* out_margin, const float* q_err, const float* g_stats, int32_t* fb_state, int32_t* fb_count,
                       int32_t* fb_list, float* fb_bound, uint8_t* certified, void* stream);
/* Exact top-k of the listed queries by a scan of ALL gallery rows -- what the reference's per-query loop does
 * (pmath.dist one-vs-all + torch.topk, src/train.py:3259, src/auxiliary.py:374; cosine: retrieval.ipynb:368,202) --
 * with the arithmetic and ordering of hypret_rerank; overwrites rows q_list[0 .. *q_count) of out_score / out_idx
 * ([Q,k], + idx_offset).  *q_count is read on the device (no host synchronisation; an empty list costs one launch of
 * CTAs that exit at once).  g_sqnorm64 [N] from hypret_row_sqnorm64; fb_state [2Q] as left by hypret_rerank_cert.
 * k <= 32. 
 *   init_bound [Q] fp32 or NULL: per query a score that is no better than its true k-th best (distance: >=, similarity:
 *         <=; +-inf = none).  The scan then starts warm -- only rows at or inside the bound are contenders and CTAs
 *         that see none skip the merge -- and still returns the exact top-k. */
int hypret_exact_topk(const float* q32, const float* g32, const double* g_sqnorm64, int64_t Q, int64_t N, int d,
                      float c, int metric, int k, int64_t idx_offset, const int32_t* q_list, const int32_t* q_count,
                      int32_t* fb_state, const float* init_bound, float* out_score, int64_t* out_idx, void* stream);
/* Exact-top-k certificate of a MERGED result, for a row-sharded gallery whose queries are owned by one rank each
 * (dist.ShardedGalleryIndex.search_sharded): score/idx [Q,k] = the merged exact lists, thr [Q] = the global k'-th best
 * filter score of each query (hypret_kth_smallest over every shard's list), q_err [Q] = the rounding residuals of the
 * owner's query operands (hypret_project_rows_cert / _peers), g_stats [4] = the MAXIMA over all shards of the gallery
 * statistics.  flags [Q] int32 out: 0 = proven exact (as hypret_rerank_cert: thr - S_kth > E), 1 = not proven -- those
 * queries go through hypret_exact_topk on every shard.  out_margin [Q] or NULL. */
int hypret_cert_merged(const float* q32, int64_t Q, int d, float c, int metric, const float* score, const int64_t* idx,
                       int k, const float* thr, const float* q_err, const float* g_stats, int32_t* flags,
                       float* out_margin, void* stream);

/* flags [n] int32 (non-zero = listed) -> q_list (the flagged positions, any order), *q_count, and the fb_state words of
 * the flagged queries reset: the inputs hypret_exact_topk expects, built on the device. */
int hypret_flag_compact(const int32_t* flags, int64_t n, int32_t* q_list, int32_t* q_count, int32_t* fb_state,
                        void* stream);

/* The next page of the same exact ranking: only rows whose key (ordered fp32 score << 32 | local row id, the order of the
 * result lists) lies ABOVE after[q] of the listed query q (after is indexed by query id, [Q]) are taken.  Paging through the ranking 32 rows at a time gives the
 * exact top-k for any k (notebooks/retrieval.ipynb:202 with k beyond the 128 the filtered path serves). */
int hypret_exact_topk_after(const float* q32, const float* g32, const double* g_sqnorm64, int64_t Q, int64_t N, int d,
                            float c, int metric, int k, int64_t idx_offset, const int32_t* q_list,
                            const int32_t* q_count, int32_t* fb_state, const uint64_t* after, float* out_score,
                            int64_t* out_idx, void* stream);
/* out[i] = ||x_i||^2 accumulated in fp64 (x [n,d] fp32, d % 4 == 0), in the summation order of the rerank kernels. */
int hypret_row_sqnorm64(const float* x, int64_t n, int d, double* out, void* stream);

/* Multi-GPU pruning (sharded serving; the reference has no distributed path).  A query's exact
 * rescoring needs only its GLOBAL approximate top-kprime, of which a shard holds kprime / n_shards on average:
 *   hypret_cand_select    the kprime best candidates of each query by surrogate score, ascending (surrogate,
 *                         index), padded with (+inf, -1): sel_score [Q,kprime] fp32, sel_idx [Q,kprime] int32.
 *                         Surrogates of different shards are comparable (same query operand, same formula).
 *   hypret_kth_smallest   out[q] = kth smallest (1-based) of the n_parts*m values vals[n_parts, Q, m] (after the
 *                         owner of a query has received every shard's sel_score); +inf if fewer exist.
 *   hypret_rerank_pruned  hypret_rerank that skips every candidate whose surrogate exceeds prune_thr[q]
 *                         (the global kprime-th best): results of the shards then merge to exactly the list a
 *                         single-GPU hypret_rerank over the whole gallery returns (ties at the threshold kept).
 *                         k <= kprime <= 32 only. */
int hypret_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                       int n_lists, int kprime, float* sel_score, int32_t* sel_idx, void* stream);
int hypret_kth_smallest(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out, void* stream);
/* The same three with their exchange fused in (hypret_peer_route above):
 *   hypret_cand_select_route    sel_score additionally lands in the query owner's receive region (recv_off) -- the
 *                               all_to_all of the [Q,kprime] surrogates
 *   hypret_kth_smallest_route   out[q] (Q == route->ql own queries) additionally lands at [me*ql + q] of the [n_ranks*ql]
 *                               region (out_off) of EVERY rank -- the all_gather of the thresholds
 *   hypret_rerank_pruned_route  the [Q,k] lists go ONLY to the query owners' receive regions (score_off: fp32,
 *                               idx_off: int64) -- the two all_to_alls of the result lists */
int hypret_cand_select_route(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                             int n_lists, int kprime, float* sel_score, int32_t* sel_idx,
                             const hypret_peer_route* route, int64_t recv_off, void* stream);
int hypret_kth_smallest_route(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out,
                              const hypret_peer_route* route, int64_t out_off, void* stream);
int hypret_rerank_pruned_route(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                               const float* cand_score, const int32_t* cand_idx, int n_lists, int kprime, int k,
                               int64_t idx_offset, const float* prune_thr, const hypret_peer_route* route,
                               int64_t score_off, int64_t idx_off, void* stream);
int hypret_rerank_pruned(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, int n_lists, int kprime, int k,
                         int64_t idx_offset, const float* prune_thr, float* out_score, int64_t* out_idx,
                         void* stream);

/* Multi-GPU exchange step: merge the per-shard top-k lists of every query (after an
 * all-gather of [Q,k] scores + global indices) into the global top-k, with the ordering of the
 * reference's single-device ranking (notebooks/retrieval.ipynb:383, src/auxiliary.py:374):
 * ascending (descending != 0: descending) score, ties -> lower index.
 *   scores [n_shards, Q, k] fp32, idx [n_shards, Q, k] int64 (-1 = empty)
 *   out_score [Q,k], out_idx [Q,k];  n_shards * k <= 256. */
int hypret_merge_topk(const float* scores, const int64_t* idx, int n_shards, int64_t Q, int k, int descending,
                      float* out_score, int64_t* out_idx, void* stream);

/* Fused epilogue of one MobiusLinear layer, applied to mx = x W^T (the GEMM itself is left to the
 * caller).  Replaces, in one pass over [n,d]:
 *   hyperbolic_input == 0:  pmath.expmap0(mx)                              src/models.py:309-310
 *   hyperbolic_input != 0:  the rescale of pmath.mobius_matvec(W, x)       src/models.py:307  (needs xsq = ||x||^2 [n])
 *   bias != NULL:           pmath.mobius_add(., bias)  (bias on the ball)  src/models.py:314
 *   n_project (0..2):       pmath.project, repeated                        src/models.py:317, 504
 *   post_tanh != 0:         pmath.mobius_fn_apply(tanh, .)                 src/models.py:491
 * y [n,d] fp32 out, sqnorm [n] fp32 out or NULL (||y||^2, the next layer's xsq).  d % 4 == 0, d <= 2048. */
int hypret_mobius_epilogue(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                           int hyperbolic_input, int post_tanh, int n_project, float* y, float* sqnorm, void* stream);

/* One MobiusLinear layer as ONE kernel (csrc/headgemm.cu): tcgen05 GEMM from 2-way fp16 split operands + the whole
 * hyperbolic epilogue of hypret_mobius_epilogue on the accumulator in TMEM (src/models.py:291-318, 481-505).
 *   x_row_op [n, kpad(d_in)] fp16   row operand [hi|lo|hi] of the layer input (hypret_flash_prep, or the op_out of the
 *                                   previous layer); w_col_op [n_out, kpad(d_in)] column operand [hi|hi|lo] of the weight
 *   xsq [n] or NULL    ||x||^2 of the input: given = hyperbolic input (mobius_matvec), NULL = Euclidean input (expmap0)
 *   bias [n_out] or NULL (on the ball), post_tanh, n_project: as hypret_mobius_epilogue
 *   outputs, each optional: mx_out [n,n_out] fp32 raw product (the backward pass needs it), y_out [n,n_out] fp32,
 *   ysq_out [n], op_out [n, kpad(n_out)] fp16 = the NEXT layer's row operand (no fp32 round trip of the activations)
 * n_out % 16 == 0, n_out <= 256; d_in % 4 == 0; kpad = hypret_flash_kpad. */
int hypret_mobius_gemm(const void* x_row_op, const void* w_col_op, int64_t n, int d_in, int n_out, const float* xsq,
                       const float* bias, float c, int post_tanh, int n_project, float* mx_out, float* y_out,
                       float* ysq_out, void* op_out, void* stream);

/* Backward of the epilogue of one MobiusLinear layer (csrc/headbwd.cu): what autograd does through the ~35 elementwise
 * nodes of src/models.py:291-318 (+ 491, 504), in closed form, a warp per row.  The forward scalars are recomputed from
 * the saved raw product mx (hypret_mobius_gemm's mx_out).
 *   gy [n,d]      dL/d(layer output)
 *   gmx [n,d]     out: dL/d(mx)           (dW = gmx^T x_in, dx_in = gmx W: hypret_sgemm_strided)
 *   gbias [d]     +=: dL/d(bias), accumulated with atomics -- zero it first; NULL or bias == NULL: skipped
 *   gxn [n]       out: (dL/d||x_in||) / ||x_in|| through the mobius_matvec rescale, so that dL/dx_in = gmx W + gxn x_in;
 *                 needs xsq; NULL: skipped
 * Flags as hypret_mobius_gemm.  d % 4 == 0, d <= 512. */
int hypret_mobius_epilogue_bwd(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                               int post_tanh, int n_project, const float* gy, float* gmx, float* gbias, float* gxn,
                               void* stream);

/* out[m,n] (row-major, fp32) = sum_k a[m*a_row_stride + k*a_col_stride] * b[k*b_row_stride + n*b_col_stride]
 *                               (+ row_scale[m] * addend[m,n] when both are given): the batch-sized dense products of the head's backward pass (torch.nn.functional.linear's backward in the reference,
 * src/models.py:305-310), FP32 FMA tiles. */
int hypret_sgemm_strided(const float* a, int64_t a_row_stride, int64_t a_col_stride, const float* b,
                         int64_t b_row_stride, int64_t b_col_stride, int m, int n, int k, const float* row_scale,
                         const float* addend, float* out, void* stream);

/* Exact pairwise Poincare distance matrix out[i,j] = dist(a_i, p_j), fp32 [n,m].
 * Replaces the Python loops of 1x1 / 1xN pmath.dist calls (src/train.py:1832-1840, 2304-2320,
 * 3259, 1033).  Differences formed explicitly in fp32, transcendental tail in fp64. */
int hypret_pairdist(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float* out, void* stream);

/* Backward of hypret_pairdist (the reference differentiates through ~40 elementwise autograd
 * nodes per pair, src/train.py:1846).  Given grad_out = dL/dD [n,m] and the forward matrix dmat, ONE fp32 pass:
 *   w_out                    W (see csrc/pairdist.cu): w_format 0 = fp32 [n,m]; 1 = three bf16 planes [3,n,m] (hi, mid,
 *                            lo; hi+mid+lo = W to fp32 accuracy) for split-bf16 tensor-core GEMMs of the two products
 *   row_partial [n_row_partial,n]  partial row sums of W_ij (1 + c s_ij / alpha_i): the columns are cut into
 *                            n_row_partial chunks (grid = ceil(n/16) x n_row_partial CTAs; pick it so that the grid
 *                            is several waves of 2 CTAs per SM); sum over dim 0
 *   col_partial [n_partial,m]  partial column sums of W_ij (1 + c s_ij / beta_j), one row per HYPRET_BWD_ROWS (16)
 *                            matrix rows: n_partial >= ceil(n/16); sum over dim 0
 * so that  dA = a * row_sum[:,None] - W P,  dP = p * col_sum[:,None] - W^T A  (two plain GEMMs
 * left to the caller).  asq / psq = squared norms of the rows of a / p. */
#define HYPRET_BWD_ROWS 16
int hypret_pairdist_bwd(const float* grad_out, const float* dmat, const float* asq, const float* psq, int64_t n,
                        int64_t m, float c, void* w_out, int w_format, float* row_partial, int n_row_partial,
                        float* col_partial, int n_partial, void* stream);

/* In-batch InfoNCE over the distance matrix, forward and backward, without torch passes over [n,m]
 * (src/train.py:1832-1844 rows only; 2304-2334 symmetric).  sim = -D * inv_tau.
 * hypret_pairdist_ce_fwd: dmat [n,m] = hypret_pairdist with an fp32 arccosh tail (training accuracy), row_lse [n] =
 *   logsumexp_j sim_ij, and (want_col_lse) col_lse [m] = logsumexp_i sim_ij; scratch [2*n_part*m] fp32.  The loss is
 *   mean_i(row_lse_i - sim_ii) (+ mean_j(col_lse_j - sim_jj), halved) -- O(n) work left to the caller.
 * hypret_pairdist_ce_bwd: like hypret_pairdist_bwd, but the upstream gradient is formed on the fly:
 *   g_ij = -(gs * inv_tau / n) [ w_rows (exp(sim_ij - row_lse_i) - [i==j]) + w_cols (exp(sim_ij - col_lse_j) - [i==j]) ]
 *   with gs = *grad_scale (device scalar, NULL = 1).  col_partial [ceil(n/16), m]; row_partial [n_row_partial, n] as above.
 *   Row block of a larger batch (negatives sharded across ranks): the target of row i is column i + diag_offset and
 *   the loss is a mean over n_total >= n rows (single process: diag_offset = 0, n_total = n). */
int hypret_pairdist_ce_fwd(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float inv_tau,
                           int want_col_lse, float* dmat, float* row_lse, float* col_lse, float* scratch,
                           int n_part, void* stream);
int hypret_pairdist_ce_bwd(const float* dmat, const float* asq, const float* psq, int64_t n, int64_t m, float c,
                           const float* row_lse, const float* col_lse, float inv_tau, float w_rows, float w_cols,
                           const float* grad_scale, void* w_out, int w_format, float* row_partial, int n_row_partial,
                           float* col_partial, int64_t diag_offset, int64_t n_total, void* stream);

/* The same n x m distance matrix on the TENSOR CORES (training path, csrc/gramdist.cu): one tcgen05 GEMM over
 * operands that carry a 3-way bf16 split of the fp32 rows along K (six cross products = the fp32 inner product),
 * distance formed in the epilogue, near pairs (|a-p|^2 < (|a|^2+|p|^2)/4, i.e. where |a|^2+|p|^2-2<a,p> cancels)
 * recomputed exactly from the fp32 rows.  Accuracy ~1e-6 relative (fp32 accumulation; fp32 arccosh tail).
 *   hypret_gram_kpad(d)   operand row length = roundup(6 d, 64) bf16 elements
 *   hypret_gram_split     x [n,d] fp32 -> operand [n,kpad] bf16 for side 0 (rows of a) / 1 (rows of p), sqnorm [n]
 *   hypret_gram_dist      out [n,m] fp32 from the two operands, the fp32 rows and the squared norms
 *   hypret_neg_lse        row / column log-sum-exps of -dmat * inv_tau (the second half of hypret_pairdist_ce_fwd) */
int64_t hypret_gram_kpad(int d);
/* x [count] fp32 -> out [3, count] bf16 planes (hi, mid, lo; hi + mid + lo = x to fp32 accuracy): the W format
 * (w_format 1) of the two functions above, cut from an fp32 W in one streaming pass -- faster than emitting the
 * planes from inside the backward pass.  count % 4 == 0. */
int hypret_split3(const float* x, int64_t count, void* out_bf16, void* stream);
int hypret_gram_split(const float* x, int64_t n, int d, int side, void* out_bf16, float* sqnorm, void* stream);
int hypret_gram_dist(const void* a_op, const void* p_op, const float* a32, const float* p32, const float* asq,
                     const float* psq, int64_t n, int64_t m, int d, float c, float* out, void* stream);
int hypret_neg_lse(const float* dmat, int64_t n, int64_t m, float inv_tau, int want_col_lse, float* row_lse,
                   float* col_lse, float* scratch, int n_part, void* stream);

/* Metrics of ranked lists, restating the per-query loops of notebooks/retrieval.ipynb:310-324
 * (MRR@k, Precision@k), :411-420 (AP), :430-437 (nDCG), :439-443 (Recall@k), :446-456 (means).
 *   ranked       [Q,K] int64 gallery ids, best first, -1 padded
 *   pos_offsets  [Q+1] int64, pos_items [nnz] int64: CSR of each query's positives in the gallery
 *   n_pos_total  [Q] int32 or NULL: |P| including positives absent from the gallery (NULL: CSR count)
 *   ks_host      HOST array of n_ks (<= 8) cut-offs, e.g. {5, 10, 20}
 *   per_query    [Q, 3 + 3*n_ks] fp64: mrr, ap, ndcg, then (mrr@k, precision@k, recall@k) per k
 *   means        [3 + 3*n_ks] fp64 or NULL: column means in a fixed (deterministic) order */
int hypret_retrieval_metrics(const int64_t* ranked, int64_t Q, int K, const int64_t* pos_offsets,
                             const int64_t* pos_items, const int32_t* n_pos_total, const int32_t* ks_host, int n_ks,
                             double* per_query, double* means, void* stream);

/* Average precision over FULL score rows (higher = better) by rank counting, no sort.
 * grouped_ties != 0: sklearn.average_precision_score semantics (src/train.py:3285,
 * src/auxiliary.py:200-224); == 0: ranking order with lower-index tie-break
 * (notebooks/retrieval.ipynb:411-420).  Rows without an in-range positive or with a NaN/inf
 * score get valid = 0 and are left out of the mean (src/train.py:3244-3262).
 *   scores [Q,N] fp32, ap [Q] fp64, valid [Q] int32, mean_ap [1] fp64 or NULL */
int hypret_ap_full(const float* scores, int64_t Q, int64_t N, const int64_t* pos_offsets, const int64_t* pos_items,
                   int grouped_ties, double* ap, int32_t* valid, double* mean_ap, void* stream);

/* Exact full-ranking AP WITHOUT the [Q,N] score matrix, shardable over gallery rows (multi-GPU "collective 2").
 * Same reference semantics as hypret_ap_full (src/train.py:3259-3293; notebooks/retrieval.ipynb:383,411-420), from
 * rank counts per (query, positive) pair.  key = Poincare distance (metric HYPERBOLIC, the arithmetic of
 * hypret_pairdist bit for bit) or minus cosine similarity (metric COSINE): smaller = better.
 *   q32 [Q,d] fp32 (on-ball points / raw features), g32 [n_local,d] fp32: this shard's gallery rows, whose global
 *   ids are idx_offset .. idx_offset+n_local-1;  pos_offsets [Q+1], pos_items [nnz]: CSR of GLOBAL positive ids.
 * 1. hypret_pair_keys: keys[t] = key(query of t, positive t) if this shard owns the positive, else 0
 *    -> sum over shards (all-reduce) gives every shard all keys.
 * 2. hypret_rank_count: ADDS to counts[t*3 + {0,1,2}] the number of this shard's rows that score strictly better
 *    than positive t / tie with it and have a lower global id / tie with it (itself included on the owner), and to
 *    bad[q] the number of non-finite scores of query q (queries with >= 1 positive).  Caller zeroes counts [nnz*3]
 *    (uint64) and bad [Q] (int32) first -> sum over shards (all-reduce).
 * 3. hypret_ap_from_counts: ap [Q] fp64, valid [Q] int32 (0: no in-range positive or a non-finite score), mean_ap
 *    [1] fp64 or NULL over the valid queries.  grouped_ties as in hypret_ap_full. */
int hypret_pair_keys(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                     const int64_t* pos_offsets, const int64_t* pos_items, int64_t idx_offset, float* keys,
                     void* stream);
int hypret_rank_count(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                      const int64_t* pos_offsets, const int64_t* pos_items, const float* pos_keys, int64_t idx_offset,
                      uint64_t* counts, int32_t* bad, void* stream);
int hypret_ap_from_counts(const int64_t* pos_offsets, const int64_t* pos_items, const float* pos_keys,
                          const uint64_t* counts, const int32_t* bad, int64_t Q, int64_t n_total, int grouped_ties,
                          double* ap, int32_t* valid, double* mean_ap, void* stream);

/* ---- Flash-style train_hyp step (csrc/flash.cu): the in-batch InfoNCE over the n x m Poincare distance matrix, forward
 * and backward, WITHOUT ever writing the matrix (or any other [n,m] array) to memory, all dense products on tcgen05.
 * Replaces the O(n^2) double loop of 1x1 pmath.dist + autograd (src/train.py:1832-1846; symmetric: 2304-2334).
 * hypret_flash_lse / hypret_flash_grad: 16 <= d <= 128, d % 16 == 0 (larger d: hypret_gram_dist / hypret_pairdist_ce_*);
 * hypret_flash_prep itself serves any d % 4 == 0 (its operands also feed hypret_mobius_gemm).
 *   hypret_flash_kpad(d)      Gram operand row length: roundup(3 d, 64) fp16 elements
 *   hypret_flash_workspace    fp32 elements of workspace hypret_flash_lse / hypret_flash_grad need for n x m
 *   hypret_flash_prep         x [n,d] fp32 -> row_op [n,kpad] fp16 ([hi|lo|hi]), col_op [n,kpad] fp16 ([hi|hi|lo]),
 *                             t_planes [2, d, t_cols] bf16 (transposed hi / mid planes; t_cols >= n, % 8 == 0, columns
 *                             >= n zero), sqnorm [n]; any output may be NULL
 *   hypret_flash_lse          lse_out[i] = logsumexp_j(-dist(x_i, y_j) * inv_tau), i < n, j < m
 *   hypret_flash_grad         dx_out [n,d] = d loss / d x for
 *                               loss = (gs / n_total) sum_i [ w_rows (x_lse_i - sim_i,t(i)) + w_cols (y_lse_t(i) - sim_i,t(i)) ],
 *                             sim = -dist * inv_tau, t(i) = i + diag_offset (rows without a target column in [0,m)
 *                             contribute only their softmax terms), gs = *grad_scale (NULL = 1); x_lse [n] / y_lse [m]
 *                             natural-log log-sum-exps over the columns of row i / over the rows of column j.
 *                             One call per operand: d loss / d y is the same call with the roles (and w_rows / w_cols,
 *                             the two lse arrays, the sign of diag_offset) swapped. */
/* out[j] = ln sum_r exp(parts[r, j]), parts [n_parts, n] fp32: the column log-sum-exps of a batch whose rows are
 * sharded across ranks, from the per-rank partials (train.ShardedInBatchInfoNCE). */
int hypret_lse_combine(const float* parts, int n_parts, int64_t n, float* out, void* stream);
/* out[j] = sum_r parts[r, j] in a fixed order: folds the per-CTA row / column partial sums of hypret_pairdist_bwd /
 * hypret_pairdist_ce_bwd (the terms the reference's autograd adds up in its diagonal corrections). */
int hypret_sum_parts(const float* parts, int n_parts, int64_t n, float* out, void* stream);
int64_t hypret_flash_kpad(int d);
int64_t hypret_flash_workspace(int64_t n, int64_t m, int d);
int hypret_flash_prep(const float* x, int64_t n, int d, void* row_op, void* col_op, void* t_planes, int64_t t_cols,
                      float* sqnorm, void* stream);
int hypret_flash_lse(const void* x_row_op, const void* y_col_op, const float* x32, const float* y32, const float* xsq,
                     const float* ysq, int64_t n, int64_t m, int d, float c, float inv_tau, float* workspace,
                     float* lse_out, void* stream);
int hypret_flash_grad(const void* x_row_op, const void* y_col_op, const void* y_t_planes, int64_t t_cols,
                      const float* x32, const float* y32, const float* xsq, const float* ysq, const float* x_lse,
                      const float* y_lse, int64_t n, int64_t m, int d, float c, float inv_tau, float w_rows,
                      float w_cols, const float* grad_scale, int64_t diag_offset, int64_t n_total, float* workspace,
                      float* dx_out, void* stream);

/* ---- Row-local manifold kernels of the train_hyp step (csrc/manifold.cu) ------------------------------------------
 * hypret_rowpair_dist      out[t] = pmath.dist(x[ia[t]], y[ib[t]]): the per-pair Python loops of
 *                          calculate_pair_loss (src/models.py:712-719, 824-829), the re-encode loop
 *                          (src/train.py:1433-1443) and the negatives of sample_to_prototype_loss (src/train.py:1036),
 *                          batched.  x [nx,d], y [ny,d] fp32 on-ball points, ia / ib [n_pairs] int64.
 * hypret_rowpair_dist_bwd  ADDS grad_out[t] * dd_t/dx into grad_x[ia[t]] (and /dy into grad_y[ib[t]]); either may be
 *                          NULL; caller zeroes them.
 * hypret_hmi_pairs         HMI insideness (mode 0) / disjointedness (mode 1) of label pairs, pairs [n_pairs,2] int64
 *                          rows of emb [L,d] (src/models.py:630-674), and the hinge loss around them (:550-604);
 *                          proj_eps = geoopt's projx margin for the caller's dtype (4e-3 fp32, 1e-5 fp64 parameters):
 *                          values [n_pairs] or NULL; *loss_sum += sum_t relu(margin - v_t) (fp64, caller zeroes; the
 *                          loss is loss_sum / n_pairs) or NULL; grad_emb [L,d] or NULL: ADDS *grad_scale (NULL = 1)
 *                          times d(mean hinge)/d emb.
 * hypret_dist0_reg         dist0 regulariser (src/models.py:606-628): *loss_sum += sum_i relu(lo - d0_i) + relu(d0_i - hi)
 *                          (lo < 0: upper hinge only), loss = loss_sum / n; grad_x [n,d] or NULL: ADDS the gradient of
 *                          the mean.
 * hypret_radam_ball_step   one geoopt.optim.RiemannianAdam step (src/train.py:1362) on a [n,d] ManifoldParameter of the
 *                          Poincare ball, in place: egrad2rgrad, moments, retraction project(x - lr dir), parallel
 *                          transport of exp_avg.  exp_avg_sq [n,d] holds one value per row (geoopt's broadcast layout).
 *                          step >= 1 is the step count AFTER this update. */
int hypret_rowpair_dist(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs, int d,
                        float c, float* out, void* stream);
int hypret_rowpair_dist_bwd(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs,
                            int d, float c, const float* grad_out, float* grad_x, float* grad_y, void* stream);
int hypret_hmi_pairs(const float* emb, const int64_t* pairs, int64_t n_pairs, int d, float c, int mode, float margin,
                     float proj_eps, float* values, double* loss_sum, const float* grad_scale, float* grad_emb,
                     void* stream);
int hypret_dist0_reg(const float* x, int64_t n, int d, float c, float lo, float hi, double* loss_sum,
                     const float* grad_scale, float* grad_x, void* stream);
int hypret_radam_ball_step(float* x, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int d, float c,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYPRET_H_ */
