// extern "C" boundary of libhypret.so -- argument validation and dispatch only.
// Signatures and the reference call sites they replace are documented in include/hypret.h.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

int check_device() {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  // Driver-API calls (cuTensorMapEncodeTiled) need the primary context CURRENT on the calling thread; a thread that
  // has only queried the runtime so far (autograd's backward thread) has none yet -> CUDA_ERROR_INVALID_CONTEXT.
  static thread_local int bound_dev = -1;
  if (bound_dev != dev) {
    e = cudaFree(nullptr);
    if (e != cudaSuccess) return (int)e;
    bound_dev = dev;
  }
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return (int)e;
  return major == 10 ? HYPRET_OK : HYPRET_ENOTSM100;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

const char* hypret_strerror(int rc) {
  switch (rc) {
    case HYPRET_OK: return "ok";
    case HYPRET_EINVAL: return "hypret: invalid argument (shape, alignment or enum)";
    case HYPRET_EUNSUPPORTED: return "hypret: request not supported by this build (size or dimension limit)";
    case HYPRET_EDRIVER: return "hypret: CUDA driver entry point cuTensorMapEncodeTiled unavailable";
    case HYPRET_ENOTSM100: return "hypret: device is not compute capability 10.x (sm_100a kernels only; no fallback)";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
  return "hypret: unknown error";
}

int hypret_version(void) { return 100; }

int64_t hypret_operand_kpad(int d) { return d > 0 ? (int64_t)hypret_kpad(d) : 0; }

static int project_rows_common(const float* u, int64_t n, int d, float c, int mode, int side, float* y32,
                               void* op_f16, float* sqnorm, float* op_err, float* stats, void* stream);

int hypret_project_rows(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, void* op_f16,
                        float* sqnorm, void* stream) {
  return project_rows_common(u, n, d, c, mode, side, y32, op_f16, sqnorm, nullptr, nullptr, stream);
}

int hypret_project_rows_cert(const float* u, int64_t n, int d, float c, int mode, int side, float* y32, void* op_f16,
                             float* sqnorm, float* op_err, float* stats, void* stream) {
  if ((op_err != nullptr || stats != nullptr) && op_f16 == nullptr) return HYPRET_EINVAL;
  return project_rows_common(u, n, d, c, mode, side, y32, op_f16, sqnorm, op_err, stats, stream);
}

static int project_rows_common(const float* u, int64_t n, int d, float c, int mode, int side, float* y32,
                               void* op_f16, float* sqnorm, float* op_err, float* stats, void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 2048) return HYPRET_EINVAL;
  if (mode < HYPRET_MODE_EXPMAP0 || mode > HYPRET_MODE_COSINE) return HYPRET_EINVAL;
  if (side != HYPRET_SIDE_QUERY && side != HYPRET_SIDE_GALLERY) return HYPRET_EINVAL;
  if (mode != HYPRET_MODE_COSINE && !(c > 0.f)) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (u == nullptr || !aligned16(u) || !aligned16(y32) || !aligned16(op_f16)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_project_rows(u, n, d, c, mode, side, y32, op_f16, sqnorm, op_err, stats,
                                    static_cast<cudaStream_t>(stream));
}

int hypret_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out_host) {
  if (bytes == 0 || dev_ptr == nullptr || handle_out_host == nullptr) return HYPRET_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == HYPRET_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle_out_host, &h, sizeof(h));
  *dev_ptr = p;
  return HYPRET_OK;
}

int hypret_peer_free(void* dev_ptr) { return dev_ptr == nullptr ? HYPRET_OK : (int)cudaFree(dev_ptr); }

int hypret_peer_open(const void* handle_host, void** peer_ptr) {
  if (handle_host == nullptr || peer_ptr == nullptr) return HYPRET_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  return (int)cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int hypret_peer_close(void* peer_ptr) { return peer_ptr == nullptr ? HYPRET_OK : (int)cudaIpcCloseMemHandle(peer_ptr); }

int hypret_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return HYPRET_OK;
  if (dst == nullptr || src == nullptr) return HYPRET_EINVAL;
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
}

int hypret_project_rows_peers(const float* u, int64_t n, int d, float c, int mode, float* y32,
                              void* const* op_dsts_host, int n_dst, float* op_err, void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 2048 || n_dst < 1 || n_dst > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  if (mode < HYPRET_MODE_EXPMAP0 || mode > HYPRET_MODE_COSINE) return HYPRET_EINVAL;
  if (mode != HYPRET_MODE_COSINE && !(c > 0.f)) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (u == nullptr || op_dsts_host == nullptr || !aligned16(u) || !aligned16(y32)) return HYPRET_EINVAL;
  for (int i = 0; i < n_dst; ++i)
    if (op_dsts_host[i] == nullptr || !aligned16(op_dsts_host[i])) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_project_rows_peers(u, n, d, c, mode, y32, op_dsts_host, n_dst, op_err,
                                          static_cast<cudaStream_t>(stream));
}

int hypret_peer_signal(void* const* flags_host, int n, uint32_t value, void* stream) {
  if (flags_host == nullptr || n < 1 || n > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  for (int i = 0; i < n; ++i)
    if (flags_host[i] == nullptr) return HYPRET_EINVAL;
  return hypret_launch_peer_signal(flags_host, n, value, static_cast<cudaStream_t>(stream));
}

int hypret_peer_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, void* stream) {
  if (flags == nullptr || n < 1 || n > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  return hypret_launch_peer_wait(flags, n, value, err, static_cast<cudaStream_t>(stream));
}

int hypret_score_topk_bound(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int kbound,
                            int n_lists, int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx,
                            uint32_t* thr_workspace, int32_t* list_count, float* debug_scores, void* stream) {
  if (Q < 1 || N < 1 || d < 4 || (d & 3) || d > 2048) return HYPRET_EINVAL;
  if (kprime < 1 || kprime > 64 || n_lists < 1 || max_ctas < 0 || min_lists < 0) return HYPRET_EINVAL;
  if (kbound < kprime || (kbound > kprime && (kprime > 16 || kbound > 32 || thr_workspace == nullptr)))
    return HYPRET_EINVAL;
  if (q_op == nullptr || g_op == nullptr || cand_score == nullptr || cand_idx == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_score_topk(q_op, Q, g_op, N, d, kprime, kbound, n_lists, max_ctas, min_lists, cand_score,
                                  cand_idx, thr_workspace, list_count, debug_scores, static_cast<cudaStream_t>(stream));
}

int hypret_score_topk(const void* q_op, int64_t Q, const void* g_op, int64_t N, int d, int kprime, int n_lists,
                      int max_ctas, int min_lists, float* cand_score, int32_t* cand_idx, uint32_t* thr_workspace,
                      int32_t* list_count, float* debug_scores, void* stream) {
  return hypret_score_topk_bound(q_op, Q, g_op, N, d, kprime, kprime, n_lists, max_ctas, min_lists, cand_score,
                                 cand_idx, thr_workspace, list_count, debug_scores, stream);
}

static int rerank_common(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_lists,
                         int kprime, int k,
                         int64_t idx_offset, const float* prune_thr, float* out_score, int64_t* out_idx,
                         float* out_margin, void* stream, const hypret_peer_route* route = nullptr,
                         int64_t score_off = 0, int64_t idx_off = 0, const double* g_sq64 = nullptr,
                         const CertArgs& cert = no_cert()) {
  if (Q < 0 || N < 1 || d < 4 || (d & 3)) return HYPRET_EINVAL;
  const bool routed = route != nullptr && route->n_ranks > 0;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (kprime < 1 || kprime > 64 || k < 1 || k > 128 || n_lists < 1) return HYPRET_EINVAL;
  if (k > kprime * n_lists || (int64_t)n_lists * kprime > 16384) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (q32 == nullptr || g32 == nullptr || cand_score == nullptr || cand_idx == nullptr ||
      (!routed && (out_score == nullptr || out_idx == nullptr)) || !aligned16(q32) || !aligned16(g32))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_rerank(q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_lists * kprime,
                              kprime, k, idx_offset, prune_thr, out_score, out_idx, out_margin, route, score_off,
                              idx_off, g_sq64, cert, static_cast<cudaStream_t>(stream));
}

static int check_route(const hypret_peer_route* route, int64_t Q, bool rows_are_own) {
  if (route == nullptr || route->n_ranks < 1 || route->n_ranks > HYPRET_MAX_PEERS || route->me < 0 ||
      route->me >= route->n_ranks || route->ql < 1)
    return HYPRET_EINVAL;
  if (Q != (rows_are_own ? route->ql : route->ql * route->n_ranks)) return HYPRET_EINVAL;
  for (int i = 0; i < route->n_ranks; ++i)
    if (route->base[i] == nullptr) return HYPRET_EINVAL;
  return HYPRET_OK;
}

int hypret_rerank(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                  const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_lists, int kprime,
                  int k, int64_t idx_offset, float* out_score, int64_t* out_idx, float* out_margin,
                  const double* g_sqnorm64, void* stream) {
  return rerank_common(q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_lists, kprime, k, idx_offset,
                       nullptr, out_score, out_idx, out_margin, stream, nullptr, 0, 0, g_sqnorm64);
}

int hypret_rerank_cert(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                       const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int n_lists,
                       int kprime, int ksel, int k, int64_t idx_offset, float* out_score, int64_t* out_idx,
                       float* out_margin, const float* q_err, const float* g_stats, int32_t* fb_state, int32_t* fb_count,
                       int32_t* fb_list, float* fb_bound, uint8_t* certified, void* stream) {
  if (kprime > 32 || k > 32 || k > kprime) return HYPRET_EUNSUPPORTED;
  if (ksel != 0 && (ksel < kprime || ksel > 32)) return HYPRET_EINVAL;
  if (q_err == nullptr || g_stats == nullptr || fb_state == nullptr || fb_count == nullptr || fb_list == nullptr)
    return HYPRET_EINVAL;
  CertArgs cert;
  cert.q_err = q_err; cert.g_stats = g_stats; cert.state = fb_state; cert.count = fb_count; cert.list = fb_list;
  cert.flags = certified;
  cert.ksel = ksel;
  cert.bound = fb_bound;
  // fp32 accumulation in the tensor core (truncating: up to 2 ulps of the running sum per 16-deep MMA step) and the
  // 2^-24 tails of the 3-way splits of the extension columns
  cert.slack = (float)(hypret_kpad(d) / 16 + 8) * 2.384185791015625e-07f;
  cudaError_t e = cudaMemsetAsync(fb_count, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return (int)e;
  return rerank_common(q32, g32, Q, N, d, c, metric, cand_score, cand_idx, list_count, n_lists, kprime, k, idx_offset,
                       nullptr, out_score, out_idx, out_margin, stream, nullptr, 0, 0, nullptr, cert);
}

int hypret_exact_topk(const float* q32, const float* g32, const double* g_sqnorm64, int64_t Q, int64_t N, int d,
                      float c, int metric, int k, int64_t idx_offset, const int32_t* q_list, const int32_t* q_count,
                      int32_t* fb_state, const float* init_bound, float* out_score, int64_t* out_idx, void* stream) {
  if (Q < 0 || N < 1 || d < 4 || (d & 3) || k < 1 || k > 32) return HYPRET_EINVAL;
  if (N > 0x7fffffffll) return HYPRET_EUNSUPPORTED;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (q32 == nullptr || g32 == nullptr || g_sqnorm64 == nullptr || q_list == nullptr || q_count == nullptr ||
      fb_state == nullptr || out_score == nullptr || out_idx == nullptr || !aligned16(q32) || !aligned16(g32))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_exact_topk(q32, g32, g_sqnorm64, Q, N, d, c, metric, k, idx_offset, q_list, q_count, fb_state,
                                  out_score, out_idx, nullptr, init_bound, static_cast<cudaStream_t>(stream));
}

int hypret_exact_topk_after(const float* q32, const float* g32, const double* g_sqnorm64, int64_t Q, int64_t N, int d,
                            float c, int metric, int k, int64_t idx_offset, const int32_t* q_list,
                            const int32_t* q_count, int32_t* fb_state, const uint64_t* after, float* out_score,
                            int64_t* out_idx, void* stream) {
  if (Q < 0 || N < 1 || d < 4 || (d & 3) || k < 1 || k > 32 || N > 0x7fffffffll) return HYPRET_EINVAL;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (q32 == nullptr || g32 == nullptr || g_sqnorm64 == nullptr || q_list == nullptr || q_count == nullptr ||
      fb_state == nullptr || after == nullptr || out_score == nullptr || out_idx == nullptr || !aligned16(q32) ||
      !aligned16(g32))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_exact_topk(q32, g32, g_sqnorm64, Q, N, d, c, metric, k, idx_offset, q_list, q_count, fb_state,
                                  out_score, out_idx, reinterpret_cast<const unsigned long long*>(after), nullptr,
                                  static_cast<cudaStream_t>(stream));
}

int hypret_cert_merged(const float* q32, int64_t Q, int d, float c, int metric, const float* score, const int64_t* idx,
                       int k, const float* thr, const float* q_err, const float* g_stats, int32_t* flags,
                       float* out_margin, void* stream) {
  if (Q < 0 || d < 4 || (d & 3) || k < 1) return HYPRET_EINVAL;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (q32 == nullptr || score == nullptr || idx == nullptr || thr == nullptr || q_err == nullptr ||
      g_stats == nullptr || flags == nullptr || !aligned16(q32))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  const float slack = (float)(hypret_kpad(d) / 16 + 8) * 2.384185791015625e-07f;     // as hypret_rerank_cert
  return hypret_launch_cert_merged(q32, Q, d, c, metric, score, idx, k, thr, q_err, g_stats, slack, flags, out_margin,
                                   static_cast<cudaStream_t>(stream));
}

int hypret_flag_compact(const int32_t* flags, int64_t n, int32_t* list, int32_t* count, int32_t* fb_state,
                        void* stream) {
  if (n < 0 || n > 0x7fffffffll) return HYPRET_EINVAL;
  if (count == nullptr || (n > 0 && (flags == nullptr || list == nullptr || fb_state == nullptr))) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_flag_compact(flags, n, list, count, fb_state, static_cast<cudaStream_t>(stream));
}

int hypret_row_sqnorm64(const float* x, int64_t n, int d, double* out, void* stream) {
  if (n < 0 || d < 4 || (d & 3)) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x == nullptr || out == nullptr || !aligned16(x)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_row_sqnorm64(x, n, d, out, static_cast<cudaStream_t>(stream));
}

int hypret_rerank_pruned(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                         const float* cand_score, const int32_t* cand_idx, int n_lists, int kprime, int k,
                         int64_t idx_offset, const float* prune_thr, float* out_score, int64_t* out_idx,
                         void* stream) {
  if (prune_thr == nullptr || kprime > 32 || k > 32 || k > kprime) return HYPRET_EINVAL;
  return rerank_common(q32, g32, Q, N, d, c, metric, cand_score, cand_idx, nullptr, n_lists, kprime, k, idx_offset,
                       prune_thr, out_score, out_idx, nullptr, stream);
}

static int cand_select_common(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                              int n_lists, int kprime, float* sel_score, int32_t* sel_idx,
                              const hypret_peer_route* route, int64_t recv_off, void* stream) {
  if (Q < 0 || n_lists < 1 || kprime < 1 || kprime > 32 || (int64_t)n_lists * kprime > 16384) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (cand_score == nullptr || cand_idx == nullptr || sel_score == nullptr || sel_idx == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_cand_select(cand_score, cand_idx, list_count, Q, n_lists * kprime, kprime, sel_score, sel_idx,
                                   route, recv_off, static_cast<cudaStream_t>(stream));
}

int hypret_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                       int n_lists, int kprime, float* sel_score, int32_t* sel_idx, void* stream) {
  return cand_select_common(cand_score, cand_idx, list_count, Q, n_lists, kprime, sel_score, sel_idx, nullptr, 0,
                            stream);
}

int hypret_cand_select_route(const float* cand_score, const int32_t* cand_idx, const int32_t* list_count, int64_t Q,
                             int n_lists, int kprime, float* sel_score, int32_t* sel_idx,
                             const hypret_peer_route* route, int64_t recv_off, void* stream) {
  if (Q > 0 && (check_route(route, Q, false) != HYPRET_OK || recv_off < 0 || (recv_off & 3))) return HYPRET_EINVAL;
  return cand_select_common(cand_score, cand_idx, list_count, Q, n_lists, kprime, sel_score, sel_idx, route, recv_off,
                            stream);
}

static int kth_smallest_common(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out,
                               const hypret_peer_route* route, int64_t out_off, void* stream) {
  if (n_parts < 1 || Q < 0 || m < 1 || kth < 1 || (int64_t)n_parts * m > 2048) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (vals == nullptr || out == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_kth_smallest(vals, n_parts, Q, m, kth, out, route, out_off, static_cast<cudaStream_t>(stream));
}

int hypret_kth_smallest(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out, void* stream) {
  return kth_smallest_common(vals, n_parts, Q, m, kth, out, nullptr, 0, stream);
}

int hypret_kth_smallest_route(const float* vals, int n_parts, int64_t Q, int m, int kth, float* out,
                              const hypret_peer_route* route, int64_t out_off, void* stream) {
  if (Q > 0 && (check_route(route, Q, true) != HYPRET_OK || out_off < 0 || (out_off & 3))) return HYPRET_EINVAL;
  return kth_smallest_common(vals, n_parts, Q, m, kth, out, route, out_off, stream);
}

int hypret_rerank_pruned_route(const float* q32, const float* g32, int64_t Q, int64_t N, int d, float c, int metric,
                               const float* cand_score, const int32_t* cand_idx, int n_lists, int kprime, int k,
                               int64_t idx_offset, const float* prune_thr, const hypret_peer_route* route,
                               int64_t score_off, int64_t idx_off, void* stream) {
  if (prune_thr == nullptr || kprime > 32 || k > 32 || k > kprime) return HYPRET_EINVAL;
  if (Q > 0 && (check_route(route, Q, false) != HYPRET_OK || score_off < 0 || idx_off < 0 || (score_off & 3) ||
                (idx_off & 7)))
    return HYPRET_EINVAL;
  return rerank_common(q32, g32, Q, N, d, c, metric, cand_score, cand_idx, nullptr, n_lists, kprime, k, idx_offset,
                       prune_thr, nullptr, nullptr, nullptr, stream, route, score_off, idx_off);
}

int hypret_merge_topk(const float* scores, const int64_t* idx, int n_shards, int64_t Q, int k, int descending,
                      float* out_score, int64_t* out_idx, void* stream) {
  if (n_shards < 1 || Q < 0 || k < 1 || k > 32 || n_shards * k > 256) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (scores == nullptr || idx == nullptr || out_score == nullptr || out_idx == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_merge_topk(scores, idx, n_shards, Q, k, descending != 0, out_score, out_idx,
                                  static_cast<cudaStream_t>(stream));
}

int hypret_mobius_epilogue(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                           int hyperbolic_input, int post_tanh, int n_project, float* y, float* sqnorm, void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 2048 || !(c > 0.f) || n_project < 0 || n_project > 2) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (mx == nullptr || y == nullptr || !aligned16(mx) || !aligned16(y) || !aligned16(bias)) return HYPRET_EINVAL;
  if (hyperbolic_input && xsq == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_mobius_epilogue(mx, n, d, xsq, bias, c, hyperbolic_input != 0, post_tanh != 0, n_project, y,
                                       sqnorm, static_cast<cudaStream_t>(stream));
}

int hypret_pairdist(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float* out, void* stream) {
  if (n < 0 || m < 0 || d < 4 || (d & 3) || !(c > 0.f)) return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (a == nullptr || p == nullptr || out == nullptr || !aligned16(a) || !aligned16(p)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_pairdist(a, p, n, m, d, c, out, static_cast<cudaStream_t>(stream));
}

int hypret_pairdist_bwd(const float* grad_out, const float* dmat, const float* asq, const float* psq, int64_t n,
                        int64_t m, float c, void* w_out, int w_format, float* row_partial, int n_row_partial,
                        float* col_partial, int n_partial, void* stream) {
  if (n < 0 || m < 0 || !(c > 0.f) || n_partial < 1 || n_partial > 65535 || (w_format != 0 && w_format != 1) ||
      n_row_partial < 1 || n_row_partial > 65535)
    return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (grad_out == nullptr || dmat == nullptr || asq == nullptr || psq == nullptr || w_out == nullptr ||
      row_partial == nullptr || col_partial == nullptr)
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_pairdist_bwd(grad_out, dmat, asq, psq, n, m, c, w_out, w_format, row_partial, n_row_partial,
                                    col_partial, n_partial, static_cast<cudaStream_t>(stream));
}

int hypret_pairdist_ce_fwd(const float* a, const float* p, int64_t n, int64_t m, int d, float c, float inv_tau,
                           int want_col_lse, float* dmat, float* row_lse, float* col_lse, float* scratch,
                           int n_part, void* stream) {
  if (n < 0 || m < 0 || d < 4 || (d & 3) || !(c > 0.f) || !(inv_tau > 0.f) || n_part < 1 || n_part > 65535)
    return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (a == nullptr || p == nullptr || dmat == nullptr || row_lse == nullptr || !aligned16(a) || !aligned16(p) ||
      !aligned16(dmat))
    return HYPRET_EINVAL;
  if (want_col_lse && (col_lse == nullptr || scratch == nullptr)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_pairdist_ce_fwd(a, p, n, m, d, c, inv_tau, want_col_lse, dmat, row_lse, col_lse, scratch, n_part,
                                       static_cast<cudaStream_t>(stream));
}

int hypret_split3(const float* x, int64_t count, void* out_bf16, void* stream) {
  if (count < 0) return HYPRET_EINVAL;
  if (count == 0) return HYPRET_OK;
  if (x == nullptr || out_bf16 == nullptr || !aligned16(x) || !aligned16(out_bf16) || (count & 3)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_split3(x, count, out_bf16, static_cast<cudaStream_t>(stream));
}

int64_t hypret_gram_kpad(int d) { return d > 0 ? hypret_gram_kpad_impl(d) : 0; }

int hypret_gram_split(const float* x, int64_t n, int d, int side, void* out_bf16, float* sqnorm, void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 1024 || (side != 0 && side != 1)) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x == nullptr || out_bf16 == nullptr || sqnorm == nullptr || !aligned16(out_bf16)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_gram_split(x, n, d, side, out_bf16, sqnorm, static_cast<cudaStream_t>(stream));
}

int hypret_gram_dist(const void* a_op, const void* p_op, const float* a32, const float* p32, const float* asq,
                     const float* psq, int64_t n, int64_t m, int d, float c, float* out, void* stream) {
  if (n < 0 || m < 0 || d < 4 || (d & 3) || d > 1024 || !(c > 0.f)) return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (a_op == nullptr || p_op == nullptr || a32 == nullptr || p32 == nullptr || asq == nullptr || psq == nullptr ||
      out == nullptr || !aligned16(a_op) || !aligned16(p_op) || !aligned16(a32) || !aligned16(p32) || !aligned16(out))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_gram_dist(a_op, p_op, a32, p32, asq, psq, n, m, d, c, out, static_cast<cudaStream_t>(stream));
}

int hypret_neg_lse(const float* dmat, int64_t n, int64_t m, float inv_tau, int want_col_lse, float* row_lse,
                   float* col_lse, float* scratch, int n_part, void* stream) {
  if (n < 0 || m < 0 || !(inv_tau > 0.f) || n_part < 1 || n_part > 65535) return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (dmat == nullptr || row_lse == nullptr || (want_col_lse && (col_lse == nullptr || scratch == nullptr)))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_neg_lse(dmat, n, m, inv_tau, want_col_lse, row_lse, col_lse, scratch, n_part,
                               static_cast<cudaStream_t>(stream));
}

int hypret_pairdist_ce_bwd(const float* dmat, const float* asq, const float* psq, int64_t n, int64_t m, float c,
                           const float* row_lse, const float* col_lse, float inv_tau, float w_rows, float w_cols,
                           const float* grad_scale, void* w_out, int w_format, float* row_partial, int n_row_partial,
                           float* col_partial, int64_t diag_offset, int64_t n_total, void* stream) {
  if (n < 0 || m < 0 || !(c > 0.f) || !(inv_tau > 0.f) || (w_format != 0 && w_format != 1) || n_row_partial < 1 ||
      n_row_partial > 65535 || diag_offset < 0 || n_total < n)
    return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (dmat == nullptr || asq == nullptr || psq == nullptr || row_lse == nullptr || w_out == nullptr ||
      row_partial == nullptr || col_partial == nullptr || (w_cols != 0.f && col_lse == nullptr))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_pairdist_ce_bwd(dmat, asq, psq, n, m, c, row_lse, col_lse, inv_tau, w_rows, w_cols, grad_scale,
                                       w_out, w_format, row_partial, n_row_partial, col_partial, diag_offset, n_total,
                                       static_cast<cudaStream_t>(stream));
}

int hypret_retrieval_metrics(const int64_t* ranked, int64_t Q, int K, const int64_t* pos_offsets,
                             const int64_t* pos_items, const int32_t* n_pos_total, const int32_t* ks_host, int n_ks,
                             double* per_query, double* means, void* stream) {
  if (Q < 0 || K < 1 || n_ks < 0 || n_ks > 8 || (n_ks > 0 && ks_host == nullptr)) return HYPRET_EINVAL;
  for (int i = 0; i < n_ks; ++i)
    if (ks_host[i] < 1) return HYPRET_EINVAL;
  if (ranked == nullptr || pos_offsets == nullptr || per_query == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_retrieval_metrics(ranked, Q, K, pos_offsets, pos_items, n_pos_total, ks_host, n_ks, per_query,
                                         means, static_cast<cudaStream_t>(stream));
}

int hypret_ap_full(const float* scores, int64_t Q, int64_t N, const int64_t* pos_offsets, const int64_t* pos_items,
                   int grouped_ties, double* ap, int32_t* valid, double* mean_ap, void* stream) {
  if (Q < 0 || N < 1) return HYPRET_EINVAL;
  if (scores == nullptr || pos_offsets == nullptr || ap == nullptr || valid == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_ap_full(scores, Q, N, pos_offsets, pos_items, grouped_ties != 0, ap, valid, mean_ap,
                               static_cast<cudaStream_t>(stream));
}

int hypret_pair_keys(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                     const int64_t* pos_offsets, const int64_t* pos_items, int64_t idx_offset, float* keys,
                     void* stream) {
  if (Q < 0 || n_local < 0 || d < 4 || (d & 3)) return HYPRET_EINVAL;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (Q == 0) return HYPRET_OK;
  if (q32 == nullptr || (g32 == nullptr && n_local > 0) || pos_offsets == nullptr || keys == nullptr)
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_pair_keys(q32, g32, Q, n_local, d, c, metric, pos_offsets, pos_items, idx_offset, keys,
                                 static_cast<cudaStream_t>(stream));
}

int hypret_rank_count(const float* q32, const float* g32, int64_t Q, int64_t n_local, int d, float c, int metric,
                      const int64_t* pos_offsets, const int64_t* pos_items, const float* pos_keys, int64_t idx_offset,
                      uint64_t* counts, int32_t* bad, void* stream) {
  if (Q < 0 || n_local < 0 || d < 4 || (d & 3)) return HYPRET_EINVAL;
  if (metric != HYPRET_METRIC_COSINE && metric != HYPRET_METRIC_HYPERBOLIC) return HYPRET_EINVAL;
  if (metric == HYPRET_METRIC_HYPERBOLIC && !(c > 0.f)) return HYPRET_EINVAL;
  if (Q == 0 || n_local == 0) return HYPRET_OK;
  if (q32 == nullptr || g32 == nullptr || pos_offsets == nullptr || pos_keys == nullptr || counts == nullptr ||
      bad == nullptr || !aligned16(q32) || !aligned16(g32))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_rank_count(q32, g32, Q, n_local, d, c, metric, pos_offsets, pos_items, pos_keys, idx_offset,
                                  reinterpret_cast<unsigned long long*>(counts), bad,
                                  static_cast<cudaStream_t>(stream));
}

int hypret_ap_from_counts(const int64_t* pos_offsets, const int64_t* pos_items, const float* pos_keys,
                          const uint64_t* counts, const int32_t* bad, int64_t Q, int64_t n_total, int grouped_ties,
                          double* ap, int32_t* valid, double* mean_ap, void* stream) {
  if (Q < 0 || n_total < 1) return HYPRET_EINVAL;
  if (pos_offsets == nullptr || ap == nullptr || valid == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_ap_from_counts(pos_offsets, pos_items, pos_keys,
                                      reinterpret_cast<const unsigned long long*>(counts), bad, Q, n_total,
                                      grouped_ties != 0, ap, valid, mean_ap, static_cast<cudaStream_t>(stream));
}

int64_t hypret_flash_kpad(int d) { return d > 0 ? hypret_flash_kpad_impl(d) : 0; }

int64_t hypret_flash_workspace(int64_t n, int64_t m, int d) {
  if (n < 1 || m < 1 || d < 1) return 0;
  return hypret_flash_workspace_floats(n, m, d);
}

int hypret_flash_prep(const float* x, int64_t n, int d, void* row_op, void* col_op, void* t_planes, int64_t t_cols,
                      float* sqnorm, void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 4096) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x == nullptr || !aligned16(row_op) || !aligned16(col_op) || !aligned16(t_planes)) return HYPRET_EINVAL;
  if (t_planes != nullptr && (t_cols < n || (t_cols & 7))) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_flash_prep(x, n, d, row_op, col_op, t_planes, t_cols, sqnorm, static_cast<cudaStream_t>(stream));
}

static int flash_common(int bwd, const void* x_row_op, const void* y_col_op, const void* y_t_planes, int64_t t_cols,
                        const float* x32, const float* y32, const float* xsq, const float* ysq, const float* x_lse,
                        const float* y_lse, int64_t n, int64_t m, int d, float c, float inv_tau, float wx, float wy,
                        const float* grad_scale, int64_t diag_offset, int64_t n_total, float* workspace, float* out,
                        void* stream) {
  if (n < 0 || m < 0 || d < 16 || (d & 15) || d > 128 || !(c > 0.f) || !(inv_tau > 0.f)) return HYPRET_EINVAL;
  if (n == 0 || m == 0) return HYPRET_OK;
  if (x_row_op == nullptr || y_col_op == nullptr || x32 == nullptr || y32 == nullptr || xsq == nullptr ||
      ysq == nullptr || workspace == nullptr || out == nullptr || !aligned16(x_row_op) || !aligned16(y_col_op) ||
      !aligned16(x32) || !aligned16(y32) || !aligned16(workspace) || !aligned16(out))
    return HYPRET_EINVAL;
  if (bwd && (y_t_planes == nullptr || !aligned16(y_t_planes) || t_cols < m || (t_cols & 7) || n_total < 1 ||
              (wx != 0.f && x_lse == nullptr) || (wy != 0.f && y_lse == nullptr)))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_flash(bwd, x_row_op, y_col_op, y_t_planes, t_cols, x32, y32, xsq, ysq, x_lse, y_lse, n, m, d, c,
                             inv_tau, wx, wy, grad_scale, diag_offset, n_total, workspace, out,
                             static_cast<cudaStream_t>(stream));
}

int hypret_flash_lse(const void* x_row_op, const void* y_col_op, const float* x32, const float* y32, const float* xsq,
                     const float* ysq, int64_t n, int64_t m, int d, float c, float inv_tau, float* workspace,
                     float* lse_out, void* stream) {
  return flash_common(0, x_row_op, y_col_op, nullptr, 0, x32, y32, xsq, ysq, nullptr, nullptr, n, m, d, c, inv_tau, 0.f,
                      0.f, nullptr, 0, 1, workspace, lse_out, stream);
}

int hypret_flash_grad(const void* x_row_op, const void* y_col_op, const void* y_t_planes, int64_t t_cols,
                      const float* x32, const float* y32, const float* xsq, const float* ysq, const float* x_lse,
                      const float* y_lse, int64_t n, int64_t m, int d, float c, float inv_tau, float w_rows,
                      float w_cols, const float* grad_scale, int64_t diag_offset, int64_t n_total, float* workspace,
                      float* dx_out, void* stream) {
  return flash_common(1, x_row_op, y_col_op, y_t_planes, t_cols, x32, y32, xsq, ysq, x_lse, y_lse, n, m, d, c, inv_tau,
                      w_rows, w_cols, grad_scale, diag_offset, n_total, workspace, dx_out, stream);
}

int hypret_mobius_gemm(const void* x_row_op, const void* w_col_op, int64_t n, int d_in, int n_out, const float* xsq,
                       const float* bias, float c, int post_tanh, int n_project, float* mx_out, float* y_out,
                       float* ysq_out, void* op_out, void* stream) {
  if (n < 0 || d_in < 4 || (d_in & 3) || d_in > 4096 || n_out < 16 || (n_out & 15) || n_out > 256 || !(c > 0.f) ||
      n_project < 0 || n_project > 2)
    return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x_row_op == nullptr || w_col_op == nullptr || !aligned16(x_row_op) || !aligned16(w_col_op) ||
      !aligned16(mx_out) || !aligned16(y_out) || !aligned16(op_out) ||
      (mx_out == nullptr && y_out == nullptr && op_out == nullptr && ysq_out == nullptr))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_mobius_gemm(x_row_op, w_col_op, n, d_in, n_out, xsq, bias, c, post_tanh != 0, n_project, mx_out,
                                   y_out, ysq_out, op_out, static_cast<cudaStream_t>(stream));
}

int hypret_mobius_epilogue_bwd(const float* mx, int64_t n, int d, const float* xsq, const float* bias, float c,
                               int post_tanh, int n_project, const float* gy, float* gmx, float* gbias, float* gxn,
                               void* stream) {
  if (n < 0 || d < 4 || (d & 3) || d > 512 || !(c > 0.f) || n_project < 0 || n_project > 2) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (mx == nullptr || gy == nullptr || gmx == nullptr || !aligned16(mx) || !aligned16(gy) || !aligned16(gmx) ||
      !aligned16(bias) || (gbias != nullptr && bias == nullptr) || (gxn != nullptr && xsq == nullptr))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_mobius_epilogue_bwd(mx, n, d, xsq, bias, c, post_tanh, n_project, gy, gmx, gbias, gxn,
                                           static_cast<cudaStream_t>(stream));
}

int hypret_sgemm_strided(const float* a, int64_t a_row_stride, int64_t a_col_stride, const float* b,
                         int64_t b_row_stride, int64_t b_col_stride, int m, int n, int k, const float* row_scale,
                         const float* addend, float* out, void* stream) {
  if (m < 0 || n < 0 || k < 0 || ((row_scale == nullptr) != (addend == nullptr))) return HYPRET_EINVAL;
  if (m == 0 || n == 0) return HYPRET_OK;
  if (out == nullptr || (k > 0 && (a == nullptr || b == nullptr))) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_sgemm_strided(a, a_row_stride, a_col_stride, b, b_row_stride, b_col_stride, m, n, k, row_scale,
                                     addend, out, static_cast<cudaStream_t>(stream));
}

int hypret_lse_combine(const float* parts, int n_parts, int64_t n, float* out, void* stream) {
  if (n_parts < 1 || n < 0) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (parts == nullptr || out == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_lse_combine(parts, n_parts, n, out, static_cast<cudaStream_t>(stream));
}

int hypret_sum_parts(const float* parts, int n_parts, int64_t n, float* out, void* stream) {
  if (n_parts < 1 || n < 0) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (parts == nullptr || out == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_sum_parts(parts, n_parts, n, out, static_cast<cudaStream_t>(stream));
}

int hypret_rowpair_dist(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs, int d,
                        float c, float* out, void* stream) {
  if (n_pairs < 0 || d < 1 || !(c > 0.f)) return HYPRET_EINVAL;
  if (n_pairs == 0) return HYPRET_OK;
  if (x == nullptr || y == nullptr || ia == nullptr || ib == nullptr || out == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_rowpair_dist(x, y, ia, ib, n_pairs, d, c, out, nullptr, nullptr, nullptr,
                                    static_cast<cudaStream_t>(stream));
}

int hypret_rowpair_dist_bwd(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs,
                            int d, float c, const float* grad_out, float* grad_x, float* grad_y, void* stream) {
  if (n_pairs < 0 || d < 1 || !(c > 0.f)) return HYPRET_EINVAL;
  if (n_pairs == 0) return HYPRET_OK;
  if (x == nullptr || y == nullptr || ia == nullptr || ib == nullptr || grad_out == nullptr ||
      (grad_x == nullptr && grad_y == nullptr))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_rowpair_dist(x, y, ia, ib, n_pairs, d, c, nullptr, grad_out, grad_x, grad_y,
                                    static_cast<cudaStream_t>(stream));
}

int hypret_hmi_pairs(const float* emb, const int64_t* pairs, int64_t n_pairs, int d, float c, int mode, float margin,
                     float proj_eps, float* values, double* loss_sum, const float* grad_scale, float* grad_emb,
                     void* stream) {
  if (n_pairs < 0 || d < 1 || !(c > 0.f) || (mode != 0 && mode != 1) || !(proj_eps > 0.f && proj_eps < 1.f))
    return HYPRET_EINVAL;
  if (n_pairs == 0) return HYPRET_OK;
  if (emb == nullptr || pairs == nullptr || (values == nullptr && loss_sum == nullptr && grad_emb == nullptr))
    return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_hmi_pairs(emb, pairs, n_pairs, d, c, mode, margin, proj_eps, values, loss_sum, grad_scale,
                                 grad_emb, static_cast<cudaStream_t>(stream));
}

int hypret_dist0_reg(const float* x, int64_t n, int d, float c, float lo, float hi, double* loss_sum,
                     const float* grad_scale, float* grad_x, void* stream) {
  if (n < 0 || d < 1 || !(c > 0.f) || !(hi > 0.f)) return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x == nullptr || (loss_sum == nullptr && grad_x == nullptr)) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  return hypret_launch_dist0_reg(x, n, d, c, lo, hi, loss_sum, grad_scale, grad_x, static_cast<cudaStream_t>(stream));
}

int hypret_radam_ball_step(float* x, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int d, float c,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  if (n < 0 || d < 1 || !(c > 0.f) || step < 1 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f))
    return HYPRET_EINVAL;
  if (n == 0) return HYPRET_OK;
  if (x == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr) return HYPRET_EINVAL;
  int rc = check_device();
  if (rc != HYPRET_OK) return rc;
  const float bc1 = (float)(1.0 - pow((double)beta1, step)), bc2 = (float)(1.0 - pow((double)beta2, step));
  return hypret_launch_radam_ball(x, grad, exp_avg, exp_avg_sq, n, d, c, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                  static_cast<cudaStream_t>(stream));
}

}  // extern "C"
