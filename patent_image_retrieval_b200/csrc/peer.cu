// Peer-memory exchange over NVLink / NVSwitch (multi-GPU serving, one process per GPU on one box).
//
// The reference is single-process, single-GPU: nothing here replaces a reference call site.  It replaces the
// NCCL all_gather of the per-rank query batches in the sharded-serving protocol (dist.py): every rank owns an
// exchange buffer that all ranks of the box map (CUDA IPC), the projection kernel stores each operand row into
// all of them (csrc/project.cu, hypret_project_rows_peers), the exact fp32 points follow through the copy
// engines while the scoring kernel runs, and arrival is announced with per-source step counters:
//   peer_signal_kernel   one thread per destination: st.release.sys of the step number into flag[src] there
//   peer_wait_kernel     one thread per source: ld.acquire.sys until flag[src] >= step (bounded spin)
// Stream order does the rest: a signal is launched behind the kernel / copies it announces, a wait in front of
// the kernel that consumes them.  Counters only grow, so a peer that is already one step ahead still satisfies
// the wait; two buffer slots suffice (a rank cannot start step i+2 before every peer has finished step i, because
// its step i+1 waited for their step-i+1 signal, which they issue behind their step i).
#include "common.cuh"

namespace {

struct FlagDsts {
  uint32_t* p[HYPRET_MAX_PEERS];
};

__global__ void peer_signal_kernel(const FlagDsts dsts, int n, uint32_t value) {
  if ((int)threadIdx.x < n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dsts.p[threadIdx.x]), "r"(value) : "memory");
  }
}

__global__ void peer_wait_kernel(const uint32_t* __restrict__ flags, int n, uint32_t value, uint32_t* err) {
  if ((int)threadIdx.x < n) {
    const uint32_t* f = flags + threadIdx.x;
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v >= value) break;
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {          // 20 s: a peer died; report instead of hanging the GPU
        if (err != nullptr) atomicExch(err, 1u + threadIdx.x);
        break;
      }
    }
  }
}

}  // namespace

int hypret_launch_peer_signal(void* const* flags_host, int n, uint32_t value, cudaStream_t stream) {
  if (n < 1 || n > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  FlagDsts d;
  for (int i = 0; i < n; ++i) d.p[i] = reinterpret_cast<uint32_t*>(flags_host[i]);
  peer_signal_kernel<<<1, 32, 0, stream>>>(d, n, value);
  return (int)cudaGetLastError();
}

int hypret_launch_peer_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, cudaStream_t stream) {
  if (n < 1 || n > HYPRET_MAX_PEERS) return HYPRET_EINVAL;
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, n, value, err);
  return (int)cudaGetLastError();
}
