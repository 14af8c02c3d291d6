// One-shot all-reduce of a few doubles over peer memory (NVLink 5 / NVSwitch P2P stores), callable from inside
// a kernel: the data-parallel step's only exchange is the 16-double [parameter gradient | loss sum] buffer
// (SURVEY.md section 8(e)), so instead of a host-launched collective after the reduction kernel, the
// reduction kernel itself pushes its result into every peer's buffer and sums the world's rows.
//
// Every rank owns one CommBuf in its own HBM, mapped into all peers by CUDA IPC (cn_comm.cu).  Epoch e:
//   1. store my n values into data[e & 1][my_rank][*] of EVERY rank's buffer (mine included), fence (system scope)
//   2. release-store e into flag[my_rank] of every rank's buffer
//   3. spin (acquire loads, bounded by a timeout) until my own buffer's flag[r] >= e for all r
//   4. sum the world's rows of my buffer in rank order -> identical bits on every rank
// A rank can run at most one epoch ahead of the slowest one (it cannot pass step 3 of epoch e + 1 before
// every peer has sent flag e + 1, which a peer does only after finishing step 4 of epoch e), so two data
// buffers alternate safely and flags only ever grow.  The epoch counter lives in device memory: a CUDA-graph
// replay of the kernel needs no new arguments.  One communicator serves one stream at a time.
#pragma once
#include <cstdint>

namespace cn {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_MAX_ELEMS = 32;
constexpr unsigned long long COMM_TIMEOUT_NS = 4000000000ull;   // a missing peer must not hang the GPU

struct CommBuf {                                   // peer-visible (one per rank, IPC-mapped everywhere)
  unsigned long long flag[COMM_MAX_WORLD];         // flag[r]: last epoch whose row rank r has completely written here
  double data[2][COMM_MAX_WORLD][COMM_MAX_ELEMS];
};

struct CommDev {                                   // rank-local device state handed to the kernels
  int32_t rank, world;
  unsigned long long epoch;                        // last completed epoch
  int32_t error;                                   // 1 after a timeout (results are then undefined)
  int32_t pad;
  CommBuf* peer[COMM_MAX_WORLD];                   // peer[r] = rank r's buffer in this process's address space
};

#if defined(__CUDACC__)
__device__ __forceinline__ void comm_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long comm_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void comm_st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double comm_ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long comm_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Called by ALL threads of ONE block (blockDim.x >= 32).  v: n (<= COMM_MAX_ELEMS) values in shared memory, valid
// before the call; on return (after the trailing barrier) v[i] = sum over ranks of their v[i], in rank order.
__device__ __forceinline__ void comm_allreduce_block(CommDev* C, double* v, int n) {
  const int t = threadIdx.x, world = C->world, rank = C->rank;
  const unsigned long long e = C->epoch + 1;
  const int p = (int)(e & 1ull);
  __syncthreads();                                               // v complete; everyone has read the epoch
  for (int idx = t; idx < world * n; idx += blockDim.x) {
    const int r = idx / n, i = idx - r * n;
    comm_st_relaxed_sys(&C->peer[r]->data[p][rank][i], v[i]);
  }
  __threadfence_system();
  __syncthreads();
  if (t < world) {
    comm_st_release_sys(&C->peer[t]->flag[rank], e);
    const unsigned long long* f = &C->peer[rank]->flag[t];
    const unsigned long long t0 = comm_globaltimer();
    while (comm_ld_acquire_sys(f) < e) {
      if (comm_globaltimer() - t0 > COMM_TIMEOUT_NS) { C->error = 1; break; }
    }
  }
  __syncthreads();
  if (t < n) {
    const CommBuf* mine = C->peer[rank];
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += comm_ld_relaxed_sys(&mine->data[p][r][t]);
    v[t] = s;
  }
  if (t == 0) C->epoch = e;
  __syncthreads();
}
#endif

}  // namespace cn
