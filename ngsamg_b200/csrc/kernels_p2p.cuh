// kernels_p2p.cuh -- DIS2CO / CO2CU halo exchange (DCCMap, src/base/linalg/dcc_map.cpp:225-300) over NVLink PEER MEMORY.
//
// The NCCL path costs pack kernel -> ncclGroup{Send/Recv} -> unpack kernel, ~80 us per exchange at 2 GPUs for < 1 MB of payload
// (profiles/r02_bench_2gpu_n311.json: 0.82 ms of a 9.9 ms V-cycle in ten exchanges).  Here every rank exports one receive arena per
// distributed level (cudaIpcGetMemHandle; the handles travel once at setup through the host callbacks), and an exchange is two kernels:
//   push : gathers the shared dofs of the vector and stores them DIRECTLY into the neighbour's receive buffer (remote stores over
//          NVLink), then publishes the exchange number in the neighbour's flag word (last-block-done + st.release.sys);
//   pull : waits for the flag words of the neighbours it receives from (ld.acquire.sys), adds (DIS2CO, master side; neighbours in
//          ascending order like the reference) or overwrites (CO2CU, ghost side) from the local receive buffer, then acknowledges.
// Receive buffers are double-buffered by the parity of the exchange number; a sender waits for the acknowledgement of exchange e-2
// before it overwrites a buffer (never blocks in practice).  Plain kernels: the exchanges are captured into the V-cycle CUDA graph.
#pragma once
#include "kernels.cuh"

namespace ngb {

constexpr int P2P_CPB = 8;        // CTAs per neighbour in the push kernel
constexpr int P2P_THREADS = 256;

struct P2PPeer {
  double *remote_recv;            // neighbour's receive arena [2][remote_cap]
  int *remote_flag, *remote_ack;  // this rank's slot in the neighbour's flag / acknowledgement words
  i64 remote_cap;                 // doubles per parity in the neighbour's arena
  i64 remote_m_off, remote_g_off; // first dof of this rank's segment in the neighbour's buffer: DIS2CO lands in its m-layout, CO2CU in its g-layout
};

struct P2PView {
  int npeers;
  const P2PPeer *peer;
  double *recv;                   // local arena [2][cap]
  i64 cap;
  int *flag, *ack, *cnt, *seq;    // local control words: flag[np], ack[np], cnt[np + 1] (block counters), seq (exchanges completed)
  const i64 *m_off, *g_off;       // [np + 1], in dofs
  int *err;
};

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// dir 0: DIS2CO, ghost side -- send the g-list values (and zero them: BufferG); dir 1: CO2CU, master side -- send the m-list values
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_push(P2PView V, int dir, int b, const i32 *__restrict__ idx, double *v)
{
  const int k = blockIdx.x / P2P_CPB, part = blockIdx.x % P2P_CPB;
  const i64 *soff = dir == 0 ? V.g_off : V.m_off;
  const i64 cnt = (soff[k + 1] - soff[k]) * b;
  if (cnt == 0) return;
  const int e = *(volatile int *)V.seq + 1;
  const P2PPeer P = V.peer[k];
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire_sys(V.ack + k) < e - 2)
      if (spin_fail(spins, V.err)) break;
  }
  __syncthreads();
  double *dst = P.remote_recv + (i64)(e & 1) * P.remote_cap + (dir == 0 ? P.remote_m_off : P.remote_g_off) * b;
  const i32 *ix = idx + soff[k];
  for (i64 t = (i64)part * P2P_THREADS + threadIdx.x; t < cnt; t += (i64)P2P_CPB * P2P_THREADS) {
    const i64 kk = t / b;
    const int q = (int)(t - kk * b);
    const i64 p = (i64)ix[kk] * b + q;
    dst[t] = v[p];
    if (dir == 0) v[p] = 0.0;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(V.cnt + k, 1);
    if (old == P2P_CPB - 1) {
      V.cnt[k] = 0;
      __threadfence_system();
      st_release_sys(P.remote_flag, e);
    }
  }
}

// dir 0: DIS2CO, master side: v[dof] += received values (k_halo_add); dir 1: CO2CU, ghost side: v[dof] = received value (k_halo_set)
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_pull(P2PView V, int dir, int b, i64 nitems, const i32 *__restrict__ dof,
                                                        const i64 *__restrict__ src_ptr, const i64 *__restrict__ src_pos, double *v)
{
  const int e = *(volatile int *)V.seq + 1;
  const i64 *roff = dir == 0 ? V.m_off : V.g_off;
  if (threadIdx.x < V.npeers && roff[threadIdx.x + 1] > roff[threadIdx.x]) {
    unsigned spins = 0;
    while (ld_acquire_sys(V.flag + threadIdx.x) < e)
      if (spin_fail(spins, V.err)) break;
  }
  for (int k = P2P_THREADS + threadIdx.x; k < V.npeers; k += P2P_THREADS)      // more neighbours than threads (never in practice)
    if (roff[k + 1] > roff[k]) {
      unsigned spins = 0;
      while (ld_acquire_sys(V.flag + k) < e)
        if (spin_fail(spins, V.err)) break;
    }
  __syncthreads();
  const double *buf = V.recv + (i64)(e & 1) * V.cap;
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nitems * b) {
    const i64 k = t / b;
    const int q = (int)(t - k * b);
    if (dir == 0) {
      double s = v[(i64)dof[k] * b + q];
      for (i64 j = src_ptr[k]; j < src_ptr[k + 1]; j++) s += __ldcg(buf + src_pos[j] * b + q);
      v[(i64)dof[k] * b + q] = s;
    } else {
      v[(i64)dof[k] * b + q] = __ldcg(buf + t);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(V.cnt + V.npeers, 1);
    if (old == (int)gridDim.x - 1) {
      V.cnt[V.npeers] = 0;
      for (int k = 0; k < V.npeers; k++)
        if (roff[k + 1] > roff[k]) st_release_sys(V.peer[k].remote_ack, e);
      *(volatile int *)V.seq = e;
      __threadfence();
    }
  }
}

}  // namespace ngb
