// micro-benchmarks behind the small-level sweep design: instruction latencies in ONE warp (clock64), globaltimer cost/granularity,
// and the store -> poll hop between two warps on different SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double ld_relaxed(const double *p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

__global__ void k_lat(double *out, long long *cyc, unsigned long long *gt)
{
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = lane; sm[lane + 32] = 1.0;
  __syncwarp();
  double a = out[lane];
  long long t0 = clock64();
  // (a) dependent DADD chain x64
#pragma unroll
  for (int i = 0; i < 64; i++) a = a + 1.0000001;
  long long t1 = clock64();
  // (b) butterfly: 5 x (2 SHFL + DADD), x8
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  long long t2 = clock64();
  // (c) dependent LDS chain x32
  int idx = lane;
#pragma unroll
  for (int i = 0; i < 32; i++) idx = (int)sm[idx & 31] ;
  long long t3 = clock64();
  // (d) DFMA chain x64
#pragma unroll
  for (int i = 0; i < 64; i++) a = fma(a, 1.0000001, 0.5);
  long long t4 = clock64();
  // (e) globaltimer read x16
  unsigned long long g[17];
#pragma unroll
  for (int i = 0; i < 17; i++) g[i] = gtimer();
  long long t5 = clock64();
  out[lane] = a + idx;
  if (lane == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    for (int i = 0; i < 17; i++) gt[i] = g[i];
  }
}

// hop: CTA b waits for flag[b-1] (the data is the flag: sentinel -1 NaN), then publishes flag[b].  One warp per CTA, CTAs on different SMs.
__global__ void k_hop(double *flag, int nhops, int mode, unsigned sleep_ns, unsigned long long *stamps)
{
  const int b = blockIdx.x, nb = gridDim.x, lane = threadIdx.x;
  for (int h = b; h < nhops; h += nb) {
    if (h > 0) {
      if (mode == 0) {          // lane 0 polls
        if (lane == 0) { while (__double_as_longlong(ld_relaxed(flag + h - 1)) == -1LL) if (sleep_ns) __nanosleep(sleep_ns); }
        __syncwarp();
      } else {                  // all lanes poll (same address)
        while (__double_as_longlong(ld_relaxed(flag + h - 1)) == -1LL) if (sleep_ns) __nanosleep(sleep_ns);
        __syncwarp();
      }
    }
    double a = (double)lane;
    if (mode == 2) {
#pragma unroll
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if (lane == 0) { __stcg(flag + h, a); if (stamps) stamps[h] = gtimer(); }
  }
}

int main()
{
  double *out; long long *cyc; unsigned long long *gt;
  cudaMalloc(&out, 64 * 8); cudaMemset(out, 0, 64 * 8);
  cudaMallocManaged(&cyc, 8 * 8); cudaMallocManaged(&gt, 32 * 8);
  for (int rep = 0; rep < 2; rep++) { k_lat<<<1, 32>>>(out, cyc, gt); cudaDeviceSynchronize(); }
  printf("DADD chain: %.1f cyc/op   butterfly(2 SHFL + DADD): %.1f cyc/stage   LDS chain (+cvt): %.1f cyc/op   DFMA chain: %.1f cyc/op   globaltimer read: %.1f cyc/op\n",
         cyc[0] / 64.0, cyc[1] / 40.0, cyc[2] / 32.0, cyc[3] / 64.0, cyc[4] / 17.0);
  printf("globaltimer deltas (ns):");
  for (int i = 1; i < 17; i++) printf(" %llu", gt[i] - gt[i - 1]);
  printf("\n");
  const int nhops = 4096;
  double *flag; unsigned long long *st;
  cudaMalloc(&flag, nhops * 8); cudaMallocManaged(&st, nhops * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int nb : {2, 8, 148, 296, 1184})
    for (int mode : {0, 1, 2})
      for (unsigned sl : {0u, 100u}) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          cudaMemset(flag, 0xFF, nhops * 8);
          cudaEventRecord(e0);
          k_hop<<<nb, 32>>>(flag, nhops, mode, sl, nullptr);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("hop: %4d CTAs (1 warp each) mode %d (0 lane0 polls, 1 all lanes poll, 2 all + butterfly) sleep %3u ns: %.3f us/hop\n", nb, mode, sl, best * 1e3 / nhops);
      }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
