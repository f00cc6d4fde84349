// micro-benchmarks behind the design of the CTA-per-tile sweep (kernels_ctile.cuh): what does one tile-local level cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o levelloop levelloop.cu && ./levelloop
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, long long *cyc, int n)
{
  double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / (4LL * n);
}
__global__ void k_lds(int *out, long long *cyc, int n)
{
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { p = s[p]; p = s[p]; p = s[p]; p = s[p]; }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / (4LL * n);
}
__global__ void k_bar(int *out, long long *cyc, int n)
{
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  long long t1 = clock64();
  out[threadIdx.x] = (int)t1;
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / (4LL * n);
}
// the level loop of the tile kernel in miniature: 22 "levels", at each one warp is active: 8 (col, val) pairs from shared memory, 8 gathered
// x values from shared memory, FMA chains, one store; a CTA barrier after every level.  mode: 0 = as in the kernel, 1 = no barrier
// (__syncwarp only), 2 = no FMA work (loads + barrier), 3 = barrier only with one warp doing a single LDS+STS
__global__ void k_levels(double *out, long long *cyc, int n, int mode)
{
  __shared__ double xs[512], vals[8 * 32 * 8];
  __shared__ int cols[8 * 32 * 8];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < 512; i += blockDim.x) xs[i] = 1.0 / (i + 1);
  for (int i = tid; i < 8 * 32 * 8; i += blockDim.x) { vals[i] = 1e-3 * (i & 15); cols[i] = (i * 7) & 511; }
  __syncthreads();
  double acc = 0.5 + tid;
  long long t0 = clock64();
  for (int r = 0; r < n; r++) {
    for (int s = 0; s < 22; s++) {
      if ((s & 7) == w) {
        if (mode == 3) { xs[(s * 32 + lane) & 511] = xs[(s * 29 + lane) & 511] + 1.0; }
        else {
          int c[8]; double v[8], x[8];
#pragma unroll
          for (int e = 0; e < 8; e++) { c[e] = cols[(w * 8 + e) * 32 + lane]; v[e] = vals[(w * 8 + e) * 32 + lane]; }
#pragma unroll
          for (int e = 0; e < 8; e++) x[e] = xs[c[e]];
          double a = acc, a2 = 0.0;
          if (mode != 2) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) { a = fma(-v[e], x[e], a); a2 = fma(-v[e + 1], x[e + 1], a2); }
            a += a2;
            a = a * 0.999 + 1e-3;
          } else a = x[0] + x[7] + v[3];
          xs[(s * 32 + lane) & 511] = a;
          acc = a;
        }
      }
      if (mode == 1) __syncwarp(); else __syncthreads();
    }
  }
  long long t1 = clock64();
  out[tid] = acc;
  if (tid == 0) cyc[3 + mode] = (t1 - t0) / (22LL * n);
}

// one SOLVER warp walks all levels of the tile (no CTA barrier inside, __syncwarp between levels); the other warps wait at the end.
// pre = 1: the (col, val) pairs of the next level are loaded before the current level's gather (software pipeline in registers)
__global__ void k_solver(double *out, long long *cyc, int n, int pre)
{
  __shared__ double xs[512], vals[8 * 32 * 8];
  __shared__ int cols[8 * 32 * 8];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < 512; i += blockDim.x) xs[i] = 1.0 / (i + 1);
  for (int i = tid; i < 8 * 32 * 8; i += blockDim.x) { vals[i] = 1e-3 * (i & 15); cols[i] = (i * 7) & 511; }
  __syncthreads();
  double acc = 0.5 + tid;
  long long t0 = clock64();
  for (int r = 0; r < n; r++) {
    if (w == 0) {
      int c[8]; double v[8];
      if (pre) {
#pragma unroll
        for (int e = 0; e < 8; e++) { c[e] = cols[e * 32 + lane]; v[e] = vals[e * 32 + lane]; }
      }
      for (int s = 0; s < 22; s++) {
        double x[8];
        if (!pre) {
#pragma unroll
          for (int e = 0; e < 8; e++) { c[e] = cols[((s & 7) * 8 + e) * 32 + lane]; v[e] = vals[((s & 7) * 8 + e) * 32 + lane]; }
        }
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = xs[c[e]];
        double vv[8];
#pragma unroll
        for (int e = 0; e < 8; e++) vv[e] = v[e];
        if (pre) {
#pragma unroll
          for (int e = 0; e < 8; e++) { c[e] = cols[(((s + 1) & 7) * 8 + e) * 32 + lane]; v[e] = vals[(((s + 1) & 7) * 8 + e) * 32 + lane]; }
        }
        double a = acc, a2 = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e += 2) { a = fma(-vv[e], x[e], a); a2 = fma(-vv[e + 1], x[e + 1], a2); }
        a += a2;
        a = a * 0.999 + 1e-3;
        xs[(s * 32 + lane) & 511] = a;
        acc = a;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[tid] = acc;
  if (tid == 0) cyc[pre] = (t1 - t0) / (22LL * n);
}

int main()
{
  double *d; int *di; long long *c, h[8] = {0};
  cudaMalloc(&d, 8 * 1024); cudaMalloc(&di, 4 * 1024); cudaMalloc(&c, 64); cudaMemset(c, 0, 64); cudaMemset(d, 0, 8 * 1024);
  k_dfma<<<1, 32>>>(d, c, 1000);
  k_lds<<<1, 32>>>(di, c, 1000);
  k_bar<<<1, 256>>>(di, c, 1000);
  for (int m = 0; m < 4; m++) k_levels<<<1, 256>>>(d, c, 200, m);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("dependent DFMA: %lld cycles\ndependent LDS : %lld cycles\n__syncthreads (8 warps, all there): %lld cycles\n", h[0], h[1], h[2]);
  printf("level loop, per level: kernel-like %lld | __syncwarp instead of barrier %lld | loads + barrier, no FMA %lld | one LDS+STS + barrier %lld cycles\n", h[3], h[4], h[5], h[6]);
  // the same with every SM busy (148 CTAs x 3): contention for nothing but the SM's own resources
  cudaMemset(c, 0, 64);
  k_levels<<<444, 256>>>(d, c, 200, 0);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("level loop with 3 CTAs on every SM: %lld cycles per level\n", h[3]);
  cudaMemset(c, 0, 64);
  k_solver<<<1, 256>>>(d, c, 200, 0);
  k_solver<<<1, 256>>>(d, c, 200, 1);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("solver warp (one warp walks all levels, __syncwarp): %lld cycles per level; with the next level's matrix entries prefetched: %lld\n", h[0], h[1]);
  cudaMemset(c, 0, 64);
  k_solver<<<444, 256>>>(d, c, 200, 1);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("solver warp, prefetch, 3 CTAs on every SM: %lld cycles per level\n", h[1]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
