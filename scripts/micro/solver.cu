// what bounds one tile-local level when ONE warp walks the levels?  variants of the inner block, cycles per level
#include <cstdio>
#include <cuda_runtime.h>

template <class T, int NE, int NCHAIN, bool SYNC, bool GATHER>
__global__ void k(T *out, long long *cyc, int n, int slot)
{
  __shared__ T xs[512], vals[8 * 32 * 8];
  __shared__ int cols[8 * 32 * 8];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 512; i += blockDim.x) xs[i] = T(1.0) / T(i + 1);
  for (int i = tid; i < 8 * 32 * 8; i += blockDim.x) { vals[i] = T(1e-3) * T(i & 15); cols[i] = (i * 7) & 511; }
  __syncthreads();
  T acc = T(0.5) + T(tid);
  long long t0 = clock64();
  for (int r = 0; r < n; r++) {
    for (int s = 0; s < 22; s++) {
      int c[NE]; T v[NE], x[NE];
#pragma unroll
      for (int e = 0; e < NE; e++) { c[e] = cols[((s & 7) * 8 + e) * 32 + lane]; v[e] = vals[((s & 7) * 8 + e) * 32 + lane]; }
#pragma unroll
      for (int e = 0; e < NE; e++) x[e] = GATHER ? xs[c[e]] : xs[(e * 32 + lane + s) & 511];
      T a[NCHAIN];
#pragma unroll
      for (int q = 0; q < NCHAIN; q++) a[q] = q ? T(0) : acc;
#pragma unroll
      for (int e = 0; e < NE; e++) a[e % NCHAIN] = fma(-v[e], x[e], a[e % NCHAIN]);
      T t = a[0];
#pragma unroll
      for (int q = 1; q < NCHAIN; q++) t += a[q];
      t = t * T(0.999) + T(1e-3);
      xs[(s * 32 + lane) & 511] = t;
      acc = t;
      if (SYNC) __syncwarp();
    }
  }
  long long t1 = clock64();
  out[tid] = acc;
  if (tid == 0) cyc[slot] = (t1 - t0) / (22LL * n);
}

int main()
{
  double *d; float *f; long long *c, h[16] = {0};
  cudaMalloc(&d, 8 * 1024); cudaMalloc(&f, 4 * 1024); cudaMalloc(&c, 128); cudaMemset(c, 0, 128);
  k<double, 8, 2, true, true><<<1, 32>>>(d, c, 300, 0);
  k<double, 8, 4, true, true><<<1, 32>>>(d, c, 300, 1);
  k<double, 8, 8, true, true><<<1, 32>>>(d, c, 300, 2);
  k<double, 8, 2, false, true><<<1, 32>>>(d, c, 300, 3);
  k<double, 8, 2, true, false><<<1, 32>>>(d, c, 300, 4);
  k<float, 8, 2, true, true><<<1, 32>>>(f, c, 300, 5);
  k<double, 4, 2, true, true><<<1, 32>>>(d, c, 300, 6);
  k<double, 1, 1, true, true><<<1, 32>>>(d, c, 300, 7);
  k<float, 1, 1, true, true><<<1, 32>>>(f, c, 300, 8);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 128, cudaMemcpyDeviceToHost);
  printf("f64 8 entries: 2 chains %lld | 4 chains %lld | 8 chains %lld | no syncwarp %lld | no indirect gather %lld cycles/level\n", h[0], h[1], h[2], h[3], h[4]);
  printf("f32 8 entries 2 chains %lld | f64 4 entries %lld | f64 1 entry %lld | f32 1 entry %lld cycles/level\n", h[5], h[6], h[7], h[8]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
