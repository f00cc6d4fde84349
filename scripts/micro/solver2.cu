// solver-warp inner loop with prepared operands: per row a 16-byte word of 8 u16 gather indices (index 512 = a zero slot), a row base
// into the slot-major value slab, and {acc, dinv, aux}; software pipeline: the operands of item i+1 are loaded while item i is computed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Ops { uint4 idx; double v[8]; double a, dv, ax; };

template <bool PIPE, int NCH>
__global__ void __launch_bounds__(256) k(double *out, long long *cyc, int n, int slot)
{
  __shared__ double xs[520], vals[8 * 32 * 8], acc[512], dvs[512], aux[512];
  __shared__ uint4 idx[512];
  __shared__ int rbase[512];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < 520; i += blockDim.x) xs[i] = i < 512 ? 1.0 / (i + 1) : 0.0;
  for (int i = tid; i < 8 * 32 * 8; i += blockDim.x) vals[i] = 1e-3 * (i & 15);
  for (int i = tid; i < 512; i += blockDim.x) {
    acc[i] = 0.5 + i; dvs[i] = 0.999; aux[i] = 1e-3; rbase[i] = (i & 31) + ((i >> 5) & 7) * 256 * 0;
    unsigned short h[8];
    for (int e = 0; e < 8; e++) h[e] = (unsigned short)(e == 7 ? 512 : ((i * 7 + e * 13) & 511));
    idx[i] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
  }
  __syncthreads();
  long long t0 = clock64();
  double keep = 0.0;
  for (int r = 0; r < n; r++) {
    if (w == 0) {
      auto load = [&](int item, Ops &o) {
        const int row = (item * 17 + lane) & 511;
        o.idx = idx[row];
        const int rb = rbase[row];
#pragma unroll
        for (int e = 0; e < 8; e++) o.v[e] = vals[rb + e * 32];
        o.a = acc[row]; o.dv = dvs[row]; o.ax = aux[row];
      };
      auto compute = [&](int item, const Ops &o) {
        const unsigned iw[4] = {o.idx.x, o.idx.y, o.idx.z, o.idx.w};
        double x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = xs[(iw[e >> 1] >> ((e & 1) * 16)) & 0xffff];
        double a[NCH];
#pragma unroll
        for (int q = 0; q < NCH; q++) a[q] = q ? 0.0 : o.a;
#pragma unroll
        for (int e = 0; e < 8; e++) a[e % NCH] = fma(-o.v[e], x[e], a[e % NCH]);
        double t = a[0];
#pragma unroll
        for (int q = 1; q < NCH; q++) t += a[q];
        const double d = o.dv * t;
        const int row = (item * 17 + lane) & 511;
        xs[row] = o.ax + d;
        acc[row] = fma(-o.ax, d, t);
        keep += d;
        __syncwarp();
      };
      if (PIPE) {
        Ops o[2];
        load(0, o[0]);
#pragma unroll 1
        for (int it = 0; it < 30; it += 2) {
          load(it + 1, o[1]);
          compute(it, o[0]);
          load(it + 2, o[0]);
          compute(it + 1, o[1]);
        }
      } else {
#pragma unroll 1
        for (int it = 0; it < 30; it++) { Ops o; load(it, o); compute(it, o); }
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[tid] = keep;
  if (tid == 0) cyc[slot] = (t1 - t0) / (30LL * n);
}

int main()
{
  double *d; long long *c, h[16] = {0};
  cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 128); cudaMemset(c, 0, 128);
  k<false, 2><<<1, 256>>>(d, c, 300, 0);
  k<true, 2><<<1, 256>>>(d, c, 300, 1);
  k<true, 4><<<1, 256>>>(d, c, 300, 2);
  k<true, 4><<<444, 256>>>(d, c, 300, 3);
  cudaDeviceSynchronize();
  cudaMemcpy(h, c, 128, cudaMemcpyDeviceToHost);
  printf("prepared operands, cycles per item: no pipeline %lld | pipelined, 2 chains %lld | pipelined, 4 chains %lld | the same with 3 CTAs on every SM %lld\n", h[0], h[1], h[2], h[3]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
