// device.hpp -- host-visible declarations of the CUDA side (no torch, no NGSolve types).
#pragma once
#include <cuda_runtime.h>

#include "common.hpp"

namespace ngb {

#define NGB_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      throw ::ngb::Error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                         std::to_string(__LINE__) + " (" #call ")");                                     \
  } while (0)

template <class T>
inline T *dev_alloc(size_t n)
{
  T *p = nullptr;
  NGB_CUDA(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}
template <class T>
inline void dev_free(T *&p)
{
  if (p) cudaFree((void *)p);
  p = nullptr;
}

// plain block-CSR on the device, original DOF numbering (setup path: RAP, layout construction)
struct DevCsr {
  i64 nrows = 0, ncols = 0, nnz = 0;
  int bh = 1, bw = 1;
  i64 *rowptr = nullptr;
  i32 *col = nullptr;
  double *val = nullptr;
  int bs() const { return bh * bw; }
};

void dev_csr_upload(const HostBsr &h, DevCsr &d, cudaStream_t st);
void dev_csr_upload_raw(i64 nrows, i64 ncols, int bh, int bw, const i64 *rowptr, const i32 *col, const double *val,
                        DevCsr &d, cudaStream_t st);
void dev_csr_download(const DevCsr &d, HostBsr &h, cudaStream_t st, bool with_values = true);
void dev_csr_free(DevCsr &d);

// C = A * B.  Pattern = structural sorted union per row (bit-exact with MergeArrays), values accumulated in
// A-row order then B-row order without FMA contraction (MatMultABImpl, utils_sparseMM.cpp:107-238).
void dev_spgemm(const DevCsr &A, const DevCsr &B, DevCsr &C, cudaStream_t st, i64 *launches);

// T = A^T with every block transposed; rows of T ascending (TransposeSPMImpl, utils_sparseMM.cpp:54-93) -- bit-identical integers
void dev_transpose(const DevCsr &A, DevCsr &T, cudaStream_t st, i64 *launches);

}  // namespace ngb
