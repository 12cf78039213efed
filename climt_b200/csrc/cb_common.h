// climt_b200 native engine -- common definitions shared by the CUDA kernels and the
// (test-only) host emulation of the same per-thread code.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_D __device__ __forceinline__
#else
#define CB_HD inline
#define CB_D inline
#endif

#if defined(__CUDA_ARCH__)
#define CB_LDG(p) __ldg(p)
#else
#define CB_LDG(p) (*(p))
#endif

namespace cb {

// Fortran real->integer assignment / int(): truncation toward zero.
CB_HD int f2i(double x) { return (int)x; }
CB_HD int imin(int a, int b) { return a < b ? a : b; }
CB_HD int imax(int a, int b) { return a > b ? a : b; }

}  // namespace cb
