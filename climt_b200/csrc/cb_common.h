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

#if defined(__CUDA_ARCH__)
// a*b + c with two roundings (never contracted to an FMA): used wherever the result is truncated to a table index,
// so that the device picks the same table slot as the reference's separately-rounded multiply and add.
#define CB_MULADD_2R(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#else
#define CB_MULADD_2R(a, b, c) (((a) * (b)) + (c))
#endif

// Branch-coverage hooks of the taumol evaluators (tools/taumol_coverage.py): compiled in only for the host emulation built with
// -DCB_COVERAGE; in every other build -- the CUDA kernels included -- the macro is empty.
#if defined(CB_COVERAGE) && !defined(__CUDA_ARCH__)
extern "C" void cb_cov_hit(int engine, int band, int lower, int id);
#define CB_COV(engine, band, lower, id) cb_cov_hit((engine), (band), (lower) ? 1 : 0, (id))
#else
#define CB_COV(engine, band, lower, id) ((void)0)
#endif
namespace cb {
// what a coverage run records per (band, lower | upper atmosphere)
enum CovId {
  COV_REGION = 0,       // a layer was evaluated in this region
  COV_KEY_NONZERO,      // key-species term > 0
  COV_S0_LOW, COV_S0_MID, COV_S0_HIGH,   // binary-species stencil at the lower pressure level: specparm < 0.125 | between | > 0.875
  COV_S1_LOW, COV_S1_MID, COV_S1_HIGH,   // ... at the upper pressure level
  COV_SELF_NONZERO, COV_FOR_NONZERO,     // water-vapour self / foreign continuum contributes (> 0)
  COV_MINOR0_NONZERO, COV_MINOR1_NONZERO, COV_MINOR2_NONZERO,   // minor gas k contributes
  COV_MINOR0_ADJ, COV_MINOR1_ADJ, COV_MINOR2_ADJ,               // ... through the "too abundant to be minor" adjustment
  COV_XSEC0_NONZERO, COV_XSEC1_NONZERO,  // halocarbon cross-section (LW) / extra absorber (SW) contributes
  COV_PLANCK_INTERP,    // Planck fraction (LW) / solar source (SW) interpolated in the key-species ratio
  COV_N
};

// U consecutive doubles of a table row (16-byte aligned by construction: even table offsets, even g-point counts,
// unit starts that are multiples of 4) fetched with 128-bit read-only loads.
template <int U>
struct Row {
  double v[U];
  CB_HD double operator[](int i) const { return v[i]; }
};
template <int U>
CB_HD Row<U> ldrow(const double* __restrict__ p) {
  Row<U> r;
#if defined(__CUDA_ARCH__)
  if (U == 4) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  } else if (U == 2) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    r.v[0] = a.x; r.v[1] = a.y;
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) r.v[u] = __ldg(p + u);
  }
#else
  for (int u = 0; u < U; ++u) r.v[u] = p[u];
#endif
  return r;
}

// The same row through ordinary (coherent) loads: for tables staged in shared memory, where ld.global.nc is not allowed.
template <int U>
CB_HD Row<U> ldrow_plain(const double* __restrict__ p) {
  Row<U> r;
#if defined(__CUDA_ARCH__)
  if (U == 4) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  } else if (U == 2) {
    const double2 a = reinterpret_cast<const double2*>(p)[0];
    r.v[0] = a.x; r.v[1] = a.y;
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) r.v[u] = p[u];
  }
#else
  for (int u = 0; u < U; ++u) r.v[u] = p[u];
#endif
  return r;
}

// Division / reciprocal for well-conditioned operands (finite, normal, non-zero divisor): MUFU.RCP64H seed, two Newton
// steps, one residual correction -- 1 MUFU + 7 (4) fp64 FMAs and none of the ~10 integer/branch instructions per site that
// nvcc's IEEE division spends on denormal / inf / NaN operands (r01 ncu: two thirds of the transfer kernels' instructions
// were not fp64 arithmetic).  The quotient is within 1 ulp of the correctly rounded one (normally equal to it); it is NOT
// used where a quotient is truncated to a table index.  The host build (tests/emul) uses the plain operators.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double frcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fdiv(double a, double b) {
  const double r = frcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
#else
inline double frcp(double b) { return 1.0 / b; }
inline double fdiv(double a, double b) { return a / b; }
#endif

// Cache-residency hints of the slab form of the transfer kernels (lw_engine.cu / sw_engine.cu): the rows a warp carries from one
// vertical sweep to the other live in a slab that is meant to stay in the L2, so
//   * what streams through once (the taumol rows, the partial-flux rows) is loaded / stored "evict first" (ld/st.global.cs) and
//     fetched a few layers ahead into the L2 only (prefetch.global.L2: no registers held while the line travels);
//   * slab rows are stored with an L2 evict_last policy and read back for the last time with evict_first.
// Plain loads and stores on the host (tests/emul).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// The 128-byte line at p (128-byte aligned) holds scratch that has just been read for the last time: if it is still dirty in the
// L2 it need not be written back to HBM.  (Its content is undefined afterwards -- the next launch writes it before reading it.)
__device__ __forceinline__ void discard_l2(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }
__device__ __forceinline__ unsigned long long policy_keep() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long policy_drop() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_policy(double* p, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_policy(const double* p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
#else
inline double ld_stream(const double* p) { return *p; }
inline void st_stream(double* p, double v) { *p = v; }
inline void prefetch_l2(const void*) {}
inline void discard_l2(const void*) {}
inline unsigned long long policy_keep() { return 0; }
inline unsigned long long policy_drop() { return 0; }
inline void st_policy(double* p, double v, unsigned long long) { *p = v; }
inline double ld_policy(const double* p, unsigned long long) { return *p; }
#endif

// Fortran real->integer assignment / int(): truncation toward zero.
// Slot of the exp / tau-transition tables, int(tblint * x / (bpade + x) + 0.5) (rrtmg_lw_rtrn.f90:357-360, rrtmg_sw_reftra.f90:
// 208-211), WITHOUT the IEEE division (a ~35-instruction subroutine on the device, and the longest dependent chain of a cell):
// fdiv's quotient is within 1 ulp of the correctly rounded one, so the truncated index can differ only when the argument of the
// truncation lies within ~1e-12 of an integer -- those (one in ~1e9) take the IEEE division.  Same slot as the reference, always.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int tbl_slot(double x, double bpade, double tblint) {
  const double v = __dadd_rn(__dmul_rn(tblint, fdiv(x, bpade + x)), 0.5);
  if (fabs(v - rint(v)) < 1.e-9) return (int)__dadd_rn(__dmul_rn(tblint, x / (bpade + x)), 0.5);
  return (int)v;
}
#else
inline int tbl_slot(double x, double bpade, double tblint) { return (int)((tblint * (x / (bpade + x))) + 0.5); }
#endif

CB_HD int f2i(double x) { return (int)x; }
CB_HD int imin(int a, int b) { return a < b ? a : b; }
CB_HD int imax(int a, int b) { return a > b ? a : b; }

}  // namespace cb
