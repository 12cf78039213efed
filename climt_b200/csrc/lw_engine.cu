// climt_b200 -- RRTMG longwave engine: CUDA kernels (sm_100a), launcher and C ABI (include/climt_b200.h).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <future>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/climt_b200.h"
#include "engine_common.h"
#include "cb_async.cuh"
#include "lw_tables.h"
#include "mcica_host.h"
#include "mcica_compat.h"

using namespace cb::lw;

namespace {

using cb::kBlock;
#ifndef CB_LW_RT_MIN_BLOCKS
#define CB_LW_RT_MIN_BLOCKS 6  // transfer kernel: 80 registers, no spills (ptxas), 24 warps per SM
#endif

// inatm + setcoef: one thread per (column, layer)
__global__ void __launch_bounds__(kBlock) k_prep_layer(const __grid_constant__ Tables T, const __grid_constant__ In in, const Flags fl,
                                                       const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (c < n) prep_column<true, false>(T, in, fl, W, c0, c, l, l + 1);
}
// what couples the layers of a column (pwvcm, laytrop, cloud optics): one thread per column
__global__ void __launch_bounds__(kBlock) k_prep(const __grid_constant__ Tables T, const __grid_constant__ In in, const Flags fl,
                                                 const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) prep_column<false, true>(T, in, fl, W, c0, c, 0, in.nlay);
}

// rtrn cloud prologue: one thread per (column, layer), after k_prep (needs pwvcm and ncbands of the column)
__global__ void __launch_bounds__(kBlock) k_cld_scale(const __grid_constant__ In in, const Flags fl, const __grid_constant__ Work W,
                                                      int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) prep_cloud_scale(in, fl, W, c0, c, blockIdx.y);
}

struct UnitList {
  Unit u[kMaxUnits];
  int n;
};

#ifndef CB_LW_TAU_MIN_BLOCKS
#define CB_LW_TAU_MIN_BLOCKS 5  // r02 B200 sweep, 8192 x 60 | McICA 16384 x 72: 3 -> 0.824 | 2.053 ms, 4 -> 0.702 | 1.693, 5 -> 0.669 | 1.541, 6 -> 0.671 | 1.540, 8 -> 0.847 | 1.897
#endif
#ifndef CB_LW_LAYER_CHUNKS
#define CB_LW_LAYER_CHUNKS 4  // taumol: layers are independent -> blockIdx.z cuts them into chunks for more threads in flight
#endif

#ifndef CB_LW_STAGE
#define CB_LW_STAGE 1  // stage the table block of single-key-species bands in shared memory (0: every band reads HBM through L1/L2)
#endif
// Bands whose lower- AND upper-atmosphere regions have at most one key species: their whole (band, 4-g-point group) table block
// (k-distribution for both regions, continua, minor gases, Planck fractions: 8-14 KB) fits beside 4 resident blocks per SM.
__host__ __device__ constexpr bool lw_band_staged(int B) {
  return CB_LW_STAGE && (B == 1 || B == 2 || B == 6 || B == 8 || B == 10 || B == 11 || B == 14);
}

// taumol: one block = 128 adjacent columns x one unit (<= 4 g-points of one band) x one chunk of layers; every branch on
// the band is block-uniform, all global accesses are column-contiguous.  For the single-key-species bands the block first pulls
// its table block into shared memory with ONE bulk asynchronous copy (cp.async.bulk -> UBLKCP, completion on an mbarrier): the
// 4-12 table gathers per (layer, unit) then hit shared memory instead of L1/L2 (r01 ncu: L1 hit 47 %, L2 hit 45 %).
__global__ void __launch_bounds__(kBlock, CB_LW_TAU_MIN_BLOCKS)
    k_lw_taumol(const __grid_constant__ Tables T, const __grid_constant__ In in, const __grid_constant__ Work W,
                const __grid_constant__ UnitList UL, int c0, int n) {
  extern __shared__ __align__(128) double s_tab[];
  __shared__ __align__(8) unsigned long long s_bar;
  const Unit un = UL.u[blockIdx.y];
  if (lw_band_staged(un.band)) {  // block-uniform
    const BandOff& O = T.b[un.band - 1];
    const unsigned bytes = (unsigned)O.rows * TGW * sizeof(double);
    if (threadIdx.x == 0) cb::bulk::mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      cb::bulk::mbar_arrive_expect_tx(&s_bar, bytes);
      cb::bulk::copy_g2s(s_tab, T.base + O.base + (size_t)(un.g0 / TGW) * O.rows * TGW, bytes, &s_bar);
    }
    cb::bulk::mbar_wait(&s_bar, 0);
  }
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int per = (in.nlay + gridDim.z - 1) / gridDim.z;
  const int l0 = blockIdx.z * per, l1 = min(in.nlay, l0 + per);
#define CB_CASE(B)                                                                                   \
  case B:                                                                                            \
    if (un.u == 4) lw_taumol_unit<B, 4, lw_band_staged(B)>(T, in, W, c0, c, un.g0, l0, l1, s_tab);    \
    else lw_taumol_unit<B, 2, lw_band_staged(B)>(T, in, W, c0, c, un.g0, l0, l1, s_tab);              \
    break;
  switch (un.band) {
    CB_CASE(1) CB_CASE(2) CB_CASE(3) CB_CASE(4) CB_CASE(5) CB_CASE(6) CB_CASE(7) CB_CASE(8)
    CB_CASE(9) CB_CASE(10) CB_CASE(11) CB_CASE(12) CB_CASE(13) CB_CASE(14) CB_CASE(15) CB_CASE(16)
  }
#undef CB_CASE
}

constexpr int kPartK = 4;  // interfaces staged in shared memory between two flushes of LwPartSmem

// Device-side sink of the sweeps' radiance sums (interface: LwPartDirect in lw_core.cuh).  The CB_LW_GROUP warps of a block are
// the units of one group, the same 32 columns in every warp.  A put parks the thread's values in shared memory; every kPartK
// interfaces the block meets, each warp sums a share of the (interface, row) pairs over the warps IN UNIT ORDER (deterministic,
// = the serial emulation's order) and writes one coalesced 256-byte row per pair.
struct LwPartSmem {
  double* buf;   // shared: [kPartK][4][nthreads]
  double* part;  // rows of this block's group at this thread's column
  size_t pstride, lev_first;
  int ncc, nthreads, nw, tid, warp, lane, count;
  int nv;   // values per put of the current sweep: down 2 (rows 1, 3), up 2 or 4 (rows 0, 2, 4, 5); odd ones are clear-sky rows
  int dir;  // +1: up sweep, interfaces arrive bottom-up; -1: down sweep, top-down
  bool valid, cloudy;
  bool stream = false;  // slab form of the kernel: the rows are written "evict first" so that they do not displace the slabs in the L2
  __device__ void stage(size_t lev, double v0, double v1, double v2, double v3) {
    if (count == 0) lev_first = lev;
    double* b = buf + (size_t)count * 4 * nthreads + tid;
    b[0] = v0; b[nthreads] = v1;
    if (nv == 4) { b[2 * nthreads] = v2; b[3 * nthreads] = v3; }
    if (++count == kPartK) flush();
  }
  __device__ void put_dn(size_t lev, bool cloudy_col, double dn, double dnc) {
    if (count == 0) { nv = 2; dir = -1; }
    cloudy = cloudy_col;
    stage(lev, dn, dnc, 0., 0.);
  }
  __device__ void put_up(size_t lev, bool cloudy_col, double up, double upc, bool drv, double dup, double dupc) {
    if (count == 0) { nv = drv ? 4 : 2; dir = 1; }
    cloudy = cloudy_col;
    stage(lev, up, upc, dup, dupc);
  }
  __device__ void flush() {
    cb::barrier_unaligned(nthreads);
    for (int p = warp; p < count * nv; p += nw) {
      const int k = p / nv, q = p - k * nv;
      const double* b = buf + ((size_t)k * 4 + q) * nthreads + lane;
      double s = b[0];
      for (int w = 1; w < nw; ++w) s = s + b[w * 32];
      const int row = dir < 0 ? 1 + 2 * q : (q < 2 ? 2 * q : 2 + q);
      // (column validity and cloudiness are properties of the lane's column: the same in every warp of the block)
      if (valid && (!(q & 1) || cloudy)) {
        double* o = part + row * pstride + (size_t)((long)lev_first + (long)dir * k) * ncc;
        if (stream) cb::st_stream(o, s); else *o = s;
      }
    }
    cb::barrier_unaligned(nthreads);
    count = 0;
  }
  __device__ void end_sweep() {
    if (count) flush();
  }
};

// rtrn / rtrnmc / rtrnmr: one block = 32 adjacent columns x CB_LW_GROUP units (one warp each, <= CB_LW_UMAX g-points of one band);
// one code body for all bands
template <bool MC, bool MR, bool DRV = false>
__global__ void __launch_bounds__(32 * CB_LW_GROUP, (CB_LW_RT_MIN_BLOCKS * 4 + CB_LW_GROUP - 1) / CB_LW_GROUP)
    k_units(const __grid_constant__ Tables T, const __grid_constant__ In in, const __grid_constant__ Work W,
            const __grid_constant__ UnitList UL, int c0, int n) {
  __shared__ double s_part[kPartK * 4 * 32 * CB_LW_GROUP];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int group = blockIdx.y;
  const int k = group * CB_LW_GROUP + threadIdx.y;
  LwPartSmem sink;
  sink.buf = s_part;
  sink.pstride = (size_t)(in.nlay + 1) * W.ncc;
  sink.part = W.part + (size_t)group * W.npart * sink.pstride + c;
  sink.ncc = W.ncc;
  sink.nthreads = 32 * CB_LW_GROUP; sink.nw = CB_LW_GROUP;
  sink.tid = threadIdx.y * 32 + threadIdx.x; sink.warp = threadIdx.y; sink.lane = threadIdx.x;
  sink.count = 0; sink.lev_first = 0; sink.nv = 2; sink.dir = 1;
  sink.valid = c < n; sink.cloudy = false;
  if (c >= n || k >= UL.n) {  // no work, but the block's meetings need every thread: the same sequence of puts
    const bool cl = c < n && W.ncbands[c] > 0;
    for (int lev = in.nlay; lev >= 1; --lev) sink.put_dn((size_t)(lev - 1), cl, 0., 0.);
    sink.end_sweep();
    for (int lev = 0; lev <= in.nlay; ++lev) sink.put_up((size_t)lev, cl, 0., 0., DRV, 0., 0.);
    sink.end_sweep();
    return;
  }
  const Unit un = UL.u[k];
  if (CB_LW_UMAX >= 4 && un.u == 4) lw_transfer_unit<4, MC, MR, DRV>(T, in, W, c0, c, un.band - 1, un.g0, sink);
  else if (CB_LW_UMAX == 1) lw_transfer_unit<1, MC, MR, DRV>(T, in, W, c0, c, un.band - 1, un.g0, sink);
  else lw_transfer_unit<2, MC, MR, DRV>(T, in, W, c0, c, un.band - 1, un.g0, sink);
}

// The slab form of the transfer kernel: a persistent grid (blocks per SM chosen by the host) walks the (column tile, group of
// units) work items; each warp keeps the rows it carries from the down sweep to the up sweep in a slab of its own that it reuses
// item after item, laid out [layer][g-point][row][lane].  The rows are read back in the reverse order of their writing, so with the
// slabs of all resident warps inside the L2 they never reach HBM.
#ifndef CB_LW_SLAB_MIN_BLOCKS
#define CB_LW_SLAB_MIN_BLOCKS 4
#endif
template <bool MC, bool MR, bool DRV = false>
__global__ void __launch_bounds__(32 * CB_LW_GROUP, CB_LW_SLAB_MIN_BLOCKS)
    k_units_slab(const __grid_constant__ Tables T, const __grid_constant__ In in, const __grid_constant__ Work W,
                 const __grid_constant__ UnitList UL, double* __restrict__ slabs, int c0, int n) {
  __shared__ double s_part[kPartK * 4 * 32 * CB_LW_GROUP];
  const int ntiles = (n + 31) / 32, ngroups = (UL.n + CB_LW_GROUP - 1) / CB_LW_GROUP;
  const int nlay = in.nlay;
  Carry cy;
  cy.rs = 32; cy.us = 4 * 32; cy.ls = CB_LW_UMAX * 4 * 32;
  cy.p = slabs + ((size_t)blockIdx.x * CB_LW_GROUP + threadIdx.y) * ((size_t)nlay * CB_LW_UMAX * 4 * 32) + threadIdx.x;
  for (int w = blockIdx.x; w < ntiles * ngroups; w += gridDim.x) {
    const int tile = w % ntiles, group = w / ntiles;
    const int c = tile * 32 + threadIdx.x;
    const int k = group * CB_LW_GROUP + threadIdx.y;
    LwPartSmem sink;
    sink.buf = s_part;
    sink.pstride = (size_t)(nlay + 1) * W.ncc;
    sink.part = W.part + (size_t)group * W.npart * sink.pstride + c;
    sink.ncc = W.ncc;
    sink.nthreads = 32 * CB_LW_GROUP; sink.nw = CB_LW_GROUP;
    sink.tid = threadIdx.y * 32 + threadIdx.x; sink.warp = threadIdx.y; sink.lane = threadIdx.x;
    sink.count = 0; sink.lev_first = 0; sink.nv = 2; sink.dir = 1;
    sink.valid = c < n; sink.cloudy = false; sink.stream = true;
    if (c >= n || k >= UL.n) {
      const bool cl = c < n && W.ncbands[c] > 0;
      for (int lev = nlay; lev >= 1; --lev) sink.put_dn((size_t)(lev - 1), cl, 0., 0.);
      sink.end_sweep();
      for (int lev = 0; lev <= nlay; ++lev) sink.put_up((size_t)lev, cl, 0., 0., DRV, 0., 0.);
      sink.end_sweep();
      continue;
    }
    const Unit un = UL.u[k];
    if (CB_LW_UMAX >= 4 && un.u == 4) lw_transfer_unit<4, MC, MR, DRV, LwPartSmem, true>(T, in, W, c0, c, un.band - 1, un.g0, sink, cy);
    else if (CB_LW_UMAX == 1) lw_transfer_unit<1, MC, MR, DRV, LwPartSmem, true>(T, in, W, c0, c, un.band - 1, un.g0, sink, cy);
    else lw_transfer_unit<2, MC, MR, DRV, LwPartSmem, true>(T, in, W, c0, c, un.band - 1, un.g0, sink, cy);
  }
}


// ---- column-tile form of the transfer (lw_core.cuh: lw_tile_cell / lw_tile_sweeps) ---------------------------------------------
// One block = one tile of TW adjacent columns x one band group (lw_tile_group_bands); CB_LW_TILE_THREADS / 32 warps, specialised
// (r02 B200: with 8 warps the 7 producer warps were latency-bound on their own dependent arithmetic -- 4.2 ms at 11 % issue
// utilisation for 8192 x 60 -- the cell evaluation needs as many warps per scheduler as the unit form has):
//   warps 1.. (producers)  evaluate the (layer, column) cells of one g-point -- warp-wide rows of TW columns, so the taumol rows
//                          are read as coalesced 256-byte (128-byte) segments -- into one of two row buffers in shared memory;
//   warp 0 (consumer)      runs the down sweep, the surface and the up sweep of its lane's column over a filled buffer and adds the
//                          band-weighted radiances to the per-level sums, also in shared memory, while the producers fill the other
//                          buffer with the next g-point.
// Hand-over through named barriers (FULL / EMPTY per buffer: producers arrive, the consumer syncs, and the reverse).  At the end the
// block writes the sums of its group: `part[group][row][level][column]` -- 4 groups instead of 18 and nothing else: the rows the two
// sweeps exchange never leave the SM.  Shared memory: (2 NR nlay + NA (nlay + 1) + 2) TW doubles, NR / NA = 3 / 2 (cloud-free
// call, TW = 32) or 6 / 4 (TW = 16): 123 KB at 60 layers, 148 KB at 72.
#ifndef CB_LW_TILE_THREADS
#define CB_LW_TILE_THREADS 512
#endif
constexpr int kTileThreads = CB_LW_TILE_THREADS;
constexpr int kTileProducers = kTileThreads / 32 - 1;  // producer warps
__device__ __forceinline__ void bar_sync(int id) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(kTileThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(kTileThreads) : "memory"); }

// KC: cells per producer thread, >= ceil(nlay / (kTileProducers * 32 / TW)) (chosen by the launcher)
template <int TW, bool CLOUDY, bool MC, int KC, bool SELECT>
__global__ void __launch_bounds__(kTileThreads, 1)
    k_lw_tile(const __grid_constant__ Tables T, const __grid_constant__ In in, const __grid_constant__ Work W, int c0, int n) {
  extern __shared__ double sm[];
  constexpr int NR = CLOUDY ? kTileRowsCloudy : kTileRowsClear;
  constexpr int NA = CLOUDY ? 4 : 2;
  constexpr int FULL0 = 1, EMPTY0 = 3;  // barrier ids: FULL0 + b, EMPTY0 + b (0 is __syncthreads')
  const int nlay = in.nlay;
  const size_t prows = (size_t)nlay * TW;        // one row of a buffer
  const size_t arows = (size_t)(nlay + 1) * TW;  // one row of the sums
  double* const Pbuf = sm;                        // [2][NR][nlay][TW]
  double* const acc = sm + 2 * NR * prows;        // [NA][nlay+1][TW]
  double* const frac1 = acc + NA * arows;         // [2][TW]  Planck fraction of the lowest layer
  const int tile = blockIdx.x, group = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (SELECT) {
    // a call with clouds launches both forms; each 32-column supertile is taken by exactly one of them: the cloud-free form when
    // none of its columns has a cloudy layer (k_prep's ncbands), else the cloudy form (two 16-column tiles)
    const int c32 = (tile * TW / 32) * 32 + lane;
    const bool anyc = __any_sync(0xffffffffu, c32 < n && W.ncbands[c32] > 0);
    if (anyc != CLOUDY) return;
  }
  for (size_t i = threadIdx.x; i < NA * arows; i += kTileThreads) acc[i] = 0.0;
  __syncthreads();
  int ib0, ib1;
  lw_tile_group_bands(group, ib0, ib1);
  const int nit = band_gstart(ib1 - 1) + band_ngpt(ib1 - 1) - band_gstart(ib0);  // g-points of the group
  int it = 0;
  if (warp == 0) {
    // ---- consumer: cloud-free form lane = column; cloudy form lanes 0-15 the total-sky stream of the 16 columns, lanes 16-31 their
    // clear-sky stream (the same instruction stream: the update is selected, not branched on)
    const int col = lane % TW, stream = lane / TW;
    const int c = tile * TW + col;
    const bool live = c < n;
    const size_t gc = (size_t)c0 + (live ? c : 0);
    double* const acc_up = acc + (size_t)(2 * stream) * arows + col;
    double* const acc_dn = acc + (size_t)(2 * stream + 1) * arows + col;
    for (int ib = ib0; ib < ib1; ++ib) {
      const double* __restrict__ tp = T.base + T.totplnk + (size_t)ib * 181;
      double semiss = 1.0, plankbnd = 0.0;
      if (live) {
        semiss = in.emis[(size_t)ib * in.ncol + gc];
        plankbnd = semiss * planck_band(tp, in.tsfc[gc]);
      }
      const double wband = 0.5 * CB_LDG(T.base + T.delwave + ib);
      const int ng = band_ngpt(ib);
      for (int g = 0; g < ng; ++g, ++it) {
        const int b = it & 1;
        bar_sync(FULL0 + b);
#ifndef CB_TILE_SKIP_CONSUMER  // (timing experiments only)
        if (live)
#else
        if (live && it < 0)
#endif
          lw_tile_sweeps(Pbuf + ((size_t)b * NR + (CLOUDY && stream == 0 ? TR_TT : 0)) * prows + col, prows, TW, nlay,
                         frac1[b * TW + col] * plankbnd, 1. - semiss, wband, acc_dn, acc_up, TW);
        __threadfence_block();
        if (it + 2 < nit) bar_arrive(EMPTY0 + b);  // (nobody waits for the last two)
      }
    }
  } else {
    // ---- producers: a warp covers 32 / TW layers at a time, lanes = columns.  A thread owns the same <= KC cells for every g-point:
    // their per-band terms live in registers, and the two taumol rows of the NEXT g-point are loaded before the cells of the
    // current one are evaluated (the rows stream from HBM once; nothing else hides their latency at 8 warps per SM).
    constexpr int LPW = 32 / TW;  // layers per warp pass
    const int col = lane % TW, lsub = lane / TW;
    const int c = tile * TW + col;
    const bool live = c < n;
    const int pw = warp - 1, npw = kTileThreads / 32 - 1;
    const int lstep = npw * LPW, lfirst = pw * LPW + lsub;
    double tau_n[KC], frac_n[KC];
    TileCellCol cck[KC];
    if (live) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int l = lfirst + k * lstep;
        if (l < nlay) {
          lw_tile_cell_load(in, W, c, l, band_gstart(ib0), tau_n[k], frac_n[k]);
          cck[k] = lw_tile_cell_col<MC>(in, W, c0, c, l);
        }
      }
    }
    for (int ib = ib0; ib < ib1; ++ib) {
      TileBandCol bc{0, 0, 1.66};
      TileCellBand cbk[KC];
      if (live) {
        bc = lw_tile_band_col(W, c, ib);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int l = lfirst + k * lstep;
          if (l < nlay) cbk[k] = lw_tile_cell_band<CLOUDY>(T, in, W, c0, c, l, ib, bc, cck[k].cloudy);
        }
      }
      const int ng = band_ngpt(ib), gs = band_gstart(ib);
      for (int g = 0; g < ng; ++g, ++it) {
        const int b = it & 1;
        double tau_c[KC], frac_c[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) { tau_c[k] = tau_n[k]; frac_c[k] = frac_n[k]; }
        if (live && it + 1 < nit) {
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int l = lfirst + k * lstep;
            if (l < nlay) lw_tile_cell_load(in, W, c, l, gs + g + 1, tau_n[k], frac_n[k]);
          }
        }
        if (it >= 2) bar_sync(EMPTY0 + b);
        double* __restrict__ Pb = Pbuf + (size_t)b * NR * prows + col;
#ifndef CB_TILE_SKIP_PRODUCER  // (timing experiments only)
        if (live) {
#else
        if (live && it < 0) {
#endif
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int l = lfirst + k * lstep;
            if (l < nlay) {
              lw_tile_cell<MC, CLOUDY>(T, gs + g, bc, cck[k], cbk[k], tau_c[k], frac_c[k], Pb + (size_t)l * TW, prows);
              if (l == 0) frac1[b * TW + col] = frac_c[k];
            }
          }
        }
        __threadfence_block();
        bar_arrive(FULL0 + b);
      }
    }
  }
  __syncthreads();
  // the group's sums -> part[group][row][level][column]; the clear-sky rows only for columns with clouds (lw_reduce_level)
  const size_t pstride = (size_t)(nlay + 1) * W.ncc;
  for (size_t i = threadIdx.x; i < NA * arows; i += kTileThreads) {
    const int col = (int)(i % TW);
    const size_t rest = i / TW;
    const int lev = (int)(rest % (nlay + 1)), q = (int)(rest / (nlay + 1));
    const int c = tile * TW + col;
    if (c >= n) continue;
    if (q >= 2 && !(W.ncbands[c] > 0)) continue;
    cb::st_stream(W.part + ((size_t)group * W.npart + q) * pstride + (size_t)lev * W.ncc + c, acc[i]);
  }
}
template <int TW, bool CLOUDY>
constexpr size_t lw_tile_smem(int nlay) {
  return sizeof(double) * ((size_t)2 * (CLOUDY ? kTileRowsCloudy : kTileRowsClear) * nlay + (size_t)(CLOUDY ? 4 : 2) * (nlay + 1) + 2) * TW;
}

// McICA cloud mask with the per-column kissvec generator: one thread per column
__global__ void __launch_bounds__(kBlock) k_mask_kiss(const __grid_constant__ In in, const __grid_constant__ Work W,
                                                      int icld, int seed, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (cb::mcica::mask_column_kiss(in.play, in.cldfr, in.ncol, in.nlay, 140, 5, icld, seed, W.mask, W.ncc, c0, c)) *W.err = 9;
}

__global__ void __launch_bounds__(kBlock) k_reduce(const __grid_constant__ Tables T, const __grid_constant__ Work W,
                                                   int ngroups, const Out out, int nlay, int ncol, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int lev = blockIdx.y;
  if (c < n) lw_reduce_level(T, W, ngroups, nlay, c0, c, lev, ncol, out);
}

__global__ void __launch_bounds__(kBlock) k_heat(const __grid_constant__ Tables T, const __grid_constant__ In in,
                                                 const Out out, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (c < n) lw_heating(T, in, out, c0 + c, l);
}

#define CUDA_OK(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      e->error = std::string(#call) + ": " + cudaGetErrorString(_e);                          \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

}  // namespace

struct cb200_lw_engine {
  int device = 0;
  Tables T;
  double* d_tables = nullptr;
  Flags fl{1, 0, 2, 1, 1, 0};
  int irng = 1, permuteseed = 0;
  unsigned* d_mask_full = nullptr;
  size_t mask_full_cap = 0;
  std::vector<unsigned> ext_mask;  // caller-supplied sub-column mask [nlay][5][ncol] (cb200_lw_set_subcolumn_mask), else empty
  int ext_ncol = 0, ext_nlay = 0;
  UnitList UL;      // transfer kernel units (<= CB_LW_UMAX g-points); `part` holds one flux set per unit
  UnitList UL_tau;  // taumol kernel units (<= CB_LW_TAU_UMAX g-points)
  size_t tau_smem = 0;  // dynamic shared memory of k_lw_taumol: the largest staged (band, group) table block
  // workspace (grown on demand)
  int cap_ncc = 0, cap_nlay = 0, cap_npart = 0;
  Work W{};
  double *drv_up = nullptr, *drv_upc = nullptr;  // idrv = 1 outputs of the next run call (host or device pointers, like its outputs)
  int max_chunk = 16384;
  bool tile = true;          // column-tile form of the transfer kernel where it applies (CLIMT_B200_LW_TILE=0: the unit form)
  size_t smem_optin = 0;     // opt-in shared memory per block of the device
  int slab_bps = 0;          // > 0: the slab form of the transfer kernel with this many 4-warp blocks per SM (CLIMT_B200_LW_SLAB)
  double* d_slabs = nullptr;
  size_t slabs_cap = 0;
  int n_sm = 148;
  // host-pointer path
  cb::HostPipe pipe;
  size_t h2d_bytes = 0, d2h_bytes = 0;  // moved by the last host-pointer call
  bool skip_zero_inputs = true;         // CLIMT_B200_SKIP_ZERO_INPUTS=0 turns the all-zero scan of the host call off
  cb::ScanGuard scan_guard;             // ... and so does a host on which the scan is slower than the copy it saves
  int host_marshal = 0;       // host-pointer calls: bit 0 the h2ovmr argument is specific humidity, bit 1 tlev is computed (set_host_marshal)
  bool host_pending = false;
  std::future<int> enqueue;  // the chunk loop of a run_host_async call, running on its own host thread
  int* h_err = nullptr;
  std::string error;
  int launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;
  double unit_ms = 0.0;    // transfer kernel (the dominant one) of the last timed call
  double taumol_ms = 0.0;  // taumol kernel

  void free_work() {
    cudaFree(W.ws); cudaFree(W.idx); cudaFree(W.laytrop); cudaFree(W.ncbands); cudaFree(W.pwvcm);
    cudaFree(W.cld); cudaFree(W.scr); cudaFree(W.part); cudaFree(W.ovl); cudaFree(W.err); cudaFree(W.mask);
    W = Work{};
    cap_ncc = cap_nlay = cap_npart = 0;
  }
  int ensure_work(int ncc, int nlay) {
    cb200_lw_engine* e = this;
    const int npart = fl.idrv == 1 ? 6 : 4;
    if (ncc <= cap_ncc && nlay <= cap_nlay && npart <= cap_npart && W.ws) { W.npart = npart; return 0; }
    free_work();
    const size_t n = (size_t)ncc, L = (size_t)nlay;
    CUDA_OK(cudaMalloc(&W.ws, sizeof(double) * NF * L * n));
    CUDA_OK(cudaMalloc(&W.idx, sizeof(int) * L * n));
    CUDA_OK(cudaMalloc(&W.laytrop, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&W.ncbands, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&W.pwvcm, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&W.cld, sizeof(double) * 32 * L * n));
    CUDA_OK(cudaMalloc(&W.scr, sizeof(double) * 140 * NSCR * L * n));
    CUDA_OK(cudaMalloc(&W.part, sizeof(double) * ((UL.n + CB_LW_GROUP - 1) / CB_LW_GROUP) * npart * (L + 1) * n));
    W.npart = npart;
    cap_npart = npart;
    CUDA_OK(cudaMalloc(&W.ovl, sizeof(double) * OV_NROWS * (L + 2) * n));
    CUDA_OK(cudaMalloc(&W.mask, sizeof(unsigned) * 5 * L * n));
    CUDA_OK(cudaMalloc(&W.err, sizeof(int)));
    CUDA_OK(cudaMemset(W.err, 0, sizeof(int)));
    cap_ncc = ncc;
    cap_nlay = nlay;
    return 0;
  }
};

extern "C" const char* cb200_global_error(void) { return cb::g_error.c_str(); }

extern "C" int cb200_lw_create(cb200_lw_engine** out, const char* table_blob, const double constants[11], int device) {
  *out = nullptr;
  auto* e = new cb200_lw_engine();
  try {
    Constants k;
    std::memcpy(&k, constants, sizeof(k));
    std::vector<double> img;
    build_tables(table_blob, k, img, e->T);
    e->device = device;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) throw std::runtime_error(std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    ce = cudaMalloc(&e->d_tables, img.size() * sizeof(double));
    if (ce != cudaSuccess) throw std::runtime_error(std::string("cudaMalloc(tables): ") + cudaGetErrorString(ce));
    ce = cudaMemcpy(e->d_tables, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) throw std::runtime_error(std::string("cudaMemcpy(tables): ") + cudaGetErrorString(ce));
    e->T.base = e->d_tables;
    e->UL.n = build_units(e->UL.u, CB_LW_UMAX);
    e->UL_tau.n = build_units(e->UL_tau.u, CB_LW_TAU_UMAX);
    for (int b = 1; b <= 16; ++b)
      if (lw_band_staged(b)) e->tau_smem = std::max(e->tau_smem, (size_t)e->T.b[b - 1].rows * TGW * sizeof(double));
    if (const char* mc = std::getenv("CLIMT_B200_MAX_CHUNK")) e->max_chunk = std::max(128, std::atoi(mc));
    if (const char* z = std::getenv("CLIMT_B200_SKIP_ZERO_INPUTS")) e->skip_zero_inputs = std::atoi(z) != 0;
    if (const char* sb = std::getenv("CLIMT_B200_LW_SLAB")) e->slab_bps = std::max(0, std::atoi(sb));
    cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (const char* tl = std::getenv("CLIMT_B200_LW_TILE")) e->tile = std::atoi(tl) != 0;
    {
      int optin = 0;
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
      e->smem_optin = (size_t)optin;
    }
    cudaMallocHost(&e->h_err, sizeof(int));
    cudaEventCreate(&e->ev0);
    cudaEventCreate(&e->ev1);
    cudaEventCreate(&e->evm);
  } catch (std::exception& ex) {
    cb::set_global_error(ex.what());
    delete e;
    return -1;
  }
  *out = e;
  return 0;
}

extern "C" void cb200_lw_destroy(cb200_lw_engine* e) {
  if (!e) return;
  if (e->enqueue.valid()) e->enqueue.wait();  // a host call still being enqueued
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  e->free_work();
  cudaFree(e->d_tables);
  cudaFree(e->d_mask_full);
  cudaFree(e->d_slabs);
  e->pipe.destroy();
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->evm) cudaEventDestroy(e->evm);
  delete e;
}

extern "C" int cb200_lw_set_options(cb200_lw_engine* e, int icld, int idrv, int inflag, int iceflag, int liqflag) {
  if (icld < 0 || icld > 3) icld = 2;  // rrtmg_lw_rad.nomcica.f90:437
  if (idrv != 0 && idrv != 1) { e->error = "idrv must be 0 or 1 (rrtmg_lw_rad.nomcica.f90:192)"; return -2; }
  e->fl = Flags{icld, idrv, inflag, iceflag, liqflag, e->fl.mcica};
  return 0;
}

// idrv = 1 (calculate_change_up_flux): where the next run call writes d(upward flux)/d(surface temperature), total and clear
// sky, (nlay+1, ncol) each -- duflx_dt / duflxc_dt of rrtmg_lw_c_binder.f90:176-256.  Host pointers for run_host, device
// pointers for run_device, like that call's own outputs.  Cleared by passing nulls.
extern "C" int cb200_lw_set_derivative_outputs(cb200_lw_engine* e, double* duflx_dt, double* duflxc_dt) {
  e->drv_up = duflx_dt;
  e->drv_upc = duflxc_dt;
  return 0;
}

extern "C" int cb200_lw_set_mcica(cb200_lw_engine* e, int enabled, int irng, int permuteseed) {
  e->fl.mcica = enabled ? 1 : 0;
  e->irng = irng != 0 ? 1 : 0;  // mcica_subcol_gen_lw.f90:303
  e->permuteseed = permuteseed;
  return 0;
}

extern "C" const char* cb200_lw_last_error(cb200_lw_engine* e) { return e ? e->error.c_str() : cb::g_error.c_str(); }
extern "C" int cb200_lw_last_launches(cb200_lw_engine* e) { return e->launches; }
extern "C" int cb200_lw_enable_timing(cb200_lw_engine* e, int on) { e->timing = on != 0; return 0; }
extern "C" double cb200_lw_last_unit_kernel_ms(cb200_lw_engine* e) { return e->unit_ms; }
extern "C" double cb200_lw_last_taumol_kernel_ms(cb200_lw_engine* e) { return e->taumol_ms; }

static In make_in(int ncol, int nlay, const cb200_lw_inputs* p) {
  In in;
  in.ncol = ncol; in.nlay = nlay;
  in.play = p->play; in.plev = p->plev; in.tlay = p->tlay; in.tlev = p->tlev; in.tsfc = p->tsfc;
  in.h2o = p->h2ovmr; in.o3 = p->o3vmr; in.co2 = p->co2vmr; in.ch4 = p->ch4vmr; in.n2o = p->n2ovmr; in.o2 = p->o2vmr;
  in.cfc11 = p->cfc11vmr; in.cfc12 = p->cfc12vmr; in.cfc22 = p->cfc22vmr; in.ccl4 = p->ccl4vmr;
  in.emis = p->emis; in.cldfr = p->cldfr; in.taucld = p->taucld; in.cicewp = p->cicewp; in.cliqwp = p->cliqwp;
  in.reice = p->reice; in.reliq = p->reliq; in.tauaer = p->tauaer;
  return in;
}

// One chunk of columns [c0, c0+n) of `in` through the kernels on stream `st` (the engine's single workspace).
static int launch_chunk(cb200_lw_engine* e, const In& in, const Out& out, Work& W, int c0, int n, int out_ncol, bool mc,
                        cudaStream_t st) {
  const int nlay = in.nlay;
  const int gx = (n + kBlock - 1) / kBlock;
  if (mc && e->irng == 0) { k_mask_kiss<<<(n + 31) / 32, 32, 0, st>>>(in, W, e->fl.icld, e->permuteseed, c0, n); e->launches += 1; }
  k_prep_layer<<<dim3(gx, nlay), kBlock, 0, st>>>(e->T, in, e->fl, W, c0, n);
  // column-serial kernels: one warp per block so that even 8 192 columns spread over every SM
  const int gw = (n + 31) / 32;
  k_prep<<<gw, 32, 0, st>>>(e->T, in, e->fl, W, c0, n);
  if (e->fl.icld >= 1) { k_cld_scale<<<dim3(gx, nlay), kBlock, 0, st>>>(in, e->fl, W, c0, n); e->launches += 1; }
  if (e->timing) cudaEventRecord(e->ev0, st);
  k_lw_taumol<<<dim3(gx, e->UL_tau.n, CB_LW_LAYER_CHUNKS), kBlock, e->tau_smem, st>>>(e->T, in, W, e->UL_tau, c0, n);
  if (e->timing) cudaEventRecord(e->evm, st);
  // non-McICA: icld = 1 -> rtrn (random overlap); icld = 2, 3 -> rtrnmr (maximum-random), rrtmg_lw_rad.nomcica.f90:527-541
  const dim3 gu((n + 31) / 32, (e->UL.n + CB_LW_GROUP - 1) / CB_LW_GROUP), bu(32, CB_LW_GROUP);
  int ngroups = (int)gu.y;
  // the column-tile form (rows carried in shared memory) covers rtrn and rtrnmc without the surface-temperature derivative
  const bool cloudy_call = e->fl.icld >= 1;
  const size_t tile_smem = cloudy_call ? lw_tile_smem<16, true>(nlay) : lw_tile_smem<32, false>(nlay);
  // cells per producer thread of the two forms (cloud-free: 32-column tiles, one layer per warp pass; cloudy: 16 columns, two layers)
  const int kc32 = (nlay + kTileProducers - 1) / kTileProducers, kc16 = (nlay + 2 * kTileProducers - 1) / (2 * kTileProducers);
  const size_t smem32 = lw_tile_smem<32, false>(nlay), smem16 = lw_tile_smem<16, true>(nlay);
  if (e->tile && W.npart == 4 && (mc || e->fl.icld < 2) && smem32 <= e->smem_optin && smem16 <= e->smem_optin && kc32 <= 12) {
#define CB_TILE(TW, CL, MCF, KC, SEL, SMEM)                                                                                        \
    do {                                                                                                                             \
      CUDA_OK(cudaFuncSetAttribute(k_lw_tile<TW, CL, MCF, KC, SEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));      \
      k_lw_tile<TW, CL, MCF, KC, SEL><<<dim3((n + TW - 1) / TW, kTileGroups), kTileThreads, SMEM, st>>>(e->T, in, W, c0, n);         \
    } while (0)
#define CB_TILE_KC(TW, CL, MCF, SEL, SMEM, KCV)                                                                     \
    do {                                                                                                             \
      if (KCV <= 3) CB_TILE(TW, CL, MCF, 3, SEL, SMEM); else if (KCV <= 5) CB_TILE(TW, CL, MCF, 5, SEL, SMEM);       \
      else if (KCV <= 9) CB_TILE(TW, CL, MCF, 9, SEL, SMEM); else CB_TILE(TW, CL, MCF, 12, SEL, SMEM);               \
    } while (0)
    if (!cloudy_call) {
      CB_TILE_KC(32, false, false, false, smem32, kc32);
    } else {
      // both forms; every block checks which of the two owns its 32-column supertile (k_lw_tile, SELECT)
      CB_TILE_KC(32, false, false, true, smem32, kc32);
      if (mc) CB_TILE_KC(16, true, true, true, smem16, kc16);
      else CB_TILE_KC(16, true, false, true, smem16, kc16);
      e->launches += 1;
    }
#undef CB_TILE_KC
#undef CB_TILE
    ngroups = kTileGroups;
  } else if (e->slab_bps > 0) {
    const int nb = (int)std::min<size_t>((size_t)gu.x * gu.y, (size_t)e->n_sm * e->slab_bps);
    const size_t need = (size_t)nb * CB_LW_GROUP * nlay * CB_LW_UMAX * 4 * 32;
    if (need > e->slabs_cap) {
      cudaStreamSynchronize(st);
      cudaFree(e->d_slabs);
      e->d_slabs = nullptr;
      e->slabs_cap = 0;
      CUDA_OK(cudaMalloc(&e->d_slabs, need * sizeof(double)));
      e->slabs_cap = need;
    }
    double* sl = e->d_slabs;
    if (W.npart == 6) {
      if (mc) k_units_slab<true, false, true><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
      else if (e->fl.icld >= 2) k_units_slab<false, true, true><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
      else k_units_slab<false, false, true><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
    } else {
      if (mc) k_units_slab<true, false><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
      else if (e->fl.icld >= 2) k_units_slab<false, true><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
      else k_units_slab<false, false><<<nb, bu, 0, st>>>(e->T, in, W, e->UL, sl, c0, n);
    }
  } else if (W.npart == 6) {
    if (mc) k_units<true, false, true><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
    else if (e->fl.icld >= 2) k_units<false, true, true><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
    else k_units<false, false, true><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
  } else {
    if (mc) k_units<true, false><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
    else if (e->fl.icld >= 2) k_units<false, true><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
    else k_units<false, false><<<gu, bu, 0, st>>>(e->T, in, W, e->UL, c0, n);
  }
  if (e->timing) cudaEventRecord(e->ev1, st);
  k_reduce<<<dim3(gx, nlay + 1), kBlock, 0, st>>>(e->T, W, ngroups, out, nlay, out_ncol, c0, n);
  k_heat<<<dim3(gx, nlay), kBlock, 0, st>>>(e->T, in, out, c0, n);
  e->launches += 6;
  if (e->timing) {
    CUDA_OK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->evm, e->ev1);
    e->unit_ms += ms;
    cudaEventElapsedTime(&ms, e->ev0, e->evm);
    e->taumol_ms += ms;
  }
  return 0;
}

// Mersenne-twister McICA mask: one serial stream for the whole call (bit parity with climt's default RNG), so it is
// drawn on the host from the host copy of the cloud fraction and uploaded once, [lay][word][ncol].
static int upload_mt_mask(cb200_lw_engine* e, const double* h_cldfr, int ncol, int nlay, cudaStream_t st) {
  std::vector<unsigned> h_mask;
  if (!e->ext_mask.empty()) {
    if (e->ext_ncol != ncol || e->ext_nlay != nlay) { e->error = "sub-column mask was set for a different ncol/nlay"; return -3; }
    h_mask = e->ext_mask;
  } else {
    cb::mcica::mask_mt_host(h_cldfr, ncol, nlay, 140, 5, e->fl.icld, e->permuteseed, h_mask);
  }
  if (h_mask.size() > e->mask_full_cap) {
    cudaFree(e->d_mask_full);
    e->d_mask_full = nullptr;
    e->mask_full_cap = 0;
    CUDA_OK(cudaMalloc(&e->d_mask_full, h_mask.size() * sizeof(unsigned)));
    e->mask_full_cap = h_mask.size();
  }
  CUDA_OK(cudaMemcpyAsync(e->d_mask_full, h_mask.data(), h_mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int cb200_lw_run_device(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* pin,
                                   const cb200_lw_outputs* pout, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrtm.f90:31)"; return -3; }
  CUDA_OK(cudaSetDevice(e->device));
  int chunk = ncol < e->max_chunk ? ncol : e->max_chunk;
  chunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(chunk, nlay)) return -1;
  Work W = e->W;
  W.ncc = chunk;
  const In in = make_in(ncol, nlay, pin);
  Out out{pout->uflx, pout->dflx, pout->hr, pout->uflxc, pout->dflxc, pout->hrc};
  if (e->fl.idrv == 1) {
    if (!e->drv_up || !e->drv_upc) { e->error = "idrv = 1 needs cb200_lw_set_derivative_outputs before the run call"; return -3; }
    out.duflx_dt = e->drv_up;
    out.duflxc_dt = e->drv_upc;
  }
  e->launches = 0;
  e->unit_ms = 0.0;
  e->taumol_ms = 0.0;
  const bool mc = e->fl.mcica && e->fl.icld >= 1;
  W.mstride = chunk;
  W.moff = 0;
  if (mc && e->irng == 1) {
    std::vector<double> h_cld((size_t)nlay * ncol);
    CUDA_OK(cudaMemcpyAsync(h_cld.data(), in.cldfr, h_cld.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (upload_mt_mask(e, h_cld.data(), ncol, nlay, st)) return -1;
    W.mask = e->d_mask_full;
    W.mstride = ncol;
  }
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int n = (ncol - c0) < chunk ? (ncol - c0) : chunk;
    if (mc && e->irng == 1) W.moff = c0;
    if (launch_chunk(e, in, out, W, c0, n, ncol, mc, st)) return -1;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cb200_lw_check(cb200_lw_engine* e) {
  if (!e->W.err) return 0;
  CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaMemcpy(e->h_err, e->W.err, sizeof(int), cudaMemcpyDeviceToHost));
  const int code = *e->h_err;
  if (code) {
    static const char* msg[] = {"", "ICE RADIUS TOO SMALL", "ICE RADIUS OUT OF BOUNDS",
                                "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS", "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS",
                                "", "", "", "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA",
                                "MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS."};
    e->error = msg[code < 10 ? code : 2];  // message text of the Fortran `stop` (rrtmg_lw_cldprop.f90:193-243)
    cudaMemset(e->W.err, 0, sizeof(int));
  }
  return code;
}

// Host-pointer call (what the reference-named wrapper and the Python component use).  Column chunks flow through
// the three-stream pipeline of cb::HostPipe; arrays the option flags make dead are not transferred at all
// (cloud inputs when icld = 0, taucld unless inflag = 0 -- see DESIGN.md "host path").
static int lw_host_enqueue(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* hin, const cb200_lw_outputs* hout) {
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrtm.f90:31)"; return -3; }
  CUDA_OK(cudaSetDevice(e->device));
  cb::HostPipe& P = e->pipe;
  CUDA_OK(P.init());
  const int L = nlay;
  // rows (of ncol doubles) of the 23 inputs in cb200_lw_inputs order, then of the 6 outputs
  const int irows[23] = {L, L + 1, L, L + 1, 1, L, L, L, L, L, L, L, L, L, L, 16, L, L, L, L, L, L, 16 * L};
  int inner[23];
  for (int i = 0; i < 23; ++i) inner[i] = 1;
  inner[17] = 16;  // taucld(nbndlw, ncol, nlay): band-fastest
  const bool drv = e->fl.idrv == 1;
  if (drv && (!e->drv_up || !e->drv_upc)) { e->error = "idrv = 1 needs cb200_lw_set_derivative_outputs before the run call"; return -3; }
  const int nout = drv ? 8 : 6;
  const int orows[8] = {L + 1, L + 1, L, L + 1, L + 1, L, L + 1, L + 1};
  bool used[23];
  for (int i = 0; i < 23; ++i) used[i] = true;
  const bool clouds = e->fl.icld >= 1;
  const bool mc = e->fl.mcica && clouds;
  // 16 cldfr, 17 taucld, 18 cicewp, 19 cliqwp, 20 reice, 21 reliq
  if (!clouds) for (int i = 16; i <= 21; ++i) used[i] = false;
  if (e->fl.inflag != 0) used[17] = false;
  const double* const* hp = reinterpret_cast<const double* const*>(hin);
  const bool mar_q = (e->host_marshal & 1) != 0, mar_t = (e->host_marshal & 2) != 0;  // the components' marshal arithmetic on the device
  // Inputs the reference ABI always carries but that are all zero in most model states -- the four halocarbons (11..14), the
  // 16-band aerosol optical depth (22; LW `iaer = 10` is hard-wired, rrtmg_lw_rad.nomcica.f90:442) and the cloud fraction (16) --
  // are scanned on the host chunk by chunk (cb::all_zero_parallel) and, when zero, set in HBM by a memset instead of crossing PCIe.
  bool zero[23];
  for (int i = 0; i < 23; ++i) zero[i] = false;
  size_t irow_tot = 0, orow_tot = 0;
  for (int i = 0; i < 23; ++i) if (used[i]) irow_tot += (size_t)irows[i] * inner[i];
  for (int i = 0; i < nout; ++i) orow_tot += (size_t)orows[i];
  e->h2d_bytes = 0;
  e->d2h_bytes = orow_tot * (size_t)ncol * sizeof(double);
  int chunk = ncol < P.chunk ? ncol : P.chunk;
  const int wchunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(wchunk, nlay)) return -1;
  CUDA_OK(P.ensure(irow_tot * (size_t)chunk, orow_tot * (size_t)chunk));
  double* hop[8];
  for (int i = 0; i < 6; ++i) hop[i] = reinterpret_cast<double* const*>(hout)[i];
  hop[6] = e->drv_up;
  hop[7] = e->drv_upc;
  Work W = e->W;
  W.ncc = wchunk;
  W.mstride = wchunk;
  W.moff = 0;
  e->launches = 0;
  e->unit_ms = 0.0;
  e->taumol_ms = 0.0;
  if (mc && e->irng == 1) {
    if (upload_mt_mask(e, hin->cldfr, ncol, nlay, P.s_cmp)) return -1;
    W.mask = e->d_mask_full;
    W.mstride = ncol;
  }
  // pageable caller arrays go through the pipe's pinned staging slots (engine_common.h); page-locked ones are read / written in place
  bool pg_in[23], pg_out[8], any_pg_in = false, any_pg_out = false;
  for (int i = 0; i < 23; ++i) { pg_in[i] = used[i] && hp[i] && !cb::HostPipe::dma_able(hp[i]); any_pg_in |= pg_in[i]; }
  for (int i = 0; i < nout; ++i) { pg_out[i] = !cb::HostPipe::dma_able(hop[i]); any_pg_out |= pg_out[i]; }
  if (any_pg_in || any_pg_out) CUDA_OK(P.ensure_staging(any_pg_in ? irow_tot * (size_t)chunk : 0, any_pg_out ? orow_tot * (size_t)chunk : 0));
  struct { int s, c0, n; bool valid; } prev{0, 0, 0, false};
  auto drain_outputs = [&](int ps, int pc0, int pn) -> cudaError_t {  // staged outputs of a finished chunk -> the caller's arrays
    cudaError_t ce = cudaEventSynchronize(P.out_done[ps]);
    if (ce != cudaSuccess) return ce;
    size_t o = 0;
    for (int i = 0; i < nout; ++i) {
      if (pg_out[i]) cb::HostPipe::scatter_staged_finish(hop[i], P.h_out[ps] + o, orows[i], ncol, pc0, pn);
      o += (size_t)orows[i] * pn;
    }
    return cudaSuccess;
  };
  int k = 0;
  for (int c0 = 0, n = 0; c0 < ncol; c0 += n, ++k) {
    n = P.chunk_size(k, ncol - c0);
    const int s = k & 1;
    if (e->skip_zero_inputs) {
      const int cand[6] = {11, 12, 13, 14, 22, 16};
      const int nc = clouds ? 6 : 5;
      cb::ZeroView zv[6];
      bool zz[6];
      size_t zbytes = 0;
      for (int j = 0; j < nc; ++j) {
        zv[j] = cb::ZeroView{hp[cand[j]], (size_t)irows[cand[j]], (size_t)ncol, (size_t)c0, (size_t)n};
        zbytes += (size_t)irows[cand[j]] * n * sizeof(double);
      }
      const double t0 = cb::HostPipe::now_ms();
      cb::all_zero_parallel(zv, nc, zz);
      if (!e->scan_guard.note(zbytes, (cb::HostPipe::now_ms() - t0) * 1e-3)) e->skip_zero_inputs = false;
      for (int j = 0; j < nc; ++j) zero[cand[j]] = zz[j];
      // no cloud in these columns: water paths and particle sizes are never read (cldprop / cldprmc skip layers below cldmin)
      for (int i = 17; i <= 21; ++i) zero[i] = clouds && zero[16];
    } else {
      for (int i = 0; i < 23; ++i) zero[i] = false;  // (the guard may have stopped the scan in the middle of a call)
    }
    // H2D: the slot is free once the chunk that last used it has been computed
    CUDA_OK(cudaStreamWaitEvent(P.s_in, P.cmp_done[s], 0));
    if (any_pg_in) CUDA_OK(cudaEventSynchronize(P.in_done[s]));  // the copies that last read this staging slot have left the host
    P.mark(P.s_in, k, 0);
    cb200_lw_inputs din;
    const double** dp = reinterpret_cast<const double**>(&din);
    size_t off = 0;
    for (int i = 0; i < 23; ++i) {
      if (!used[i]) { dp[i] = nullptr; continue; }
      if (i == 3 && mar_t) {
        // tlev: computed on the device from this chunk's tlay, tsfc, play, plev (below), nothing crosses PCIe
      } else if (zero[i]) {
        CUDA_OK(cudaMemsetAsync(P.d_in[s] + off, 0, (size_t)irows[i] * inner[i] * n * sizeof(double), P.s_in));
      } else {
        if (pg_in[i]) CUDA_OK(P.gather_staged(P.d_in[s] + off, P.h_in[s] + off, hp[i], irows[i], ncol, c0, n, inner[i]));
        else CUDA_OK(P.gather(P.d_in[s] + off, hp[i], irows[i], ncol, c0, n, inner[i]));
        e->h2d_bytes += (size_t)irows[i] * inner[i] * n * sizeof(double);
      }
      dp[i] = P.d_in[s] + off;
      off += (size_t)irows[i] * inner[i] * n;
    }
    CUDA_OK(cudaEventRecord(P.in_done[s], P.s_in));
    P.mark(P.s_in, k, 1);
    double* dop[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    off = 0;
    for (int i = 0; i < nout; ++i) { dop[i] = P.d_out[s] + off; off += (size_t)orows[i] * n; }
    // compute: inputs landed, and the output slot has been drained
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.in_done[s], 0));
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.out_done[s], 0));
    P.mark(P.s_cmp, k, 2);
    if (mar_q || mar_t)  // the components' marshal arithmetic (util.py:47-142) on this chunk, before anything reads it
      CUDA_OK(cb::marshal_launch(n, nlay, din.h2ovmr, din.tlay, din.tsfc, din.play, din.plev, mar_q ? const_cast<double*>(din.h2ovmr) : nullptr,
                                 mar_t ? const_cast<double*>(din.tlev) : nullptr, P.s_cmp));
    const In in = make_in(n, nlay, &din);
    Out out{dop[0], dop[1], dop[2], dop[3], dop[4], dop[5]};
    out.duflx_dt = dop[6];
    out.duflxc_dt = dop[7];
    if (mc && e->irng == 1) W.moff = c0;
    if (launch_chunk(e, in, out, W, 0, n, n, mc, P.s_cmp)) return -1;
    CUDA_OK(cudaEventRecord(P.cmp_done[s], P.s_cmp));
    P.mark(P.s_cmp, k, 3);
    // D2H
    CUDA_OK(cudaStreamWaitEvent(P.s_out, P.cmp_done[s], 0));
    {
      size_t o = 0;
      for (int i = 0; i < nout; ++i) {
        if (pg_out[i]) CUDA_OK(P.scatter_staged_issue(P.h_out[s] + o, dop[i], orows[i], n));
        else CUDA_OK(P.scatter(hop[i], dop[i], orows[i], ncol, c0, n));
        o += (size_t)orows[i] * n;
      }
    }
    CUDA_OK(cudaEventRecord(P.out_done[s], P.s_out));
    P.mark(P.s_out, k, 4);
    // the previous chunk's staged outputs are copied out while this chunk runs
    if (any_pg_out && prev.valid) CUDA_OK(drain_outputs(prev.s, prev.c0, prev.n));
    prev = {s, c0, n, true};
  }
  if (any_pg_out && prev.valid) CUDA_OK(drain_outputs(prev.s, prev.c0, prev.n));
  CUDA_OK(cudaGetLastError());
  return 0;
}

// The asynchronous form returns at once: the chunk loop (zero scans, copies, launches) runs on a host thread of its own, so that a
// caller that starts the longwave and the shortwave call back to back has both pipelines feeding the GPU from the first chunk on
// (r01 trace, 8192 x 60: issued from one thread the second engine started 1.2 ms late and finished 1.8 ms after the first).
// The host-pointer calls can do the components' marshal arithmetic themselves, on the device, chunk by chunk: flags bit 0 -- the
// h2ovmr argument holds SPECIFIC HUMIDITY (kg/kg) and is converted to a volume mixing ratio (climt/_core/util.py:47-86); bit 1 --
// tlev is ignored (may be NULL) and computed from tlay, tsfc, play, plev by ln-p interpolation (util.py:89-142).  0 = the plain
// reference ABI.  (The components spend more host time in those two numpy expressions than the whole call takes on the GPU.)
extern "C" int cb200_lw_set_host_marshal(cb200_lw_engine* e, int flags) {
  e->host_marshal = flags & 3;
  return 0;
}

extern "C" int cb200_lw_run_host_async(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* hin,
                                       const cb200_lw_outputs* hout) {
  if (e->host_pending) { e->error = "a previous run_host_async call has not been waited for"; return -3; }
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrtm.f90:31)"; return -3; }
  const cb200_lw_inputs in = *hin;
  const cb200_lw_outputs out = *hout;
  e->host_pending = true;
  e->enqueue = std::async(std::launch::async, [e, ncol, nlay, in, out] { return lw_host_enqueue(e, ncol, nlay, &in, &out); });
  return 0;
}

// Completes the call started by cb200_lw_run_host_async: outputs are in the caller's buffers on return.
extern "C" int cb200_lw_wait(cb200_lw_engine* e) {
  if (!e->host_pending) return 0;
  e->host_pending = false;
  int rc = 0;
  if (e->enqueue.valid()) rc = e->enqueue.get();
  CUDA_OK(cudaSetDevice(e->device));
  if (rc) {
    // the chunk loop failed half-way: chunks already enqueued still copy from / into the caller's buffers -- drain them
    // before the caller is told (and frees or reuses those buffers)
    if (e->pipe.s_in) { cudaStreamSynchronize(e->pipe.s_in); cudaStreamSynchronize(e->pipe.s_cmp); cudaStreamSynchronize(e->pipe.s_out); }
    return rc;
  }
  CUDA_OK(cudaStreamSynchronize(e->pipe.s_out));
  e->pipe.trace_dump("LW");
  return cb200_lw_check(e);
}

extern "C" int cb200_lw_run_host(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* hin,
                                 const cb200_lw_outputs* hout) {
  if (e->host_pending) { e->error = "a previous run_host_async call has not been waited for"; return -3; }
  if (int rc = lw_host_enqueue(e, ncol, nlay, hin, hout)) {
    if (e->pipe.s_in) { cudaStreamSynchronize(e->pipe.s_in); cudaStreamSynchronize(e->pipe.s_cmp); cudaStreamSynchronize(e->pipe.s_out); }
    return rc;
  }
  e->host_pending = true;
  return cb200_lw_wait(e);
}

// 1: the host calls scan always-present-but-usually-zero inputs and memset them in HBM; 0: turned off by
// CLIMT_B200_SKIP_ZERO_INPUTS=0; -1: turned off by the engine itself because the scan ran slower than the copy it saves
// (cb::ScanGuard) -- sticky for the life of the engine, so a benchmark can report which regime it measured.
extern "C" int cb200_lw_zero_scan_state(cb200_lw_engine* e) {
  if (e->skip_zero_inputs) return 1;
  return e->scan_guard.slow >= 3 ? -1 : 0;
}

extern "C" void cb200_lw_last_transfer_bytes(cb200_lw_engine* e, double* h2d, double* d2h) {
  *h2d = (double)e->h2d_bytes;
  *d2h = (double)e->d2h_bytes;
}

// ---------------------------------------------------------------------------------------------
// Reference-named entry points bound to one process-global engine.
namespace {
double g_consts[11] = {0};
cb200_lw_engine* g_engine = nullptr;

// The reduced-table blob of the reference-named init symbol: $CLIMT_B200_LW_TABLES, else $CLIMT_B200_CACHE/<name>, else the copy
// the Python side regenerates next to the library (data/_cache), else the one shipped with the package (data/).
std::string default_blob() { return cb::find_table_blob("CLIMT_B200_LW_TABLES", "rrtmg_lw_reduced.blob", (void*)&cb200_lw_create); }
}  // namespace

extern "C" void rrtmg_set_constants(double* pi, double* grav, double* planck, double* boltz, double* clight,
                                    double* avogad, double* alosmt, double* gascon, double* sbcnst, double* secdy) {
  double v[10] = {*pi, *grav, *planck, *boltz, *clight, *avogad, *alosmt, *gascon, *sbcnst, *secdy};
  std::memcpy(g_consts, v, sizeof v);
}

extern "C" void rrtmg_lw_ini_wrapper(double* cpdair) {
  g_consts[10] = *cpdair;
  if (g_engine) { cb200_lw_destroy(g_engine); g_engine = nullptr; }
  int dev = 0;
  if (const char* d = std::getenv("CLIMT_B200_DEVICE")) dev = std::atoi(d);
  if (cb200_lw_create(&g_engine, default_blob().c_str(), g_consts, dev)) {
    std::fprintf(stderr, "climt_b200: rrtmg_lw_ini_wrapper failed: %s\n", cb200_global_error());
    g_engine = nullptr;  // every later wrapper call NaN-fills its outputs and reports
  }
}

// The reference wrappers are void and the Fortran under them ends the process with `stop` on invalid input.  Ending the
// caller's interpreter is not acceptable for a library, returning stale buffers is worse: on ANY failure every output array
// is filled with NaN, the message goes to stderr and stays readable through cb200_global_error().
namespace {
void lw_wrapper_fail(const std::string& msg, int ncol, int nlay, double* uflx, double* dflx, double* hr, double* uflxc,
                     double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt, bool derivs) {
  std::fprintf(stderr, "climt_b200: %s\n", msg.c_str());
  cb::set_global_error(msg);
  const size_t n1 = (size_t)ncol * (nlay + 1), n0 = (size_t)ncol * nlay;
  cb::mcica::nan_fill(uflx, n1); cb::mcica::nan_fill(dflx, n1); cb::mcica::nan_fill(hr, n0);
  cb::mcica::nan_fill(uflxc, n1); cb::mcica::nan_fill(dflxc, n1); cb::mcica::nan_fill(hrc, n0);
  if (derivs) { cb::mcica::nan_fill(duflx_dt, n1); cb::mcica::nan_fill(duflxc_dt, n1); }
}
int lw_ngb(int g) {  // 0-based band of 0-based g-point g (ngb, rrtmg_lw_init.f90:306-311)
  int b = 0;
  while (b < 15 && g >= kGS[b + 1]) ++b;
  return b;
}
}  // namespace

extern "C" void rrtmg_lw_nomcica_wrapper(int* ncol, int* nlay, int* icld, int* idrv, double* play, double* plev,
                                         double* tlay, double* tlev, double* tsfc, double* h2ovmr, double* o3vmr,
                                         double* co2vmr, double* ch4vmr, double* n2ovmr, double* o2vmr,
                                         double* cfc11vmr, double* cfc12vmr, double* cfc22vmr, double* ccl4vmr,
                                         double* emis, int* inflglw, int* iceflglw, int* liqflglw, double* cldfr,
                                         double* taucld, double* cicewp, double* cliqwp, double* reice, double* reliq,
                                         double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc,
                                         double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt) {
  const bool derivs = *idrv == 1;
  auto fail = [&](const std::string& m) {
    lw_wrapper_fail(m, *ncol, *nlay, uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, derivs);
  };
  if (!g_engine) return fail("rrtmg_lw_ini_wrapper has not been called (or failed)");
  if (*icld < 0 || *icld > 3) *icld = 2;
  cb200_lw_set_mcica(g_engine, 0, 1, 0);
  if (cb200_lw_set_options(g_engine, *icld, *idrv, *inflglw, *iceflglw, *liqflglw)) return fail(cb200_lw_last_error(g_engine));
  cb200_lw_inputs in{play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr,
                     cfc22vmr, ccl4vmr, emis, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer};
  cb200_lw_outputs out{uflx, dflx, hr, uflxc, dflxc, hrc};
  if (derivs) cb200_lw_set_derivative_outputs(g_engine, duflx_dt, duflxc_dt);
  const int rc = cb200_lw_run_host(g_engine, *ncol, *nlay, &in, &out);
  cb200_lw_set_derivative_outputs(g_engine, nullptr, nullptr);
  if (rc) fail(cb200_lw_last_error(g_engine));
}

// Sub-column generator of the reference ABI (rrtmg_lw_c_binder.f90:50-92 -> mcica_subcol_gen_lw.f90:48-154): fills the
// caller's (ngptlw, ncol, nlay) arrays.  Both generators run on the host here -- the Mersenne twister is one serial stream
// by definition, and for kissvec the 4 x 140 x ncol x nlay doubles this ABI asks for dominate any generator cost.  (The
// engine's own McICA path, cb200_lw_set_mcica + cb200_lw_run_*, keeps the mask as bits in HBM and never builds these arrays.)
extern "C" void mcica_subcol_lw_wrapper(int* iplon, int* ncol, int* nlay, int* icld, int* permuteseed, int* irng,
                                        double* play, double* cldfrac, double* ciwp, double* clwp, double* rei,
                                        double* rel, double* tauc, double* cldfmcl, double* ciwpmcl, double* clwpmcl,
                                        double* reicmcl, double* relqmcl, double* taucmcl) {
  (void)iplon;
  const int nc = *ncol, nl = *nlay;
  if (*icld == 0) return;  // mcica_subcol_gen_lw.f90:119
  const size_t n3 = (size_t)140 * nc * nl, n2 = (size_t)nc * nl;
  auto fail = [&](const std::string& m) {
    std::fprintf(stderr, "climt_b200: %s\n", m.c_str());
    cb::set_global_error(m);
    cb::mcica::nan_fill(cldfmcl, n3); cb::mcica::nan_fill(ciwpmcl, n3); cb::mcica::nan_fill(clwpmcl, n3);
    cb::mcica::nan_fill(taucmcl, n3); cb::mcica::nan_fill(reicmcl, n2); cb::mcica::nan_fill(relqmcl, n2);
  };
  if (*icld < 0 || *icld > 3) return fail("MCICA_SUBCOL: INVALID ICLD");
  if (*irng != 0) *irng = 1;  // :303
  std::vector<unsigned> mask;
  if (*irng == 1) {
    cb::mcica::mask_mt_host(cldfrac, nc, nl, 140, 5, *icld, *permuteseed, mask);
  } else {
    mask.assign((size_t)nl * 5 * nc, 0u);
    std::atomic<int> bad{0};
    cb::WorkerPool::get().parallel_for((nc + 63) / 64, [&](int t) {
      for (int c = t * 64; c < nc && c < (t + 1) * 64; ++c)
        if (cb::mcica::mask_column_kiss(play, cldfrac, nc, nl, 140, 5, *icld, *permuteseed, mask.data(), nc, 0, c)) bad.store(1);
    });
    if (bad.load()) return fail("MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS.");
  }
  int ngb[140];
  for (int g = 0; g < 140; ++g) ngb[g] = lw_ngb(g);
  cb::mcica::SubcolIn si{ciwp, clwp, rei, rel, {tauc, nullptr, nullptr, nullptr}, {0., 0., 0., 0.}};
  cb::mcica::SubcolOut so{cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, {taucmcl, nullptr, nullptr, nullptr}};
  cb::mcica::expand(mask.data(), nc, nl, 140, 5, 16, ngb, 1, si, so);
}

// rrtmg_lw_c_binder.f90:94-174 -> rrtmg_lw_rad.f90:80 (rtrnmc).  The per-g-point arrays are folded back into the engine's
// form (one mask bit per sub-column + the layer's water paths / band optics, mcica_compat.h) and run through the same
// kernels as cb200_lw_run_host with McICA enabled.
extern "C" void rrtmg_lw_mcica_wrapper(int* ncol, int* nlay, int* icld, int* idrv, double* play, double* plev,
                                       double* tlay, double* tlev, double* tsfc, double* h2ovmr, double* o3vmr,
                                       double* co2vmr, double* ch4vmr, double* n2ovmr, double* o2vmr, double* cfc11vmr,
                                       double* cfc12vmr, double* cfc22vmr, double* ccl4vmr, double* emis, int* inflglw,
                                       int* iceflglw, int* liqflglw, double* cldfmcl, double* taucmcl, double* ciwpmcl,
                                       double* clwpmcl, double* reicmcl, double* relqmcl, double* tauaer, double* uflx,
                                       double* dflx, double* hr, double* uflxc, double* dflxc, double* hrc,
                                       double* duflx_dt, double* duflxc_dt) {
  const int nc = *ncol, nl = *nlay;
  const bool derivs = *idrv == 1;
  auto fail = [&](const std::string& m) {
    lw_wrapper_fail(m, nc, nl, uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, derivs);
  };
  if (!g_engine) return fail("rrtmg_lw_ini_wrapper has not been called (or failed)");
  if (*icld < 0 || *icld > 3) *icld = 2;  // rrtmg_lw_rad.f90:451
  if (cb200_lw_set_options(g_engine, *icld, *idrv, *inflglw, *iceflglw, *liqflglw)) return fail(cb200_lw_last_error(g_engine));
  int ngb[140];
  for (int g = 0; g < 140; ++g) ngb[g] = lw_ngb(g);
  cb::mcica::CollapseOut co;
  const double* const bandmcl[4] = {taucmcl, nullptr, nullptr, nullptr};
  const std::string why = cb::mcica::collapse(nc, nl, 140, 5, 16, ngb, 1, cldfmcl, ciwpmcl, clwpmcl, bandmcl, co);
  if (!why.empty()) return fail(why);
  cb200_lw_set_mcica(g_engine, 1, 1, 0);
  g_engine->ext_mask.swap(co.mask);
  g_engine->ext_ncol = nc;
  g_engine->ext_nlay = nl;
  cb200_lw_inputs in{play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr,
                     cfc22vmr, ccl4vmr, emis, co.cldfr.data(), co.band[0].data(), co.ciwp.data(), co.clwp.data(), reicmcl,
                     relqmcl, tauaer};
  cb200_lw_outputs out{uflx, dflx, hr, uflxc, dflxc, hrc};
  if (derivs) cb200_lw_set_derivative_outputs(g_engine, duflx_dt, duflxc_dt);
  const int rc = cb200_lw_run_host(g_engine, nc, nl, &in, &out);
  cb200_lw_set_derivative_outputs(g_engine, nullptr, nullptr);
  g_engine->ext_mask.clear();
  cb200_lw_set_mcica(g_engine, 0, 1, 0);
  if (rc) fail(cb200_lw_last_error(g_engine));
}
