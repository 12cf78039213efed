// climt_b200 -- RRTMG shortwave engine: CUDA kernels (sm_100a), launcher and C ABI (include/climt_b200.h).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <future>
#include <string>
#include <vector>

#include "../../include/climt_b200.h"
#include "engine_common.h"
#include "cb_async.cuh"
#include "sw_tables.h"
#include "mcica_host.h"
#include "mcica_compat.h"

using namespace cb::sw;

namespace {
using cb::kBlock;
#ifndef CB_SW_RT_MIN_BLOCKS
#define CB_SW_RT_MIN_BLOCKS 8  // transfer kernel, one g-point per thread: 32 warps per SM (r01 B200 sweep at 2 g-points per thread: 4 -> 2.55 ms,
                               // 5 -> 2.30, 6 -> 2.25; at 1 g-point: 6 -> 1.83, 8 -> 1.66, 10 -> 1.93 ms for the transfer kernel alone)
#endif

struct UnitList {
  Unit u[kMaxUnits];
  int n;
};

// inatm_sw + setcoef_sw (+ ECMWF aerosol mix): one thread per (column, layer)
__global__ void __launch_bounds__(kBlock) k_sw_prep_layer(const __grid_constant__ Tables T, const __grid_constant__ In in, const Flags fl,
                                                          const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (c < n) sw_prep_column<true, false, false>(T, in, fl, W, c0, c, l, l + 1);
}
// ECMWF aerosol mix (iaer = 6 only): one thread per (column, layer)
__global__ void __launch_bounds__(kBlock) k_sw_aer_mix(const __grid_constant__ Tables T, const __grid_constant__ In in,
                                                       const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) sw_aerosol_mix(T, in, W, c0, c, blockIdx.y);
}
// what couples the layers of a column (laytrop, cloud optics, solar-source layers): one thread per column
__global__ void __launch_bounds__(kBlock) k_sw_prep(const __grid_constant__ Tables T, const __grid_constant__ In in, const Flags fl,
                                                    const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) sw_prep_column<false, true>(T, in, fl, W, c0, c, 0, in.nlay);
}

#ifndef CB_SW_TAU_MIN_BLOCKS
#define CB_SW_TAU_MIN_BLOCKS 5  // r02 B200 sweep, 8192 x 60 | McICA 16384 x 72: 3 -> 0.401 | 0.971 ms, 4 -> 0.352 | 0.865, 5 -> 0.335 | 0.808, 6 -> 0.338 | 0.827, 8 -> 0.356 | 0.879
#endif
#ifndef CB_SW_LAYER_CHUNKS
#define CB_SW_LAYER_CHUNKS 4  // taumol: layers are independent -> blockIdx.z cuts them into chunks for more threads in flight
#endif

// taumol_sw: one block = 128 adjacent columns x one unit (<= 4 g-points of one band) x one chunk of layers;
// every branch on the band is block-uniform.
__global__ void __launch_bounds__(kBlock, CB_SW_TAU_MIN_BLOCKS)
    k_sw_taumol(const __grid_constant__ Tables T, const __grid_constant__ Solar sol, const __grid_constant__ In in,
                const __grid_constant__ Work W, const __grid_constant__ UnitList UL, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const Unit un = UL.u[blockIdx.y];
  const int per = (in.nlay + gridDim.z - 1) / gridDim.z;
  const int l0 = blockIdx.z * per, l1 = min(in.nlay, l0 + per);
#define CB_CASE(B)                                                           \
  case B:                                                                    \
    if (un.u == 4) sw_taumol_unit<B, 4>(T, sol, in, W, c0, c, un.g0, l0, l1); \
    else sw_taumol_unit<B, 2>(T, sol, in, W, c0, c, un.g0, l0, l1);           \
    break;
  switch (un.band) {
    CB_CASE(16) CB_CASE(17) CB_CASE(18) CB_CASE(19) CB_CASE(20) CB_CASE(21) CB_CASE(22)
    CB_CASE(23) CB_CASE(24) CB_CASE(25) CB_CASE(26) CB_CASE(27) CB_CASE(28) CB_CASE(29)
  }
#undef CB_CASE
}

constexpr int kPartK = 4;  // interfaces staged in shared memory between two flushes of SwPartSmem

// Device-side sink of the downward sweep's flux contributions (interface: SwPartDirect in sw_core.cuh).  The CB_SW_GROUP warps of
// a block are the units of one group, the same 32 columns in every warp.  A put() parks the thread's four values in shared
// memory; every kPartK interfaces the block meets, each warp sums a share of the (interface, quantity) pairs over the warps IN
// UNIT ORDER (deterministic, = the serial emulation's order) and writes one coalesced 256-byte row per pair.  The warps of a block
// may drift apart by up to kPartK interfaces between two meetings.
struct SwPartSmem {
  double* buf;   // shared: [kPartK][4][nthreads]
  double* part;  // rows of this block's group at this thread's column
  size_t pstride, lev_first;
  int ncc, nthreads, nw, tid, warp, lane, count;
  bool valid, cloudy;
  bool stream = false;  // slab form of the kernel: the rows are written "evict first" so that they do not displace the slabs in the L2
  __device__ void put(size_t lev, bool cloudy_col, double sfu, double sfd, double scu, double scd) {
    if (count == 0) lev_first = lev;  // interfaces arrive top-down: lev, lev - 1, ...
    cloudy = cloudy_col;
    double* b = buf + (size_t)count * 4 * nthreads + tid;
    b[0] = sfu; b[nthreads] = sfd; b[2 * nthreads] = scu; b[3 * nthreads] = scd;
    if (++count == kPartK) flush();
  }
  __device__ void flush() {
    cb::barrier_unaligned(nthreads);
    for (int p = warp; p < count * 4; p += nw) {
      const int k = p >> 2, q = p & 3;
      const double* b = buf + ((size_t)k * 4 + q) * nthreads + lane;
      double s = b[0];
      for (int w = 1; w < nw; ++w) s = s + b[w * 32];
      // (column validity and cloudiness are properties of the lane's column: the same in every warp of the block)
      if (valid && (q >= 2 || cloudy)) {
        double* o = part + q * pstride + (lev_first - k) * ncc;
        if (stream) cb::st_stream(o, s); else *o = s;
      }
    }
    cb::barrier_unaligned(nthreads);
    count = 0;
  }
  __device__ void finish() {
    if (count) flush();
  }
};

// spcvrt_sw / spcvmc_sw: one block = 32 adjacent columns x CB_SW_GROUP units (one warp each, <= CB_SW_UMAX g-points of one band);
// the same code for every band, so the instruction working set of an SM is one function body.
template <bool MC>
__global__ void __launch_bounds__(32 * CB_SW_GROUP, CB_SW_RT_MIN_BLOCKS * 4 / CB_SW_GROUP)
    k_sw_transfer(const __grid_constant__ Tables T, const __grid_constant__ Solar sol, const __grid_constant__ In in, const Flags fl,
                  const __grid_constant__ Work W, const __grid_constant__ UnitList UL, int c0, int n, int skip_clear) {
  __shared__ double s_part[kPartK * 4 * 32 * CB_SW_GROUP];
  const int c = blockIdx.x * 32 + threadIdx.x;
  // the cloud-free 32-column supertiles of a call with clouds belong to k_sw_scan (every warp of the block holds the same 32
  // columns: the vote is block-uniform)
  if (skip_clear && !__any_sync(0xffffffffu, c < n && W.anycld[c] != 0)) return;
  const int group = blockIdx.y;
  const int k = group * CB_SW_GROUP + threadIdx.y;
  SwPartSmem sink;
  sink.buf = s_part;
  sink.pstride = (size_t)(in.nlay + 1) * W.ncc;
  sink.part = W.part + (size_t)group * 4 * sink.pstride + c;
  sink.ncc = W.ncc;
  sink.nthreads = 32 * CB_SW_GROUP; sink.nw = CB_SW_GROUP;
  sink.tid = threadIdx.y * 32 + threadIdx.x; sink.warp = threadIdx.y; sink.lane = threadIdx.x;
  sink.count = 0; sink.lev_first = 0;
  sink.valid = c < n; sink.cloudy = false;
  if (c >= n || k >= UL.n) {  // no work, but the block's meetings need every thread: same number of put() calls
    sink.cloudy = c < n && W.anycld[c] != 0;
    for (int l = in.nlay; l >= 0; --l) sink.put((size_t)l, sink.cloudy, 0., 0., 0., 0.);
    sink.finish();
    return;
  }
  const Unit un = UL.u[k];
  if (CB_SW_UMAX >= 4 && un.u == 4) sw_transfer_unit<4, MC>(T, sol, in, fl, W, c0, c, un.band - 16, un.g0, sink);
  else if (CB_SW_UMAX == 1) sw_transfer_unit<1, MC>(T, sol, in, fl, W, c0, c, un.band - 16, un.g0, sink);
  else sw_transfer_unit<2, MC>(T, sol, in, fl, W, c0, c, un.band - 16, un.g0, sink);
}

// The slab form of the transfer kernel: a persistent grid (blocks per SM chosen by the host) walks the (column tile, group of
// units) work items; each warp keeps the rows it carries from the upward to the downward sweep in a slab of its own that it
// reuses item after item, laid out [layer][row][lane].  The rows are read back in the reverse order of their writing, so with
// the slabs of all resident warps inside the L2 the rows never reach HBM: the in-flight footprint is
// (resident warps) x nlay x (7 | 14 rows) x 256 B instead of the whole chunk's 112 x 14 x nlay x ncc x 8 B.
#ifndef CB_SW_SLAB_MIN_BLOCKS
#define CB_SW_SLAB_MIN_BLOCKS 4
#endif
template <bool MC>
__global__ void __launch_bounds__(32 * CB_SW_GROUP, CB_SW_SLAB_MIN_BLOCKS)
    k_sw_transfer_slab(const __grid_constant__ Tables T, const __grid_constant__ Solar sol, const __grid_constant__ In in, const Flags fl,
                       const __grid_constant__ Work W, const __grid_constant__ UnitList UL, double* __restrict__ slabs, int c0, int n) {
  __shared__ double s_part[kPartK * 4 * 32 * CB_SW_GROUP];
  const int ntiles = (n + 31) / 32, ngroups = (UL.n + CB_SW_GROUP - 1) / CB_SW_GROUP;
  const int nlay = in.nlay;
  Carry cy;
  cy.rs = 32; cy.ls = 14 * 32; cy.us = 0;  // (one g-point per unit in this form)
  cy.p = slabs + ((size_t)blockIdx.x * CB_SW_GROUP + threadIdx.y) * ((size_t)nlay * 14 * 32) + threadIdx.x;
  for (int w = blockIdx.x; w < ntiles * ngroups; w += gridDim.x) {
    const int tile = w % ntiles, group = w / ntiles;
    const int c = tile * 32 + threadIdx.x;
    const int k = group * CB_SW_GROUP + threadIdx.y;
    SwPartSmem sink;
    sink.buf = s_part;
    sink.pstride = (size_t)(nlay + 1) * W.ncc;
    sink.part = W.part + (size_t)group * 4 * sink.pstride + c;
    sink.ncc = W.ncc;
    sink.nthreads = 32 * CB_SW_GROUP; sink.nw = CB_SW_GROUP;
    sink.tid = threadIdx.y * 32 + threadIdx.x; sink.warp = threadIdx.y; sink.lane = threadIdx.x;
    sink.count = 0; sink.lev_first = 0;
    sink.valid = c < n; sink.cloudy = false; sink.stream = true;
    if (c >= n || k >= UL.n) {
      sink.cloudy = c < n && W.anycld[c] != 0;
      for (int l = nlay; l >= 0; --l) sink.put((size_t)l, sink.cloudy, 0., 0., 0., 0.);
      sink.finish();
      continue;
    }
    const Unit un = UL.u[k];
    sw_transfer_unit<1, MC, SwPartSmem, false, true>(T, sol, in, fl, W, c0, c, un.band - 16, un.g0, sink, cy);
  }
}


// ---- column-tile form of the transfer (sw_core.cuh: sw_tile_cell / sw_tile_sweeps; the longwave twin is k_lw_tile) -----------------
// One block = one tile of TW adjacent columns x one band group (sw_tile_group_bands); CB_SW_TILE_THREADS / 32 warps, specialised:
//   warps 1.. (producers)  evaluate the (layer, column) cells of one g-point -- delta scaling and reftra_sw, all the arithmetic of the
//                          path; warp-wide rows of TW columns, the taumol rows of the NEXT g-point loaded before the cells of the
//                          current one -- into one of two row buffers in shared memory;
//   warp 0 (consumer)      runs the upward and the downward adding sweep of its lane's (stream, column) over a filled buffer and adds
//                          the g-point's fluxes to the per-level sums, also in shared memory, while the producers fill the other buffer.
// Cloud-free form: lane = column (TW = 32), clear-sky stream only.  Cloudy form: lanes 0..TW-1 the clear-sky stream of the TW columns,
// lanes TW..2TW-1 their total-sky stream.  A call with clouds launches both forms and every 32-column supertile is taken by exactly
// one of them (SELECT).  Hand-over through named barriers as in k_lw_tile.  Shared memory: (2 NR nlay + 2 nlay NS + NA (nlay + 1)) TW
// doubles with NR / NS / NA = 5 / 1 / 2 (cloud-free) or 10 / 2 / 4: 216 KB at 60 layers for TW = 32 / 16.
#ifndef CB_SW_TILE_THREADS
#define CB_SW_TILE_THREADS 512
#endif
constexpr int kTileThreads = CB_SW_TILE_THREADS;
constexpr int kTileProducers = kTileThreads / 32 - 1;
__device__ __forceinline__ void bar_sync(int id) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(kTileThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(kTileThreads) : "memory"); }

template <int TW, bool CLOUDY>
constexpr size_t sw_tile_smem(int nlay) {
  return sizeof(double) * ((size_t)2 * (CLOUDY ? kSwTileRowsCloudy : kSwTileRowsClear) * nlay * TW + (size_t)2 * nlay * (CLOUDY ? 2 : 1) * TW +
                           (size_t)(CLOUDY ? 4 : 2) * (nlay + 1) * TW);
}

// KC: cells per producer thread, >= ceil(nlay / (kTileProducers * 32 / TW)) (chosen by the launcher)
template <int TW, bool CLOUDY, bool MC, int KC, bool SELECT>
__global__ void __launch_bounds__(kTileThreads, 1)
    k_sw_tile(const __grid_constant__ Tables T, const __grid_constant__ Solar sol, const __grid_constant__ In in, const Flags fl,
              const __grid_constant__ Work W, int c0, int n) {
  extern __shared__ double sm[];
  constexpr int NR = CLOUDY ? kSwTileRowsCloudy : kSwTileRowsClear;
  constexpr int NS = CLOUDY ? 2 : 1;   // streams
  constexpr int NA = 2 * NS;           // rows of sums: cloudy fu, fd, cu, cd; cloud-free cu, cd
  constexpr int FULL0 = 1, EMPTY0 = 3;
  const int nlay = in.nlay;
  const size_t prows = (size_t)nlay * TW, arows = (size_t)(nlay + 1) * TW, rrows = (size_t)nlay * NS * TW;
  double* const Pbuf = sm;                      // [2][NR][nlay][TW]
  double* const Rbuf = sm + 2 * NR * prows;     // [2][nlay][NS * TW]   rup, rupd of every consumer lane
  double* const acc = Rbuf + 2 * rrows;         // [NA][nlay+1][TW]
  const int tile = blockIdx.x, group = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (SELECT) {
    const int c32 = (tile * TW / 32) * 32 + lane;
    const bool anyc = __any_sync(0xffffffffu, c32 < n && W.anycld[c32] != 0);
    if (anyc != CLOUDY) return;
  }
  for (size_t i = threadIdx.x; i < NA * arows; i += kTileThreads) acc[i] = 0.0;
  __syncthreads();
  int ib0, ib1;
  sw_tile_group_bands(group, ib0, ib1);
  const int nit = band_gstart(ib1 - 1) + band_ngpt(ib1 - 1) - band_gstart(ib0);
  int it = 0;
  if (warp == 0) {
    // ---- consumer
    const int col = lane % TW, stream = lane / TW;  // stream 0: clear sky, 1: total sky
    const int c = tile * TW + col;
    const bool live = stream < NS && c < n;
    const size_t gc = (size_t)c0 + (live ? c : 0);
    double prmu0 = live ? in.coszen[gc] : 1.0;
    if (prmu0 < 1.e-10) prmu0 = 1.e-10;
    // rows of the sums: cloudy form [fu, fd, cu, cd], cloud-free form [cu, cd]
    double* const acc_up = acc + (size_t)(CLOUDY ? 2 * (1 - stream) : 0) * arows + col;
    double* const acc_dn = acc_up + arows;
    double* const R = Rbuf + lane;
    for (int ib = ib0; ib < ib1; ++ib) {
      const bool nir = (ib <= 8) || ib == 13;  // rad.nomcica.f90:648-659
      double albdir = 0., albdif = 0.;
      if (live) { albdir = nir ? in.aldir[gc] : in.asdir[gc]; albdif = nir ? in.aldif[gc] : in.asdif[gc]; }
      const int ng = band_ngpt(ib), gs = band_gstart(ib);
      for (int g = 0; g < ng; ++g, ++it) {
        const int b = it & 1;
        const double zinc = live ? sol.adjflux[ib] * W.src[(size_t)(gs + g) * W.ncc + c] * prmu0 : 0.;
        bar_sync(FULL0 + b);
#ifndef CB_TILE_SKIP_CONSUMER  // (timing experiments only)
        if (live)
#else
        if (live && it < 0)
#endif
          sw_tile_sweeps(Pbuf + ((size_t)b * NR + (size_t)stream * 5) * prows + col, prows, TW, nlay, albdir, albdif, zinc, R, rrows,
                         (size_t)NS * TW, acc_up, acc_dn, TW);
        __threadfence_block();
        if (it + 2 < nit) bar_arrive(EMPTY0 + b);
      }
    }
  } else {
    // ---- producers
    constexpr int LPW = 32 / TW;
    const int col = lane % TW, lsub = lane / TW;
    const int c = tile * TW + col;
    const bool live = c < n;
    const int pw = warp - 1;
    const int lstep = kTileProducers * LPW, lfirst = pw * LPW + lsub;
    double prmu0 = live ? in.coszen[(size_t)c0 + c] : 1.0;
    if (prmu0 < 1.e-10) prmu0 = 1.e-10;
    const bool cloudy_col = CLOUDY && live && W.anycld[c] != 0;
    double tg_n[KC], tr_n[KC];
    if (live) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int l = lfirst + k * lstep;
        if (l < nlay) sw_tile_cell_load(in, W, c, l, band_gstart(ib0), tg_n[k], tr_n[k]);
      }
    }
    for (int ib = ib0; ib < ib1; ++ib) {
      const int ng = band_ngpt(ib), gs = band_gstart(ib);
      for (int g = 0; g < ng; ++g, ++it) {
        const int b = it & 1;
        double tg_c[KC], tr_c[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) { tg_c[k] = tg_n[k]; tr_c[k] = tr_n[k]; }
        if (live && it + 1 < nit) {
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int l = lfirst + k * lstep;
            if (l < nlay) sw_tile_cell_load(in, W, c, l, gs + g + 1, tg_n[k], tr_n[k]);
          }
        }
        if (it >= 2) bar_sync(EMPTY0 + b);
        double* __restrict__ Pb = Pbuf + (size_t)b * NR * prows + col;
#ifndef CB_TILE_SKIP_PRODUCER  // (timing experiments only)
        if (live) {
#else
        if (live && it < 0) {
#endif
#pragma unroll 1
          for (int k = 0; k < KC; ++k) {
            const int l = lfirst + k * lstep;
            if (l < nlay)
              sw_tile_cell<MC, CLOUDY>(T, in, fl, W, c0, c, l, ib, gs + g, prmu0, cloudy_col, tg_c[k], tr_c[k], Pb + (size_t)l * TW, prows);
          }
        }
        __threadfence_block();
        bar_arrive(FULL0 + b);
      }
    }
  }
  __syncthreads();
  // the group's sums -> part[group][fu, fd, cu, cd][level][column]; the cloud-free form writes the clear-sky rows only (sw_reduce_level)
  const size_t pstride = (size_t)(nlay + 1) * W.ncc;
  for (size_t i = threadIdx.x; i < NA * arows; i += kTileThreads) {
    const int col = (int)(i % TW);
    const size_t rest = i / TW;
    const int lev = (int)(rest % (nlay + 1)), q = (int)(rest / (nlay + 1));
    const int c = tile * TW + col;
    if (c >= n) continue;
    cb::st_stream(W.part + ((size_t)group * 4 + (CLOUDY ? q : q + 2)) * pstride + (size_t)lev * W.ncc + c, acc[i]);
  }
}

// ---- scan form (sw_core.cuh: scan_up_local / scan_dn_local / scan_walk) ----------------------------------------------------------------
// One block = a tile of 32 adjacent cloud-free columns x one band group, 8 warps, two blocks per SM; per g-point two phases:
//   cells   every thread evaluates its (layer, column) cells (lanes = columns: the taumol rows are read as coalesced 256-byte rows,
//           those of the next g-point loaded ahead) into ONE row buffer in shared memory;
//   sweeps  every warp takes four columns, eight lanes per column: local matrix products over a lane's own layers, a three-step
//           shuffle scan up and one down, then each lane walks its layers and adds the fluxes of its interfaces to REGISTER sums.
// Nothing but the layer rows is parked, no sums in shared memory, no hand-over protocol: two __syncthreads per g-point, and while
// one block of the SM sweeps the other evaluates cells.  Row buffer layout [row][l % KL][l / KL][36]: lanes of a column group read
// different bank groups, a warp of the cells phase writes 32 consecutive doubles.  Shared memory 5 x KL x 8 x 36 doubles: 92 KB for
// up to 63 layers.  Columns with clouds are left to k_sw_transfer (SELECT; sw_reduce picks the right group count per supertile).
constexpr int kScanThreads = 256, kScanLS = 36;
template <int KLMAX>
constexpr size_t sw_scan_smem(int KL) { return sizeof(double) * (size_t)kSwTileRowsClear * KL * kScanLanes * kScanLS; }
__device__ __forceinline__ UpMap shfl_up_map(const UpMap& m, int d) {
  UpMap o;
  o.a11 = __shfl_up_sync(0xffffffffu, m.a11, d, kScanLanes); o.a12 = __shfl_up_sync(0xffffffffu, m.a12, d, kScanLanes);
  o.a21 = __shfl_up_sync(0xffffffffu, m.a21, d, kScanLanes); o.a22 = __shfl_up_sync(0xffffffffu, m.a22, d, kScanLanes);
  o.c1 = __shfl_up_sync(0xffffffffu, m.c1, d, kScanLanes); o.c2 = __shfl_up_sync(0xffffffffu, m.c2, d, kScanLanes);
  o.a33 = __shfl_up_sync(0xffffffffu, m.a33, d, kScanLanes);
  return o;
}
__device__ __forceinline__ DnMap shfl_down_map(const DnMap& m, int d) {
  DnMap o;
  o.a11 = __shfl_down_sync(0xffffffffu, m.a11, d, kScanLanes); o.a12 = __shfl_down_sync(0xffffffffu, m.a12, d, kScanLanes);
  o.a21 = __shfl_down_sync(0xffffffffu, m.a21, d, kScanLanes); o.a22 = __shfl_down_sync(0xffffffffu, m.a22, d, kScanLanes);
  o.beta = __shfl_down_sync(0xffffffffu, m.beta, d, kScanLanes); o.t = __shfl_down_sync(0xffffffffu, m.t, d, kScanLanes);
  o.g1 = __shfl_down_sync(0xffffffffu, m.g1, d, kScanLanes); o.g2 = __shfl_down_sync(0xffffffffu, m.g2, d, kScanLanes);
  return o;
}

template <int KLMAX, bool SELECT>
__global__ void __launch_bounds__(kScanThreads, 2)
    k_sw_scan(const __grid_constant__ Tables T, const __grid_constant__ Solar sol, const __grid_constant__ In in, const Flags fl,
              const __grid_constant__ Work W, int c0, int n) {
  extern __shared__ double P[];
  constexpr int TW = 32, NW = kScanThreads / 32;
  const int nlay = in.nlay;
  const int KL = (nlay + 1 + kScanLanes - 1) / kScanLanes;  // interfaces (and layers) per lane of a column group
  const int tile = blockIdx.x, group = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (SELECT) {
    const int c32 = tile * TW + lane;
    if (__any_sync(0xffffffffu, c32 < n && W.anycld[c32] != 0)) return;  // a supertile with clouds: k_sw_transfer's
  }
  const size_t rstride = (size_t)KL * kScanLanes * kScanLS;
  // cells phase: lane = column, warp w owns layers w, w + NW, ...
  const int ccol = tile * TW + lane;
  const bool clive = ccol < n;
  double cmu0 = clive ? in.coszen[(size_t)c0 + ccol] : 1.0;
  if (cmu0 < 1.e-10) cmu0 = 1.e-10;
  // sweeps phase: lanes 8 gi .. 8 gi + 7 share column 4 warp + gi
  const int gi = lane / kScanLanes, li = lane % kScanLanes;
  const int scol_t = NW == 8 ? 4 * warp + gi : 0;  // column of the tile
  const int scol = tile * TW + scol_t;
  const bool slive = scol < n;
  const size_t sgc = (size_t)c0 + (slive ? scol : 0);
  double smu0 = slive ? in.coszen[sgc] : 1.0;
  if (smu0 < 1.e-10) smu0 = 1.e-10;
  double acc_up[KLMAX], acc_dn[KLMAX];
#pragma unroll
  for (int k = 0; k < KLMAX; ++k) { acc_up[k] = 0.; acc_dn[k] = 0.; }
  // property r of this lane's k-th layer (layer li KL + k) of its column
  const double* const Pl = P + (size_t)li * kScanLS + scol_t;
  auto row = [&](int r, int k) { return Pl[(size_t)r * rstride + (size_t)k * (kScanLanes * kScanLS)]; };
  int ib0, ib1;
  sw_tile_group_bands(group, ib0, ib1);
  const int nit = band_gstart(ib1 - 1) + band_ngpt(ib1 - 1) - band_gstart(ib0);
  // the taumol rows of the next g-point are pulled into the L2 while the cells of the current one are evaluated (no registers held)
  const size_t wstride = (size_t)nlay * W.ncc;
  int it = 0;
  for (int ib = ib0; ib < ib1; ++ib) {
    const bool nir = (ib <= 8) || ib == 13;  // rad.nomcica.f90:648-659
    double albdir = 0., albdif = 0.;
    if (slive) { albdir = nir ? in.aldir[sgc] : in.asdir[sgc]; albdif = nir ? in.aldif[sgc] : in.asdif[sgc]; }
    const int ng = band_ngpt(ib), gs = band_gstart(ib);
    for (int g = 0; g < ng; ++g, ++it) {
      const double zinc = slive ? sol.adjflux[ib] * W.src[(size_t)(gs + g) * W.ncc + scol] * smu0 : 0.;
      if (clive) {
        int kk = warp % KL, ii = warp / KL;  // slot of layer l = (l % KL, l / KL), advanced without dividing
#pragma unroll 1
        for (int l = warp; l < nlay; l += NW) {
          const double* __restrict__ scr = W.scr + (((size_t)(gs + g) * NSCR) * nlay + l) * W.ncc + ccol;
          if (it + 1 < nit) {
            cb::prefetch_l2(scr + (size_t)NSCR * wstride + R_TAUG * wstride);
            cb::prefetch_l2(scr + (size_t)NSCR * wstride + R_TAUR * wstride);
          }
          const double taug = cb::ld_stream(scr + R_TAUG * wstride), taur = cb::ld_stream(scr + R_TAUR * wstride);
          sw_tile_cell<false, false>(T, in, fl, W, c0, ccol, l, ib, gs + g, cmu0, false, taug, taur,
                                     P + ((size_t)kk * kScanLanes + ii) * kScanLS + lane, rstride);
          kk += NW;
          while (kk >= KL) { kk -= KL; ++ii; }
        }
      }
      __syncthreads();
      {
        // every lane takes part in the shuffles (a ragged tile's dead column groups sweep whatever the buffer holds; nothing of
        // theirs is written)
        UpMap um;
        DnMap dm;
        scan_local<KLMAX>(row, nlay, li, KL, um, dm);
#pragma unroll
        for (int d = 1; d < kScanLanes; d <<= 1) {
          const UpMap uo = shfl_up_map(um, d);
          const DnMap dn_o = shfl_down_map(dm, d);
          if (li >= d) um = up_compose(um, uo);
          if (li + d < kScanLanes) dm = dn_compose(dm, dn_o);
        }
        UpMap below = shfl_up_map(um, 1);
        DnMap above = shfl_down_map(dm, 1);
        if (li == 0) below = up_identity();
        if (li == kScanLanes - 1) above = dn_identity();
        scan_walk<KLMAX>(row, nlay, li, KL, below, above, albdir, albdif, zinc, acc_up, acc_dn);
      }
      __syncthreads();
    }
  }
  // the group's sums of this lane's interfaces -> part[group][cu, cd][level][column] (cloud-free columns: clear-sky rows only)
  if (slive) {
    const size_t pstride = (size_t)(nlay + 1) * W.ncc;
#pragma unroll
    for (int k = 0; k < KLMAX; ++k) {
      const int lev = li * KL + k;
      if (k < KL && lev <= nlay) {
        cb::st_stream(W.part + ((size_t)group * 4 + 2) * pstride + (size_t)lev * W.ncc + scol, acc_up[k]);
        cb::st_stream(W.part + ((size_t)group * 4 + 3) * pstride + (size_t)lev * W.ncc + scol, acc_dn[k]);
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_sw_mask_kiss(const __grid_constant__ In in, const __grid_constant__ Work W,
                                                         int icld, int seed, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (cb::mcica::mask_column_kiss(in.play, in.cldfr, in.ncol, in.nlay, 112, 4, icld, seed, W.mask, W.ncc, c0, c)) *W.err = 9;
}

// ngroups_clear > 0: the cloud-free 32-column supertiles hold that many groups (k_sw_scan), the others `ngroups` (a warp of this
// kernel is one supertile)
__global__ void __launch_bounds__(kBlock) k_sw_reduce(const __grid_constant__ Work W, int ngroups, int ngroups_clear, const Out out,
                                                      int nlay, int ncol, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int lev = blockIdx.y;
  const bool anyc = __any_sync(0xffffffffu, c < n && W.anycld[c] != 0);
  if (c < n) sw_reduce_level(W, ngroups_clear > 0 && !anyc ? ngroups_clear : ngroups, nlay, c0, c, lev, ncol, out);
}

__global__ void __launch_bounds__(kBlock) k_sw_heat(const __grid_constant__ Tables T, const __grid_constant__ In in,
                                                    const Out out, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (c < n) sw_heating(T, in, out, c0 + c, l);
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

struct cb200_sw_engine {
  int device = 0;
  Tables T;
  double* d_tables = nullptr;
  Flags fl{1, 0, 2, 1, 1, 0};
  int irng = 1, permuteseed = 0;
  unsigned* d_mask_full = nullptr;
  size_t mask_full_cap = 0;
  std::vector<unsigned> ext_mask;  // caller-supplied sub-column mask [nlay][4][ncol] (cb200_sw_set_subcolumn_mask), else empty
  int ext_ncol = 0, ext_nlay = 0;
  SolarOptions solar;
  UnitList UL;      // transfer kernel units (<= CB_SW_UMAX g-points); `part` holds one flux set per unit
  UnitList UL_tau;  // taumol kernel units (<= CB_SW_TAU_UMAX g-points)
  int cap_ncc = 0, cap_nlay = 0;
  Work W{};
  int max_chunk = 8192;
  bool scan = false;         // scan form of the transfer (k_sw_scan) for cloud-free supertiles (CLIMT_B200_SW_SCAN=1).  Off by default:
                             // r02 B200, 8192 x 60 clear sky, 0.91 GB of DRAM traffic instead of 7.2 GB but 2.06 ms against 1.83 ms
                             // (16 warps per SM at 128 registers, 8 % more instructions than the unit form; profiles/r02_tile_kernels.md)
  bool tile = false;         // column-tile form of the transfer kernel (CLIMT_B200_SW_TILE=1).  Off by default: r02 B200, 8192 x 60
                             // clear sky, it moves 0.89 GB instead of 7.2 GB but takes 2.79 ms against 1.83 ms -- the one consumer warp
                             // of a tile issues ~60 fp64 instructions per level on one scheduler (profiles/r02_tile_kernels.md)
  int tile_min_tw = 16;      // narrowest cloudy-form tile worth using (CLIMT_B200_SW_TILE_MIN_TW; 8 lets 72-layer cloudy calls through)
  size_t smem_optin = 0;     // opt-in shared memory per block of the device
  int slab_bps = 0;          // > 0: the slab form of the transfer kernel with this many 4-warp blocks per SM (CLIMT_B200_SW_SLAB)
  double* d_slabs = nullptr;
  size_t slabs_cap = 0;
  int n_sm = 148;
  cb::HostPipe pipe;
  size_t h2d_bytes = 0, d2h_bytes = 0;  // moved by the last host-pointer call
  bool skip_zero_inputs = true;         // CLIMT_B200_SKIP_ZERO_INPUTS=0 turns the all-zero scan of the host call off
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
    cudaFree(W.ws); cudaFree(W.idx); cudaFree(W.laytrop); cudaFree(W.laysolfr); cudaFree(W.anycld);
    cudaFree(W.cld); cudaFree(W.aer); cudaFree(W.scr); cudaFree(W.src); cudaFree(W.part); cudaFree(W.err); cudaFree(W.mask);
    W = Work{};
    cap_ncc = cap_nlay = 0;
  }
  int ensure_work(int ncc, int nlay) {
    cb200_sw_engine* e = this;
    if (ncc <= cap_ncc && nlay <= cap_nlay && W.ws) return 0;
    free_work();
    const size_t n = (size_t)ncc, L = (size_t)nlay;
    CUDA_OK(cudaMalloc(&W.ws, sizeof(double) * NF * L * n));
    CUDA_OK(cudaMalloc(&W.idx, sizeof(int) * L * n));
    CUDA_OK(cudaMalloc(&W.laytrop, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&W.laysolfr, sizeof(int) * 14 * n));
    CUDA_OK(cudaMalloc(&W.anycld, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&W.cld, sizeof(double) * 42 * L * n));
    CUDA_OK(cudaMalloc(&W.aer, sizeof(double) * 42 * L * n));
    CUDA_OK(cudaMalloc(&W.scr, sizeof(double) * 112 * NSCR * L * n));
    CUDA_OK(cudaMalloc(&W.src, sizeof(double) * 112 * n));
    CUDA_OK(cudaMalloc(&W.part, sizeof(double) * ((UL.n + CB_SW_GROUP - 1) / CB_SW_GROUP) * 4 * (L + 1) * n));
    CUDA_OK(cudaMalloc(&W.mask, sizeof(unsigned) * 4 * L * n));
    CUDA_OK(cudaMalloc(&W.err, sizeof(int)));
    CUDA_OK(cudaMemset(W.err, 0, sizeof(int)));
    cap_ncc = ncc;
    cap_nlay = nlay;
    return 0;
  }
};

extern "C" int cb200_sw_create(cb200_sw_engine** out, const char* table_blob, const double constants[11], int device) {
  *out = nullptr;
  auto* e = new cb200_sw_engine();
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
    e->UL.n = build_units(e->UL.u, CB_SW_UMAX);
    e->UL_tau.n = build_units(e->UL_tau.u, CB_SW_TAU_UMAX);
    if (const char* mc = std::getenv("CLIMT_B200_MAX_CHUNK")) e->max_chunk = std::max(128, std::atoi(mc));
    if (const char* z = std::getenv("CLIMT_B200_SKIP_ZERO_INPUTS")) e->skip_zero_inputs = std::atoi(z) != 0;
    if (const char* sb = std::getenv("CLIMT_B200_SW_SLAB")) e->slab_bps = CB_SW_UMAX == 1 ? std::max(0, std::atoi(sb)) : 0;
    cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (const char* tl = std::getenv("CLIMT_B200_SW_TILE")) e->tile = std::atoi(tl) != 0;
    if (const char* tl = std::getenv("CLIMT_B200_SW_SCAN")) e->scan = std::atoi(tl) != 0;
    if (const char* tl = std::getenv("CLIMT_B200_SW_TILE_MIN_TW")) e->tile_min_tw = std::atoi(tl);
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

extern "C" void cb200_sw_destroy(cb200_sw_engine* e) {
  if (!e) return;
  if (e->enqueue.valid()) e->enqueue.wait();  // a host call still being enqueued
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  e->free_work();
  cudaFree(e->d_tables);
  e->pipe.destroy();
  cudaFree(e->d_mask_full);
  cudaFree(e->d_slabs);
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->evm) cudaEventDestroy(e->evm);
  delete e;
}

extern "C" int cb200_sw_set_options(cb200_sw_engine* e, int icld, int iaer, int inflag, int iceflag, int liqflag) {
  if (icld < 0 || icld > 3) icld = 2;                       // rrtmg_sw_rad.nomcica.f90:558
  if (iaer != 0 && iaer != 6 && iaer != 10) iaer = 0;       // :566
  e->fl = Flags{icld, iaer, inflag, iceflag, liqflag, e->fl.mcica};
  return 0;
}
extern "C" int cb200_sw_set_mcica(cb200_sw_engine* e, int enabled, int irng, int permuteseed) {
  e->fl.mcica = enabled ? 1 : 0;
  e->irng = irng != 0 ? 1 : 0;
  e->permuteseed = permuteseed;
  return 0;
}
extern "C" int cb200_sw_set_solar(cb200_sw_engine* e, int isolvar, double scon, const double indsolvar[2],
                                  const double bndsolvar[14]) {
  e->solar.isolvar = isolvar;
  e->solar.scon = scon;
  if (indsolvar) { e->solar.indsolvar[0] = indsolvar[0]; e->solar.indsolvar[1] = indsolvar[1]; }
  if (bndsolvar) for (int i = 0; i < 14; ++i) e->solar.bndsolvar[i] = bndsolvar[i];
  return 0;
}
extern "C" const char* cb200_sw_last_error(cb200_sw_engine* e) { return e ? e->error.c_str() : cb::g_error.c_str(); }
extern "C" int cb200_sw_last_launches(cb200_sw_engine* e) { return e->launches; }
extern "C" int cb200_sw_enable_timing(cb200_sw_engine* e, int on) { e->timing = on != 0; return 0; }
extern "C" double cb200_sw_last_unit_kernel_ms(cb200_sw_engine* e) { return e->unit_ms; }
extern "C" double cb200_sw_last_taumol_kernel_ms(cb200_sw_engine* e) { return e->taumol_ms; }

static In make_in(int ncol, int nlay, const cb200_sw_inputs* p) {
  In in;
  in.ncol = ncol; in.nlay = nlay;
  const double** d = &in.play;
  const double* const* s = reinterpret_cast<const double* const*>(p);
  for (int i = 0; i < 29; ++i) d[i] = s[i];  // identical field order (checked by the static_assert below)
  return in;
}
static_assert(sizeof(cb200_sw_inputs) == 29 * sizeof(double*), "cb200_sw_inputs layout");

// One chunk of columns [c0, c0+n) of `in` through the kernels on stream `st` (the engine's single workspace).
static int launch_chunk(cb200_sw_engine* e, const Solar& sol, const In& in, const Out& out, Work& W, int c0, int n,
                        int out_ncol, bool mc, cudaStream_t st) {
  const int nlay = in.nlay;
  const int gx = (n + kBlock - 1) / kBlock;
  if (mc && e->irng == 0) { k_sw_mask_kiss<<<(n + 31) / 32, 32, 0, st>>>(in, W, e->fl.icld, e->permuteseed, c0, n); e->launches += 1; }
  k_sw_prep_layer<<<dim3(gx, nlay), kBlock, 0, st>>>(e->T, in, e->fl, W, c0, n);
  if (e->fl.iaer == 6) { k_sw_aer_mix<<<dim3(gx, nlay), kBlock, 0, st>>>(e->T, in, W, c0, n); e->launches += 1; }
  // column-serial kernels: one warp per block so that even 8 192 columns spread over every SM
  const int gw = (n + 31) / 32;
  k_sw_prep<<<gw, 32, 0, st>>>(e->T, in, e->fl, W, c0, n);
  if (e->timing) cudaEventRecord(e->ev0, st);
  k_sw_taumol<<<dim3(gx, e->UL_tau.n, CB_SW_LAYER_CHUNKS), kBlock, 0, st>>>(e->T, sol, in, W, e->UL_tau, c0, n);
  if (e->timing) cudaEventRecord(e->evm, st);
  const dim3 gt((n + 31) / 32, (e->UL.n + CB_SW_GROUP - 1) / CB_SW_GROUP), bt(32, CB_SW_GROUP);
  int ngroups = (int)gt.y;
  int ngroups_clear = 0;  // > 0: the cloud-free 32-column supertiles were summed into this many groups (k_sw_scan)
  // the column-tile form (layer properties and adding sweeps in shared memory): the widest tiles that fit, the cloudy form half as wide
  const bool cloudy_call = e->fl.icld >= 1;
  const int tw_clear = sw_tile_smem<32, false>(nlay) <= e->smem_optin ? 32 : (sw_tile_smem<16, false>(nlay) <= e->smem_optin ? 16 : 0);
  const int tw_cloudy = tw_clear / 2;
  if (e->tile && tw_clear >= (cloudy_call ? e->tile_min_tw * 2 : 16) && (nlay + kTileProducers - 1) / kTileProducers <= 8) {
#define CB_TILE(TW, CL, MCF, KC, SEL)                                                                                                \
    do {                                                                                                                               \
      const size_t smem = sw_tile_smem<TW, CL>(nlay);                                                                                  \
      CUDA_OK(cudaFuncSetAttribute(k_sw_tile<TW, CL, MCF, KC, SEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
      k_sw_tile<TW, CL, MCF, KC, SEL><<<dim3((n + TW - 1) / TW, kTileGroups), kTileThreads, smem, st>>>(e->T, sol, in, e->fl, W, c0, n); \
    } while (0)
#define CB_TILE_KC(TW, CL, MCF, SEL)                                                                   \
    do {                                                                                                \
      const int kc = (nlay + kTileProducers * (32 / TW) - 1) / (kTileProducers * (32 / TW));            \
      if (kc <= 3) CB_TILE(TW, CL, MCF, 3, SEL); else if (kc <= 5) CB_TILE(TW, CL, MCF, 5, SEL);        \
      else CB_TILE(TW, CL, MCF, 8, SEL);                                                                \
    } while (0)
    if (!cloudy_call) {
      if (tw_clear == 32) CB_TILE_KC(32, false, false, false); else CB_TILE_KC(16, false, false, false);
    } else {
      if (tw_clear == 32) CB_TILE_KC(32, false, false, true); else CB_TILE_KC(16, false, false, true);
      if (tw_cloudy == 16) { if (mc) CB_TILE_KC(16, true, true, true); else CB_TILE_KC(16, true, false, true); }
      else { if (mc) CB_TILE_KC(8, true, true, true); else CB_TILE_KC(8, true, false, true); }
      e->launches += 1;
    }
#undef CB_TILE_KC
#undef CB_TILE
    ngroups = kTileGroups;
  } else if (e->slab_bps > 0) {
    const int nblocks = (int)std::min<size_t>((size_t)gt.x * gt.y, (size_t)e->n_sm * e->slab_bps);
    const size_t need = (size_t)nblocks * CB_SW_GROUP * nlay * 14 * 32;
    if (need > e->slabs_cap) {
      cudaStreamSynchronize(st);
      cudaFree(e->d_slabs);
      e->d_slabs = nullptr;
      e->slabs_cap = 0;
      CUDA_OK(cudaMalloc(&e->d_slabs, need * sizeof(double)));
      e->slabs_cap = need;
    }
    if (mc) k_sw_transfer_slab<true><<<nblocks, bt, 0, st>>>(e->T, sol, in, e->fl, W, e->UL, e->d_slabs, c0, n);
    else k_sw_transfer_slab<false><<<nblocks, bt, 0, st>>>(e->T, sol, in, e->fl, W, e->UL, e->d_slabs, c0, n);
  } else {
    // scan form for the cloud-free supertiles (all of them when the call has no clouds), unit form for the rest
    const int KL = (nlay + 1 + kScanLanes - 1) / kScanLanes;
    const bool scan = e->scan && KL <= 16 && nlay <= 8 * 16;
    if (scan) {
      const dim3 gs((n + 31) / 32, kTileGroups);
#define CB_SCAN(KLMAX, SEL)                                                                                                 \
      do {                                                                                                                    \
        const size_t smem = sw_scan_smem<KLMAX>(KL);                                                                          \
        CUDA_OK(cudaFuncSetAttribute(k_sw_scan<KLMAX, SEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        k_sw_scan<KLMAX, SEL><<<gs, kScanThreads, smem, st>>>(e->T, sol, in, e->fl, W, c0, n);                                \
      } while (0)
#define CB_SCAN_KL(SEL)                                                                             \
      do {                                                                                           \
        if (KL <= 8) CB_SCAN(8, SEL); else if (KL <= 10) CB_SCAN(10, SEL); else CB_SCAN(16, SEL);    \
      } while (0)
      if (cloudy_call) CB_SCAN_KL(true); else CB_SCAN_KL(false);
#undef CB_SCAN_KL
#undef CB_SCAN
      ngroups_clear = kTileGroups;
      e->launches += cloudy_call ? 1 : 0;
    }
    if (!scan || cloudy_call) {
      if (mc) k_sw_transfer<true><<<gt, bt, 0, st>>>(e->T, sol, in, e->fl, W, e->UL, c0, n, scan ? 1 : 0);
      else k_sw_transfer<false><<<gt, bt, 0, st>>>(e->T, sol, in, e->fl, W, e->UL, c0, n, scan ? 1 : 0);
    }
  }
  if (e->timing) cudaEventRecord(e->ev1, st);
  k_sw_reduce<<<dim3(gx, nlay + 1), kBlock, 0, st>>>(W, ngroups, ngroups_clear, out, nlay, out_ncol, c0, n);
  k_sw_heat<<<dim3(gx, nlay), kBlock, 0, st>>>(e->T, in, out, c0, n);
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

// Mersenne-twister McICA mask: serial stream, generated on the host for bit parity, uploaded once [lay][word][ncol]
static int upload_mt_mask(cb200_sw_engine* e, const double* h_cldfr, int ncol, int nlay, cudaStream_t st) {
  std::vector<unsigned> h_mask;
  if (!e->ext_mask.empty()) {
    if (e->ext_ncol != ncol || e->ext_nlay != nlay) { e->error = "sub-column mask was set for a different ncol/nlay"; return -3; }
    h_mask = e->ext_mask;
  } else {
    cb::mcica::mask_mt_host(h_cldfr, ncol, nlay, 112, 4, e->fl.icld, e->permuteseed, h_mask);
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

extern "C" int cb200_sw_run_device(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                                   const cb200_sw_inputs* pin, const cb200_sw_outputs* pout, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrsw.f90:27)"; return -3; }
  CUDA_OK(cudaSetDevice(e->device));
  int chunk = ncol < e->max_chunk ? ncol : e->max_chunk;
  chunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(chunk, nlay)) return -1;
  Work W = e->W;
  W.ncc = chunk;
  const In in = make_in(ncol, nlay, pin);
  Out out{pout->uflx, pout->dflx, pout->hr, pout->uflxc, pout->dflxc, pout->hrc};
  const Solar sol = compute_solar(e->solar, adjes, dyofyr, solcycfrac);
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
    if (launch_chunk(e, sol, in, out, W, c0, n, ncol, mc, st)) return -1;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cb200_sw_check(cb200_sw_engine* e) {
  if (!e->W.err) return 0;
  CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaMemcpy(e->h_err, e->W.err, sizeof(int), cudaMemcpyDeviceToHost));
  const int code = *e->h_err;
  if (code) {
    // message text of the Fortran `stop` statements (rrtmg_sw_cldprop.f90:166-294, rrtmg_sw_rad.nomcica.f90:618)
    switch (code) {
      case 2: e->error = "ICE RADIUS OUT OF BOUNDS"; break;
      case 3: e->error = "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; break;
      case 4: e->error = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; break;
      case 5: e->error = "ICE OPTICAL PROPERTY OUT OF RANGE"; break;
      case 6: e->error = "FDELTA OUT OF RANGE"; break;
      case 7: e->error = "LIQUID OPTICAL PROPERTY OUT OF RANGE"; break;
      case 8: e->error = "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA"; break;
      case 9: e->error = "MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS."; break;
      case 10: e->error = "PARTIAL CLOUD NOT ALLOWED"; break;
      default: e->error = "invalid input"; break;
    }
    cudaMemset(e->W.err, 0, sizeof(int));
  }
  return code;
}

// Host-pointer call: column chunks through the three-stream pipeline of cb::HostPipe.  Arrays the option flags make
// dead are not transferred (cloud inputs when icld = 0; direct cloud optics unless inflag = 0; aerosol arrays by iaer).
static int sw_host_enqueue(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                           const cb200_sw_inputs* hin, const cb200_sw_outputs* hout) {
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrsw.f90:27)"; return -3; }
  CUDA_OK(cudaSetDevice(e->device));
  cb::HostPipe& P = e->pipe;
  CUDA_OK(P.init());
  const int L = nlay;
  const int irows[29] = {L, L + 1, L, L + 1, 1, L, L, L, L, L, L, 1, 1, 1, 1, 1, L, L, L, L, L,
                         L, L, L, L, 14 * L, 14 * L, 14 * L, 6 * L};
  int inner[29];
  for (int i = 0; i < 29; ++i) inner[i] = 1;
  for (int i = 17; i <= 20; ++i) inner[i] = 14;  // taucld/ssacld/asmcld/fsfcld(nbndsw, ncol, nlay): band-fastest
  const int orows[6] = {L + 1, L + 1, L, L + 1, L + 1, L};
  bool used[29];
  for (int i = 0; i < 29; ++i) used[i] = true;
  const bool clouds = e->fl.icld >= 1;
  const bool mc = e->fl.mcica && clouds;
  // 16 cldfr | 17-20 taucld ssacld asmcld fsfcld | 21-24 cicewp cliqwp reice reliq | 25-27 tau/ssa/asm aer | 28 ecaer
  if (!clouds) for (int i = 16; i <= 24; ++i) used[i] = false;
  if (e->fl.inflag != 0) for (int i = 17; i <= 20; ++i) used[i] = false;
  if (e->fl.iaer != 10) for (int i = 25; i <= 27; ++i) used[i] = false;
  if (e->fl.iaer != 6) used[28] = false;
  // a cloud fraction that is zero in a chunk's columns is set in HBM by a memset instead of crossing PCIe, and so are the cloud
  // arrays that are dead without cloud (same scheme as the longwave call, lw_engine.cu)
  bool zero[29];
  for (int i = 0; i < 29; ++i) zero[i] = false;
  size_t irow_tot = 0, orow_tot = 0;
  for (int i = 0; i < 29; ++i) if (used[i]) irow_tot += (size_t)irows[i] * inner[i];
  for (int i = 0; i < 6; ++i) orow_tot += (size_t)orows[i];
  e->h2d_bytes = 0;
  e->d2h_bytes = orow_tot * (size_t)ncol * sizeof(double);
  int chunk = ncol < P.chunk ? ncol : P.chunk;
  const int wchunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(wchunk, nlay)) return -1;
  CUDA_OK(P.ensure(irow_tot * (size_t)chunk, orow_tot * (size_t)chunk));
  const double* const* hp = reinterpret_cast<const double* const*>(hin);
  const bool mar_q = (e->host_marshal & 1) != 0, mar_t = (e->host_marshal & 2) != 0;  // the components' marshal arithmetic on the device
  double* const* hop = reinterpret_cast<double* const*>(hout);
  const Solar sol = compute_solar(e->solar, adjes, dyofyr, solcycfrac);
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
  bool pg_in[29], pg_out[8], any_pg_in = false, any_pg_out = false;
  for (int i = 0; i < 29; ++i) { pg_in[i] = used[i] && hp[i] && !cb::HostPipe::dma_able(hp[i]); any_pg_in |= pg_in[i]; }
  for (int i = 0; i < 6; ++i) { pg_out[i] = !cb::HostPipe::dma_able(hop[i]); any_pg_out |= pg_out[i]; }
  if (any_pg_in || any_pg_out) CUDA_OK(P.ensure_staging(any_pg_in ? irow_tot * (size_t)chunk : 0, any_pg_out ? orow_tot * (size_t)chunk : 0));
  struct { int s, c0, n; bool valid; } prev{0, 0, 0, false};
  auto drain_outputs = [&](int ps, int pc0, int pn) -> cudaError_t {  // staged outputs of a finished chunk -> the caller's arrays
    cudaError_t ce = cudaEventSynchronize(P.out_done[ps]);
    if (ce != cudaSuccess) return ce;
    size_t o = 0;
    for (int i = 0; i < 6; ++i) {
      if (pg_out[i]) cb::HostPipe::scatter_staged_finish(hop[i], P.h_out[ps] + o, orows[i], ncol, pc0, pn);
      o += (size_t)orows[i] * pn;
    }
    return cudaSuccess;
  };
  int k = 0;
  for (int c0 = 0, n = 0; c0 < ncol; c0 += n, ++k) {
    n = P.chunk_size(k, ncol - c0);
    const int s = k & 1;
    if (e->skip_zero_inputs && clouds) {
      const cb::ZeroView zv{reinterpret_cast<const double* const*>(hin)[16], (size_t)L, (size_t)ncol, (size_t)c0, (size_t)n};
      bool zz;
      cb::all_zero_parallel(&zv, 1, &zz);
      for (int i = 16; i <= 24; ++i) zero[i] = zz;
    }
    CUDA_OK(cudaStreamWaitEvent(P.s_in, P.cmp_done[s], 0));
    if (any_pg_in) CUDA_OK(cudaEventSynchronize(P.in_done[s]));  // the copies that last read this staging slot have left the host
    P.mark(P.s_in, k, 0);
    cb200_sw_inputs din;
    const double** dp = reinterpret_cast<const double**>(&din);
    size_t off = 0;
    for (int i = 0; i < 29; ++i) {
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
    cb200_sw_outputs dout;
    double** dop = reinterpret_cast<double**>(&dout);
    off = 0;
    for (int i = 0; i < 6; ++i) { dop[i] = P.d_out[s] + off; off += (size_t)orows[i] * n; }
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.in_done[s], 0));
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.out_done[s], 0));
    P.mark(P.s_cmp, k, 2);
    if (mar_q || mar_t)  // the components' marshal arithmetic (util.py:47-142) on this chunk, before anything reads it
      CUDA_OK(cb::marshal_launch(n, nlay, din.h2ovmr, din.tlay, din.tsfc, din.play, din.plev, mar_q ? const_cast<double*>(din.h2ovmr) : nullptr,
                                 mar_t ? const_cast<double*>(din.tlev) : nullptr, P.s_cmp));
    const In in = make_in(n, nlay, &din);
    Out out{dout.uflx, dout.dflx, dout.hr, dout.uflxc, dout.dflxc, dout.hrc};
    if (mc && e->irng == 1) W.moff = c0;
    if (launch_chunk(e, sol, in, out, W, 0, n, n, mc, P.s_cmp)) return -1;
    CUDA_OK(cudaEventRecord(P.cmp_done[s], P.s_cmp));
    P.mark(P.s_cmp, k, 3);
    CUDA_OK(cudaStreamWaitEvent(P.s_out, P.cmp_done[s], 0));
    {
      size_t o = 0;
      for (int i = 0; i < 6; ++i) {
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

// The asynchronous form returns at once: the chunk loop runs on a host thread of its own (see cb200_lw_run_host_async).
// The host-pointer calls can do the components' marshal arithmetic themselves, on the device, chunk by chunk: flags bit 0 -- the
// h2ovmr argument holds SPECIFIC HUMIDITY (kg/kg) and is converted to a volume mixing ratio (climt/_core/util.py:47-86); bit 1 --
// tlev is ignored (may be NULL) and computed from tlay, tsfc, play, plev by ln-p interpolation (util.py:89-142).  0 = the plain
// reference ABI.  (The components spend more host time in those two numpy expressions than the whole call takes on the GPU.)
extern "C" int cb200_sw_set_host_marshal(cb200_sw_engine* e, int flags) {
  e->host_marshal = flags & 3;
  return 0;
}

extern "C" int cb200_sw_run_host_async(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                                       const cb200_sw_inputs* hin, const cb200_sw_outputs* hout) {
  if (e->host_pending) { e->error = "a previous run_host_async call has not been waited for"; return -3; }
  if (ncol <= 0 || nlay <= 0 || nlay > 203) { e->error = "bad ncol/nlay (1 <= nlay <= 203, parrrsw.f90:27)"; return -3; }
  const cb200_sw_inputs in = *hin;
  const cb200_sw_outputs out = *hout;
  e->host_pending = true;
  e->enqueue = std::async(std::launch::async, [e, ncol, nlay, adjes, dyofyr, solcycfrac, in, out] {
    return sw_host_enqueue(e, ncol, nlay, adjes, dyofyr, solcycfrac, &in, &out);
  });
  return 0;
}

// Completes the call started by cb200_sw_run_host_async: outputs are in the caller's buffers on return.
extern "C" int cb200_sw_wait(cb200_sw_engine* e) {
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
  e->pipe.trace_dump("SW");
  return cb200_sw_check(e);
}

extern "C" int cb200_sw_run_host(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                                 const cb200_sw_inputs* hin, const cb200_sw_outputs* hout) {
  if (e->host_pending) { e->error = "a previous run_host_async call has not been waited for"; return -3; }
  if (int rc = sw_host_enqueue(e, ncol, nlay, adjes, dyofyr, solcycfrac, hin, hout)) {
    if (e->pipe.s_in) { cudaStreamSynchronize(e->pipe.s_in); cudaStreamSynchronize(e->pipe.s_cmp); cudaStreamSynchronize(e->pipe.s_out); }
    return rc;
  }
  e->host_pending = true;
  return cb200_sw_wait(e);
}

extern "C" void cb200_sw_last_transfer_bytes(cb200_sw_engine* e, double* h2d, double* d2h) {
  *h2d = (double)e->h2d_bytes;
  *d2h = (double)e->d2h_bytes;
}

// ---- reference-named entry points, one process-global engine
namespace {
double g_consts[11] = {0};
cb200_sw_engine* g_engine = nullptr;
std::string default_blob() { return cb::find_table_blob("CLIMT_B200_SW_TABLES", "rrtmg_sw_reduced.blob", (void*)&cb200_sw_create); }
}  // namespace

extern "C" void rrtmg_sw_set_constants(double* pi, double* grav, double* planck, double* boltz, double* clight,
                                       double* avogad, double* alosmt, double* gascon, double* sbcnst, double* secdy) {
  double v[10] = {*pi, *grav, *planck, *boltz, *clight, *avogad, *alosmt, *gascon, *sbcnst, *secdy};
  std::memcpy(g_consts, v, sizeof v);
}
extern "C" void rrtmg_sw_ini_wrapper(double* cpdair) {
  g_consts[10] = *cpdair;
  if (g_engine) { cb200_sw_destroy(g_engine); g_engine = nullptr; }
  int dev = 0;
  if (const char* d = std::getenv("CLIMT_B200_DEVICE")) dev = std::atoi(d);
  if (cb200_sw_create(&g_engine, default_blob().c_str(), g_consts, dev)) {
    std::fprintf(stderr, "climt_b200: rrtmg_sw_ini_wrapper failed: %s\n", cb200_global_error());
    g_engine = nullptr;  // every later wrapper call NaN-fills its outputs and reports
  }
}
// Failure policy of the void reference-named wrappers: see lw_engine.cu (NaN-filled outputs + stderr + cb200_global_error()).
namespace {
void sw_wrapper_fail(const std::string& msg, int ncol, int nlay, double* uflx, double* dflx, double* hr, double* uflxc,
                     double* dflxc, double* hrc) {
  std::fprintf(stderr, "climt_b200: %s\n", msg.c_str());
  cb::set_global_error(msg);
  const size_t n1 = (size_t)ncol * (nlay + 1), n0 = (size_t)ncol * nlay;
  cb::mcica::nan_fill(uflx, n1); cb::mcica::nan_fill(dflx, n1); cb::mcica::nan_fill(hr, n0);
  cb::mcica::nan_fill(uflxc, n1); cb::mcica::nan_fill(dflxc, n1); cb::mcica::nan_fill(hrc, n0);
}
int sw_ngb(int g) {  // 0-based band (0 = band 16) of 0-based g-point g (ngb, rrtmg_sw_init.f90:286-293)
  int b = 0;
  while (b < 13 && g >= kGS[b + 1]) ++b;
  return b;
}
}  // namespace

extern "C" void rrtmg_sw_nomcica_wrapper(int* ncol, int* nlay, int* icld, int* iaer, double* play, double* plev,
                                         double* tlay, double* tlev, double* tsfc, double* h2ovmr, double* o3vmr,
                                         double* co2vmr, double* ch4vmr, double* n2ovmr, double* o2vmr, double* asdir,
                                         double* asdif, double* aldir, double* aldif, double* coszen, double* adjes,
                                         int* dyofyr, double* scon, int* isolvar, int* inflgsw, int* iceflgsw,
                                         int* liqflgsw, double* cldfr, double* taucld, double* ssacld, double* asmcld,
                                         double* fsfcld, double* cicewp, double* cliqwp, double* reice, double* reliq,
                                         double* tauaer, double* ssaaer, double* asmaer, double* ecaer, double* swuflx,
                                         double* swdflx, double* swhr, double* swuflxc, double* swdflxc, double* swhrc,
                                         double* bndsolvar, double* indsolvar, double* solcycfrac) {
  auto fail = [&](const std::string& m) { sw_wrapper_fail(m, *ncol, *nlay, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc); };
  if (!g_engine) return fail("rrtmg_sw_ini_wrapper has not been called (or failed)");
  if (*icld < 0 || *icld > 3) *icld = 2;
  if (*iaer != 0 && *iaer != 6 && *iaer != 10) *iaer = 0;
  cb200_sw_set_mcica(g_engine, 0, 1, 0);
  cb200_sw_set_options(g_engine, *icld, *iaer, *inflgsw, *iceflgsw, *liqflgsw);
  cb200_sw_set_solar(g_engine, *isolvar, *scon, indsolvar, bndsolvar);
  cb200_sw_inputs in{play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif,
                     coszen, cldfr, taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, reliq, tauaer, ssaaer, asmaer, ecaer};
  cb200_sw_outputs out{swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc};
  if (cb200_sw_run_host(g_engine, *ncol, *nlay, *adjes, *dyofyr, solcycfrac ? *solcycfrac : 0.0, &in, &out))
    fail(cb200_sw_last_error(g_engine));
}

// Sub-column generator of the reference ABI (rrtmg_sw_c_binder.f90:59-107 -> mcica_subcol_gen_sw.f90:66-180): fills the
// caller's (ngptsw, ncol, nlay) arrays; clear sub-columns get tau 0, ssa 1, asm 0, fsf 0 (:517-525).  Host side for the
// reasons given at mcica_subcol_lw_wrapper (lw_engine.cu).
extern "C" void mcica_subcol_sw_wrapper(int* iplon, int* ncol, int* nlay, int* icld, int* permuteseed, int* irng,
                                        double* play, double* cldfrac, double* ciwp, double* clwp, double* rei,
                                        double* rel, double* tauc, double* ssac, double* asmc, double* fsfc,
                                        double* cldfmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl,
                                        double* relqmcl, double* taucmcl, double* ssacmcl, double* asmcmcl,
                                        double* fsfcmcl) {
  (void)iplon;
  const int nc = *ncol, nl = *nlay;
  if (*icld == 0) return;  // mcica_subcol_gen_sw.f90:143
  const size_t n3 = (size_t)112 * nc * nl, n2 = (size_t)nc * nl;
  auto fail = [&](const std::string& m) {
    std::fprintf(stderr, "climt_b200: %s\n", m.c_str());
    cb::set_global_error(m);
    for (double* p : {cldfmcl, ciwpmcl, clwpmcl, taucmcl, ssacmcl, asmcmcl, fsfcmcl}) cb::mcica::nan_fill(p, n3);
    cb::mcica::nan_fill(reicmcl, n2); cb::mcica::nan_fill(relqmcl, n2);
  };
  if (*icld < 0 || *icld > 3) return fail("MCICA_SUBCOL: INVALID ICLD");
  if (*irng != 0) *irng = 1;
  std::vector<unsigned> mask;
  if (*irng == 1) {
    cb::mcica::mask_mt_host(cldfrac, nc, nl, 112, 4, *icld, *permuteseed, mask);
  } else {
    mask.assign((size_t)nl * 4 * nc, 0u);
    std::atomic<int> bad{0};
    cb::WorkerPool::get().parallel_for((nc + 63) / 64, [&](int t) {
      for (int c = t * 64; c < nc && c < (t + 1) * 64; ++c)
        if (cb::mcica::mask_column_kiss(play, cldfrac, nc, nl, 112, 4, *icld, *permuteseed, mask.data(), nc, 0, c)) bad.store(1);
    });
    if (bad.load()) return fail("MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS.");
  }
  int ngb[112];
  for (int g = 0; g < 112; ++g) ngb[g] = sw_ngb(g);
  cb::mcica::SubcolIn si{ciwp, clwp, rei, rel, {tauc, ssac, asmc, fsfc}, {0., 1., 0., 0.}};
  cb::mcica::SubcolOut so{cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, {taucmcl, ssacmcl, asmcmcl, fsfcmcl}};
  cb::mcica::expand(mask.data(), nc, nl, 112, 4, 14, ngb, 4, si, so);
}

// rrtmg_sw_c_binder.f90:109-201 -> rrtmg_sw_rad.f90:97 (spcvmc_sw): per-g-point arrays folded back into mask bits + layer
// values (mcica_compat.h), then the same kernels as cb200_sw_run_host with McICA enabled.
extern "C" void rrtmg_sw_mcica_wrapper(int* ncol, int* nlay, int* icld, int* iaer, double* play, double* plev,
                                       double* tlay, double* tlev, double* tsfc, double* h2ovmr, double* o3vmr,
                                       double* co2vmr, double* ch4vmr, double* n2ovmr, double* o2vmr, double* asdir,
                                       double* asdif, double* aldir, double* aldif, double* coszen, double* adjes,
                                       int* dyofyr, double* scon, int* isolvar, int* inflgsw, int* iceflgsw,
                                       int* liqflgsw, double* cldfmcl, double* taucmcl, double* ssacmcl, double* asmcmcl,
                                       double* fsfcmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl,
                                       double* relqmcl, double* tauaer, double* ssaaer, double* asmaer, double* ecaer,
                                       double* swuflx, double* swdflx, double* swhr, double* swuflxc, double* swdflxc,
                                       double* swhrc, double* bndsolvar, double* indsolvar, double* solcycfrac) {
  const int nc = *ncol, nl = *nlay;
  auto fail = [&](const std::string& m) { sw_wrapper_fail(m, nc, nl, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc); };
  if (!g_engine) return fail("rrtmg_sw_ini_wrapper has not been called (or failed)");
  if (*icld < 0 || *icld > 3) *icld = 2;                  // rrtmg_sw_rad.f90:587
  if (*iaer != 0 && *iaer != 6 && *iaer != 10) *iaer = 0;
  cb200_sw_set_options(g_engine, *icld, *iaer, *inflgsw, *iceflgsw, *liqflgsw);
  cb200_sw_set_solar(g_engine, *isolvar, *scon, indsolvar, bndsolvar);
  int ngb[112];
  for (int g = 0; g < 112; ++g) ngb[g] = sw_ngb(g);
  cb::mcica::CollapseOut co;
  const double* const bandmcl[4] = {taucmcl, ssacmcl, asmcmcl, fsfcmcl};
  const std::string why = cb::mcica::collapse(nc, nl, 112, 4, 14, ngb, 4, cldfmcl, ciwpmcl, clwpmcl, bandmcl, co);
  if (!why.empty()) return fail(why);
  cb200_sw_set_mcica(g_engine, 1, 1, 0);
  g_engine->ext_mask.swap(co.mask);
  g_engine->ext_ncol = nc;
  g_engine->ext_nlay = nl;
  cb200_sw_inputs in{play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif,
                     coszen, co.cldfr.data(), co.band[0].data(), co.band[1].data(), co.band[2].data(), co.band[3].data(),
                     co.ciwp.data(), co.clwp.data(), reicmcl, relqmcl, tauaer, ssaaer, asmaer, ecaer};
  cb200_sw_outputs out{swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc};
  const int rc = cb200_sw_run_host(g_engine, nc, nl, *adjes, *dyofyr, solcycfrac ? *solcycfrac : 0.0, &in, &out);
  g_engine->ext_mask.clear();
  cb200_sw_set_mcica(g_engine, 0, 1, 0);
  if (rc) fail(cb200_sw_last_error(g_engine));
}
