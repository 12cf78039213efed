// climt_b200 -- CORK correlated-k longwave / shortwave engine: table re-layout, CUDA kernels (sm_100a), C ABI.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/climt_b200.h"
#include "engine_common.h"
#include "cork_tables.h"

using namespace cb::cork;

namespace {
using cb::kBlock;
#ifndef CB_CORK_MIN_BLOCKS
#define CB_CORK_MIN_BLOCKS 3
#endif
#ifndef CB_CORK_UMAX
#define CB_CORK_UMAX 8  // g-points per thread: largest of {8, 4, 2, 1} <= CB_CORK_UMAX dividing ngpt (r01 B200: U=8 12.8 ms, U=4 14.4 ms per 65536 x 60 LW call)
#endif

__global__ void __launch_bounds__(kBlock) k_cork_prep(const __grid_constant__ Table Tb, const Consts K, const __grid_constant__ In in,
                                                      const __grid_constant__ Work W, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) prep_cell(Tb, K, in, W, c0, c, blockIdx.y);
}

// One block = 128 adjacent columns x one unit (U g-points of one band)
template <int U, bool LW, typename KT, int OPT = 0, bool DIAG = false>
__global__ void __launch_bounds__(kBlock, CB_CORK_MIN_BLOCKS)
    k_cork_units(const __grid_constant__ Table Tb, const Consts K, const __grid_constant__ In in, const __grid_constant__ Work W,
                 int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int unit = blockIdx.y;
  const int band = unit / Tb.nchunk, chunk = unit - band * Tb.nchunk;
  if (LW) lw_unit<U, KT, OPT, DIAG>(Tb, K, in, W, c0, c, band, chunk, unit);
  else sw_unit<U, KT, OPT, DIAG>(Tb, K, in, W, c0, c, band, chunk, unit);
}

__global__ void __launch_bounds__(kBlock) k_cork_reduce(const __grid_constant__ Table Tb, const __grid_constant__ Work W,
                                                        const __grid_constant__ Out out, int nlev, int ncol, int c0, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) reduce_level(Tb, W, nlev, ncol, c0, c, blockIdx.y, out);
}
// diagnostics_level >= 1 only
__global__ void __launch_bounds__(kBlock) k_cork_reduce_diag(const __grid_constant__ Table Tb, const __grid_constant__ Work W,
                                                             const __grid_constant__ Out out, int nlev, int n, int lw) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) reduce_diag_level(Tb, W, nlev, c, blockIdx.y, lw != 0, out);
}

__global__ void __launch_bounds__(kBlock) k_cork_heat(const __grid_constant__ Table Tb, const Consts K, const __grid_constant__ In in,
                                                      const __grid_constant__ Out out, int c0, int n, int lw) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) heat_layer(Tb, K, in, out, c0, c, blockIdx.y, lw != 0);
}

#define CUDA_OK(call)                                                                         \
  do {                                                                                        \
    cudaError_t err__ = (call);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      e->error = std::string(#call) + ": " + cudaGetErrorString(err__);                       \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

}  // namespace

struct cb200_cork_engine {
  int device = 0;
  Table T{};
  Consts K{};
  bool premixed = true, is_lw = false, is_sw = false, k_f64 = false;
  void* d_blob = nullptr;  // every table array, one allocation
  double* d_solar = nullptr;  // (nband, ngpt) solar flux of the current call
  int nunits = 0;
  int cap_ncc = 0, cap_nlev = 0;
  Work W{};
  int max_chunk = 16384;
  cb::HostPipe pipe;
  std::string error;
  int launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double unit_ms = 0.0;
  cb200_cork_diagnostics diag{};  // level 0 = off (cb200_cork_set_diagnostics)
  double* d_wsum = nullptr;      // (nband) the caller's weight sums, or null
  double* d_diag = nullptr;      // host calls: the diagnostics of the whole call, copied back at its end
  size_t d_diag_cap = 0;
  size_t dpart_cap = 0;

  void free_work() {
    cudaFree(W.ws); cudaFree(W.idx); cudaFree(W.scr); cudaFree(W.part); cudaFree(W.dpart);
    W = Work{};
    cap_ncc = cap_nlev = 0;
    dpart_cap = 0;
  }
  // diagnostics_level >= 1: the per-unit sums of the diagnostics (sized like `part`)
  int ensure_dpart(bool lw, int ncc, int nlev) {
    cb200_cork_engine* e = this;
    W.ndiag = lw ? (int)DL_N : (int)DS_N;
    W.diag_level = diag.level;
    W.wsum = diag.weight_sum ? d_wsum : nullptr;
    const size_t need = (size_t)nunits * W.ndiag * (nlev + 1) * ncc;
    if (need > dpart_cap) {
      cudaDeviceSynchronize();
      cudaFree(W.dpart);
      W.dpart = nullptr;
      dpart_cap = 0;
      CUDA_OK(cudaMalloc(&W.dpart, need * sizeof(double)));
      dpart_cap = need;
    }
    return 0;
  }
  int ensure_work(int ncc, int nlev) {
    cb200_cork_engine* e = this;
    if (ncc <= cap_ncc && nlev <= cap_nlev && W.ws) return 0;
    free_work();
    const size_t n = (size_t)ncc, L = (size_t)nlev;
    const int nscr = 7 * T.U;  // SW needs 7 rows per g-point, LW 2
    CUDA_OK(cudaMalloc(&W.ws, sizeof(double) * (F_AMT0 + T.ngas) * L * n));
    CUDA_OK(cudaMalloc(&W.idx, sizeof(int) * L * n));
    CUDA_OK(cudaMalloc(&W.scr, sizeof(double) * (size_t)nunits * (is_sw ? nscr : 2 * T.U) * L * n));
    CUDA_OK(cudaMalloc(&W.part, sizeof(double) * (size_t)nunits * 3 * (L + 1) * n));
    cap_ncc = ncc;
    cap_nlev = nlev;
    return 0;
  }
};

extern "C" int cb200_cork_create(cb200_cork_engine** out, const cb200_cork_table* t, double g, double cpd, double sigma,
                                 int device) {
  *out = nullptr;
  const std::string bad = check_table(t);
  if (!bad.empty()) { cb::set_global_error(bad); return -1; }
  auto* e = new cb200_cork_engine();
  e->device = device;
  e->K = Consts{g, cpd, sigma, 1.66};
  e->premixed = t->premixed != 0;
  int umax = CB_CORK_UMAX;
  if (const char* u = std::getenv("CLIMT_B200_CORK_U")) {
    const int req = std::atoi(u);
    if (req == 1 || req == 2 || req == 4 || req == 8) umax = req;
  }
  TableImage im;
  build_images(t, umax, e->T, im);
  e->is_lw = im.is_lw;
  e->is_sw = im.is_sw;
  e->nunits = e->T.nband * e->T.nchunk;
  e->k_f64 = im.k_f64;
  auto up256 = [](size_t b) { return (b + 255) / 256 * 256; };
  const void* ksrc = im.k_f64 ? static_cast<const void*>(im.k64.data()) : static_cast<const void*>(im.k32.data());
  const size_t kbytes = im.k_f64 ? im.k64.size() * sizeof(double) : im.k32.size() * sizeof(float);
  const size_t bk = up256(kbytes), bp = up256(im.planck.size() * sizeof(double)), bd = im.d.size() * sizeof(double);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&e->d_blob, bk + bp + bd);
  char* base = static_cast<char*>(e->d_blob);
  if (ce == cudaSuccess) ce = cudaMemcpy(base, ksrc, kbytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !im.planck.empty())
    ce = cudaMemcpy(base + bk, im.planck.data(), im.planck.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(base + bk + bp, im.d.data(), bd, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    cb::set_global_error(std::string("cork create: ") + cudaGetErrorString(ce));
    cudaFree(e->d_blob);
    delete e;
    return -1;
  }
  bind(e->T, im, base, reinterpret_cast<const double*>(base + bk), reinterpret_cast<const double*>(base + bk + bp));
  if (const char* mc = std::getenv("CLIMT_B200_MAX_CHUNK")) e->max_chunk = std::max(128, std::atoi(mc));
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  *out = e;
  return 0;
}

// ---- the engine's own on-disk table format (writer and format description: climt_b200/table_store.py) ----
namespace {
struct FileEntry {
  char name[48];
  int32_t dtype, ndim;
  int64_t shape[7], offset, nbytes;
};
static_assert(sizeof(FileEntry) == 128, "container entry is 128 bytes");

struct KTableFile {
  std::vector<char> raw;
  std::vector<FileEntry> entries;
  std::vector<std::vector<double>> promoted;  // float32 grids promoted to double (exact)
  const FileEntry* find(const char* name) const {
    for (const auto& e : entries)
      if (std::strncmp(e.name, name, sizeof(e.name)) == 0) return &e;
    return nullptr;
  }
  size_t count(const FileEntry* e) const {
    size_t n = 1;
    for (int i = 0; i < e->ndim; ++i) n *= (size_t)e->shape[i];
    return n;
  }
  const double* f64(const char* name) {
    const FileEntry* e = find(name);
    if (!e || e->dtype > 1) return nullptr;
    if (e->dtype == 0) return reinterpret_cast<const double*>(raw.data() + e->offset);
    const float* f = reinterpret_cast<const float*>(raw.data() + e->offset);
    promoted.emplace_back(f, f + count(e));
    return promoted.back().data();
  }
  int i32(const char* name, int dflt) const {
    const FileEntry* e = find(name);
    return (e && e->dtype == 2 && e->nbytes >= 4) ? *reinterpret_cast<const int32_t*>(raw.data() + e->offset) : dflt;
  }
};

std::string read_ktable_file(const char* path, KTableFile& f) {
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return std::string("cork: cannot open ") + path;
  std::fseek(fp, 0, SEEK_END);
  const long size = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  f.raw.resize(size > 0 ? (size_t)size : 0);
  const size_t got = f.raw.empty() ? 0 : std::fread(f.raw.data(), 1, f.raw.size(), fp);
  std::fclose(fp);
  if (got != f.raw.size() || f.raw.size() < 16 || std::memcmp(f.raw.data(), "CB2KTB01", 8) != 0)
    return std::string("cork: ") + path + " is not a CB2KTB01 table container";
  int64_t n = 0;
  std::memcpy(&n, f.raw.data() + 8, 8);
  if (n < 0 || (uint64_t)n > (f.raw.size() - 16) / sizeof(FileEntry)) return std::string("cork: truncated header in ") + path;
  f.entries.resize((size_t)n);
  std::memcpy(f.entries.data(), f.raw.data() + 16, (size_t)n * sizeof(FileEntry));
  for (auto& e : f.entries) {
    e.name[sizeof(e.name) - 1] = 0;
    static const size_t item[4] = {8, 4, 4, 1};
    if (e.dtype < 0 || e.dtype > 3 || e.ndim < 0 || e.ndim > 7 || e.offset < 0 || e.nbytes < 0 || (e.offset % 64) != 0 ||
        (size_t)e.offset + (size_t)e.nbytes > f.raw.size() || f.count(&e) * item[e.dtype] != (size_t)e.nbytes)
      return std::string("cork: bad entry '") + e.name + "' in " + path;
  }
  return "";
}
}  // namespace

extern "C" int cb200_cork_create_from_file(cb200_cork_engine** out, const char* path, double g, double cpd, double sigma, int device) {
  *out = nullptr;
  KTableFile f;
  std::string bad;
  try {  // (nothing may unwind through the C ABI: a hostile size field ends here as an error string)
    bad = read_ktable_file(path, f);
  } catch (const std::exception& ex) {
    bad = std::string("cork: cannot read ") + path + ": " + ex.what();
  }
  if (!bad.empty()) { cb::set_global_error(bad); return -1; }
  if (f.i32("_overlap_additive", 1) == 0) {
    cb::set_global_error("cork: this container holds an ESFT-overlap table; the engine evaluates those as additive tables on the "
                         "combined g-points -- expand it first (climt_b200.cork.expand_esft_table, then cb200_cork_create)");
    return -1;
  }
  const FileEntry* k = f.find("k_coefficients");
  if (!k || k->dtype > 1 || k->ndim < 5) { cb::set_global_error("cork: k_coefficients missing or not (gas, band, g, T, P[, X[, C]])"); return -1; }
  cb200_cork_table t;
  std::memset(&t, 0, sizeof(t));
  t.ngas = (int)k->shape[0]; t.nband = (int)k->shape[1]; t.ngpt = (int)k->shape[2]; t.nT = (int)k->shape[3]; t.nP = (int)k->shape[4];
  t.nX = k->ndim >= 6 ? (int)k->shape[5] : 0;
  t.nC = k->ndim == 7 ? (int)k->shape[6] : 0;
  if (k->dtype == 1) t.k_coefficients_f32 = reinterpret_cast<const float*>(f.raw.data() + k->offset);
  else t.k_coefficients_f64 = reinterpret_cast<const double*>(f.raw.data() + k->offset);
  t.temperature_grid = f.f64("temperature_grid");
  t.pressure_grid_log = f.f64("pressure_grid_log");
  t.h2o_vmr_grid = t.nX ? f.f64("h2o_vmr_grid") : nullptr;
  t.co2_vmr_grid = t.nC ? f.f64("co2_vmr_grid") : nullptr;
  t.gpoint_weights = f.f64("gpoint_weights");
  if (const FileEntry* pf = f.find("planck_fraction")) {
    if (pf->ndim != 3) { cb::set_global_error("cork: planck_fraction must be (band, g, T)"); return -1; }
    t.planck_fraction = f.f64("planck_fraction");
    t.nband_pf = (int)pf->shape[0];
    t.ngpt_pf = (int)pf->shape[1];
  }
  if (const FileEntry* ck = f.find("continuum_kappa"))
    if (ck->ndim == 4 && t.nX) t.continuum_kappa = f.f64("continuum_kappa");
  t.solar_source_per_gpoint = f.f64("solar_source_per_gpoint");
  t.rayleigh_coefficient = f.f64("rayleigh_coefficient");
  t.co2_logk = f.i32("_co2_logk", 1);
  t.premixed = f.i32("_premixed", 0);
  return cb200_cork_create(out, &t, g, cpd, sigma, device);
}

extern "C" int cb200_cork_create_picket(cb200_cork_engine** out, const cb200_picket_coeffs* c, int longwave, double g, double cpd,
                                        double sigma, int device) {
  *out = nullptr;
  static_assert(sizeof(cb200_picket_coeffs) == sizeof(Picket) && CB200_PICKET_MAX_REGIONS == kPicketMaxRegions, "picket layout");
  if (!c || c->nregion < 1 || c->nregion > kPicketMaxRegions) {
    cb::set_global_error("cork picket: need 1.." + std::to_string(kPicketMaxRegions) + " T_eff regions");
    return -1;
  }
  auto* e = new cb200_cork_engine();
  e->device = device;
  e->K = Consts{g, cpd, sigma, 1.66};
  e->premixed = true;
  e->is_lw = longwave != 0;
  e->is_sw = !e->is_lw;
  Table& T = e->T;
  T = Table{};
  T.optics = 1;
  T.ngas = 1; T.nband = e->is_lw ? 2 : 3; T.ngpt = 1; T.U = 1; T.nchunk = 1;
  T.nT = T.nP = T.nX = T.nC = 1;
  std::memcpy(&T.pk, c, sizeof(Picket));
  e->nunits = T.nband;
  const double ones[3] = {1.0, 1.0, 1.0};  // weights = np.ones((nband, 1)) (cork/lw/component.py:241, sw/component.py:265)
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&e->d_blob, sizeof(ones));
  if (ce == cudaSuccess) ce = cudaMemcpy(e->d_blob, ones, sizeof(ones), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    cb::set_global_error(std::string("cork picket create: ") + cudaGetErrorString(ce));
    cudaFree(e->d_blob);
    delete e;
    return -1;
  }
  T.weights = static_cast<const double*>(e->d_blob);
  if (const char* mc = std::getenv("CLIMT_B200_MAX_CHUNK")) e->max_chunk = std::max(128, std::atoi(mc));
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  *out = e;
  return 0;
}

extern "C" void cb200_cork_destroy(cb200_cork_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  e->free_work();
  e->pipe.destroy();
  cudaFree(e->d_blob);
  cudaFree(e->d_solar);
  cudaFree(e->d_wsum);
  cudaFree(e->d_diag);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  delete e;
}
extern "C" const char* cb200_cork_last_error(cb200_cork_engine* e) { return e ? e->error.c_str() : cb::g_error.c_str(); }
extern "C" int cb200_cork_last_launches(cb200_cork_engine* e) { return e->launches; }
extern "C" int cb200_cork_enable_timing(cb200_cork_engine* e, int on) { e->timing = on != 0; return 0; }
extern "C" double cb200_cork_last_unit_kernel_ms(cb200_cork_engine* e) { return e->unit_ms; }

namespace {

In make_in(int ncol, int nlev, const cb200_cork_inputs* p, const double* d_solar) {
  In in{};
  in.ncol = ncol; in.nlev = nlev;
  in.T = p->T; in.p = p->p; in.p_int = p->p_int; in.T_surf = p->T_surf;
  in.q_h2o = p->q_h2o; in.co2_vmr = p->co2_vmr; in.gas_q = p->gas_q;
  in.emissivity = p->emissivity; in.tau_cloud = p->tau_cloud;
  in.zenith = p->zenith; in.albedo = p->albedo; in.ssa_cloud = p->ssa_cloud; in.g_cloud = p->g_cloud;
  in.solar_flux = d_solar;
  in.T_irr = p->T_irr; in.T_int = p->T_int; in.bond_albedo = p->bond_albedo;
  return in;
}

Out make_out(const cb200_cork_outputs* p) {
  Out o{};
  o.up_broad = p->up_broad; o.down_broad = p->down_broad; o.heating = p->heating_rate; o.up_band = p->up_band;
  o.down_band = p->down_band; o.tau_band = p->tau_band; o.trans_band = p->trans_band; o.hr_band = p->hr_band;
  return o;
}
// the diagnostics fields a level provides: LW 3 at level >= 1; SW 4 at level 1, all 10 at level >= 2 (sw/kernels.py:381-394)
bool diag_field_on(bool lw, int level, int j) {
  if (level <= 0) return false;
  if (lw) return j < (int)DL_N;
  return level >= 2 ? j < (int)DS_N : j <= (int)DS_DIRECT;
}
size_t diag_rows(bool lw, int j, int nband, int nlev) {
  const bool iface = lw ? (j != (int)DL_TRANS) : diag_sw_is_interface(j);
  return (size_t)nband * (iface ? nlev + 1 : nlev);
}

int validate(cb200_cork_engine* e, bool lw, int ncol, int nlev, const cb200_cork_inputs* in, const cb200_cork_outputs* out) {
  if (ncol <= 0 || nlev <= 0) { e->error = "cork: bad ncol/nlev"; return -3; }
  if (lw && !e->is_lw) { e->error = "cork: this table has no planck_fraction (not a longwave table)"; return -3; }
  if (!lw && !e->is_sw) { e->error = "cork: this table has no solar_source_per_gpoint (not a shortwave table)"; return -3; }
  if (!in->T || !in->p || !in->p_int || !out->up_broad || !out->down_broad || !out->heating_rate) { e->error = "cork: missing required array"; return -3; }
  if (lw && (!in->T_surf || !in->emissivity)) { e->error = "cork lw: T_surf and emissivity are required"; return -3; }
  if (!lw && (!in->zenith || !in->albedo)) { e->error = "cork sw: zenith and albedo are required"; return -3; }
  if (!lw && in->tau_cloud && (!in->ssa_cloud || !in->g_cloud)) { e->error = "cork sw: ssa_cloud and g_cloud must accompany tau_cloud"; return -3; }
  if (e->T.optics == 1 && (!in->T_irr || !in->T_int)) { e->error = "cork picket: T_irr and T_int are required"; return -3; }
  if (e->T.hasX && !in->q_h2o) { e->error = "k-table has an h2o_vmr_grid axis but specific humidity was not provided"; return -3; }  // correlated_k.py:262-265
  if (e->T.hasC && !in->co2_vmr) { e->error = "k-table has a co2_vmr_grid axis but co2_vmr was not provided"; return -3; }   // :272-275
  if (!e->premixed && !in->gas_q) { e->error = "cork: a non-premixed table needs the gas mass mixing ratios"; return -3; }
  if (out->hr_band && (!out->up_band || !out->down_band)) { e->error = "cork: hr_band needs up_band and down_band"; return -3; }
  if (out->trans_band && !out->tau_band) { e->error = "cork: trans_band needs tau_band"; return -3; }
  return 0;
}

int launch_chunk(cb200_cork_engine* e, bool lw, const Consts& K, const In& in_, const Out& out, Work& W, int c0, int n, int out_ncol,
                 cudaStream_t st) {
  In in = in_;
  if (e->premixed) in.gas_q = nullptr;
  const int nlev = in.nlev;
  const int gx = (n + kBlock - 1) / kBlock;
  k_cork_prep<<<dim3(gx, nlev), kBlock, 0, st>>>(e->T, K, in, W, c0, n);
  if (e->timing) cudaEventRecord(e->ev0, st);
  const dim3 grid(gx, e->nunits);
  const bool dg = W.diag_level > 0;
#define CB_LAUNCH4(U, D)                                                                                         \
    if (lw && !e->k_f64) k_cork_units<U, true, float, 0, D><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);     \
    else if (lw) k_cork_units<U, true, double, 0, D><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);            \
    else if (!e->k_f64) k_cork_units<U, false, float, 0, D><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);     \
    else k_cork_units<U, false, double, 0, D><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);
#define CB_LAUNCH(U)                                                                                          \
  case U:                                                                                                     \
    if (dg) { CB_LAUNCH4(U, true) } else { CB_LAUNCH4(U, false) }                                             \
    break;
  if (e->T.optics == 1) {
    if (lw && dg) k_cork_units<1, true, float, 1, true><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);
    else if (lw) k_cork_units<1, true, float, 1><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);
    else if (dg) k_cork_units<1, false, float, 1, true><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);
    else k_cork_units<1, false, float, 1><<<grid, kBlock, 0, st>>>(e->T, K, in, W, c0, n);
  } else switch (e->T.U) {
    CB_LAUNCH(1) CB_LAUNCH(2) CB_LAUNCH(4) CB_LAUNCH(8)
  }
#undef CB_LAUNCH
#undef CB_LAUNCH4
  if (e->timing) cudaEventRecord(e->ev1, st);
  k_cork_reduce<<<dim3(gx, nlev + 1), kBlock, 0, st>>>(e->T, W, out, nlev, out_ncol, c0, n);
  if (dg) { k_cork_reduce_diag<<<dim3(gx, nlev + 1), kBlock, 0, st>>>(e->T, W, out, nlev, n, lw ? 1 : 0); e->launches += 1; }
  k_cork_heat<<<dim3(gx, nlev), kBlock, 0, st>>>(e->T, K, in, out, c0, n, lw ? 1 : 0);
  e->launches += 4;
  if (e->timing) {
    CUDA_OK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->unit_ms += ms;
  }
  return 0;
}

// SW: the (nband, ngpt) solar flux of this call -> device (tiny; ordered on `st` before the kernels)
int upload_solar(cb200_cork_engine* e, const double* h_solar, cudaStream_t st) {
  const size_t n = (size_t)e->T.nband * e->T.ngpt;
  if (!e->d_solar) CUDA_OK(cudaMalloc(&e->d_solar, n * sizeof(double)));
  if (!h_solar && !e->T.solar) { e->error = "cork picket sw: the solar_flux argument is required"; return -3; }
  if (h_solar) CUDA_OK(cudaMemcpyAsync(e->d_solar, h_solar, n * sizeof(double), cudaMemcpyHostToDevice, st));
  else CUDA_OK(cudaMemcpyAsync(e->d_solar, e->T.solar, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int run_device(cb200_cork_engine* e, bool lw, int ncol, int nlev, double scalar, const double* h_solar, const cb200_cork_inputs* pin,
               const cb200_cork_outputs* pout, cudaStream_t st) {
  if (int rc = validate(e, lw, ncol, nlev, pin, pout)) return rc;
  CUDA_OK(cudaSetDevice(e->device));
  if (!lw) if (int rc = upload_solar(e, h_solar, st)) return rc;
  int chunk = ncol < e->max_chunk ? ncol : e->max_chunk;
  chunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(chunk, nlev)) return -1;
  Work W = e->W;
  W.ncc = chunk;
  W.nscr = lw ? 2 * e->T.U : 7 * e->T.U;
  Consts K = e->K;
  if (lw) K.D = scalar;
  const In in = make_in(ncol, nlev, pin, e->d_solar);
  Out out = make_out(pout);
  W.diag_level = 0;
  if (e->diag.level > 0) {  // device pointers of the caller
    if (e->ensure_dpart(lw, chunk, nlev)) return -1;
    W.dpart = e->W.dpart; W.ndiag = e->W.ndiag; W.diag_level = e->W.diag_level; W.wsum = e->W.wsum;
    for (int j = 0; j < W.ndiag; ++j) out.diag[j] = diag_field_on(lw, e->diag.level, j) ? e->diag.field[j] : nullptr;
    out.diag_ncol = ncol;
  }
  e->launches = 0;
  e->unit_ms = 0.0;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int n = (ncol - c0) < chunk ? (ncol - c0) : chunk;
    out.diag_c0 = c0;
    if (launch_chunk(e, lw, K, in, out, W, c0, n, ncol, st)) return -1;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

// host-pointer call through the 3-stream chunk pipeline
int run_host(cb200_cork_engine* e, bool lw, int ncol, int nlev, double scalar, const double* h_solar, const cb200_cork_inputs* hin,
             const cb200_cork_outputs* hout) {
  if (int rc = validate(e, lw, ncol, nlev, hin, hout)) return rc;
  CUDA_OK(cudaSetDevice(e->device));
  cb::HostPipe& P = e->pipe;
  CUDA_OK(P.init());
  if (!lw) if (int rc = upload_solar(e, h_solar, P.s_cmp)) return rc;
  const int L = nlev, nb = e->T.nband;
  // inputs in cb200_cork_inputs order: T p p_int T_surf q_h2o co2_vmr gas_q emissivity tau_cloud zenith albedo ssa_cloud g_cloud
  //                                    T_irr T_int bond_albedo
  constexpr int NI = 16;
  const int irows[NI] = {L, L, L + 1, 1, L, L, e->T.ngas * L, nb, L, 1, 1, L, L, 1, 1, 1};
  const int inner[NI] = {1, 1, 1, 1, 1, 1, 1, 1, nb, 1, 1, nb, nb, 1, 1, 1};
  const double* const* hp = reinterpret_cast<const double* const*>(hin);
  bool used[NI];
  for (int i = 0; i < NI; ++i) used[i] = hp[i] != nullptr;
  used[13] = used[13] && e->T.optics == 1; used[14] = used[14] && e->T.optics == 1; used[15] = used[15] && e->T.optics == 1 && !lw;
  used[3] = used[3] && lw; used[7] = used[7] && lw;
  used[9] = used[9] && !lw; used[10] = used[10] && !lw; used[11] = used[11] && !lw; used[12] = used[12] && !lw;
  used[4] = used[4] && e->T.hasX; used[5] = used[5] && e->T.hasC; used[6] = used[6] && !e->premixed;
  if (!lw && !used[8]) { used[11] = used[12] = false; }
  // outputs in cb200_cork_outputs order
  const int orows[8] = {L + 1, L + 1, L, nb * (L + 1), nb * (L + 1), nb * L, nb * L, nb * L};
  double* const* hop = reinterpret_cast<double* const*>(hout);
  size_t irow_tot = 0, orow_tot = 0;
  for (int i = 0; i < NI; ++i) if (used[i]) irow_tot += (size_t)irows[i] * inner[i];
  for (int i = 0; i < 8; ++i) if (hop[i]) orow_tot += (size_t)orows[i];
  int chunk = ncol < P.chunk ? ncol : P.chunk;
  const int wchunk = (chunk + kBlock - 1) / kBlock * kBlock;
  if (e->ensure_work(wchunk, nlev)) return -1;
  CUDA_OK(P.ensure(irow_tot * (size_t)chunk, orow_tot * (size_t)chunk));
  Work W = e->W;
  W.ncc = wchunk;
  W.nscr = lw ? 2 * e->T.U : 7 * e->T.U;
  W.diag_level = 0;
  // diagnostics_level >= 1: the fields of the whole call are kept on the device and copied back once the chunks are done (a
  // debugging mode -- it does not go through the chunk pipeline)
  double* ddiag[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (e->diag.level > 0) {
    if (e->ensure_dpart(lw, wchunk, nlev)) return -1;
    W.dpart = e->W.dpart; W.ndiag = e->W.ndiag; W.diag_level = e->W.diag_level; W.wsum = e->W.wsum;
    size_t tot = 0;
    for (int j = 0; j < W.ndiag; ++j)
      if (diag_field_on(lw, e->diag.level, j) && e->diag.field[j]) tot += diag_rows(lw, j, nb, L) * (size_t)ncol;
    if (tot > e->d_diag_cap) {
      cudaFree(e->d_diag);
      e->d_diag = nullptr;
      e->d_diag_cap = 0;
      CUDA_OK(cudaMalloc(&e->d_diag, tot * sizeof(double)));
      e->d_diag_cap = tot;
    }
    size_t off = 0;
    for (int j = 0; j < W.ndiag; ++j)
      if (diag_field_on(lw, e->diag.level, j) && e->diag.field[j]) { ddiag[j] = e->d_diag + off; off += diag_rows(lw, j, nb, L) * (size_t)ncol; }
  }
  Consts K = e->K;
  if (lw) K.D = scalar;
  e->launches = 0;
  e->unit_ms = 0.0;
  int k = 0;
  for (int c0 = 0; c0 < ncol; c0 += chunk, ++k) {
    const int n = (ncol - c0) < chunk ? (ncol - c0) : chunk;
    const int s = k & 1;
    CUDA_OK(cudaStreamWaitEvent(P.s_in, P.cmp_done[s], 0));
    cb200_cork_inputs din;
    const double** dp = reinterpret_cast<const double**>(&din);
    size_t off = 0;
    for (int i = 0; i < NI; ++i) {
      if (!used[i]) { dp[i] = nullptr; continue; }
      CUDA_OK(P.gather(P.d_in[s] + off, hp[i], irows[i], ncol, c0, n, inner[i]));
      dp[i] = P.d_in[s] + off;
      off += (size_t)irows[i] * inner[i] * n;
    }
    CUDA_OK(cudaEventRecord(P.in_done[s], P.s_in));
    cb200_cork_outputs dout;
    double** dop = reinterpret_cast<double**>(&dout);
    off = 0;
    for (int i = 0; i < 8; ++i) {
      dop[i] = hop[i] ? P.d_out[s] + off : nullptr;
      if (hop[i]) off += (size_t)orows[i] * n;
    }
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.in_done[s], 0));
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.out_done[s], 0));
    const In in = make_in(n, nlev, &din, e->d_solar);
    Out out = make_out(&dout);
    for (int j = 0; j < 10; ++j) out.diag[j] = ddiag[j];
    out.diag_ncol = ncol; out.diag_c0 = c0;
    if (launch_chunk(e, lw, K, in, out, W, 0, n, n, P.s_cmp)) return -1;
    CUDA_OK(cudaEventRecord(P.cmp_done[s], P.s_cmp));
    CUDA_OK(cudaStreamWaitEvent(P.s_out, P.cmp_done[s], 0));
    for (int i = 0; i < 8; ++i)
      if (hop[i]) CUDA_OK(P.scatter(hop[i], dop[i], orows[i], ncol, c0, n));
    CUDA_OK(cudaEventRecord(P.out_done[s], P.s_out));
  }
  CUDA_OK(cudaStreamSynchronize(P.s_out));
  if (e->diag.level > 0) {
    CUDA_OK(cudaStreamSynchronize(P.s_cmp));
    for (int j = 0; j < W.ndiag; ++j)
      if (ddiag[j]) CUDA_OK(cudaMemcpy(e->diag.field[j], ddiag[j], diag_rows(lw, j, nb, L) * (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost));
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int cb200_cork_set_diagnostics(cb200_cork_engine* e, const cb200_cork_diagnostics* d) {
  if (!d || d->level <= 0) { e->diag = cb200_cork_diagnostics{}; return 0; }
  e->diag = *d;
  if (d->weight_sum) {
    CUDA_OK(cudaSetDevice(e->device));
    if (!e->d_wsum) CUDA_OK(cudaMalloc(&e->d_wsum, sizeof(double) * e->T.nband));
    CUDA_OK(cudaMemcpy(e->d_wsum, d->weight_sum, sizeof(double) * e->T.nband, cudaMemcpyHostToDevice));
  }
  return 0;
}

static_assert(sizeof(cb200_cork_inputs) == 16 * sizeof(double*), "cb200_cork_inputs layout");
static_assert(sizeof(cb200_cork_outputs) == 8 * sizeof(double*), "cb200_cork_outputs layout");

extern "C" int cb200_cork_lw_run_device(cb200_cork_engine* e, int ncol, int nlev, double D, const cb200_cork_inputs* in,
                                        const cb200_cork_outputs* out, void* stream) {
  return run_device(e, true, ncol, nlev, D, nullptr, in, out, (cudaStream_t)stream);
}
extern "C" int cb200_cork_sw_run_device(cb200_cork_engine* e, int ncol, int nlev, const double* solar_flux, const cb200_cork_inputs* in,
                                        const cb200_cork_outputs* out, void* stream) {
  return run_device(e, false, ncol, nlev, 0.0, solar_flux, in, out, (cudaStream_t)stream);
}
extern "C" int cb200_cork_lw_run_host(cb200_cork_engine* e, int ncol, int nlev, double D, const cb200_cork_inputs* in,
                                      const cb200_cork_outputs* out) {
  return run_host(e, true, ncol, nlev, D, nullptr, in, out);
}
extern "C" int cb200_cork_sw_run_host(cb200_cork_engine* e, int ncol, int nlev, const double* solar_flux, const cb200_cork_inputs* in,
                                      const cb200_cork_outputs* out) {
  return run_host(e, false, ncol, nlev, 0.0, solar_flux, in, out);
}
