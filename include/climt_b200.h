/* climt_b200 -- C ABI of the B200-native column radiative-transfer engines.
 *
 * This is the drop-in boundary: it replaces the C symbols climt's Cython shims bind
 * (all citations relative to the reference tree, climt/):
 *
 *   rrtmg_set_constants        _lib/rrtmg_lw/rrlw_con.f90:46-71      (decl _components/rrtmg/lw/_rrtmg_lw.pyx:19-24)
 *   rrtmg_lw_ini_wrapper       _lib/rrtmg_lw/rrtmg_lw_c_binder.f90:39-48   (decl _rrtmg_lw.pyx:26)
 *   rrtmg_lw_nomcica_wrapper   _lib/rrtmg_lw/rrtmg_lw_c_binder.f90:176-256 (decl _rrtmg_lw.pyx:62-80)
 *   mcica_subcol_lw_wrapper    _lib/rrtmg_lw/rrtmg_lw_c_binder.f90:50-92   (decl _rrtmg_lw.pyx:28-40, call :261-276)
 *   rrtmg_lw_mcica_wrapper     _lib/rrtmg_lw/rrtmg_lw_c_binder.f90:94-174  (decl _rrtmg_lw.pyx:42-60, call :283-320)
 *   (the five shortwave twins are listed at the shortwave section below)
 *
 * Failure of a reference-named (void) entry point: the Fortran under the reference's symbols ends the process with `stop`;
 * here every output array of the call is filled with NaN, the message is printed to stderr and kept for
 * cb200_global_error().  Nothing is ever left stale.
 *
 * Two flavours are exported:
 *  (1) handle-based, re-entrant entry points (cb200_lw_*): explicit constants, explicit options, per-instance
 *      tables in HBM, status codes instead of Fortran `stop`.  Device-pointer and host-pointer variants.
 *  (2) the reference's own symbol names with the reference's own by-pointer signatures, bound to one
 *      process-global engine, so `_rrtmg_lw.pyx` links against libclimt_b200.so unchanged.
 *
 * Array layout everywhere = the reference ABI: Fortran (ncol, nlay[+1]) == C (nlay[+1], ncol), column
 * fastest; emis (16, ncol); taucld (nlay, ncol, 16) [Fortran (16,ncol,nlay)]; tauaer (16, nlay, ncol)
 * [Fortran (ncol,nlay,16)].  fp64.  Pressures hPa, temperatures K, gases volume mixing ratio,
 * cloud water paths g m-2, particle sizes micron.  Outputs: fluxes W m-2 on nlay+1 interfaces
 * (surface first), heating rates K day-1 on nlay layers.
 */
#ifndef CLIMT_B200_H
#define CLIMT_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb200_lw_engine cb200_lw_engine;

/* order = argument order of rrtmg_lw_nomcica_wrapper (rrtmg_lw_c_binder.f90:176-189) */
typedef struct cb200_lw_inputs {
  const double *play, *plev, *tlay, *tlev, *tsfc;
  const double *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *n2ovmr, *o2vmr, *cfc11vmr, *cfc12vmr, *cfc22vmr, *ccl4vmr;
  const double *emis;
  const double *cldfr, *taucld, *cicewp, *cliqwp, *reice, *reliq;
  const double *tauaer;
} cb200_lw_inputs;

typedef struct cb200_lw_outputs {
  double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc; /* hr/hrc: (nlay, ncol); fluxes: (nlay+1, ncol) */
} cb200_lw_outputs;

/* constants[11] = pi, grav [m s-2], planck [erg s], boltz [erg K-1], clight [cm s-1], avogadro, loschmidt [cm-3],
 * gas constant [erg mol-1 K-1], stefan-boltzmann [W cm-2 K-4], seconds per day  (rrtmg_set_constants)
 * + cp of dry air [J kg-1 K-1] (rrtmg_lw_ini_wrapper).  table_blob: file written by climt_b200/rrtmg_tables.py.
 * Returns 0 on success; on failure *out is NULL and cb200_global_error() explains. */
int cb200_lw_create(cb200_lw_engine** out, const char* table_blob, const double constants[11], int device);
void cb200_lw_destroy(cb200_lw_engine* e);
/* icld: 0 clear, 1 random (rtrn), 2 maximum-random, 3 maximum (both rtrnmr, as in rrtmg_lw_rad.nomcica.f90:527-541; with McICA the
 * overlap is applied by the sub-column generator instead); idrv: 0 only; inflag/iceflag/liqflag as RRTMG
 * (module globals in _rrtmg_lw.pyx:8-14 in the reference; per engine here). */
int cb200_lw_set_options(cb200_lw_engine* e, int icld, int idrv, int inflag, int iceflag, int liqflag);
/* McICA (rrtmg_lw_rad.f90 + mcica_subcol_gen_lw.f90): enabled 0/1; irng 0 = kissvec (per-column seeds, generated on the
 * device), 1 = Mersenne twister (one serial stream per call, generated on the host for bit parity);
 * permuteseed as in mcica_subcol_lw_wrapper (rrtmg_lw_c_binder.f90:50-92).  Call before cb200_lw_set_options. */
int cb200_lw_set_mcica(cb200_lw_engine* e, int enabled, int irng, int permuteseed);
/* idrv = 1 (calculate_change_up_flux, rrtmg_lw_rtrn.f90:279-296,458-512,546-555): where the next run call writes
 * d(upward flux)/d(surface temperature) [W m-2 K-1], total and clear sky, (nlay+1, ncol) each -- duflx_dt / duflxc_dt of
 * rrtmg_lw_c_binder.f90:176-256.  Host pointers for run_host, device pointers for run_device; nulls clear them. */
int cb200_lw_set_derivative_outputs(cb200_lw_engine* e, double* duflx_dt, double* duflxc_dt);
/* Asynchronous on `stream` (a cudaStream_t, NULL = default stream); pointers are device pointers owned by
 * the caller.  Returns 0 or a negative launch/configuration error.  Input-validation errors that the Fortran
 * turns into `stop` are reported by cb200_lw_check() after the stream has been synchronised. */
int cb200_lw_run_device(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* in,
                        const cb200_lw_outputs* out, void* stream);
/* Same with host pointers: stages H2D, runs, copies back, synchronises, checks. */
int cb200_lw_run_host(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* in,
                      const cb200_lw_outputs* out);
/* The same call split in two: _async returns once every copy and kernel of the call is enqueued (host buffers must stay
 * valid and, to overlap, be page-locked); cb200_lw_wait blocks until the outputs are in the caller's buffers and reports
 * input-validation errors like cb200_lw_check.  Lets a caller overlap the LW and SW engines' pipelines. */
/* Host-pointer calls only: let the engine do the components' marshal arithmetic on the device, chunk by chunk, instead of numpy on
 * the host (replaces mass_to_volume_mixing_ratio and get_interface_values, climt/_core/util.py:47-142, as called by
 * rrtmg/lw/component.py:373-393 and rrtmg/sw/component.py:560-600).  flags bit 0: the h2ovmr argument holds specific humidity
 * (kg/kg); bit 1: tlev is ignored (may be NULL) and computed from tlay, tsfc, play, plev.  0 (default) = the plain reference ABI. */
int cb200_lw_set_host_marshal(cb200_lw_engine* e, int flags);
int cb200_lw_run_host_async(cb200_lw_engine* e, int ncol, int nlay, const cb200_lw_inputs* in, const cb200_lw_outputs* out);
int cb200_lw_wait(cb200_lw_engine* e);
/* bytes the last host-pointer call moved over PCIe (arrays the option flags make dead are not transferred) */
void cb200_lw_last_transfer_bytes(cb200_lw_engine* e, double* h2d, double* d2h);
/* all-zero-input scan of the host calls: 1 active, 0 disabled by CLIMT_B200_SKIP_ZERO_INPUTS=0, -1 disabled by the engine
 * (the scan measured slower than the PCIe copy it saves on this host; sticky) */
int cb200_lw_zero_scan_state(cb200_lw_engine* e);
/* 0 = ok; >0 = input out of the range the reference accepts (message via cb200_lw_last_error). */
int cb200_lw_check(cb200_lw_engine* e);
const char* cb200_lw_last_error(cb200_lw_engine* e);
const char* cb200_global_error(void);
/* number of kernel launches issued by the last run call (for bench.py's gpu_launches) */
int cb200_lw_last_launches(cb200_lw_engine* e);
/* device time [ms] of the dominant kernel (g-point units) in the last run_host/run_device call when
 * timing was enabled with cb200_lw_enable_timing(e, 1); measured with CUDA events on the launch stream. */
int cb200_lw_enable_timing(cb200_lw_engine* e, int on);
double cb200_lw_last_unit_kernel_ms(cb200_lw_engine* e);   /* transfer kernel (k_units), CUDA events on its launch stream */
double cb200_lw_last_taumol_kernel_ms(cb200_lw_engine* e); /* k_lw_taumol */

/* ---- reference-named entry points (one process-global engine; tables from $CLIMT_B200_LW_TABLES or
 *      <dir of this library>/../data/_cache/rrtmg_lw_reduced.blob) ---- */
void rrtmg_set_constants(double* pi, double* grav, double* planck, double* boltz, double* clight, double* avogad,
                         double* alosmt, double* gascon, double* sbcnst, double* secdy);
void rrtmg_lw_ini_wrapper(double* cpdair);
void rrtmg_lw_nomcica_wrapper(int* ncol, int* nlay, int* icld, int* idrv, double* play, double* plev, double* tlay,
                              double* tlev, double* tsfc, double* h2ovmr, double* o3vmr, double* co2vmr,
                              double* ch4vmr, double* n2ovmr, double* o2vmr, double* cfc11vmr, double* cfc12vmr,
                              double* cfc22vmr, double* ccl4vmr, double* emis, int* inflglw, int* iceflglw,
                              int* liqflglw, double* cldfr, double* taucld, double* cicewp, double* cliqwp,
                              double* reice, double* reliq, double* tauaer, double* uflx, double* dflx, double* hr,
                              double* uflxc, double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt);
/* McICA, the reference's two-step form.  mcica_subcol_lw_wrapper (sub-column generator, mcica_subcol_gen_lw.f90:48-154) fills the
 * caller's Fortran (140, ncol, nlay) arrays cldfmcl / ciwpmcl / clwpmcl / taucmcl and (ncol, nlay) reicmcl / relqmcl from the layer
 * cloud fields; tauc is Fortran (16, ncol, nlay); irng is intent(inout) (anything but 0 becomes 1 = Mersenne twister);
 * icld = 0 returns without touching the outputs (:119).  rrtmg_lw_mcica_wrapper (rrtmg_lw_rad.f90:80, rtrnmc) consumes them:
 * sub-column cloud fractions must be 0 or 1 and the cloudy sub-columns of a layer must share one set of water paths / optical
 * depths per band -- which is what the generator produces -- otherwise the call fails as described at the top of this file. */
void mcica_subcol_lw_wrapper(int* iplon, int* ncol, int* nlay, int* icld, int* permuteseed, int* irng, double* play,
                             double* cldfrac, double* ciwp, double* clwp, double* rei, double* rel, double* tauc,
                             double* cldfmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl, double* relqmcl,
                             double* taucmcl);
void rrtmg_lw_mcica_wrapper(int* ncol, int* nlay, int* icld, int* idrv, double* play, double* plev, double* tlay,
                            double* tlev, double* tsfc, double* h2ovmr, double* o3vmr, double* co2vmr, double* ch4vmr,
                            double* n2ovmr, double* o2vmr, double* cfc11vmr, double* cfc12vmr, double* cfc22vmr,
                            double* ccl4vmr, double* emis, int* inflglw, int* iceflglw, int* liqflglw, double* cldfmcl,
                            double* taucmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl, double* relqmcl,
                            double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc, double* dflxc,
                            double* hrc, double* duflx_dt, double* duflxc_dt);


/* ============================== shortwave ==============================
 * Replaces (climt/_lib/rrtmg_sw/rrtmg_sw_c_binder.f90): rrtmg_sw_set_constants :19-46, rrtmg_sw_ini_wrapper :48-57,
 * rrtmg_sw_nomcica_wrapper :203-296, mcica_subcol_sw_wrapper :59-107, rrtmg_sw_mcica_wrapper :109-201
 * (decls in _components/rrtmg/sw/_rrtmg_sw.pyx:22-105; McICA call order :341-417).
 * Cloud arrays taucld/ssacld/asmcld/fsfcld are (nlay, ncol, 14) [Fortran (14,ncol,nlay)]; aerosol arrays
 * tauaer/ssaaer/asmaer (14, nlay, ncol); ecaer (6, nlay, ncol); albedos and coszen (ncol). */
typedef struct cb200_sw_engine cb200_sw_engine;
typedef struct cb200_sw_inputs {
  const double *play, *plev, *tlay, *tlev, *tsfc;
  const double *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *n2ovmr, *o2vmr;
  const double *asdir, *asdif, *aldir, *aldif, *coszen;
  const double *cldfr, *taucld, *ssacld, *asmcld, *fsfcld, *cicewp, *cliqwp, *reice, *reliq;
  const double *tauaer, *ssaaer, *asmaer, *ecaer;
} cb200_sw_inputs;
typedef cb200_lw_outputs cb200_sw_outputs; /* swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc */

int cb200_sw_create(cb200_sw_engine** out, const char* table_blob, const double constants[11], int device);
void cb200_sw_destroy(cb200_sw_engine* e);
/* icld 0..3 (non-McICA accepts cloud fractions 0 or 1 only), iaer 0|6|10, inflag 0|2, iceflag 1|2|3, liqflag 1 */
int cb200_sw_set_options(cb200_sw_engine* e, int icld, int iaer, int inflag, int iceflag, int liqflag);
/* McICA (rrtmg_sw_rad.f90 + mcica_subcol_gen_sw.f90), same meaning as cb200_lw_set_mcica; call before set_options. */
int cb200_sw_set_mcica(cb200_sw_engine* e, int enabled, int irng, int permuteseed);
/* isolvar -1..3, scon [W m-2] (0 = internal), indsolvar[2], bndsolvar[14] (module globals in _rrtmg_sw.pyx:8-19) */
int cb200_sw_set_solar(cb200_sw_engine* e, int isolvar, double scon, const double indsolvar[2], const double bndsolvar[14]);
int cb200_sw_run_device(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                        const cb200_sw_inputs* in, const cb200_sw_outputs* out, void* stream);
int cb200_sw_run_host(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                      const cb200_sw_inputs* in, const cb200_sw_outputs* out);
int cb200_sw_set_host_marshal(cb200_sw_engine* e, int flags); /* as cb200_lw_set_host_marshal */
int cb200_sw_run_host_async(cb200_sw_engine* e, int ncol, int nlay, double adjes, int dyofyr, double solcycfrac,
                            const cb200_sw_inputs* in, const cb200_sw_outputs* out);
int cb200_sw_wait(cb200_sw_engine* e);
void cb200_sw_last_transfer_bytes(cb200_sw_engine* e, double* h2d, double* d2h);
int cb200_sw_check(cb200_sw_engine* e);
const char* cb200_sw_last_error(cb200_sw_engine* e);
int cb200_sw_last_launches(cb200_sw_engine* e);
int cb200_sw_enable_timing(cb200_sw_engine* e, int on);
double cb200_sw_last_unit_kernel_ms(cb200_sw_engine* e);   /* k_sw_transfer */
double cb200_sw_last_taumol_kernel_ms(cb200_sw_engine* e); /* k_sw_taumol */

void rrtmg_sw_set_constants(double* pi, double* grav, double* planck, double* boltz, double* clight, double* avogad,
                            double* alosmt, double* gascon, double* sbcnst, double* secdy);
void rrtmg_sw_ini_wrapper(double* cpdair);
void rrtmg_sw_nomcica_wrapper(int* ncol, int* nlay, int* icld, int* iaer, double* play, double* plev, double* tlay,
                              double* tlev, double* tsfc, double* h2ovmr, double* o3vmr, double* co2vmr,
                              double* ch4vmr, double* n2ovmr, double* o2vmr, double* asdir, double* asdif,
                              double* aldir, double* aldif, double* coszen, double* adjes, int* dyofyr, double* scon,
                              int* isolvar, int* inflgsw, int* iceflgsw, int* liqflgsw, double* cldfr, double* taucld,
                              double* ssacld, double* asmcld, double* fsfcld, double* cicewp, double* cliqwp,
                              double* reice, double* reliq, double* tauaer, double* ssaaer, double* asmaer,
                              double* ecaer, double* swuflx, double* swdflx, double* swhr, double* swuflxc,
                              double* swdflxc, double* swhrc, double* bndsolvar, double* indsolvar, double* solcycfrac);
/* McICA, two-step form (see the longwave pair): (112, ncol, nlay) arrays; tauc / ssac / asmc / fsfc are Fortran (14, ncol, nlay);
 * clear sub-columns get tau 0, ssa 1, asm 0, fsf 0 (mcica_subcol_gen_sw.f90:517-525). */
void mcica_subcol_sw_wrapper(int* iplon, int* ncol, int* nlay, int* icld, int* permuteseed, int* irng, double* play,
                             double* cldfrac, double* ciwp, double* clwp, double* rei, double* rel, double* tauc,
                             double* ssac, double* asmc, double* fsfc, double* cldfmcl, double* ciwpmcl, double* clwpmcl,
                             double* reicmcl, double* relqmcl, double* taucmcl, double* ssacmcl, double* asmcmcl,
                             double* fsfcmcl);
void rrtmg_sw_mcica_wrapper(int* ncol, int* nlay, int* icld, int* iaer, double* play, double* plev, double* tlay,
                            double* tlev, double* tsfc, double* h2ovmr, double* o3vmr, double* co2vmr, double* ch4vmr,
                            double* n2ovmr, double* o2vmr, double* asdir, double* asdif, double* aldir, double* aldif,
                            double* coszen, double* adjes, int* dyofyr, double* scon, int* isolvar, int* inflgsw,
                            int* iceflgsw, int* liqflgsw, double* cldfmcl, double* taucmcl, double* ssacmcl,
                            double* asmcmcl, double* fsfcmcl, double* ciwpmcl, double* clwpmcl, double* reicmcl,
                            double* relqmcl, double* tauaer, double* ssaaer, double* asmaer, double* ecaer,
                            double* swuflx, double* swdflx, double* swhr, double* swuflxc, double* swdflxc, double* swhrc,
                            double* bndsolvar, double* indsolvar, double* solcycfrac);


/* ============================== gray longwave ==============================
 * Replaces the numba kernel `_gray_lw_kernel_np` and the flux-divergence arithmetic of
 * GrayLongwaveRadiation.array_call (climt/_components/radiation.py:65-109,162-190).  Stateless.
 * t (nlay, ncol) [K], tau, p_int (nlay+1, ncol) [-, Pa], t_surf (ncol) -> lw_down, lw_up (nlay+1, ncol) [W m-2],
 * tendency (nlay, ncol) [K s-1].  sigma [W m-2 K-4], g [m s-2], cpd [J kg-1 K-1]. */
int cb200_gray_lw_run_device(int device, int ncol, int nlay, const double* t, const double* p_int, const double* t_surf,
                             const double* tau, double sigma, double g, double cpd, double* lw_down, double* lw_up,
                             double* tendency, void* stream);
int cb200_gray_lw_run_host(int device, int ncol, int nlay, const double* t, const double* p_int, const double* t_surf,
                           const double* tau, double sigma, double g, double cpd, double* lw_down, double* lw_up,
                           double* tendency);

/* ============================== CORK correlated-k longwave / shortwave ==============================
 * Replaces, fused into one engine call, the reference's numba kernels and the numpy glue between them:
 *   _ck_tau_additive_co2_kernel / compute_ck_optical_depth   climt/_components/cork/optics/correlated_k.py:81-117, 378-561
 *   planck_sources_kernel, _lw_transport_kernel / lw_transport climt/_components/cork/lw/kernels.py:9-184
 *   _sw_two_stream_core / sw_two_stream                       climt/_components/cork/sw/kernels.py:18-263
 *   compute_column_amount, compute_heating_rate               climt/_components/cork/common.py:38-80
 *   CorkLongwaveRadiation.array_call  (optics="correlated_k") climt/_components/cork/lw/component.py:208-373
 *   CorkShortwaveRadiation.array_call (optics="correlated_k") climt/_components/cork/sw/component.py:223-496
 * Additive overlap only (ESFT tables are rejected at create).  Arrays are fp64, (nlev[+1], ncol) column-fastest, level 0 at
 * the surface; per-band state arrays keep the reference's state layout (nlev, ncol, nband).
 */
typedef struct cb200_cork_engine cb200_cork_engine;

/* A k-table exactly as the reference's .npz/.nc files hold it (correlated_k.py:9-21); host pointers, read at create. */
typedef struct cb200_cork_table {
  int ngas, nband, ngpt, nT, nP, nX, nC;  /* nX = 0: no H2O axis; nC = 0: no CO2 axis */
  const float* k_coefficients_f32;        /* (ngas, nband, ngpt, nT, nP[, nX[, nC]]) C order, in the table's own dtype: */
  const double* k_coefficients_f64;       /* exactly one of the two is non-NULL (float32 tables stay float32 in HBM) */
  const double* temperature_grid;         /* (nT) */
  const double* pressure_grid_log;        /* (nP) */
  const double* h2o_vmr_grid;             /* (nX) or NULL */
  const double* co2_vmr_grid;             /* (nC) or NULL */
  const double* gpoint_weights;           /* (nband, ngpt) */
  const double* planck_fraction;          /* (nband_pf, ngpt_pf, nT), longwave tables; NULL otherwise */
  int nband_pf, ngpt_pf;
  const double* continuum_kappa;          /* (nband, nT, nP, nX) or NULL */
  const double* solar_source_per_gpoint;  /* (nband, ngpt), shortwave tables; NULL otherwise */
  const double* rayleigh_coefficient;     /* (nband) or NULL */
  int co2_logk;                           /* _CO2_INTERP_LOGK (correlated_k.py:27): 1 = geometric interpolation in CO2 */
  int premixed;                           /* 1: k per kg of air -> the gas amount is the column mass of air (lw/component.py:254-267) */
} cb200_cork_table;

typedef struct cb200_cork_inputs {
  const double *T, *p, *p_int, *T_surf;   /* K, Pa, Pa (nlev+1), K (ncol) */
  const double* q_h2o;                    /* specific humidity (nlev, ncol); NULL when the table has no H2O axis */
  const double* co2_vmr;                  /* (nlev, ncol); NULL when the table has no CO2 axis */
  const double* gas_q;                    /* (ngas, nlev, ncol) mass mixing ratios, non-premixed tables only; else NULL */
  const double* emissivity;               /* LW: (nband, ncol) */
  const double* tau_cloud;                /* (nlev, ncol, nband) or NULL (= 0) */
  const double *zenith, *albedo;          /* SW: (ncol) radians / - */
  const double *ssa_cloud, *g_cloud;      /* SW: (nlev, ncol, nband), required when tau_cloud is given */
  const double *T_irr, *T_int;            /* picket-fence engines: irradiation / internal temperature (ncol), K */
  const double* bond_albedo;              /* picket-fence SW: Bond albedo (ncol) entering T_eff, or NULL (= 0) */
} cb200_cork_inputs;

typedef struct cb200_cork_outputs {
  double *up_broad, *down_broad;          /* (nlev+1, ncol) W m-2 */
  double* heating_rate;                   /* (nlev, ncol) K s-1 */
  double *up_band, *down_band;            /* (nband, nlev+1, ncol) or NULL */
  double *tau_band, *trans_band, *hr_band; /* (nband, nlev, ncol) or NULL; trans_band longwave only; hr_band K day-1 */
} cb200_cork_outputs;

/* g [m s-2], cpd [J kg-1 K-1], sigma [W m-2 K-4] as sympl's get_constant gives them (lw/component.py:224-226) */
/* The same, from the engine's own on-disk container (.cb2k: typed, 64-byte-aligned arrays under the reference's names plus the
 * table classification resolved at conversion time; format and converter: climt_b200/table_store.py, tools/convert_tables.py).
 * Replaces load_k_table + the constructor's classification (cork/optics/correlated_k.py:120-218, cork/lw/component.py:46-59) for
 * a host without numpy / scipy. */
int cb200_cork_create_from_file(cb200_cork_engine** out, const char* path, double g, double cpd, double sigma, int device);
int cb200_cork_create(cb200_cork_engine** out, const cb200_cork_table* table, double g, double cpd, double sigma, int device);

/* Picket-fence optics (optics="parmentier", the reference constructors' default) instead of a k-table.  Replaces
 *   compute_rosseland_mean_opacity, lookup_ratio_coefficients, compute_thermal_opacities   climt/_components/cork/optics/parmentier.py:8-153
 *   CorkLongwaveRadiation._parmentier_optics      climt/_components/cork/lw/component.py:375-422  (2 thermal bands x 1 g-point)
 *   CorkShortwaveRadiation._parmentier_sw_optics  climt/_components/cork/sw/component.py:498-532  (3 visible bands x 1 g-point)
 * -- scalar Python loops over (column, level) in the reference -- followed by the same transport kernels as the k-table engines.
 * The members are the arrays of climt/_data/cork/parmentier/{solar_composition,freedman2014}.npz. */
#define CB200_PICKET_MAX_REGIONS 8
typedef struct cb200_picket_coeffs {
  int nregion;                                         /* len(T_eff_boundaries) - 1, <= CB200_PICKET_MAX_REGIONS */
  double T_eff_boundaries[CB200_PICKET_MAX_REGIONS + 1];
  double log10_gamma_v1_ab[CB200_PICKET_MAX_REGIONS][2], log10_gamma_v2_ab[CB200_PICKET_MAX_REGIONS][2],
      log10_gamma_v3_ab[CB200_PICKET_MAX_REGIONS][2], beta_ab[CB200_PICKET_MAX_REGIONS][2];
  double log10_gamma_P_quad[3];
  double T_boundary, a_hi, b_hi, c_hi, a_lo, b_lo, c_lo;  /* Freedman et al. (2014) fit */
} cb200_picket_coeffs;
/* longwave != 0: 2-band thermal engine (run with cb200_cork_lw_run_*); else 3-band visible engine (cb200_cork_sw_run_*, whose
 * solar_flux argument -- (3, 1) doubles -- is then mandatory).  Inputs T_irr and T_int are required by both. */
int cb200_cork_create_picket(cb200_cork_engine** out, const cb200_picket_coeffs* coeffs, int longwave, double g, double cpd,
                             double sigma, int device);
void cb200_cork_destroy(cb200_cork_engine* e);
const char* cb200_cork_last_error(cb200_cork_engine* e);
int cb200_cork_last_launches(cb200_cork_engine* e);
int cb200_cork_enable_timing(cb200_cork_engine* e, int on);
double cb200_cork_last_unit_kernel_ms(cb200_cork_engine* e);
/* diagnostics_level >= 1 of the reference constructors (cork/lw/component.py:63,189-202,341-358; cork/sw/component.py:455-492;
 * kernels: lw/kernels.py:103-116, sw/kernels.py:299-394): per-band g-point averages (weighted by the g-point weights, divided by
 * their sum) of the layer quantities and of the longwave per-g-point fluxes, plain sums over g of the shortwave interface
 * quantities -- what the components hand back after reducing the kernels' (nband, ngpt, ...) dumps.  Band-major like up_band:
 * layer fields (nband, nlev, ncol), interface fields (nband, nlev+1, ncol).  The pointers are DEVICE pointers for the *_run_device
 * calls and HOST pointers for the *_run_host calls that follow; NULL fields are skipped.  level 0 / NULL switches it off.
 *   longwave  field[0] layer transmittance, [1] weighted upward flux per g-point (interface), [2] downward (interface)
 *   shortwave field[0] Rdif, [1] Tdif, [2] Tnoscat, [3] direct beam flux (interface)                        -- level 1
 *             field[4] Rdir, [5] Tdir, [6] tau_delta, [7] ssa_delta, [8] g_delta, [9] combined albedo (interface) -- level 2 */
typedef struct cb200_cork_diagnostics {
  int level;
  double* field[10];
  const double* weight_sum; /* HOST (nband): the sum of each band's g-point weights as the caller evaluates `weights.sum(axis=1)`
                             * (cork/lw/component.py:342) -- in the table's own dtype, so a float32 table divides by a float32 sum;
                             * NULL = summed by the engine in float64.  Read at the set call. */
} cb200_cork_diagnostics;
int cb200_cork_set_diagnostics(cb200_cork_engine* e, const cb200_cork_diagnostics* d);
/* Device-pointer calls, asynchronous on `stream`.  diffusivity_factor: D in trans = exp(-D tau) (lw/kernels.py:6). */
int cb200_cork_lw_run_device(cb200_cork_engine* e, int ncol, int nlev, double diffusivity_factor, const cb200_cork_inputs* in,
                             const cb200_cork_outputs* out, void* stream);
/* solar_flux: HOST pointer to (nband, ngpt) doubles = solar_source_per_gpoint * earth_sun_factor, evaluated by the caller in the
 * table's own dtype exactly as sw/component.py:371-372 does (a float32 table gives a float32 product); NULL = the table's
 * solar_source_per_gpoint unscaled. */
int cb200_cork_sw_run_device(cb200_cork_engine* e, int ncol, int nlev, const double* solar_flux, const cb200_cork_inputs* in,
                             const cb200_cork_outputs* out, void* stream);
/* Host-pointer calls: chunked 3-stream pipeline (H2D | kernels | D2H); NULL outputs are neither computed nor copied. */
int cb200_cork_lw_run_host(cb200_cork_engine* e, int ncol, int nlev, double diffusivity_factor, const cb200_cork_inputs* in,
                           const cb200_cork_outputs* out);
int cb200_cork_sw_run_host(cb200_cork_engine* e, int ncol, int nlev, const double* solar_flux, const cb200_cork_inputs* in,
                           const cb200_cork_outputs* out);

/* ============================== Emanuel moist convection ==============================
 * Replaces the per-column Fortran routine behind climt.EmanuelConvection and the numba port behind
 * climt.EmanuelConvectionPython, together with the column loop and the saturation-humidity pre-step around them:
 *   SUBROUTINE CONVECT (version 4.3c), TLIFT        climt/_lib/emanuel/convect43c.f90:146-1219
 *   init_emanuel_convection / module parameters     climt/_lib/emanuel/convect43c.f90:91-137
 *   convect() column loop (one Fortran call/column) climt/_components/emanuel/_emanuel_convection.pyx:96-201
 *   EmanuelConvection.array_call, bolton_q_sat      climt/_components/emanuel/component.py:279-340, climt/_core/util.py:177-180
 *   _convect_functional_np, compute_qs              climt/_components/emanuel/pure_python_v3.py:240-835, climt/_core/condensibles.py:104-123
 * IPBL = 0 (the reference's shim hard-wires it) and no tracers (the components pass NTRA = 0).  Parameters are per engine
 * instance (the reference keeps them in Fortran module variables).  Pressures in mbar, level 0 at the surface. */
typedef struct cb200_emanuel_engine cb200_emanuel_engine;

typedef struct cb200_emanuel_params {
  double minorig;                      /* MINORIG: lowest level from which convection may originate (1-based, integral) */
  double elcrit, tlcrit, entp, sigd, sigs, omtrain, omtsnow, coeffr, coeffs, cu, beta, dtmax, alpha, damp;
  double cpd, cpv, cl, rv, rd, lv0, g, rowl, delt0;
  double t_rain;                       /* rain/snow fall-speed switch: 273.0 (convect43c.f90:885) or 273.15 (pure_python_v3.py:586-587) */
} cb200_emanuel_params;

typedef struct cb200_emanuel_inputs {
  const double *t, *q, *u, *v, *p, *ph; /* K, kg/kg, m/s, m/s, mbar (nlev), mbar (nlev+1) */
  const double* qs;                     /* saturation specific humidity; read only when qs_mode = 0 */
  const double* cbmf;                   /* (ncol) cloud-base mass flux of the previous step, kg m-2 s-1 */
} cb200_emanuel_inputs;

typedef struct cb200_emanuel_outputs {
  double *ft, *fq, *fu, *fv;            /* tendencies: K s-1, kg/kg s-1, m s-2, m s-2 */
  double *precip, *wd, *tprime, *qprime; /* (ncol) mm day-1, m s-1, K, kg/kg */
  double* cbmf;                          /* (ncol) updated cloud-base mass flux; may alias inputs.cbmf */
  double* cape;                          /* (ncol) J kg-1; 0 where the routine returns before computing it */
  int* iflag;                            /* (ncol) int32 convective_state: 0 none, 1 convection, 2/3 no LCL / cloud base too high, 4 CFL */
} cb200_emanuel_outputs;

int cb200_emanuel_create(cb200_emanuel_engine** out, const cb200_emanuel_params* params, int device);
void cb200_emanuel_destroy(cb200_emanuel_engine* e);
const char* cb200_emanuel_last_error(cb200_emanuel_engine* e);
int cb200_emanuel_last_launches(cb200_emanuel_engine* e);
int cb200_emanuel_enable_timing(cb200_emanuel_engine* e, int on);
double cb200_emanuel_last_kernel_ms(cb200_emanuel_engine* e);   /* k_emanuel, CUDA events on its launch stream */
/* qs_mode: 0 = inputs.qs, 1 = bolton_q_sat(T, 100 p, rd, rv) (EmanuelConvection), 2 = compute_qs (EmanuelConvectionPython), fused
 * into the kernel.  max_conv_lev: NL, the components pass nlev - 3 (component.py:297).
 * Device-pointer call, asynchronous on `stream`.  layout 0: (nlev[+1], ncol) column-fastest, the layout of the radiation engines;
 * layout 1: (ncol, nlev[+1]) C order, the component's ("*", "mid_levels") arrays (transposed on the device). */
int cb200_emanuel_run_device(cb200_emanuel_engine* e, int ncol, int nlev, int max_conv_lev, double dt, int qs_mode, int layout,
                             const cb200_emanuel_inputs* in, const cb200_emanuel_outputs* out, void* stream);
/* Host-pointer call in the component's (ncol, nlev[+1]) layout: chunked 3-stream pipeline (H2D | kernels | D2H). */
int cb200_emanuel_run_host(cb200_emanuel_engine* e, int ncol, int nlev, int max_conv_lev, double dt, int qs_mode,
                           const cb200_emanuel_inputs* in, const cb200_emanuel_outputs* out);
/* The reference's own symbols (bind(c) names of convect43c.f90:91-150; declared in _emanuel_convection.pyx:10-42): process-global
 * parameters, ONE column per call, host pointers. */
void init_emanuel_convection_fortran(int* pbl, int* least_conv_level, double* thresh_water_level, double* crit_temp,
                                     double* entrain_coeff, double* downdraft_frac_area, double* precip_frac_outside_cloud,
                                     double* rain_speed, double* snow_speed, double* rain_evap_coeff, double* snow_evap_coeff,
                                     double* mom_tran_coeff, double* max_neg_temp_pert, double* beta, double* alpha, double* damp_amp,
                                     double* Cpd, double* Cpv, double* Cl, double* gas_const_vapour, double* gas_const_air,
                                     double* lat_heat, double* grav, double* density_water, double* reference_mass_flux_timescale);
void emanuel_convection(double* temp, double* q, double* qs, double* u, double* v, double* pmid, double* pint, int* nlevs,
                        int* max_conv_lev, int* num_tracers, double* dt, int* conv_state, double* dtemp, double* dq, double* du,
                        double* dv, double* precip, double* downdraft_vel_scale, double* downdraft_temp_scale,
                        double* downdraft_q_scale, double* cloud_base_mass_flux, double* cape, double* tracers, double* dtracers);

/* ============================== device-side marshal ==============================
 * What the components' array_call computes in numpy before calling the engines, for state that already lives in HBM:
 * h2ovmr = q * 28.964 / 18.02 (climt/_core/util.py:47-86), tlev by ln-p weights with tlev[0] = tsfc, tlev[nlay] = t[nlay-1]
 * (climt/_core/util.py:89-142), coszen = cos(zenith) (rrtmg/sw/component.py:591).  Any output pointer may be NULL. */
int cb200_marshal_device(int device, int ncol, int nlay, const double* q, const double* t, const double* tsfc, const double* p,
                         const double* p_int, const double* zenith, double* h2ovmr, double* tlev, double* coszen, void* stream);

/* ============================== column steps either side of the radiation call ==============================
 * Instellation: replaces `_instellation_kernel_np` and its scalar helpers (climt/_components/instellation/component.py:84-191).
 * julian_centuries = days since 2000-01-01 12:00 / 36525 (component.py:53, 66-76; the `fractional_day` argument of the reference
 * kernel is unused by it).  lat_deg, lon_deg (ncol) -> zenith (ncol) [rad, clamped to pi/2]; coszen (device call only, may be NULL)
 * = cos(zenith), the shortwave engine's input.  cb200_instellation_orbit is the per-call scalar part (host arithmetic, no GPU). */
void cb200_instellation_orbit(double julian_centuries, double* sin_dec, double* cos_dec, double* right_ascension, double* gmst);
int cb200_instellation_run_device(int device, int ncol, const double* lat_deg, const double* lon_deg, double julian_centuries,
                                  double* zenith, double* coszen, void* stream);
int cb200_instellation_run_host(int device, int ncol, const double* lat_deg, const double* lon_deg, double julian_centuries,
                                double* zenith);

/* BergerSolarInsolation: replaces the numba kernel `_get_solar_parameters_np` (climt/_components/berger_solar_insolation.py:635-680).
 * The four orbital parameters of the year come from the caller (the reference evaluates Berger's series in plain numpy, once per
 * year: :579-625); lat / lon as the component passes them (the reference takes sin / cos of the latitude in degrees, :673).
 * -> insolation [W m-2], zenith [rad] (ncol); *rho = normalised earth-sun distance.  cb200_berger_scalars: the per-call scalars
 * (sin / cos of the declination, 1 / rho^2, rho), host arithmetic. */
void cb200_berger_scalars(double lambda_m0, double eccentricity, double omega_tilde, double obliquity, double years_since_vernal_equinox,
                          double* out4);
int cb200_berger_run_device(int device, int ncol, const double* lat, const double* lon, double lambda_m0, double eccentricity,
                            double omega_tilde, double obliquity, double years_since_vernal_equinox, double fractional_day,
                            double solar_constant, double* insolation, double* zenith, double* rho, void* stream);
int cb200_berger_run_host(int device, int ncol, const double* lat, const double* lon, double lambda_m0, double eccentricity,
                          double omega_tilde, double obliquity, double years_since_vernal_equinox, double fractional_day,
                          double solar_constant, double* insolation, double* zenith, double* rho);

/* SlabSurface: replaces `_slab_surface_kernel_np` (climt/_components/slab_surface.py:449-517), include_ekman=False.
 * Every array has ncol entries except the four flux arrays, whose surface value for column i is element i * flux_stride
 * (component layout ("*", "interface_levels"): flux_stride = nlev + 1; the engines' (nlev + 1, ncol) outputs: flux_stride = 1).
 * area_type: int32 codes of AREA_MAP (slab_surface.py:7): 0 land, 1 land_ice, 2 sea, 3 sea_ice. */
typedef struct cb200_slab_inputs {
  const double *sw_down, *lw_down, *sw_up, *lw_up, *lh, *sh;
  const int* area_type;
  const double *up_heat_soil, *heat_flux_sea_ice, *sea_water_dens, *surf_dens, *heat_cap_soil, *surf_therm_cap, *ocean_mix_thick,
      *soil_layer_thick, *ocean_heat_transport;
} cb200_slab_inputs;
int cb200_slab_surface_run_device(int device, int ncol, long flux_stride, const cb200_slab_inputs* in, double* tend_ts,
                                  double* depth, void* stream);
int cb200_slab_surface_run_host(int device, int ncol, long flux_stride, const cb200_slab_inputs* in, double* tend_ts, double* depth);

/* SimplePhysics: replaces the Reed-Jablonowski package (climt/_lib/simple_physics/simple_physics_custom.f90:28-565) and the level
 * flip / layer thickness of its Cython shim (climt/_components/simple_physics/_simple_physics.pyx:84-180).  A sympl Stepper: the new
 * T, q, u, v after large-scale condensation, bulk surface fluxes and implicit boundary-layer diffusion over dtime, plus three
 * diagnostics.  Arrays (nlev[+1], ncol) column-fastest; order 0: level 0 is the surface (the component's arrays as they are),
 * order 1: level 0 is the model top (the Fortran's own order).  pressures Pa, latitude as the component passes it (degrees).
 * ts / qsurf / lat may be NULL when the switches make them dead.  workspace: 2 * nlev * ncol doubles (boundary layer only). */
typedef struct cb200_simple_physics_params {
  double gravit, cpair, rair, latvap, rh2o, radius, omega, rhow;  /* set_fortran_constants (simple_physics_custom.f90:28-57) */
  double pbltop, pblconst, C, Cd0, Cd1, Cm;
  int test;                    /* the component's simulate_cyclone flag (_simple_physics.pyx:168): 1 = baroclinic-wave SST */
  int do_lsc, do_pbl, do_surf_flux, use_ts_ext, use_qsurf_ext;
  int clamp_latent_heat_flux;  /* 1: negative latent heat fluxes are reported as 0 (simple_physics/component.py:257) */
} cb200_simple_physics_params;
typedef struct cb200_simple_physics_inputs {
  const double *t, *q, *u, *v, *pmid, *pint, *ps, *ts, *qsurf, *lat;
} cb200_simple_physics_inputs;
typedef struct cb200_simple_physics_outputs {
  double *t, *q, *u, *v, *precl, *sens_ht_flux, *lat_ht_flux;   /* precl [m s-1], fluxes [W m-2]: (ncol) */
} cb200_simple_physics_outputs;
int cb200_simple_physics_run_device(int device, int ncol, int nlev, int order, double dtime, const cb200_simple_physics_params* p,
                                    const cb200_simple_physics_inputs* in, const cb200_simple_physics_outputs* out,
                                    double* workspace, void* stream);
int cb200_simple_physics_run_host(int device, int ncol, int nlev, int order, double dtime, const cb200_simple_physics_params* p,
                                  const cb200_simple_physics_inputs* in, const cb200_simple_physics_outputs* out);
/* The reference's own symbols (declared in _simple_physics.pyx:6-27): process-global constants, host pointers, arrays
 * (pver, pcols) with the model top first, t / q / u / v updated in place. */
void set_fortran_constants(double* g, double* cpd, double* r_air, double* latent_heat, double* r_cond, double* radius,
                           double* rotation, double* density_cond, double* top_pbl, double* pbl_decay, double* drag_coeff_sens_lat,
                           double* Cd0_ext, double* Cd1_ext, double* Cm_ext);
void simple_physics(int* pcols, int* pver, double* dtime, double* lat, double* t, double* q, double* u, double* v, double* pmid,
                    double* pint, double* pdel, double* rpdel, double* ps, double* precl, int* test, int* do_lsc, int* do_pbl,
                    int* do_surf_flux, int* use_ts_ext, double* ts, int* use_qsurf_ext, double* qsurf, double* sens_ht_flux,
                    double* lat_ht_flux);

#ifdef __cplusplus
}
#endif
#endif
