// climt_b200 -- the small column steps either side of the radiation call, sm_100a.
//   BergerSolarInsolation (zenith angle and insolation from Berger 1978 orbital parameters)
//     _get_solar_parameters_np            climt/_components/berger_solar_insolation.py:635-680
//   Instellation  (produces the zenith angle the shortwave engine consumes)
//     _instellation_kernel_np, _obliquity_star_jit, _sun_ecliptic_longitude_jit, _gmst_jit
//                                         climt/_components/instellation/component.py:84-191
//   SlabSurface   (consumes the surface row of the four flux fields the engines produce)
//     _slab_surface_kernel_np             climt/_components/slab_surface.py:449-517
// Both are one thread per column, columns fastest, every input read once and every output written once: pure
// HBM streaming (24 B per column for the zenith angle, 144 B per column for the slab).  The orbital scalars are the
// same for every column of a call and are evaluated once on the host (cb200_instellation_orbit).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/climt_b200.h"
#include "engine_common.h"

namespace {
constexpr double kPi = 3.141592653589793;
constexpr double kDegToRad = kPi / 180.0;  // np.deg2rad(x) = x * (pi / 180)

__global__ void __launch_bounds__(128) k_instellation(int ncol, const double* __restrict__ lat_deg,
                                                      const double* __restrict__ lon_deg, double sin_dec, double cos_dec,
                                                      double right_ascension, double gmst, double* __restrict__ zenith,
                                                      double* __restrict__ coszen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const double lat = lat_deg[i] * kDegToRad;
  const double lmst = gmst + lon_deg[i] * kDegToRad;
  const double h_angle = lmst - right_ascension;
  double cos_mu = sin(lat) * sin_dec + cos(lat) * cos_dec * cos(h_angle);
  if (cos_mu > 1.0) cos_mu = 1.0;
  else if (cos_mu < -1.0) cos_mu = -1.0;
  double z = acos(cos_mu);
  if (z > kPi / 2.0) z = kPi / 2.0;  // night side: the reference clamps the angle, component.py:124-128
  zenith[i] = z;
  if (coszen) coszen[i] = cos(z);  // what RRTMGShortwave.array_call evaluates from it (rrtmg/sw/component.py:591)
}

// BergerSolarInsolation: the per-column part of _get_solar_parameters_np (climt/_components/berger_solar_insolation.py:669-679).
// The reference takes sin / cos of the latitude as it arrives, in degrees (:673); kept.
__global__ void __launch_bounds__(128) k_berger(int ncol, const double* __restrict__ lat, const double* __restrict__ lon,
                                                double fractional_day, double sin_delta, double cos_delta, double scale,
                                                double* __restrict__ insolation, double* __restrict__ zenith) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const double H = 2 * kPi * (fractional_day + lon[i] / 360.0);
  const double cos_mu = sin(lat[i]) * sin_delta - cos(lat[i]) * cos_delta * cos(H);
  zenith[i] = acos(cos_mu);
  insolation[i] = scale * cos_mu;
}

__global__ void __launch_bounds__(128) k_slab_surface(int ncol, long flux_stride, const cb200_slab_inputs in,
                                                      double* __restrict__ tend_ts, double* __restrict__ depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const size_t fi = (size_t)i * (size_t)flux_stride;
  double net = in.sw_down[fi] + in.lw_down[fi] - in.sw_up[fi] - in.lw_up[fi] - in.sh[i] - in.lh[i];
  const int at = in.area_type[i];  // 0 land, 1 land_ice, 2 sea, 3 sea_ice (AREA_MAP, slab_surface.py:7)
  const bool land = at == 0 || at == 1, sea = at == 2 || at == 3, land_ice = at == 1, sea_ice = at == 3;
  if (land_ice) net = -in.up_heat_soil[i];
  else if (sea_ice) net = in.heat_flux_sea_ice[i];
  if (sea && !sea_ice) net = net + in.ocean_heat_transport[i];
  const double dens = sea ? in.sea_water_dens[i] : in.surf_dens[i];
  const double d = sea ? in.ocean_mix_thick[i] : (land ? in.soil_layer_thick[i] : 0.0);
  const double cap = land ? in.heat_cap_soil[i] : in.surf_therm_cap[i];
  depth[i] = d;
  const double heat_cap_slab = (dens * d) * cap;
  double val = heat_cap_slab != 0.0 ? net / heat_cap_slab : 0.0;
  if (land_ice || sea_ice) val = 0.0;
  tend_ts[i] = val;
}

int fail(cudaError_t e) {
  cb::set_global_error(cudaGetErrorString(e));
  return -1;
}
}  // namespace

// Orbital scalars of one call (host arithmetic, fp64, the reference's order of operations).
extern "C" void cb200_instellation_orbit(double t, double* sin_dec, double* cos_dec, double* right_ascension, double* gmst) {
  const double t2 = t * t, t3 = t2 * t;
  const double eps = (23.0 + 26.0 / 60 + 21.406 / 3600.0 -
                      (46.836769 * t - 0.0001831 * t2 + 0.00200340 * t3 - 0.576e-6 * (t2 * t2) - 4.34e-8 * (t2 * t2 * t)) / 3600.0) *
                     kDegToRad;
  const double mean_anomaly = (357.52910 + 35999.05030 * t - 0.0001559 * t * t - 0.00000048 * t * t * t) * kDegToRad;
  const double mean_longitude = (280.46645 + 36000.76983 * t + 0.0003032 * t2) * kDegToRad;
  const double d_l = ((1.914600 - 0.004817 * t - 0.000014 * t2) * sin(mean_anomaly) +
                      (0.019993 - 0.000101 * t) * sin(2 * mean_anomaly) + 0.000290 * sin(3 * mean_anomaly)) *
                     kDegToRad;
  const double eclon = mean_longitude + d_l;
  const double x = cos(eclon), y = cos(eps) * sin(eclon), z = sin(eps) * sin(eclon);
  const double r = sqrt(1.0 - z * z);
  const double declination = atan2(z, r);
  *right_ascension = 2.0 * atan2(y, x + r);
  *sin_dec = sin(declination);
  *cos_dec = cos(declination);
  // `6.2 * 10e-6` in the reference is 6.2e-5 (component.py:183); kept
  const double theta = 67310.54841 + t * (876600.0 * 3600 + 8640184.812866 + t * (0.093104 - t * 6.2 * 10e-6));
  double th = fmod(theta / 240.0 * kDegToRad, 2.0 * kPi);
  if (th < 0) th += 2.0 * kPi;
  *gmst = th;
}

extern "C" int cb200_instellation_run_device(int device, int ncol, const double* lat_deg, const double* lon_deg,
                                             double julian_centuries, double* zenith, double* coszen, void* stream) {
  if (ncol <= 0) { cb::set_global_error("instellation: bad ncol"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  double sd, cd, ra, gm;
  cb200_instellation_orbit(julian_centuries, &sd, &cd, &ra, &gm);
  k_instellation<<<(ncol + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ncol, lat_deg, lon_deg, sd, cd, ra, gm, zenith, coszen);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(e);
}

extern "C" int cb200_instellation_run_host(int device, int ncol, const double* lat_deg, const double* lon_deg,
                                           double julian_centuries, double* zenith) {
  if (ncol <= 0) { cb::set_global_error("instellation: bad ncol"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  const size_t n = (size_t)ncol;
  double* d = nullptr;
  if ((e = cudaMalloc(&d, 3 * n * sizeof(double))) != cudaSuccess) return fail(e);
  cudaMemcpyAsync(d, lat_deg, n * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(d + n, lon_deg, n * 8, cudaMemcpyHostToDevice, 0);
  int rc = cb200_instellation_run_device(device, ncol, d, d + n, julian_centuries, d + 2 * n, nullptr, nullptr);
  if (rc == 0) {
    cudaMemcpyAsync(zenith, d + 2 * n, n * 8, cudaMemcpyDeviceToHost, 0);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) rc = fail(e);
  }
  cudaFree(d);
  return rc;
}

// Per-call scalars of _get_solar_parameters_np (:651-667), host arithmetic in the reference's order of operations.
// out: sin_delta, cos_delta, inverse_rho_squared, rho
extern "C" void cb200_berger_scalars(double lambda_m0, double eccentricity, double omega_tilde, double obliquity,
                                     double years_since_vernal_equinox, double* out) {
  const double e = eccentricity, e2 = e * e;
  const double lambda_m = lambda_m0 + years_since_vernal_equinox * 2.0 * kPi;
  const double temp = lambda_m - (omega_tilde + kPi);
  const double sin_temp = sin(temp);
  const double lmbda = lambda_m + e * (2.0 * sin_temp + e * (1.25 * sin(2 * temp) + e * ((13.0 / 12.0) * sin(3 * temp) - 0.25 * sin_temp)));
  const double inverse_rho = (1 + e * cos(lmbda - (omega_tilde + kPi))) / (1 - e2);
  const double decl = asin(sin(obliquity) * sin(lmbda));
  out[0] = sin(decl);
  out[1] = cos(decl);
  out[2] = inverse_rho * inverse_rho;
  out[3] = 1.0 / inverse_rho;
}

extern "C" int cb200_berger_run_device(int device, int ncol, const double* lat, const double* lon, double lambda_m0,
                                       double eccentricity, double omega_tilde, double obliquity,
                                       double years_since_vernal_equinox, double fractional_day, double solar_constant,
                                       double* insolation, double* zenith, double* rho, void* stream) {
  if (ncol <= 0) { cb::set_global_error("berger: bad ncol"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  double sc[4];
  cb200_berger_scalars(lambda_m0, eccentricity, omega_tilde, obliquity, years_since_vernal_equinox, sc);
  if (rho) *rho = sc[3];
  k_berger<<<(ncol + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ncol, lat, lon, fractional_day, sc[0], sc[1], solar_constant * sc[2],
                                                                 insolation, zenith);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(e);
}

extern "C" int cb200_berger_run_host(int device, int ncol, const double* lat, const double* lon, double lambda_m0,
                                     double eccentricity, double omega_tilde, double obliquity, double years_since_vernal_equinox,
                                     double fractional_day, double solar_constant, double* insolation, double* zenith, double* rho) {
  if (ncol <= 0) { cb::set_global_error("berger: bad ncol"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  const size_t n = (size_t)ncol;
  double* d = nullptr;
  if ((e = cudaMalloc(&d, 4 * n * sizeof(double))) != cudaSuccess) return fail(e);
  cudaMemcpyAsync(d, lat, n * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(d + n, lon, n * 8, cudaMemcpyHostToDevice, 0);
  int rc = cb200_berger_run_device(device, ncol, d, d + n, lambda_m0, eccentricity, omega_tilde, obliquity, years_since_vernal_equinox,
                                   fractional_day, solar_constant, d + 2 * n, d + 3 * n, rho, nullptr);
  if (rc == 0) {
    cudaMemcpyAsync(insolation, d + 2 * n, n * 8, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(zenith, d + 3 * n, n * 8, cudaMemcpyDeviceToHost, 0);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) rc = fail(e);
  }
  cudaFree(d);
  return rc;
}

extern "C" int cb200_slab_surface_run_device(int device, int ncol, long flux_stride, const cb200_slab_inputs* in,
                                             double* tend_ts, double* depth, void* stream) {
  if (ncol <= 0 || flux_stride <= 0) { cb::set_global_error("slab surface: bad ncol / flux_stride"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  k_slab_surface<<<(ncol + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ncol, flux_stride, *in, tend_ts, depth);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(e);
}

// Host pointers; the four flux arrays are the component's ("*", "interface_levels") arrays: the surface value of column i is
// element i * flux_stride (flux_stride = nlev + 1), gathered by a strided copy -- only 8 of every 8 (nlev + 1) bytes cross PCIe.
extern "C" int cb200_slab_surface_run_host(int device, int ncol, long flux_stride, const cb200_slab_inputs* in, double* tend_ts,
                                           double* depth) {
  if (ncol <= 0 || flux_stride <= 0) { cb::set_global_error("slab surface: bad ncol / flux_stride"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(e);
  const size_t n = (size_t)ncol;
  double* d = nullptr;
  if ((e = cudaMalloc(&d, 18 * n * sizeof(double))) != cudaSuccess) return fail(e);
  cb200_slab_inputs di;
  const double* flux[4] = {in->sw_down, in->lw_down, in->sw_up, in->lw_up};
  const double** dflux[4] = {&di.sw_down, &di.lw_down, &di.sw_up, &di.lw_up};
  for (int k = 0; k < 4; ++k) {
    cudaMemcpy2DAsync(d + k * n, 8, flux[k], (size_t)flux_stride * 8, 8, n, cudaMemcpyHostToDevice, 0);
    *dflux[k] = d + k * n;
  }
  const double* vec[11] = {in->lh, in->sh, in->up_heat_soil, in->heat_flux_sea_ice, in->sea_water_dens, in->surf_dens,
                           in->heat_cap_soil, in->surf_therm_cap, in->ocean_mix_thick, in->soil_layer_thick, in->ocean_heat_transport};
  const double** dvec[11] = {&di.lh, &di.sh, &di.up_heat_soil, &di.heat_flux_sea_ice, &di.sea_water_dens, &di.surf_dens,
                             &di.heat_cap_soil, &di.surf_therm_cap, &di.ocean_mix_thick, &di.soil_layer_thick, &di.ocean_heat_transport};
  for (int k = 0; k < 11; ++k) {
    cudaMemcpyAsync(d + (4 + k) * n, vec[k], n * 8, cudaMemcpyHostToDevice, 0);
    *dvec[k] = d + (4 + k) * n;
  }
  int* dat = reinterpret_cast<int*>(d + 15 * n);
  cudaMemcpyAsync(dat, in->area_type, n * sizeof(int), cudaMemcpyHostToDevice, 0);
  di.area_type = dat;
  int rc = cb200_slab_surface_run_device(device, ncol, 1, &di, d + 16 * n, d + 17 * n, nullptr);
  if (rc == 0) {
    cudaMemcpyAsync(tend_ts, d + 16 * n, n * 8, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(depth, d + 17 * n, n * 8, cudaMemcpyDeviceToHost, 0);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) rc = fail(e);
  }
  cudaFree(d);
  return rc;
}
