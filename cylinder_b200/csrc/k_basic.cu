// Library plumbing + the small O(B) / O(n) entry points: geometry evaluation, Philox known-answer hook,
// lattice initialisation, bare surface energy.
#include <stdarg.h>
#include <string.h>

#include "cyl_common.cuh"
#include "cyl_surface.cuh"

// ----------------------------------------------------------------------------------------------------
// error text (thread-local)
// ----------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void cyl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int cyl_abi_version(void) { return CYL_ABI_VERSION; }
extern "C" const char* cyl_last_error(void) { return g_err; }

extern "C" int cyl_set_device(int device) {
  CYL_CUDA(cudaSetDevice(device));
  return CYL_OK;
}

extern "C" int cyl_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin) {
  int dev = 0;
  CYL_CUDA(cudaGetDevice(&dev));
  if (sm_count) CYL_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) CYL_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) CYL_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (max_smem_optin) CYL_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return CYL_OK;
}

// ----------------------------------------------------------------------------------------------------
// geometry evaluation   system_cylinder.py:25-41
// ----------------------------------------------------------------------------------------------------
__global__ void geometry_eval_kernel(double k, double r, const double* __restrict__ a, const double* __restrict__ z,
                                     int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Geo g = geo_eval(k, r, a[i], z[i]);
  double* o = out + 6 * i;
  o[0] = geo_radius_rescaled(r, a[i]);
  o[1] = g.gth;
  o[2] = g.gth * g.gth;          // g_theta :25
  o[3] = g.gz;
  o[4] = g.gz * g.gz;            // g_z :31 (= 1 + t^2 up to one rounding)
  o[5] = g.ath;
}

extern "C" int cyl_geometry_eval_f64(double k, double r, const double* a, const double* z, int64_t n, double* out,
                                     void* stream) {
  CYL_CHECK_ARG(n >= 0 && a && z && out, "cyl_geometry_eval_f64: null buffer or n<0");
  if (n == 0) return CYL_OK;
  geometry_eval_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(k, r, a, z, n, out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

// ----------------------------------------------------------------------------------------------------
// Philox known-answer hook
// ----------------------------------------------------------------------------------------------------
__global__ void philox_kat_kernel(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, int64_t n,
                                  uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = philox4x32<10>(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" int cyl_philox4x32_10(const uint32_t* ctr, const uint32_t* key, int64_t n, uint32_t* out, void* stream) {
  CYL_CHECK_ARG(n >= 0 && ctr && key && out, "cyl_philox4x32_10: null buffer or n<0");
  if (n == 0) return CYL_OK;
  philox_kat_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(ctr, key, n, out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

__global__ void philox2_kat_kernel(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, int64_t n,
                                   uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 r = philox2x32_10(ctr[2 * i], ctr[2 * i + 1], key[i]);
  out[2 * i] = r.x; out[2 * i + 1] = r.y;
}

extern "C" int cyl_philox2x32_10(const uint32_t* ctr, const uint32_t* key, int64_t n, uint32_t* out, void* stream) {
  CYL_CHECK_ARG(n >= 0 && ctr && key && out, "cyl_philox2x32_10: null buffer or n<0");
  if (n == 0) return CYL_OK;
  philox2_kat_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(ctr, key, n, out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

// ----------------------------------------------------------------------------------------------------
// Lattice.random_initialize  On_lattice_simple.py:159-182 : psi = rect(U(0,mag_max), U(0,2pi))
// counter = (site, replica, 0, INIT tag); x0 -> magnitude, x1 -> phase
// ----------------------------------------------------------------------------------------------------
template <typename R>
__global__ void lattice_init_kernel(typename Real<R>::complex_t* __restrict__ psi, int B, const int32_t* __restrict__ Z,
                                    int Zmax, int T, double mag_max, uint64_t seed, int64_t replica0) {
  // the batch index rides in grid.x (grid.y is limited to 65535; config C5 on one GPU has 65536 replicas)
  const int per = (Zmax * T + (int)blockDim.x - 1) / (int)blockDim.x;
  const int b = blockIdx.x / per;
  const int site = (blockIdx.x - b * per) * blockDim.x + threadIdx.x;
  if (site >= Zmax * T) return;
  const int i = site / T;
  typename Real<R>::complex_t v = Real<R>::make(R(0), R(0));
  if (i < Z[b]) {
    const uint64_t rep = (uint64_t)(replica0 + b);
    const uint4 x = philox4x32<10>((uint32_t)site, (uint32_t)rep, 0u, philox_c3(CYL_TAG_INIT, 0) ^ (uint32_t)(rep >> 32),
                                   (uint32_t)seed, (uint32_t)(seed >> 32));
    const double mag = mag_max * u01_f64(x.x);
    double s, c;
    sincospi(2.0 * u01_f64(x.y), &s, &c);
    v = Real<R>::make((R)(mag * c), (R)(mag * s));
  }
  psi[((size_t)b * Zmax) * T + site] = v;
}

template <typename R>
static int lattice_init(void* psi, int B, const int32_t* Z, int Zmax, int T, double mag_max, uint64_t seed,
                        int64_t replica0, void* stream) {
  CYL_CHECK_ARG(psi && Z && B > 0 && Zmax > 0 && T > 0, "cyl_lattice_init: bad arguments");
  const long long nblk = (long long)((Zmax * T + 255) / 256) * B;
  CYL_CHECK_ARG(nblk <= 2147483647LL, "cyl_lattice_init: batch too large for one launch");
  lattice_init_kernel<R><<<(unsigned)nblk, 256, 0, as_stream(stream)>>>((typename Real<R>::complex_t*)psi, B, Z, Zmax, T, mag_max,
                                                             seed, replica0);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

extern "C" int cyl_lattice_init_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, double mag_max, uint64_t seed,
                                    int64_t replica0, void* stream) {
  return lattice_init<float>(psi, B, Z, Zmax, T, mag_max, seed, replica0, stream);
}
extern "C" int cyl_lattice_init_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, double mag_max, uint64_t seed,
                                    int64_t replica0, void* stream) {
  return lattice_init<double>(psi, B, Z, Zmax, T, mag_max, seed, replica0, stream);
}

// ----------------------------------------------------------------------------------------------------
// Cylinder.calc_surface_energy   system_cylinder.py:119-124  -- one warp per replica
// ----------------------------------------------------------------------------------------------------
__global__ void surface_energy_kernel(const CylParams* __restrict__ params, const double* __restrict__ amp, int B,
                                      double* __restrict__ E_out, double* __restrict__ terms) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const CylParams p = params[b];
  double area, bend;
  const double e = surface_energy_warp(p, amp[b], &area, &bend);
  if ((threadIdx.x & 31) == 0) {
    E_out[b] = e;
    if (terms) { terms[2 * b] = area; terms[2 * b + 1] = bend; }
  }
}

extern "C" int cyl_surface_energy_f64(const CylParams* params, const double* amp, int B, double* E_out, double* out_terms,
                                      void* stream) {
  CYL_CHECK_ARG(params && amp && E_out && B > 0, "cyl_surface_energy_f64: bad arguments");
  const int warps = 4;
  surface_energy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, as_stream(stream)>>>(params, amp, B, E_out, out_terms);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
