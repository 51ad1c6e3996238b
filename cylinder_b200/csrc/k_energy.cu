// Row moments (one warp per z-row, shuffle reduction, fp64 accumulation) and the field energy assembled
// from them (one warp per replica).  Reference: Lattice.surface_field_energy, On_lattice_simple.py:185-213 and
// On_lattice_rescaled_0thorder.py:247-278; Lattice.measure :121-131.
#include "cyl_common.cuh"
#include "cyl_energy.cuh"

// psi is addressed as base + b*batch_stride + row*pitch + col, with periodic wrap done on (row, col); the packed
// ABI layout has pitch = T, batch_stride = Zmax*T.
template <typename R>
__global__ void row_moments_kernel(const typename Real<R>::complex_t* __restrict__ psi, int B, const int32_t* __restrict__ Zs,
                                   int Zmax, int T, int64_t pitch, int64_t batch_stride,
                                   const CylParams* __restrict__ params, double* __restrict__ S) {
  using C = typename Real<R>::complex_t;
  __shared__ double red[CYL_MOM_WARPS][CYL_NMOM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpr = moments_warps_per_row(T), rows_per_cta = CYL_MOM_WARPS / wpr;
  // the batch index rides in grid.x (grid.y is limited to 65535)
  const int per = (Zmax + rows_per_cta - 1) / rows_per_cta;
  const int b = blockIdx.x / per;
  const int row_in_cta = warp / wpr, seg = warp - row_in_cta * wpr;
  const int i = (blockIdx.x - b * per) * rows_per_cta + row_in_cta;
  const int Z = Zs[b];
  double* out = (i < Zmax) ? S + ((size_t)b * Zmax + i) * CYL_NMOM : nullptr;
  double m[CYL_NMOM];
#pragma unroll
  for (int q = 0; q < CYL_NMOM; ++q) m[q] = 0.0;
  if (i < Z) {
    const CylParams p = params[b];
    const double pi = 3.14159265358979323846;
    const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;
    const double inv_lz2 = 1.0 / (lz * lz), inv_lth2 = 1.0 / (lth * lth), inv_lth = 1.0 / lth;
    const C* row = psi + (size_t)b * batch_stride + (size_t)i * pitch;
    const C* rowm = psi + (size_t)b * batch_stride + (size_t)(i == 0 ? Z - 1 : i - 1) * pitch;
#pragma unroll 4
    for (int j = seg * 32 + lane; j < T; j += wpr * 32) {
      const C pv = row[j], L = rowm[j], W = row[j == 0 ? T - 1 : j - 1];
      moments_accumulate(m, pv, L, W, inv_lz2, inv_lth2, inv_lth);
    }
  }
  moments_store(m, wpr, row_in_cta, seg, out, red);          // rows in [Z, Zmax) get zeros
}

template <typename R>
static int row_moments(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params, double* S,
                       void* stream) {
  CYL_CHECK_ARG(psi && Z && params && S && B > 0 && Zmax > 0 && T > 0, "cyl_row_moments: bad arguments");
  const int rows_per_cta = CYL_MOM_WARPS / moments_warps_per_row(T);
  const long long nblk = (long long)((Zmax + rows_per_cta - 1) / rows_per_cta) * B;
  CYL_CHECK_ARG(nblk <= 2147483647LL, "cyl_row_moments: batch too large for one launch");
  row_moments_kernel<R><<<(unsigned)nblk, CYL_MOM_WARPS * 32, 0, as_stream(stream)>>>((const typename Real<R>::complex_t*)psi, B, Z, Zmax, T,
                                                                            (int64_t)T, (int64_t)Zmax * T, params, S);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

extern "C" int cyl_row_moments_f32(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                   double* S, void* stream) {
  return row_moments<float>(psi, B, Z, Zmax, T, params, S, stream);
}
extern "C" int cyl_row_moments_f64(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                   double* S, void* stream) {
  return row_moments<double>(psi, B, Z, Zmax, T, params, S, stream);
}

// one warp per replica: lanes stride over rows
__global__ void field_energy_kernel(const double* __restrict__ S, int B, const int32_t* __restrict__ Zs, int Zmax, int T,
                                    const CylParams* __restrict__ params, const double* __restrict__ amp_eval,
                                    const double* __restrict__ amp_ref, int variant, double* __restrict__ E_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const CylParams p = params[b];
  const int Z = Zs[b];
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;
  const double ae = amp_eval[b], ar = amp_ref ? amp_ref[b] : ae;
  double e = 0.0;
  for (int i = lane; i < Z; i += 32) e += field_energy_row_eval(variant, p, lz, lth, Z, T, i, ae, ar, S + ((size_t)b * Zmax + i) * CYL_NMOM);
  e = warp_sum(e);
  if (lane == 0) E_out[b] = e * lz * lth;                       // :213
}

extern "C" int cyl_field_energy_f64(const double* S, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                    const double* amp_eval, const double* amp_ref, int variant, double* E_out,
                                    void* stream) {
  CYL_CHECK_ARG(S && Z && params && amp_eval && E_out && B > 0 && Zmax > 0 && T > 0, "cyl_field_energy_f64: bad arguments");
  CYL_CHECK_ARG(variant >= 0 && variant < CYL_NUM_VARIANTS, "cyl_field_energy_f64: unknown variant %d", variant);
  const int warps = 4;
  field_energy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, as_stream(stream)>>>(S, B, Z, Zmax, T, params, amp_eval,
                                                                                    amp_ref, variant, E_out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
