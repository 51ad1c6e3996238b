// Bare surface energy of the sinusoidally perturbed cylinder -- Cylinder.calc_surface_energy,
// system_cylinder.py:91-124 -- without SciPy: complete elliptic integral E(m) by the arithmetic-geometric
// mean, the two bending integrals by a 256-node periodic trapezoid rule (spectrally accurate for the
// analytic periodic integrands; SURVEY.md A.4 bounds the error at <= 6e-16 relative for |a| <= .99).
#pragma once
#include "cyl_common.cuh"

#define CYL_QUAD_NODES 256

// scipy.special.ellipe(m): complete elliptic integral of the 2nd kind with PARAMETER m (m <= 0 here).
// E(m) = pi/(2 a_inf) * (1 - sum_{n>=0} 2^(n-1) c_n^2),  a0 = 1, b0 = sqrt(1-m), c0^2 = m.
__host__ __device__ __forceinline__ double ellipe_agm(double m) {
  double a0 = 1.0, b0 = sqrt(1.0 - m);
  double c2sum = 0.5 * (a0 * a0 - b0 * b0), pw = 0.5;
  for (int it = 0; it < 40; ++it) {
    const double an = 0.5 * (a0 + b0), bn = sqrt(a0 * b0), cn = 0.5 * (a0 - b0);
    pw *= 2.0;
    c2sum += pw * cn * cn;
    a0 = an; b0 = bn;
    if (fabs(cn) < 1e-17 * fabs(an)) break;
  }
  return 3.14159265358979323846 / (2.0 * a0) * (1.0 - c2sum);
}

// evaluate_A_integral_0   system_cylinder.py:91-103
__host__ __device__ __forceinline__ double surface_area_integral(double k, double r, double a) {
  const double ra = geo_radius_rescaled(r, a);
  const double t = a * ra * k;
  return 4.0 * ra / k * ellipe_agm(-(t * t));
}

#ifdef __CUDACC__
// The same for the sweep kernels, where one warp evaluates the proposal's surface energy every sweep and the chain of IEEE
// square roots and divisions is the critical path of that phase: square roots from the MUFU seed + Newton steps (fast_rsqrt,
// ~2e-16 relative), the convergence test before the next square root.  The API entry points keep the IEEE version.
__device__ __forceinline__ double surface_area_integral_fast(double k, double r, double a) {
  const double ra = fast_radius_rescaled(r, a);
  const double t = a * ra * k;
  const double x0 = fma(t, t, 1.0);                             // 1 - m
  double a0 = 1.0, b0 = x0 * fast_rsqrt(x0);
  double c2sum = 0.5 * (a0 * a0 - b0 * b0), pw = 0.5;
  for (int it = 0; it < 40; ++it) {
    const double an = 0.5 * (a0 + b0), cn = 0.5 * (a0 - b0);
    pw *= 2.0;
    c2sum += pw * cn * cn;
    if (fabs(cn) < 1e-17 * fabs(an)) { a0 = an; break; }
    const double ab = a0 * b0;
    b0 = ab * fast_rsqrt(ab);
    a0 = an;
  }
  return 4.0 * ra * fast_rcp(k) * (3.14159265358979323846 * 0.5 * fast_rcp(a0) * (1.0 - c2sum));
}
#endif

// calc_bending_energy  system_cylinder.py:108-117, evaluated cooperatively by one warp (all lanes return it).
__device__ __forceinline__ double bending_energy_warp(double k, double r, double H0, double a) {
  const double ra = geo_radius_rescaled(r, a);
  const double L = 2.0 * 3.14159265358979323846 / k;
  const int lane = threadIdx.x & 31;
  double s_zz = 0.0, s_tt = 0.0;
  for (int i = lane; i < CYL_QUAD_NODES; i += 32) {
    const double z = L * i / CYL_QUAD_NODES;
    double s, c;
    sincos(k * z, &s, &c);
    const double gt = ra * (1.0 + a * s);
    const double t = ra * a * k * c;
    const double gz = sqrt(1.0 + t * t);
    const double q = ra * a * k * k * s;                         // :53
    const double gz2 = gz * gz;
    s_zz += q * q * gt / (gz2 * gz2 * gz);                       // Kzz_integrand :52-55
    s_tt += 1.0 / (gt * gz);                                     // Kthth_integrand :46-50
  }
  s_zz = warp_sum(s_zz) * (L / CYL_QUAD_NODES);
  s_tt = warp_sum(s_tt) * (L / CYL_QUAD_NODES);
  const double t = a * k * ra;
  const double kzz_lin = -L * (-1.0 + sqrt(1.0 + t * t));       // :115
  const double ktt_lin = -L;                                     // :116
  return s_zz + s_tt - 2.0 * H0 * (kzz_lin + ktt_lin);           // :117
}

// calc_surface_energy :119-124 with effective_gamma :22.  Warp-cooperative.
__device__ __forceinline__ double surface_energy_warp(const CylParams& p, double a, double* area_out = nullptr,
                                                      double* bend_out = nullptr, bool skip_zero_kappa = false) {
  const double gamma_eff = p.gamma + p.kappa / 2.0 * p.intrinsic_curvature * p.intrinsic_curvature;
  const double area = skip_zero_kappa ? surface_area_integral_fast(p.wavenumber, p.radius, a)      // sweep kernels
                                      : surface_area_integral(p.wavenumber, p.radius, a);
  double bend = 0.0;
  if (!(skip_zero_kappa && p.kappa == 0.0)) bend = bending_energy_warp(p.wavenumber, p.radius, p.intrinsic_curvature, a);
  if (area_out) *area_out = area;
  if (bend_out) *bend_out = bend;
  return 2.0 * 3.14159265358979323846 * (gamma_eff * area + 0.5 * p.kappa * bend);
}
