// Fourier-mode field model (config 2): Cylinder2D.calc_field_energy / calc_field_energy_amplitude_change,
// system_cylinder2D.py:151-198, and Cylinder1D.calc_surface_energy, system_cylinder1D.py:189-194.
//
// The reference builds Toeplitz tensors A (from 8 nz + 1 Fourier integrals of sqrt(g)), B and D with
// scipy.integrate.quad -- 184 adaptive quadratures per amplitude proposal at (nz, nth) = (2, 1) -- and contracts
// them with the coefficients by einsum.  Summing the Fourier series first turns the same number into one
// integral over the surface:
//   psi_b(z) = sum_j c[b][j] e^{i j k z},   Psi(z, th) = sum_b psi_b(z) e^{i b th}
//   E(a, c) = int_0^{2pi/k} dz { Gth Gz [ alpha sum_b |psi_b|^2 + u/2 <|Psi|^4>_th ]
//                               + C [ Gth Gz sum_b |d_z psi_b|^2 + (Gz/Gth) sum_b (b - n A_th)^2 |psi_b|^2 ] }
// evaluated with a 256-node periodic trapezoid in z and an exact (>= 4 nth + 1 node) trapezoid in theta.  No A/B/D
// tensors, no save_temporary_matrices bookkeeping: the same function serves amplitude changes and fixed-amplitude
// coefficient moves.  One warp per chain, lanes over z nodes.
#include "cyl_common.cuh"
#include "cyl_surface.cuh"

#define FOURIER_MAX_NZ 8
#define FOURIER_MAX_NTH 4

__device__ double fourier_energy_warp(const CylParams& p, int nz, int nth, double a, const double2* __restrict__ c) {
  const int lane = threadIdx.x & 31;
  const int lz = 2 * nz + 1, lth = 2 * nth + 1;
  const int nthq = 4 * nth + 2;                       // theta nodes: exact for |Psi|^4 (degree 4 nth in e^{i th})
  const double k = p.wavenumber, r = p.radius;
  const double ra = geo_radius_rescaled(r, a);
  const double Lz = 2.0 * 3.14159265358979323846 / k;
  double acc = 0.0;
  for (int node = lane; node < CYL_QUAD_NODES; node += 32) {
    const double z = Lz * node / CYL_QUAD_NODES;
    double s1, c1;
    sincospi(2.0 * node / CYL_QUAD_NODES, &s1, &c1);            // kz = 2 pi node / N exactly
    const Geo g = geo_from_sc(k, ra, a, s1, c1);
    (void)z;
    double2 psi[2 * FOURIER_MAX_NTH + 1];
    double sum_p2 = 0.0, sum_dz2 = 0.0, sum_th = 0.0;
    for (int b = 0; b < lth; ++b) {
      double pr = 0.0, pi_ = 0.0, dr = 0.0, di = 0.0;
      for (int j = 0; j < lz; ++j) {
        const int jj = j - nz;
        double sj, cj;
        sincospi(2.0 * ((jj * node) % CYL_QUAD_NODES) / CYL_QUAD_NODES, &sj, &cj);    // e^{i jj k z}
        const double2 cc = c[b * lz + j];
        const double tr = cc.x * cj - cc.y * sj, ti = cc.x * sj + cc.y * cj;
        pr += tr; pi_ += ti;
        dr += -(jj * k) * ti; di += (jj * k) * tr;              // d/dz = i jj k
      }
      psi[b] = make_double2(pr, pi_);
      const double p2 = pr * pr + pi_ * pi_;
      sum_p2 += p2;
      sum_dz2 += dr * dr + di * di;
      const double bb = (double)(b - nth) - p.n * g.ath;
      sum_th += bb * bb * p2;
    }
    double quart = 0.0;
    for (int m = 0; m < nthq; ++m) {
      double Pr = 0.0, Pi = 0.0;
      for (int b = 0; b < lth; ++b) {
        double sb, cb;
        sincospi(2.0 * (((b - nth) * m) % nthq) / nthq, &sb, &cb);
        Pr += psi[b].x * cb - psi[b].y * sb;
        Pi += psi[b].x * sb + psi[b].y * cb;
      }
      const double P2 = Pr * Pr + Pi * Pi;
      quart += P2 * P2;
    }
    quart /= nthq;
    const double sg = g.gth * g.gz;
    acc += sg * (p.alpha * sum_p2 + 0.5 * p.u * quart) + p.C * (sg * sum_dz2 + g.gz / g.gth * sum_th);
  }
  return warp_sum(acc) * (Lz / CYL_QUAD_NODES);
}

__global__ void fourier_energy_kernel(const CylParams* __restrict__ params, int nz, int nth, const double* __restrict__ amp,
                                      const double2* __restrict__ coeffs, int B, double* __restrict__ E_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const CylParams p = params[b];
  const int ncoef = (2 * nz + 1) * (2 * nth + 1);
  const double e = fourier_energy_warp(p, nz, nth, amp[b], coeffs + (size_t)b * ncoef);
  if ((threadIdx.x & 31) == 0) E_out[b] = e;
}

extern "C" int cyl_fourier_energy_f64(const CylParams* params, int nz, int nth, const double* amp, const double* coeffs,
                                      int B, double* E_out, void* stream) {
  CYL_CHECK_ARG(params && amp && coeffs && E_out && B > 0, "cyl_fourier_energy_f64: bad arguments");
  CYL_CHECK_ARG(nz >= 0 && nz <= FOURIER_MAX_NZ && nth >= 0 && nth <= FOURIER_MAX_NTH,
                "cyl_fourier_energy_f64: num_field_coeffs (%d, %d) outside (0..%d, 0..%d)", nz, nth, FOURIER_MAX_NZ, FOURIER_MAX_NTH);
  const int warps = 4;
  fourier_energy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, as_stream(stream)>>>(params, nz, nth, amp,
                                                                                      (const double2*)coeffs, B, E_out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

// Cylinder1D.calc_surface_energy: effective_gamma * int Gth Gz dz + kappa * calc_bending_energy   (system_cylinder1D.py:189-194;
// no 2 pi prefactor, kappa not halved -- a different normalisation from Cylinder.calc_surface_energy)
__global__ void fourier_surface_energy_kernel(const CylParams* __restrict__ params, const double* __restrict__ amp, int B,
                                              double* __restrict__ E_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const CylParams p = params[b];
  const double a = amp[b];
  const double gamma_eff = p.gamma + p.kappa / 2.0 * p.intrinsic_curvature * p.intrinsic_curvature;
  const double area = surface_area_integral(p.wavenumber, p.radius, a);            // = int_0^{2pi/k} Gth Gz dz
  const double bend = bending_energy_warp(p.wavenumber, p.radius, p.intrinsic_curvature, a);
  if ((threadIdx.x & 31) == 0) E_out[b] = gamma_eff * area + p.kappa * bend;
}

extern "C" int cyl_fourier_surface_energy_f64(const CylParams* params, const double* amp, int B, double* E_out, void* stream) {
  CYL_CHECK_ARG(params && amp && E_out && B > 0, "cyl_fourier_surface_energy_f64: bad arguments");
  const int warps = 4;
  fourier_surface_energy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, as_stream(stream)>>>(params, amp, B, E_out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
