// Fourier-mode field model (config 2): Cylinder2D.calc_field_energy / calc_field_energy_amplitude_change,
// system_cylinder2D.py:151-198, Cylinder1D.calc_surface_energy, system_cylinder1D.py:189-194, and the adaptive
// Metropolis engine that run.py:286-294 (method "sequential") drives over them, batched: one warp per chain.
//
// The reference builds Toeplitz tensors A (from 8 nz + 1 Fourier integrals of sqrt(g)), B and D with
// scipy.integrate.quad -- 184 adaptive quadratures per amplitude proposal at (nz, nth) = (2, 1) -- and contracts
// them with the coefficients by einsum.  Summing the Fourier series first turns the same number into one
// integral over the surface:
//   psi_b(z) = sum_j c[b][j] e^{i j k z},   Psi(z, th) = sum_b psi_b(z) e^{i b th}
//   E(a, c) = int_0^{2pi/k} dz { Gth Gz [ alpha sum_b |psi_b|^2 + u/2 <|Psi|^4>_th ]
//                               + C [ Gth Gz sum_b |d_z psi_b|^2 + (Gz/Gth) sum_b (b - n A_th)^2 |psi_b|^2 ] }
// evaluated with a 256-node periodic trapezoid in z and an exact (4 nth + 2 node) trapezoid in theta.  The phases
// e^{i j k z_n} are 256th roots of unity and e^{i b th_m} are (4 nth + 2)th roots: both come from small tables, so the
// integrand needs no trigonometry.  No A/B/D tensors, no save_temporary_matrices bookkeeping: the same function
// serves amplitude changes and fixed-amplitude coefficient moves.
#include <stdlib.h>

#include "cyl_common.cuh"
#include "cyl_site.cuh"
#include "cyl_surface.cuh"

#define FOURIER_MAX_NZ 8
#define FOURIER_MAX_NTH 4
#define FOURIER_MAX_COEF 32            // adaptive-engine kernel: 1 + ncoef <= 33 state dimensions, one per lane
#define FOURIER_NTHQ_MAX (4 * FOURIER_MAX_NTH + 2)

// roots of unity: wz[n] = e^{2 pi i n / 256}, wt[m] = e^{2 pi i m / nthq}
struct FourierTables { double2 wz[CYL_QUAD_NODES]; double2 wt[FOURIER_NTHQ_MAX]; };

__device__ __forceinline__ void fourier_tables_init(FourierTables* t, int nth) {
  const int nthq = 4 * nth + 2;
  for (int n = threadIdx.x; n < CYL_QUAD_NODES; n += blockDim.x) {
    double s, c;
    sincospi(2.0 * n / CYL_QUAD_NODES, &s, &c);
    t->wz[n] = make_double2(c, s);
  }
  for (int m = threadIdx.x; m < nthq; m += blockDim.x) {
    double s, c;
    sincospi(2.0 * m / nthq, &s, &c);
    t->wt[m] = make_double2(c, s);
  }
}

// Cooperative energy; c = coefficients [2 nth + 1][2 nz + 1] (any address space), every participating thread returns the
// value.  By default one warp shares the 256 z-nodes (lane-strided).  With cta_red != nullptr the whole CTA does (thread-
// strided; every warp must call with the same arguments): warp sums meet in cta_red[warps] and are added in warp order by
// everyone, so all threads hold the identical total.
// NZ, NTH >= 0: the mode counts are compile-time constants (loops unrolled, the small arrays in registers); -1: run time.
template <int NZ, int NTH>
__device__ __forceinline__ double fourier_energy_t(const CylParams& p, int nz_rt, int nth_rt, double a, const double2* __restrict__ c,
                                                   const FourierTables* __restrict__ tab, double* cta_red) {
  const int nz = NZ >= 0 ? NZ : nz_rt, nth = NTH >= 0 ? NTH : nth_rt;
  const int lane = cta_red ? (int)threadIdx.x : (int)(threadIdx.x & 31);
  const int nstride = cta_red ? (int)blockDim.x : 32;
  const int lz = 2 * nz + 1, lth = 2 * nth + 1;
  const int nthq = 4 * nth + 2;                       // theta nodes: exact for |Psi|^4 (degree 4 nth in e^{i th})
  const double k = p.wavenumber, r = p.radius;
  const double ra = geo_radius_rescaled(r, a);
  const double Lz = 2.0 * 3.14159265358979323846 / k;
  double acc = 0.0;
  for (int node = lane; node < CYL_QUAD_NODES; node += nstride) {
    const double2 w1 = tab->wz[node];                               // (cos kz, sin kz), kz = 2 pi node / N exactly
    const Geo g = geo_from_sc(k, ra, a, w1.y, w1.x);
    // e^{i j k z}, j = 1..nz, by repeated multiplication (the negative powers are the conjugates)
    double2 pw[(NZ >= 0 ? NZ : FOURIER_MAX_NZ) + 1];
    pw[1] = w1;
#pragma unroll
    for (int j = 2; j <= nz; ++j) pw[j] = make_double2(pw[j - 1].x * w1.x - pw[j - 1].y * w1.y, pw[j - 1].x * w1.y + pw[j - 1].y * w1.x);
    double2 psi[2 * (NTH >= 0 ? NTH : FOURIER_MAX_NTH) + 1];
    double sum_p2 = 0.0, sum_dz2 = 0.0, sum_th = 0.0;
#pragma unroll
    for (int b = 0; b < lth; ++b) {
      const double2* cb = c + b * lz + nz;                          // cb[j] = coefficient of e^{i j k z}, j = -nz..nz
      double pr = cb[0].x, pi_ = cb[0].y, dr = 0.0, di = 0.0;
#pragma unroll
      for (int j = 1; j <= nz; ++j) {
        const double2 e = pw[j], cp_ = cb[j], cm_ = cb[-j];
        const double tr = cp_.x * e.x - cp_.y * e.y, ti = cp_.x * e.y + cp_.y * e.x;     // c_{+j} e^{+i j k z}
        const double ur = cm_.x * e.x + cm_.y * e.y, ui = cm_.y * e.x - cm_.x * e.y;     // c_{-j} e^{-i j k z}
        pr += tr + ur; pi_ += ti + ui;
        const double jk = j * k;                                    // d/dz = +- i j k
        dr -= jk * (ti - ui); di += jk * (tr - ur);
      }
      psi[b] = make_double2(pr, pi_);
      const double p2 = pr * pr + pi_ * pi_;
      sum_p2 += p2;
      sum_dz2 += dr * dr + di * di;
      const double bb = (double)(b - nth) - p.n * g.ath;
      sum_th += bb * bb * p2;
    }
    // <|Psi|^4>_theta on the nthq-node grid: Psi(th_m) = psi_0 + sum_t (psi_{+t} e^{i t th_m} + psi_{-t} e^{-i t th_m})
    double quart = 0.0;
#pragma unroll
    for (int m = 0; m < nthq; ++m) {
      const double2 e1 = tab->wt[m];
      double2 e = e1;
      double Pr = psi[nth].x, Pi = psi[nth].y;
#pragma unroll
      for (int t = 1; t <= nth; ++t) {
        const double2 qp = psi[nth + t], qm = psi[nth - t];
        Pr += (qp.x * e.x - qp.y * e.y) + (qm.x * e.x + qm.y * e.y);
        Pi += (qp.x * e.y + qp.y * e.x) + (qm.y * e.x - qm.x * e.y);
        e = make_double2(e.x * e1.x - e.y * e1.y, e.x * e1.y + e.y * e1.x);
      }
      const double P2 = Pr * Pr + Pi * Pi;
      quart += P2 * P2;
    }
    quart /= nthq;
    const double sg = g.gth * g.gz;
    acc += sg * (p.alpha * sum_p2 + 0.5 * p.u * quart) + p.C * (sg * sum_dz2 + g.gz / g.gth * sum_th);
  }
  acc = warp_sum(acc);
  if (cta_red) {
    const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) cta_red[warp] = acc;
    __syncthreads();
    acc = 0.0;
    for (int w = 0; w < nwarp; ++w) acc += cta_red[w];
    __syncthreads();                                              // cta_red is free again
  }
  return acc * (Lz / CYL_QUAD_NODES);
}

__device__ __forceinline__ double fourier_energy_warp(const CylParams& p, int nz, int nth, double a, const double2* __restrict__ c,
                                                      const FourierTables* __restrict__ tab, double* cta_red = nullptr) {
  if (nz == 2 && nth == 1) return fourier_energy_t<2, 1>(p, nz, nth, a, c, tab, cta_red);      // run.py:369, the reference's default
  return fourier_energy_t<-1, -1>(p, nz, nth, a, c, tab, cta_red);
}

// Cylinder1D.calc_surface_energy: effective_gamma * int Gth Gz dz + kappa * calc_bending_energy   (system_cylinder1D.py:189-194;
// no 2 pi prefactor, kappa not halved -- a different normalisation from Cylinder.calc_surface_energy)
__device__ __forceinline__ double fourier_surface_energy_warp(const CylParams& p, double a) {
  const double gamma_eff = p.gamma + p.kappa / 2.0 * p.intrinsic_curvature * p.intrinsic_curvature;
  const double area = surface_area_integral(p.wavenumber, p.radius, a);            // = int_0^{2pi/k} Gth Gz dz
  const double bend = (p.kappa != 0.0) ? bending_energy_warp(p.wavenumber, p.radius, p.intrinsic_curvature, a) : 0.0;
  return gamma_eff * area + p.kappa * bend;
}

__global__ void fourier_energy_kernel(const CylParams* __restrict__ params, int nz, int nth, const double* __restrict__ amp,
                                      const double2* __restrict__ coeffs, int B, double* __restrict__ E_out) {
  __shared__ FourierTables tab;
  fourier_tables_init(&tab, nth);
  __syncthreads();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const CylParams p = params[b];
  const int ncoef = (2 * nz + 1) * (2 * nth + 1);
  const double e = fourier_energy_warp(p, nz, nth, amp[b], coeffs + (size_t)b * ncoef, &tab);
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

__global__ void fourier_surface_energy_kernel(const CylParams* __restrict__ params, const double* __restrict__ amp, int B,
                                              double* __restrict__ E_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const CylParams p = params[b];
  const double a = amp[b];
  const double gamma_eff = p.gamma + p.kappa / 2.0 * p.intrinsic_curvature * p.intrinsic_curvature;
  const double area = surface_area_integral(p.wavenumber, p.radius, a);
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

// ----------------------------------------------------------------------------------------------------
// The reference's tensor tables for callers that read them (tests/test_matrixA.py, tests/test_1Dvs2D.py of the reference):
//   A[d]          = int_0^L e^{i d k z} Gth Gz dz,                    d in [-4 nz, 4 nz]       system_cylinder2D.py:103-109
//   B[i, j, beta] = int_0^L e^{i (i - j) k z} ( k^2 i j Gz Gth + (beta - n A_theta)^2 Gz / Gth ) dz   :64-100, :134-147
// one warp per element, 256-node periodic trapezoid (the integrands are analytic and periodic: spectral accuracy; the
// reference runs scipy.integrate.quad twice per element).  out_A: [B][8 nz + 1] complex, out_B: [B][2nz+1][2nz+1][2nth+1] complex
// (the diagonal beta' = beta of the reference's 4-index B; its off-diagonal elements are zero by the selection rule).
// The energies never go through these tables (cyl_fourier_energy_f64 integrates the energy density directly).
// ----------------------------------------------------------------------------------------------------
__global__ void fourier_tensor_tables_kernel(const CylParams* __restrict__ params, int nz, int nth, const double* __restrict__ amp, int B,
                                             double2* __restrict__ out_A, double2* __restrict__ out_B) {
  const int nA = 8 * nz + 1, lz = 2 * nz + 1, lth = 2 * nth + 1, nB = lz * lz * lth, per = nA + nB;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * per) return;
  const int b = (int)(w / per), e = (int)(w - (long long)b * per), lane = threadIdx.x & 31;
  const CylParams p = params[b];
  const double a = amp[b], k = p.wavenumber;
  const double ra = geo_radius_rescaled(p.radius, a);
  const double L = 2.0 * 3.14159265358979323846 / k;
  int d, ii = 0, jj = 0, beta = 0;
  const bool isA = e < nA;
  if (isA) d = e - 4 * nz;
  else {
    const int q = e - nA;
    ii = q / (lz * lth) - nz; jj = (q / lth) % lz - nz; beta = q % lth - nth;
    d = ii - jj;
  }
  double sre = 0.0, sim = 0.0;
  for (int node = lane; node < CYL_QUAD_NODES; node += 32) {
    const double z = L * node / CYL_QUAD_NODES;
    double s, c;
    sincos(k * z, &s, &c);
    const double gth = ra * (1.0 + a * s);
    const double t = ra * a * k * c;
    const double gz = sqrt(1.0 + t * t);
    double body;
    if (isA) body = gth * gz;
    else {
      const double ath = t / gz, m = (double)beta - p.n * ath;
      body = k * k * ii * jj * gz * gth + m * m * gz / gth;
    }
    double sd, cd;
    sincospi(2.0 * (double)(((long long)d * node) % CYL_QUAD_NODES) / CYL_QUAD_NODES, &sd, &cd);   // e^{i d k z}, exact phase
    sre += cd * body; sim += sd * body;
  }
  sre = warp_sum(sre) * (L / CYL_QUAD_NODES);
  sim = warp_sum(sim) * (L / CYL_QUAD_NODES);
  if (lane == 0) {
    if (isA) out_A[(size_t)b * nA + e] = make_double2(sre, sim);
    else out_B[(size_t)b * nB + (e - nA)] = make_double2(sre, sim);
  }
}

extern "C" int cyl_fourier_tensor_tables_f64(const CylParams* params, int nz, int nth, const double* amp, int B, double* out_A,
                                             double* out_B, void* stream) {
  CYL_CHECK_ARG(params && amp && out_A && out_B && B > 0 && nz >= 0 && nth >= 0, "cyl_fourier_tensor_tables_f64: bad arguments");
  const long long total = (long long)B * ((8 * nz + 1) + (long long)(2 * nz + 1) * (2 * nz + 1) * (2 * nth + 1));
  const int warps = 4;
  fourier_tensor_tables_kernel<<<(unsigned)((total + warps - 1) / warps), warps * 32, 0, as_stream(stream)>>>(
      params, nz, nth, amp, B, reinterpret_cast<double2*>(out_A), reinterpret_cast<double2*>(out_B));
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

// ----------------------------------------------------------------------------------------------------
// Batched adaptive Metropolis for the Fourier model: run.py:286-294 (method "sequential") with the engine
// arithmetic of SURVEY.md Appendix B, one warp per chain, state dimension d = 1 + ncoef <= 33 (one per lane).
//   per step:  measure_every x [ real-group move ; fieldsteps x complex-group move ] ; measure()
//   real group    a' = a + sigma_r sqrt(Sigma[0][0]) N(0,1); rejected outright when |a'| >= amp_bound; energy
//                 E_field(a', c) + E_surf(a'); Robbins-Monro on sigma_r (m = 1)
//   complex group |c_i|' = |c_i| + sigma_c (L z)_i with L L^T = Sigma_c (Cholesky; the reference draws the same law with
//                 numpy's SVD-based multivariate_normal), arg c_i' = arg c_i + 0.1 N(0,1) ("magnitude-phase");
//                 energy E_field(a, c'); Robbins-Monro on sigma_c (m = ncoef)
//   measure       running mean of x = (|a|, |c_i|), and once k > 50 the recursive covariance
//                 Sigma <- Sigma (k-2)/(k-1) + mu_old mu_old^T - k/(k-1) mu mu^T + x x^T/(k-1) + sigma_r^2/k I
// Draws: Philox4x32-10 at counter (lane, chain, move index, tag).  Engine state per chain (doubles):
//   [0] sigma_r [1] sigma_c [2] real moves [3] complex moves [4] measurements [5] real accepted [6] complex accepted
//   [7] E_field [8] E_surf [9] sum a^2 [10] sum E [11..15] reserved | mean[d] | Sigma[d][d]
// ----------------------------------------------------------------------------------------------------
#define FAM_HDR 16
#define CYL_TAG_FREAL 3u
#define CYL_TAG_FCPLX 4u
#define CYL_TAG_FALL 5u

// CTA = false: one warp per chain (many chains).  CTA = true: one CTA per chain (few chains): the eight warps carry identical
// copies of the chain state and take identical decisions; only the energy quadrature is shared among them, one node per
// thread, which cuts the latency of a move about five-fold.
#ifndef FAM_MINB
#define FAM_MINB 4             // warp-per-chain layout: 128 registers, 16 warps per SM measured best (2: -10 %)
#endif
template <bool CTA>
__global__ void __launch_bounds__(CTA ? 256 : 128, CTA ? 1 : FAM_MINB)
fourier_am_kernel(const CylParams* __restrict__ params, int nz, int nth, double* __restrict__ amp, double2* __restrict__ coeffs,
                  double* __restrict__ est, int B, uint64_t seed, int64_t chain0, uint64_t step0, int n_steps, int measure_every,
                  int fieldsteps, double target, double ratio_real, double ratio_cplx, int method) {
  extern __shared__ __align__(16) unsigned char fam_smem[];
  FourierTables* tab = reinterpret_cast<FourierTables*>(fam_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int ncoef = (2 * nz + 1) * (2 * nth + 1), d = ncoef + 1;
  // The four drivers of run.py:286-317: sequential (amplitude move, then `fieldsteps` coefficient moves), simultaneous (one
  // joint move of [|a|, |c_i|]), fixed-amplitude (coefficient moves only), no-field (amplitude moves on the bare surface).
  const bool do_real = method == CYL_FOURIER_SEQUENTIAL || method == CYL_FOURIER_NO_FIELD;
  const bool do_cplx = method == CYL_FOURIER_SEQUENTIAL || method == CYL_FOURIER_FIXED_AMPLITUDE;
  const bool has_field = method != CYL_FOURIER_NO_FIELD;
  // per-warp scratch: c[ncoef], cp[ncoef] (double2), L[d][d]
  double2* c_all = reinterpret_cast<double2*>(fam_smem + sizeof(FourierTables));
  double2* c = c_all + (size_t)warp * 2 * ncoef;
  double2* cp = c + ncoef;
  double* L = reinterpret_cast<double*>(c_all + (size_t)nwarp * 2 * ncoef) + (size_t)warp * d * d;
  double* cta_red = CTA ? reinterpret_cast<double*>(c_all + (size_t)nwarp * 2 * ncoef) + (size_t)nwarp * d * d : nullptr;
  fourier_tables_init(tab, nth);
  __syncthreads();
  const int b = CTA ? (int)blockIdx.x : (int)(blockIdx.x * nwarp + warp);
  if (b >= B) return;
  const CylParams p = params[b];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t chain = (uint32_t)(chain0 + b);
  double* st = est + (size_t)b * (FAM_HDR + d + (size_t)d * d);
  double* mean = st + FAM_HDR;
  double* cov = mean + d;
  double a = amp[b];
  for (int i = lane; i < ncoef; i += 32) c[i] = coeffs[(size_t)b * ncoef + i];
  __syncwarp();
  double sig_r = st[0], sig_c = st[1], n_real = st[2], n_cplx = st[3], n_meas = st[4], acc_r = st[5], acc_c = st[6];
  double sum_a2 = st[9], sum_e = st[10];
  double e_field = has_field ? fourier_energy_warp(p, nz, nth, a, c, tab, cta_red) : 0.0;
  double e_surf = fourier_surface_energy_warp(p, a);

  // L L^T = the block of the covariance the proposal needs: cov[1:,1:] (coefficient moves) or all of it (joint move);
  // lane i owns row i
  const int coff = (method == CYL_FOURIER_SIMULTANEOUS) ? 0 : 1, cn = d - coff;
  auto cholesky = [&]() {
    if (!has_field) return;
    for (int i = lane; i < cn; i += 32)
      for (int j = 0; j < cn; ++j) L[i * cn + j] = (j <= i) ? cov[(i + coff) * d + (j + coff)] : 0.0;
    __syncwarp();
    for (int kx = 0; kx < cn; ++kx) {
      const double dk = sqrt(fmax(L[kx * cn + kx], 1e-300));
      __syncwarp();
      if (lane == kx) L[kx * cn + kx] = dk;
      if (lane > kx && lane < cn) L[lane * cn + kx] /= dk;
      __syncwarp();
      if (lane > kx && lane < cn) {
        const double lik = L[lane * cn + kx];
        for (int j = kx + 1; j <= lane; ++j) L[lane * cn + j] -= lik * L[j * cn + kx];
      }
      __syncwarp();
    }
  };
  cholesky();
  const int moves_per_iter = (do_real ? 1 : 0) + (do_cplx ? fieldsteps : 0) + (method == CYL_FOURIER_SIMULTANEOUS ? 1 : 0);

  uint64_t move = step0 * (uint64_t)measure_every * (uint64_t)moves_per_iter;         // global move index: counter word 2
  for (int s = 0; s < n_steps; ++s) {
    for (int me = 0; me < measure_every; ++me) {
      // ---- real group ----
      if (do_real) {
        const uint4 x = philox4x32<10>(0u, chain, (uint32_t)move, philox_c3(CYL_TAG_FREAL, move), k0, k1);
        double z, u;
        amp_draw(x, z, u);
        const double a_new = a + sig_r * sqrt(cov[0]) * z;
        bool accept = false;
        if (fabs(a_new) < p.amp_bound) {
          const double ef = has_field ? fourier_energy_warp(p, nz, nth, a_new, c, tab, cta_red) : 0.0;
          const double es = fourier_surface_energy_warp(p, a_new);
          const double diff = (ef + es) - (e_field + e_surf);
          accept = (diff <= 0) || (p.temperature > 0 && u <= exp(-diff / p.temperature));
          if (accept) { a = a_new; e_field = ef; e_surf = es; }
        }
        n_real += 1;
        acc_r += accept;
        const double f = fmax(n_real, 200.0), cc = sig_r * ratio_real;
        sig_r = accept ? sig_r + cc * (1 - target) / f : sig_r - cc * target / f;
        ++move;
      }
      // ---- joint move of [|a|, |c_i|] ~ N(0, sigma^2 Sigma) + phase jitter (engine step_all, predecessor :213-218) ----
      if (method == CYL_FOURIER_SIMULTANEOUS) {
        double z, u, zph, dummy;
        amp_draw(philox4x32<10>((uint32_t)lane, chain, (uint32_t)move, philox_c3(CYL_TAG_FALL, move), k0, k1), z, u);
        amp_draw(philox4x32<10>((uint32_t)lane + 64u, chain, (uint32_t)move, philox_c3(CYL_TAG_FALL, move), k0, k1), zph, dummy);
        u = __shfl_sync(0xffffffffu, u, 0);
        double dv = 0.0;
        for (int j = 0; j < d; ++j) {
          const double zj = __shfl_sync(0xffffffffu, z, j);
          if (lane < d && j <= lane) dv += L[lane * d + j] * zj;
        }
        double a_new = 0.0;
        if (lane == 0) {
          const double mag = fabs(a) + sig_r * dv;
          a_new = (a != 0.0) ? copysign(mag, a) : mag;                   // the sign of a is restored (:218)
        } else if (lane < d) {
          const double2 ci = c[lane - 1];
          const double mag = hypot(ci.x, ci.y) + sig_r * dv, ph = atan2(ci.y, ci.x) + 0.1 * zph;
          double sn, cs;
          sincos(ph, &sn, &cs);
          cp[lane - 1] = make_double2(mag * cs, mag * sn);
        }
        a_new = __shfl_sync(0xffffffffu, a_new, 0);
        __syncwarp();
        bool accept = false;
        if (fabs(a_new) < p.amp_bound) {
          const double ef = fourier_energy_warp(p, nz, nth, a_new, cp, tab, cta_red);
          const double es = fourier_surface_energy_warp(p, a_new);
          const double diff = (ef + es) - (e_field + e_surf);
          accept = (diff <= 0) || (p.temperature > 0 && u <= exp(-diff / p.temperature));
          if (accept) {
            if (lane < ncoef) c[lane] = cp[lane];
            a = a_new; e_field = ef; e_surf = es;
          }
        }
        __syncwarp();
        n_real += 1;
        acc_r += accept;
        const double f = fmax(n_real / d, 200.0), cc = sig_r * ratio_real;       // host passes the ratio for m = d
        sig_r = accept ? sig_r + cc * (1 - target) / f : sig_r - cc * target / f;
        ++move;
      }
      // ---- complex group ----
      for (int fs = 0; do_cplx && fs < fieldsteps; ++fs) {
        const uint4 x = philox4x32<10>((uint32_t)lane, chain, (uint32_t)move, philox_c3(CYL_TAG_FCPLX, move), k0, k1);
        double z, u, zph, dummy;
        amp_draw(x, z, u);                                   // z: magnitude component of this lane; u used from lane 0
        amp_draw(philox4x32<10>((uint32_t)lane + 64u, chain, (uint32_t)move, philox_c3(CYL_TAG_FCPLX, move), k0, k1), zph, dummy);  // phase jitter
        u = __shfl_sync(0xffffffffu, u, 0);
        double dm = 0.0;
        for (int j = 0; j < ncoef; ++j) {
          const double zj = __shfl_sync(0xffffffffu, z, j);
          if (lane < ncoef && j <= lane) dm += L[lane * ncoef + j] * zj;
        }
        if (lane < ncoef) {
          const double2 ci = c[lane];
          const double mag = hypot(ci.x, ci.y) + sig_c * dm, ph = atan2(ci.y, ci.x) + 0.1 * zph;
          double sn, cs;
          sincos(ph, &sn, &cs);
          cp[lane] = make_double2(mag * cs, mag * sn);
        }
        __syncwarp();
        const double ef = fourier_energy_warp(p, nz, nth, a, cp, tab, cta_red);
        const double diff = ef - e_field;
        const bool accept = (diff <= 0) || (p.temperature > 0 && u <= exp(-diff / p.temperature));
        if (accept) {
          if (lane < ncoef) c[lane] = cp[lane];
          e_field = ef;
        }
        __syncwarp();
        n_cplx += 1;
        acc_c += accept;
        const double f = fmax(n_cplx / ncoef, 200.0), cc = sig_c * ratio_cplx;
        sig_c = accept ? sig_c + cc * (1 - target) / f : sig_c - cc * target / f;
        ++move;
      }
    }
    // ---- measure ----  (CTA mode: mean and covariance live in global memory once per chain -- warp 0 updates them, the
    //                     others wait and then re-factorise from the updated covariance like warp 0 does)
    n_meas += 1;
    const double kk = n_meas;
    if (!CTA || warp == 0) {
      const int dm_ = has_field ? d : 1;                                   // no-field: the state is the amplitude alone
      const double xi = (lane == 0) ? fabs(a) : (lane < dm_ ? hypot(c[lane - 1].x, c[lane - 1].y) : 0.0);
      const double mu_old = (lane < dm_) ? mean[lane] : 0.0;
      const double mu = mu_old * (kk - 1) / kk + xi / kk;
      if (lane < dm_) mean[lane] = mu;
      if (kk > 50) {
        for (int j = 0; j < dm_; ++j) {
          const double xj = __shfl_sync(0xffffffffu, xi, j), mo = __shfl_sync(0xffffffffu, mu_old, j), mn = __shfl_sync(0xffffffffu, mu, j);
          if (lane < dm_)
            cov[lane * d + j] = cov[lane * d + j] * (kk - 2) / (kk - 1) +
                                (mu_old * mo - kk / (kk - 1) * mu * mn + xi * xj / (kk - 1) + (lane == j ? sig_r * sig_r / kk : 0.0));
        }
        __syncwarp();
      }
    }
    if (CTA) __syncthreads();
    if (kk > 50) cholesky();
    sum_a2 += a * a;
    sum_e += e_field + e_surf;
  }
  // ---- write back ----
  __syncwarp();
  if (CTA && warp != 0) return;                                    // every warp holds the same state
  for (int i = lane; i < ncoef; i += 32) coeffs[(size_t)b * ncoef + i] = c[i];
  if (lane == 0) {
    amp[b] = a;
    st[0] = sig_r; st[1] = sig_c; st[2] = n_real; st[3] = n_cplx; st[4] = n_meas; st[5] = acc_r; st[6] = acc_c;
    st[7] = e_field; st[8] = e_surf; st[9] = sum_a2; st[10] = sum_e;
  }
}

extern "C" int cyl_fourier_am_state_doubles(int nz, int nth) {
  const int d = (2 * nz + 1) * (2 * nth + 1) + 1;
  return FAM_HDR + d + d * d;
}

extern "C" int cyl_fourier_am_steps_f64(const CylParams* params, int nz, int nth, double* amp, double* coeffs, double* engine_state,
                                        int B, uint64_t seed, int64_t chain0, uint64_t step0, int n_steps, int measure_every,
                                        int fieldsteps, double target_acceptance, double ratio_real, double ratio_complex,
                                        int method, void* stream) {
  CYL_CHECK_ARG(params && amp && coeffs && engine_state && B > 0 && n_steps >= 0 && measure_every > 0 && fieldsteps >= 0,
                "cyl_fourier_am_steps_f64: bad arguments");
  CYL_CHECK_ARG(nz >= 0 && nz <= FOURIER_MAX_NZ && nth >= 0 && nth <= FOURIER_MAX_NTH, "cyl_fourier_am_steps_f64: bad num_field_coeffs");
  const int ncoef = (2 * nz + 1) * (2 * nth + 1);
  CYL_CHECK_ARG(ncoef <= FOURIER_MAX_COEF - 1, "cyl_fourier_am_steps_f64: %d coefficients > %d (one state component per lane)", ncoef,
                FOURIER_MAX_COEF - 1);
  const int layout = method & (CYL_FOURIER_LAYOUT_CTA | CYL_FOURIER_LAYOUT_WARP);
  method &= ~(CYL_FOURIER_LAYOUT_CTA | CYL_FOURIER_LAYOUT_WARP);
  CYL_CHECK_ARG(method >= CYL_FOURIER_SEQUENTIAL && method <= CYL_FOURIER_NO_FIELD, "cyl_fourier_am_steps_f64: unknown method %d", method);
  if (n_steps == 0) return CYL_OK;
  // few chains: a CTA per chain (the energy quadrature shared by eight warps); many: a warp per chain
  int dev = 0, sms = 0;
  CYL_CUDA(cudaGetDevice(&dev));
  CYL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bool cta = B <= sms;
  if (layout) cta = (layout & CYL_FOURIER_LAYOUT_CTA) != 0;                       // explicit layout (tests run both)
  const int warps = cta ? 8 : 4, dd = ncoef + 1;
  const size_t smem = sizeof(FourierTables) + (size_t)warps * (2 * ncoef * sizeof(double2) + (size_t)dd * dd * sizeof(double)) +
                      (size_t)warps * sizeof(double);
  if (cta) {
    CYL_CUDA(cudaFuncSetAttribute(fourier_am_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fourier_am_kernel<true><<<B, warps * 32, smem, as_stream(stream)>>>(
        params, nz, nth, amp, (double2*)coeffs, engine_state, B, seed, chain0, step0, n_steps, measure_every, fieldsteps,
        target_acceptance, ratio_real, ratio_complex, method);
  } else {
    CYL_CUDA(cudaFuncSetAttribute(fourier_am_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fourier_am_kernel<false><<<(B + warps - 1) / warps, warps * 32, smem, as_stream(stream)>>>(
        params, nz, nth, amp, (double2*)coeffs, engine_state, B, seed, chain0, step0, n_steps, measure_every, fieldsteps,
        target_acceptance, ratio_real, ratio_complex, method);
  }
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
