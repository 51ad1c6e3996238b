// Sequential replay of Lattice.step_lattice_loc from a recorded random tape (parity mode).
//
// This kernel restates the reference's single-site move in the reference's own operation order, in fp64,
// compiled WITHOUT fused multiply-add contraction (build flag -fmad=false for this file), so that a chain fed
// the reference's random stream reproduces its accept/reject sequence.  The cached arrays psi_squared, dz, dth
// of the reference are pure functions of psi and are recomputed on the fly.
//   simple            On_lattice_simple.py:598-686
//   rescaled_0thorder On_lattice_rescaled_0thorder.py:116-209
//   rescaled_1storder On_lattice_rescaled_1storder.py:138-227   ("1st:")
//   decoupled         On_lattice_decoupled.py:259-347           ("dec:")
//   update_sigma      On_lattice_simple.py:688-696
//   accept rule       metropolisengine.metropolis_decision (predecessor metropolis_engine.py:226-246)
// The chain is strictly sequential (sigma adapts after every proposal), so one thread walks the tape.
#include "cyl_common.cuh"

struct Cplx { double re, im; };
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cdivr(Cplx a, double s) { return {a.re / s, a.im / s}; }
__device__ __forceinline__ double csq(Cplx a) { return a.re * a.re + a.im * a.im; }        // squared() :105-106
// Im(A * conj(x) * y) for real A, evaluated as the reference does: (A*conj(x)) * y
__device__ __forceinline__ double cross_im(double A, Cplx x, Cplx y) {
  const double pr = A * x.re, pi = -(A * x.im);     // A * conj(x)
  return pr * y.im + pi * y.re;                     // imag part of (pr + i pi)(y.re + i y.im)
}

struct SiteGeo {
  double A, R, A_nb, R_nb, sg, gz_m, gz_p, gth_m, gth_p, sg_m, sg_p;
};

__device__ __forceinline__ SiteGeo site_geo(const CylParams& p, double a, double lz, int iz, int variant) {
  const double k = p.wavenumber, r = p.radius;
  const double z_loc = lz * iz;                       // :599
  const double z_m = lz * (iz - .5);                  // :600
  const double z_p = lz * (iz + .5);                  // :601
  const Geo g0 = geo_eval(k, r, a, z_loc), gm = geo_eval(k, r, a, z_m), gp = geo_eval(k, r, a, z_p);
  SiteGeo s;
  if (variant != CYL_VARIANT_RESCALED_0TH) {
    s.A = gm.ath;  s.R = 1.0 / (gm.gth * gm.gth);       // :603-604 / 1st:143-144 / dec:264-265
    s.A_nb = gp.ath; s.R_nb = 1.0 / (gp.gth * gp.gth);  // :605-606 / 1st:145-146 / dec:266-267
  } else {
    s.A = g0.ath;  s.R = 1.0 / (g0.gth * g0.gth);       // 0th:122-123
    s.A_nb = s.A;  s.R_nb = s.R;                        // 0th:164,184-188
  }
  s.sg = g0.gth * g0.gz;                                // :608
  s.gz_m = gm.gz; s.gz_p = gp.gz; s.gth_m = gm.gth; s.gth_p = gp.gth;
  s.sg_m = gm.gth * gm.gz; s.sg_p = gp.gth * gp.gz;     // 0th:128,130
  return s;
}

__device__ __forceinline__ int wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// dE of psi[iz,ith] -> nv, the quantity the reference hands to metropolis_decision(0, dE).
__device__ double delta_energy(const double2* __restrict__ psi, int Z, int T, const CylParams& p, double lz, double lth,
                               const SiteGeo& g, int variant, int iz, int ith, Cplx nv) {
  const int i0 = wrap(iz, Z), j0 = wrap(ith, T);
  const int im = wrap(iz - 1, Z), ip = wrap(iz + 1, Z), jm = wrap(ith - 1, T), jp = wrap(ith + 1, T);
  auto at = [&](int i, int j) { const double2 v = psi[(size_t)i * T + j]; return Cplx{v.x, v.y}; };
  const Cplx v = at(i0, j0), left_z = at(im, j0), right_z = at(ip, j0), left_th = at(i0, jm), right_th = at(i0, jp);
  const double Cn2 = p.C * (p.n * p.n);                                        // :32
  const double new_p2 = csq(nv), old_p2 = csq(v);                              // :618,:620
  double dE = (p.alpha * new_p2 + p.u / 2 * (new_p2 * new_p2)) - (p.alpha * old_p2 + p.u / 2 * (old_p2 * old_p2));  // :621-623
  // cached left derivatives of the current configuration (:170-182)
  const Cplx dz_c = cdivr(csub(v, left_z), lz), dth_c = cdivr(csub(v, left_th), lth);
  const Cplx dz_n = cdivr(csub(right_z, v), lz), dth_n = cdivr(csub(right_th, v), lth);
  const Cplx ndz = cdivr(csub(nv, left_z), lz), ndth = cdivr(csub(nv, left_th), lth);       // :630-631
  const Cplx nnz = cdivr(csub(right_z, nv), lz), nnth = cdivr(csub(right_th, nv), lth);     // :641-642
  double d_z, d_th, dn_z, dn_th;
  if (variant == CYL_VARIANT_SIMPLE) {
    d_z = (csq(ndz) - csq(dz_c)) / (g.gz_m * g.gz_m);                           // :633
    d_th = (csq(ndth) - csq(dth_c)) * g.R / (g.gth_m * g.gth_m);                // :634
    dn_z = (csq(nnz) - csq(dz_n)) / (g.gz_p * g.gz_p);                          // :643
    dn_th = (csq(nnth) - csq(dth_n)) * g.R_nb / (g.gth_p * g.gth_p);            // :644
  } else if (variant == CYL_VARIANT_RESCALED_1ST) {
    d_z = (csq(ndz) - csq(dz_c)) / (g.gz_m * g.gz_m);                           // 1st:171
    d_th = (csq(ndth) - csq(dth_c)) * g.R;                                      // 1st:172
    dn_z = (csq(nnz) - csq(dz_n)) / (g.gz_p * g.gz_p);                          // 1st:181
    dn_th = (csq(nnth) - csq(dth_n)) * g.R_nb;                                  // 1st:182
  } else if (variant == CYL_VARIANT_DECOUPLED) {
    d_z = (csq(ndz) - csq(dz_c));                                               // dec:294
    d_th = (csq(ndth) - csq(dth_c)) * g.R;                                      // dec:295
    dn_z = (csq(nnz) - csq(dz_n));                                              // dec:304
    dn_th = (csq(nnth) - csq(dth_n)) * g.R_nb;                                  // dec:305
  } else {
    d_z = (csq(ndz) - csq(dz_c)) / (g.gz_m * g.gz_m) * g.sg_m / g.sg;           // 0th:153
    d_th = (csq(ndth) - csq(dth_c)) * g.R;                                      // 0th:154
    dn_z = (csq(nnz) - csq(dz_n)) / (g.gz_p * g.gz_p) * g.sg_p / g.sg;          // 0th:163
    dn_th = (csq(nnth) - csq(dth_n)) * g.R;                                     // 0th:164
  }
  dE += p.C * (d_z + d_th);                                                     // :636
  dE += p.C * (dn_z + dn_th);                                                   // :646
  const double A2 = g.A * g.A;
  if (variant == CYL_VARIANT_RESCALED_1ST) {
    // the cross terms carry an extra 1/sqrt_g_theta at the interstitial row: Im(((A conj(x)) y) / G) = Im(...) / G
    const double old_cross = cross_im(g.A, v, dth_c) / g.gth_m;                 // 1st:190-191
    const double new_cross = cross_im(g.A, nv, ndth) / g.gth_m;                 // 1st:194-195
    dE += p.C * 2 * p.n * g.R * (new_cross - old_cross);                        // 1st:196
    const double old_nb = cross_im(g.A_nb, right_th, dth_n) / g.gth_p;          // 1st:199-200
    const double new_nb = cross_im(g.A_nb, right_th, nnth) / g.gth_p;           // 1st:201-202
    dE += p.C * 2 * p.n * g.R_nb * (new_nb - old_nb);                           // 1st:203
    dE += Cn2 * g.R * A2 * (new_p2 - old_p2);                                   // 1st:205-206
    dE *= g.sg;                                                                 // 1st:209
    dE *= lz * lth;                                                             // 1st:210
    return dE;
  }
  const double old_cross = cross_im(g.A, v, dth_c);                             // :651-652 / dec:312-313
  const double new_cross = cross_im(g.A, nv, ndth);                             // :654-655 / dec:315-316
  dE += p.C * 2 * p.n * g.R * (new_cross - old_cross);                          // :656 / dec:317
  if (variant == CYL_VARIANT_SIMPLE || variant == CYL_VARIANT_DECOUPLED) {      // dec:320-327 has the same shape
    const double old_nb = cross_im(g.A_nb, right_th, dth_n);                    // :659-660
    const double new_nb = cross_im(g.A_nb, right_th, nnth);                     // :661-662
    dE += p.C * 2 * p.n * g.R_nb * (new_nb - old_nb);                           // :663
    dE += Cn2 * g.R * A2 * (new_p2 - old_p2);                                   // :665-666
  } else {
    dE += Cn2 * g.R * A2 * (new_p2 - old_p2);                                   // 0th:179-180
    const double old_nb = cross_im(g.A, right_th, dth_n);                       // 0th:184-185
    const double new_nb = cross_im(g.A, right_th, nnth);                        // 0th:186-187
    dE += p.C * 2 * p.n * g.R * (new_nb - old_nb);                              // 0th:188
  }
  dE *= g.sg;                                                                   // :668
  dE *= lz * lth;                                                               // :669
  return dE;
}

__global__ void replay_kernel(double2* __restrict__ psi, int Z, int T, const CylParams* __restrict__ params, double amp,
                              int variant, const int32_t* __restrict__ iz, const int32_t* __restrict__ ith,
                              const double* __restrict__ zre, const double* __restrict__ zim,
                              const double* __restrict__ u, int64_t n, double* __restrict__ rm_state,
                              uint8_t* __restrict__ accept_out, double* __restrict__ dE_out,
                              double* __restrict__ sigma_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const CylParams p = *params;
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z;                  // :38,:47
  const double lth = (2 * pi * p.radius) / T;                     // :39,:49
  double sigma = rm_state[0];
  double step_counter = rm_state[1];
  const double target = .5, ratio = 4.0, m = 1.0;                 // :95-100 (m=1, p*=.5 => ratio = 4 exactly)
  for (int64_t s = 0; s < n; ++s) {
    const int i = iz[s], j = ith[s];
    const SiteGeo g = site_geo(p, amp, lz, i, variant);
    const double sw = (variant == CYL_VARIANT_SIMPLE) ? sigma : sigma * g.sg;     // :613 / 0th:135-136 / 1st:153-154 / dec:274-275
    const int i0 = wrap(i, Z), j0 = wrap(j, T);
    const double2 v = psi[(size_t)i0 * T + j0];
    const Cplx nv = {v.x + (0.0 + zre[s] * sw), v.y + (0.0 + zim[s] * sw)};        // :617  gauss = mu + z*sigma
    const double dE = delta_energy(psi, Z, T, p, lz, lth, g, variant, i, j, nv);
    bool accept;                                                   // engine accept rule, B.2
    if (dE <= 0) accept = true;
    else if (p.temperature == 0) accept = false;
    else accept = (u[s] <= exp(-dE / p.temperature));
    if (accept) psi[(size_t)i0 * T + j0] = make_double2(nv.re, nv.im);             // :671-685
    // update_sigma :688-696
    step_counter += 1;
    const double f = fmax(step_counter / m, 200.0);
    const double c = sigma * ratio;
    if (accept) sigma += c * (1 - target) / f;
    else sigma -= c * target / f;
    if (accept_out) accept_out[s] = accept ? 1 : 0;
    if (dE_out) dE_out[s] = dE;
    if (sigma_out) sigma_out[s] = sigma;
  }
  rm_state[0] = sigma;
  rm_state[1] = step_counter;
}

extern "C" int cyl_replay_f64(double* psi, int Z, int T, const CylParams* params, double amp, int variant,
                              const int32_t* iz, const int32_t* ith, const double* zre, const double* zim,
                              const double* u, int64_t n, double* rm_state, uint8_t* accept_out, double* dE_out,
                              double* sigma_out, void* stream) {
  CYL_CHECK_ARG(psi && params && rm_state && Z > 0 && T > 0 && n >= 0, "cyl_replay_f64: bad arguments");
  CYL_CHECK_ARG(n == 0 || (iz && ith && zre && zim && u), "cyl_replay_f64: tape buffers are null");
  CYL_CHECK_ARG(variant >= 0 && variant < CYL_NUM_VARIANTS, "cyl_replay_f64: unknown variant %d", variant);
  if (n == 0) return CYL_OK;
  replay_kernel<<<1, 32, 0, as_stream(stream)>>>((double2*)psi, Z, T, params, amp, variant, iz, ith, zre, zim, u, n,
                                                 rm_state, accept_out, dE_out, sigma_out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

__global__ void site_delta_kernel(const double2* __restrict__ psi, int Z, int T, const CylParams* __restrict__ params,
                                  double amp, int variant, int iz, int ith, double zre, double zim, double sigma,
                                  double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  const CylParams p = *params;
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;
  const SiteGeo g = site_geo(p, amp, lz, iz, variant);
  const double2 v = psi[(size_t)wrap(iz, Z) * T + wrap(ith, T)];
  const double sw = (variant == CYL_VARIANT_SIMPLE) ? sigma : sigma * g.sg;       // :613 / 0th:135-136
  const Cplx nv = {v.x + (0.0 + zre * sw), v.y + (0.0 + zim * sw)};               // :617  gauss = mu + z*sigma
  out[0] = delta_energy(psi, Z, T, p, lz, lth, g, variant, iz, ith, nv);
  out[1] = nv.re;
  out[2] = nv.im;
  out[3] = g.sg;
}

extern "C" int cyl_site_delta_f64(const double* psi, int Z, int T, const CylParams* params, double amp, int variant,
                                  int iz, int ith, double zre, double zim, double sigma, double* out, void* stream) {
  CYL_CHECK_ARG(psi && params && out && Z > 0 && T > 0, "cyl_site_delta_f64: bad arguments");
  CYL_CHECK_ARG(variant >= 0 && variant < CYL_NUM_VARIANTS, "cyl_site_delta_f64: unknown variant %d", variant);
  site_delta_kernel<<<1, 32, 0, as_stream(stream)>>>((const double2*)psi, Z, T, params, amp, variant, iz, ith, zre, zim, sigma, out);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
