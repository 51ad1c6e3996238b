// Per-row tables of the sweep kernels, evaluated in the kernel's WORKING precision A (= float for the fp32 kernels, double
// for the fp64 kernels): surface metric from cached sin/cos (system_cylinder.py:25-41), the site-move coefficients
// (cyl_site.cuh header) and the row weights of the field energy (On_lattice_simple.py:185-213 and the
// variants' overrides).  Both sweep kernels (k_sweep_smem.cu, k_sweep_hbm.cu) build their tables with THESE functions, so
// the two fp32 paths stay bit-identical to each other and the fp64 paths restate the oracle's arithmetic.
//
// Why working precision: the tables are stored in A anyway, they are rebuilt on the critical path of every accepted
// amplitude move (and the proposal's weights every sweep), and an fp64 chain of ~200 dependent operations with Newton
// reciprocals costs 3-4x the latency of the fp32 chain with MUFU reciprocals.  The chain samples exactly the model whose
// coefficients are these deterministic functions of the amplitude (a relative perturbation of ~1e-7 in fp32, the size of
// the rounding the stored tables had before).
#pragma once
#include "cyl_common.cuh"
#include "cyl_energy.cuh"
#include "cyl_site.cuh"

template <typename A> struct GeoT { A gth, gz, ath; };          // sqrt_g_theta :28, sqrt_g_z :34, A_theta :40

template <typename A> __device__ __forceinline__ A rsqrt_a(A x);
template <> __device__ __forceinline__ double rsqrt_a<double>(double x) { return fast_rsqrt(x); }
template <> __device__ __forceinline__ float rsqrt_a<float>(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <typename A> __device__ __forceinline__ A rcp_a(A x);
template <> __device__ __forceinline__ double rcp_a<double>(double x) { return fast_rcp(x); }
template <> __device__ __forceinline__ float rcp_a<float>(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double fma_a(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_a(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// sin/cos(k z) at z_i and z_{i-1/2}: amplitude-independent, tabulated once per launch (evaluated in fp64, stored in A)
template <typename A> struct __align__(16) RowTrig { A s0, c0, sm, cm; };
template <typename A>
__device__ __forceinline__ RowTrig<A> row_trig(double k, double lz, int i) {
  double s0, c0, sm, cm;
  sincos(k * (lz * i), &s0, &c0);
  sincos(k * (lz * (i - .5)), &sm, &cm);
  RowTrig<A> t;
  t.s0 = (A)s0; t.c0 = (A)c0; t.sm = (A)sm; t.cm = (A)cm;
  return t;
}

// Per-replica constants in the working precision (the few fp64 originals the rarely-taken branches need ride along).
template <typename A>
struct RepConst {
  double temperature_d, alpha_d, C_d, u_d, radius_d, wavenumber_d, lzlth_d;
  A C, n, alpha, hu, Cn2w, twoCn, Cn2c;     // Cn2w = C n^2 of the energy weights, Cn2c = (C n) n of the coefficient rows
  A lz, lth, ilz2, ilth2, ilth, lzlth;      // pixel lengths :47,:49 and their folds
  A kT;                                     // log2(e) / T (accept test in the log2 domain); unused when T == 0
  A temperature, Tlen, bg_const;            // background term of the rescaled variants (0th:66-92)
  int Z, T, variant, t_pos;
};
template <typename A>
__host__ __device__ __forceinline__ RepConst<A> make_rep_const(const CylParams& p, int Z, int T, int variant) {
  RepConst<A> c;
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;      // On_lattice_simple.py:47,:49
  c.temperature_d = p.temperature; c.alpha_d = p.alpha; c.C_d = p.C; c.u_d = p.u; c.radius_d = p.radius; c.wavenumber_d = p.wavenumber;
  c.lzlth_d = lz * lth;
  c.C = (A)p.C; c.n = (A)p.n; c.alpha = (A)p.alpha; c.hu = (A)(p.u / 2);
  c.Cn2w = (A)(p.C * p.n * p.n);
  c.twoCn = (A)(p.C * p.n * 2);
  c.Cn2c = c.C * c.n * c.n;
  c.lz = (A)lz; c.lth = (A)lth;
  c.ilz2 = (A)(1.0 / (lz * lz)); c.ilth2 = (A)(1.0 / (lth * lth)); c.ilth = (A)(1.0 / lth);
  c.lzlth = (A)(lz * lth);
  c.t_pos = p.temperature > 0;
  c.kT = c.t_pos ? (A)(1.4426950408889634 / p.temperature) : (A)0;
  c.temperature = (A)p.temperature; c.Tlen = (A)T;
  c.bg_const = (A)((p.alpha < 0) ? .5 * p.alpha * p.alpha / p.u : 0.0);
  c.Z = Z; c.T = T; c.variant = variant;
  return c;
}

// What every row shares at one amplitude: a, r(a) (:37) and r(a) a k
template <typename A> struct AmpGeo { A a, ra, t0; };
template <typename A>
__device__ __forceinline__ AmpGeo<A> amp_geo(const RepConst<A>& c, double a) {
  const double ra = fast_radius_rescaled(c.radius_d, a);
  AmpGeo<A> g;
  g.a = (A)a; g.ra = (A)ra; g.t0 = (A)(ra * a * c.wavenumber_d);
  return g;
}
template <typename A>
__device__ __forceinline__ GeoT<A> geo_sc(const AmpGeo<A>& g, A s, A c) {
  GeoT<A> o;
  o.gth = g.ra * fma_a(g.a, s, A(1));                 // :28
  const A t = g.t0 * c;
  const A x = fma_a(t, t, A(1));
  const A rs = rsqrt_a<A>(x);
  o.gz = x * rs;                                      // :34
  o.ath = t * rs;                                     // :40  k r(a) a cos(kz) / sqrt_g_z
  return o;
}

// Site-move coefficients of a row (formulas: cyl_site.cuh header) from the metric at z_i (g0) and z_{i -+ 1/2} (gm, gp).
template <typename R>
__device__ __forceinline__ RowCoef<R> coef_row(const RepConst<R>& c, const GeoT<R>& g0, const GeoT<R>& gm, const GeoT<R>& gp) {
  using A = R;
  const int variant = c.variant;
  const A C = c.C, n = c.n, ilz2 = c.ilz2, ilth = c.ilth;
  const A sg = g0.gth * g0.gz;
  const A igzm = rcp_a<A>(gm.gz), igzp = rcp_a<A>(gp.gz);
  A clz, crz, clt, crt, xl, xr, mass, sgmul;
  if (variant == CYL_VARIANT_SIMPLE) {
    const A igtm = rcp_a<A>(gm.gth), igtp = rcp_a<A>(gp.gth);
    const A Rm = igtm * igtm, Rp = igtp * igtp;
    clz = C * ilz2 * igzm * igzm;
    crz = C * ilz2 * igzp * igzp;
    clt = C * Rm * Rm * ilth * ilth;
    crt = C * Rp * Rp * ilth * ilth;
    xl = 2 * C * n * Rm * gm.ath * ilth;
    xr = 2 * C * n * Rp * gp.ath * ilth;
    mass = c.alpha + c.Cn2c * Rm * gm.ath * gm.ath;
    sgmul = A(1);
  } else if (variant == CYL_VARIANT_RESCALED_1ST || variant == CYL_VARIANT_DECOUPLED) {
    const bool first = variant == CYL_VARIANT_RESCALED_1ST;
    const A igtm = rcp_a<A>(gm.gth), igtp = rcp_a<A>(gp.gth);
    const A Rm = igtm * igtm, Rp = igtp * igtp;
    clz = first ? C * ilz2 * igzm * igzm : C * ilz2;
    crz = first ? C * ilz2 * igzp * igzp : C * ilz2;
    clt = C * Rm * ilth * ilth;
    crt = C * Rp * ilth * ilth;
    xl = 2 * C * n * Rm * gm.ath * ilth * (first ? igtm : A(1));
    xr = 2 * C * n * Rp * gp.ath * ilth * (first ? igtp : A(1));
    mass = c.alpha + c.Cn2c * Rm * gm.ath * gm.ath;
    sgmul = sg;
  } else {
    const A igt0 = rcp_a<A>(g0.gth);
    const A R0 = igt0 * igt0, isg = igt0 * rcp_a<A>(g0.gz);
    clz = C * ilz2 * igzm * gm.gth * isg;                   // C/(lz^2 Gz(z-)^2) * sg(z-)/sg(z)
    crz = C * ilz2 * igzp * gp.gth * isg;
    clt = crt = C * R0 * ilth * ilth;
    xl = xr = 2 * C * n * R0 * g0.ath * ilth;
    mass = c.alpha + c.Cn2c * R0 * g0.ath * g0.ath;
    sgmul = sg;
  }
  const A K = sg * c.lzlth;
  RowCoef<R> o;
  o.kfull = K;
  o.kacc = c.t_pos ? K * c.kT : (R)INFINITY;
  o.mass = mass;
  o.csum = clz + crz + clt + crt;
  o.hu = c.hu;
  o.clz = clz; o.crz = crz; o.clt = clt; o.crt = crt;
  o.xl = xl; o.xr = xr;
  o.sgmul = sgmul;
  return o;
}

// Row weights of the field energy at the metric g0e / gme (z_i, z_{i-1/2} at the evaluation amplitude) with the metric
// reference g0r (z_i at the reference amplitude; only the rescaled variants distinguish them), pixel lengths folded in:
//   E_row = wf[0] S1 + wf[1] S2 + wf[2] sum|dpsi_z|^2 + wf[3] sum|dpsi_th|^2 + wf[4] sum Im(conj(psi) dpsi_th) + wf[5]
// (differences NOT divided by the pixel lengths).  Same formulas as cyl_energy.cuh:field_energy_weights.
template <typename A>
__device__ __forceinline__ void weight_row(const RepConst<A>& c, const GeoT<A>& g0e, const GeoT<A>& gme, const GeoT<A>& g0r, A (&wf)[6]) {
  const int variant = c.variant;
  wf[5] = A(0);
  if (variant == CYL_VARIANT_DECOUPLED) {                      // dec:179-202: only the C n^2 |A psi|^2 coupling
    wf[0] = c.Cn2w * (gme.ath * gme.ath) * (g0e.gz * rcp_a<A>(gme.gth));
    wf[1] = wf[2] = wf[3] = wf[4] = A(0);
    return;
  }
  A w0, wir, zs, Ath;
  if (variant == CYL_VARIANT_SIMPLE) {
    zs = gme.gz;                                               // :191
    w0 = zs * g0e.gth;                                         // :193
    wir = g0e.gz * rcp_a<A>(gme.gth);                          // :194
    Ath = gme.ath;                                             // :195
  } else {
    zs = g0r.gz;                                               // 0th:253
    w0 = g0r.gz * g0e.gth;                                     // 0th:255
    wir = w0 * rcp_a<A>(g0r.gth * g0r.gth);                    // 0th:254,256
    Ath = (variant == CYL_VARIANT_RESCALED_0TH) ? g0e.ath : gme.ath;      // 0th:257 / 1st:275
    const A area = (c.lz * g0e.gz) * (c.lth * g0e.gth);        // 0th:261, :68
    A thermal = -c.temperature;                                // 0th:66-92 / 1st:86
    if (variant == CYL_VARIANT_RESCALED_1ST)                   // 1st:87-108 (cold: evaluated in fp64)
      thermal += (A)background_correction_1st(c.alpha_d, c.C_d, c.u_d, c.temperature_d, (double)area, c.Z, c.T);
    const A bg = thermal * rcp_a<A>(area) + c.bg_const;
    wf[5] = c.Tlen * bg * w0;                                  // 0th:275 / 1st:293
  }
  wf[0] = c.alpha * w0 + c.Cn2w * (Ath * Ath) * wir;           // :199 (alpha part) + :210
  wf[1] = c.hu * w0;                                           // :199
  wf[2] = c.C * rcp_a<A>(zs * zs) * w0 * c.ilz2;               // :200-202
  wf[3] = c.C * wir * c.ilth2;                                 // :206
  wf[4] = c.twoCn * Ath * wir * c.ilth;                        // :208
}

// ------------------------------------------------------------------------------------------------------
// Robbins-Monro width update of the site moves for a batch of n proposals with k accepts (cyl_site.cuh:rm_batch_update),
// split in two: the logarithms depend only on the step counter the sweep STARTS from, so the sweep kernels take them while
// the sites are being updated and apply the accept count afterwards.  rm_batch_update == rm_apply(rm_logs()).
// ------------------------------------------------------------------------------------------------------
template <typename R> struct RmLogs { R up, dn; };
template <typename R> __device__ __forceinline__ RmLogs<R> rm_logs(double step_counter, double n);
template <> __device__ __forceinline__ RmLogs<double> rm_logs<double>(double step_counter, double n) {
  const double f = fmax(step_counter + 0.5 * n, 200.0);
  RmLogs<double> l;
  l.up = log1p(2.0 / f); l.dn = log1p(-2.0 / f);
  return l;
}
template <> __device__ __forceinline__ RmLogs<float> rm_logs<float>(double step_counter, double n) {
  const float x = 2.0f / fmaxf((float)step_counter + 0.5f * (float)n, 200.0f);
  RmLogs<float> l;
  l.up = log1pf(x); l.dn = log1pf(-x);
  return l;
}
__device__ __forceinline__ double rm_apply(double sigma, const RmLogs<double>& l, double n, double k) {
  return sigma * exp(fma(k, l.up, (n - k) * l.dn));
}
__device__ __forceinline__ double rm_apply(double sigma, const RmLogs<float>& l, double n, double k) {
  return sigma * (double)expf(__fmaf_rn((float)k, l.up, __fmul_rn((float)(n - k), l.dn)));
}
template <typename R> __device__ __forceinline__ double rm_batch_update(double sigma, double step_counter, double n, double k) {
  return rm_apply(sigma, rm_logs<R>(step_counter, n), n, k);
}
