// The single-site Metropolis move of Lattice.step_lattice_loc in the form the parallel kernels use.
//
// With d = psi' - psi, every term of the reference's dE (On_lattice_simple.py:620-669; 0th-order variant
// On_lattice_rescaled_0thorder.py:142-192) is a polynomial in d whose coefficients depend only on the z-row:
//
//   dE = K [ mass*s + (u/2) s (2|p|^2 + s) + csum |d|^2 + Re(conj(d) H) ]
//     s    = |p+d|^2 - |p|^2
//     H    = 2 (csum p - clz L - crz Rz - clt W - crt E)  +  i xl W  -  i xr E        (complex)
//     L,Rz = psi[i-1,j], psi[i+1,j]     W,E = psi[i,j-1], psi[i,j+1]
//
//   simple:   clz = C/(lz^2 Gz(z-)^2)        crz = C/(lz^2 Gz(z+)^2)             (:633,:643)
//             clt = C/(lth^2 Gth(z-)^4)      crt = C/(lth^2 Gth(z+)^4)           (:634,:644)
//             xl  = 2Cn A(z-)/(lth Gth(z-)^2) xr = 2Cn A(z+)/(lth Gth(z+)^2)     (:651-663)
//             mass= alpha + C n^2 A(z-)^2/Gth(z-)^2                              (:621,:665)
//   0thorder: clz = C/(lz^2 Gz(z-)^2) sg(z-)/sg(z)   crz likewise with z+        (0th:153,:163)
//             clt = crt = C/(lth^2 Gth(z)^2)                                     (0th:154,:164)
//             xl = xr = 2Cn A(z)/(lth Gth(z)^2); mass = alpha + C n^2 A(z)^2/Gth(z)^2 (0th:168-188)
//             proposal width sigma*sg(z)                                         (0th:135-136)
//   1storder: clz = C/(lz^2 Gz(z-)^2)        crz = C/(lz^2 Gz(z+)^2)             (1st:171,:181)
//             clt = C/(lth^2 Gth(z-)^2)      crt = C/(lth^2 Gth(z+)^2)           (1st:172,:182)
//             xl  = 2Cn A(z-)/(lth Gth(z-)^3) xr = 2Cn A(z+)/(lth Gth(z+)^3)     (1st:190-203)
//             mass as simple (1st:205); proposal width sigma*sg(z) (1st:153-154)
//   decoupled: as 1storder with clz = crz = C/lz^2 (dec:294,:304) and xl,xr = 2Cn A(z-+)/(lth Gth(z-+)^2) (dec:312-324)
//   K = sg(z) lz lth (:668-669), z = lz*i, z-+ = lz*(i -+ 1/2).
//
// This is algebraically identical to the reference expression (verified against the oracle's verbatim
// restatement in tests/test_oracle_checkerboard.py) and numerically better conditioned than differencing two
// large energies, which matters for the fp32 instantiation.  The verbatim operation order lives in k_replay.cu.
#pragma once
#include "cyl_common.cuh"

template <typename R>
struct __align__(16) RowCoef {
  R kacc;    // K * log2(e) / T  (accept test in log2 domain); +inf when T == 0
  R mass;
  R csum;    // clz + crz + clt + crt
  R hu;      // u / 2
  R clz, crz, clt, crt;
  R xl, xr;
  R sgmul;   // proposal width multiplier (1 or sg)
  R kfull;   // K (for energy bookkeeping)
};

// Build the coefficients of a row from the metric at z = lz*i (g0) and z -+ lz/2 (gm, gp), all at the current
// amplitude.  The metric comes in fp64; the coefficient arithmetic runs in A = double for the fp64 kernels and in
// A = float for the fp32 kernels (the results are stored as fp32 anyway; both fp32 kernels share this code, so their
// tables are bit-identical) -- the rebuild after an accepted amplitude move is a latency-bound phase on Z threads.
template <typename A> __device__ __forceinline__ A coef_rcp(A x);
template <> __device__ __forceinline__ double coef_rcp<double>(double x) { return fast_rcp(x); }
template <> __device__ __forceinline__ float coef_rcp<float>(float x) { return __frcp_rn(x); }

template <typename R>
__device__ __forceinline__ RowCoef<R> row_coef(const CylParams& p, int variant, double lz, double lth, const Geo& g0,
                                               const Geo& gm, const Geo& gp) {
  using A = R;
  const A g0t = (A)g0.gth, g0z = (A)g0.gz, g0a = (A)g0.ath;
  const A gmt = (A)gm.gth, gmz = (A)gm.gz, gma = (A)gm.ath;
  const A gpt = (A)gp.gth, gpz = (A)gp.gz, gpa = (A)gp.ath;
  const A C = (A)p.C, n = (A)p.n, alpha = (A)p.alpha;
  const A sg = g0t * g0z;
  const A Cn2 = C * n * n;
  const A ilz2 = (A)(1.0 / (lz * lz)), ilth = (A)(1.0 / lth);
  const A igzm = coef_rcp<A>(gmz), igzp = coef_rcp<A>(gpz);
  A clz, crz, clt, crt, xl, xr, mass, sgmul;
  if (variant == CYL_VARIANT_SIMPLE) {
    const A igtm = coef_rcp<A>(gmt), igtp = coef_rcp<A>(gpt);
    const A Rm = igtm * igtm, Rp = igtp * igtp;
    clz = C * ilz2 * igzm * igzm;
    crz = C * ilz2 * igzp * igzp;
    clt = C * Rm * Rm * ilth * ilth;
    crt = C * Rp * Rp * ilth * ilth;
    xl = 2 * C * n * Rm * gma * ilth;
    xr = 2 * C * n * Rp * gpa * ilth;
    mass = alpha + Cn2 * Rm * gma * gma;
    sgmul = A(1);
  } else if (variant == CYL_VARIANT_RESCALED_1ST || variant == CYL_VARIANT_DECOUPLED) {
    const bool first = variant == CYL_VARIANT_RESCALED_1ST;
    const A igtm = coef_rcp<A>(gmt), igtp = coef_rcp<A>(gpt);
    const A Rm = igtm * igtm, Rp = igtp * igtp;
    clz = first ? C * ilz2 * igzm * igzm : C * ilz2;
    crz = first ? C * ilz2 * igzp * igzp : C * ilz2;
    clt = C * Rm * ilth * ilth;
    crt = C * Rp * ilth * ilth;
    xl = 2 * C * n * Rm * gma * ilth * (first ? igtm : A(1));
    xr = 2 * C * n * Rp * gpa * ilth * (first ? igtp : A(1));
    mass = alpha + Cn2 * Rm * gma * gma;
    sgmul = sg;
  } else {
    const A igt0 = coef_rcp<A>(g0t);
    const A R0 = igt0 * igt0, isg = igt0 * coef_rcp<A>(g0z);
    clz = C * ilz2 * igzm * gmt * isg;                      // C/(lz^2 Gz(z-)^2) * sg(z-)/sg(z)
    crz = C * ilz2 * igzp * gpt * isg;
    clt = crt = C * R0 * ilth * ilth;
    xl = xr = 2 * C * n * R0 * g0a * ilth;
    mass = alpha + Cn2 * R0 * g0a * g0a;
    sgmul = sg;
  }
  const A K = sg * (A)(lz * lth);
  RowCoef<R> c;
  c.kfull = (R)K;
  c.kacc = (p.temperature > 0) ? (R)(K * (A)(1.4426950408889634 / p.temperature)) : (R)INFINITY;
  c.mass = (R)mass;
  c.csum = (R)(clz + crz + clt + crt);
  c.hu = (R)(p.u / 2);
  c.clz = (R)clz; c.crz = (R)crz; c.clt = (R)clt; c.crt = (R)crt;
  c.xl = (R)xl; c.xr = (R)xr;
  c.sgmul = (R)sgmul;
  return c;
}

// dE / K for the move p -> p + d.
template <typename R, typename C>
__device__ __forceinline__ R site_delta_over_k(const RowCoef<R>& c, C p, C L, C Rz, C W, C E, R dre, R dim) {
  const R gre = c.csum * p.x - (c.clz * L.x + c.crz * Rz.x + c.clt * W.x + c.crt * E.x);
  const R gim = c.csum * p.y - (c.clz * L.y + c.crz * Rz.y + c.clt * W.y + c.crt * E.y);
  const R hre = R(2) * gre - c.xl * W.y + c.xr * E.y;
  const R him = R(2) * gim + c.xl * W.x - c.xr * E.x;
  const R d2 = dre * dre + dim * dim;
  const R p2 = p.x * p.x + p.y * p.y;
  const R s = d2 + R(2) * (dre * p.x + dim * p.y);
  return c.mass * s + c.hu * s * (R(2) * p2 + s) + c.csum * d2 + (dre * hre + dim * him);
}

// ------------------------------------------------------------------------------------------------------
// random draws for one proposal from one Philox2x32-10 block (x0, x1)  [bit fields: cyl_common.cuh]:
//   x0[31:9] -> Box-Muller radius, (x0[7:0] : x1[7:0]) -> Box-Muller angle (65536 directions, symmetric under
//   theta -> theta + pi, so the proposal stays symmetric), x1[31:9] -> acceptance uniform.
// Both precisions see the SAME uniforms (they are exactly representable in fp32).
// fp64: full-precision log / sincospi (the oracle restates exactly this, tests compare at 1e-11);
// fp32: MUFU lg2 / sqrt / sin / cos; uniforms built by bit tricks (no I2F).
// ------------------------------------------------------------------------------------------------------
template <typename R> struct Draw;

template <> struct Draw<double> {
  static __device__ __forceinline__ void normals(uint32_t x0, uint32_t x1, double& zre, double& zim) {
    const double rad = sqrt(-2.0 * log(u23_f64(x0)));
    double s, c;
    sincospi(2.0 * u16_f64(angle16(x0, x1)), &s, &c);
    zre = rad * c;
    zim = rad * s;
  }
  // accept iff dE <= 0, or T > 0 and u <= exp(-dE/T); den = dE/K, kacc = K log2(e)/T
  static __device__ __forceinline__ bool accept(uint32_t x1, double den, double kacc) {
    if (den <= 0.0) return true;
    return log2(u23_f64(x1)) <= -(kacc * den);          // kacc = inf => -inf => false
  }
};

// u01_f32 never returns a denormal, so the flush-to-zero MUFU form needs no range fix-up
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <> struct Draw<float> {
  static __device__ __forceinline__ void normals(uint32_t x0, uint32_t x1, float& zre, float& zim) {
    float rad;                                                                // sqrt(-2 ln u): MUFU.LG2, MUFU.SQRT
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * lg2_fast(u01_f32(x0))));
    float s, c;
    __sincosf(6.283185307179586f * u16_f32(__byte_perm(x0, x1, 0x0004) & 0xFFFFu), &s, &c);   // = angle16(x0, x1)
    zre = rad * c;
    zim = rad * s;
  }
  static __device__ __forceinline__ bool accept(uint32_t x1, float den, float kacc) {
    return (den <= 0.0f) | (lg2_fast(u01_f32(x1)) <= -(kacc * den));          // branch-free; kacc = inf when T = 0
  }
};

// draws of one global amplitude move from one Philox4x32-10 block: (x.x, x.y) -> unit normal, x.z -> acceptance uniform
__device__ __forceinline__ void amp_draw(uint4 x, double& z, double& u) {
  double s, c;
  sincospi(2.0 * u01_f64(x.y), &s, &c);
  z = sqrt(-2.0 * log(u01_f64(x.x))) * c;
  u = u01_f64(x.z);
}

// ------------------------------------------------------------------------------------------------------
// colouring.  A periodic ring of even length is 2-colourable (parity); an odd ring needs a third colour for its
// last site.  axis colour a(i) = i & 1, except a(len-1) = 2 when len is odd; site colour = (a(i) + a(j)) % ncol
// with ncol = 2 if both lengths are even, else 3.  Neighbours differ in exactly one axis colour, so they never
// share a site colour (Latin-square argument), including across the periodic seam.
// ------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int axis_colour(int i, int len) { return ((len & 1) && i == len - 1) ? 2 : (i & 1); }
__host__ __device__ __forceinline__ int num_colours(int Z, int T) { return ((Z | T) & 1) ? 3 : 2; }

// Robbins-Monro width update applied once for a batch of n proposals with k accepts and a frozen step factor
// f = max(step_counter + n/2, 200): the reference's per-proposal rule (On_lattice_simple.py:688-696, m = 1, p* = .5,
// ratio = 4) is sigma *= (1 + 2/f) on accept, (1 - 2/f) on reject; the batch applies the product.  The fp32 kernels
// evaluate the exponent in fp32 (sigma is a tuning width, and both fp32 kernels share this code so they stay
// bit-identical to each other); the fp64 kernels in fp64 (the oracle restates exactly that).
template <typename R> __device__ __forceinline__ double rm_batch_update(double sigma, double step_counter, double n, double k);
template <> __device__ __forceinline__ double rm_batch_update<double>(double sigma, double step_counter, double n, double k) {
  const double f = fmax(step_counter + 0.5 * n, 200.0);
  return sigma * exp(k * log1p(2.0 / f) + (n - k) * log1p(-2.0 / f));
}
template <> __device__ __forceinline__ double rm_batch_update<float>(double sigma, double step_counter, double n, double k) {
  const float x = 2.0f / fmaxf((float)step_counter + 0.5f * (float)n, 200.0f);
  return sigma * (double)expf((float)k * log1pf(x) + (float)(n - k) * log1pf(-x));
}
