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

// (the rows are built by cyl_rowtab.cuh:coef_row, in the kernel's working precision)

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

// u01_f32 never returns a denormal, so the flush-to-zero MUFU form needs no range fix-up
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
    const double thr = -(kacc * den);                    // accept iff log2(u) <= thr; kacc = inf => -inf => false
    // The fp32 MUFU logarithm of the same uniform (exactly representable in fp32; |error| <= 2^-22 max(1, |lg2 u|) < 6e-6)
    // decides every case that is not within 1e-4 of the threshold; only the rest evaluate the fp64 logarithm (~50 instructions).
    // The decisions are those of evaluating it always.
    const double lf = (double)lg2_fast(u01_f32(x1));
    if (lf < thr - 1e-4) return true;
    if (lf > thr + 1e-4) return false;
    return log2(u23_f64(x1)) <= thr;
  }
};

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

// ------------------------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2, PTX *.f32x2): a complex number IS a (re, im) pair, so the linear part of
// dE runs two lanes per instruction.  ptxas folds the packing moves, lane swaps, per-lane negations and scalar broadcasts
// below into operand modifiers of the packed instruction (R.F32x2.LO_HI, .NP, R.F32), so none of them costs an instruction.
// These kernels are bound by instruction issue (profiles/r02_*): the packed form of a site move is ~17 instructions shorter.
// ------------------------------------------------------------------------------------------------------
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 pk2(float x, float y) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ F2 pk2(float2 a) { return pk2(a.x, a.y); }
__device__ __forceinline__ float2 up2(F2 a) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a.v)); return r; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d; }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { F2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { F2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ F2 fma2s(float s, F2 b, F2 c) { return fma2(pk2(s, s), b, c); }       // scalar * pair + pair
__device__ __forceinline__ F2 mul2s(float s, F2 b) { return mul2(pk2(s, s), b); }
__device__ __forceinline__ F2 neg2(F2 a) { const float2 t = up2(a); return pk2(-t.x, -t.y); }
__device__ __forceinline__ float hsum2(F2 a) { const float2 t = up2(a); return t.x + t.y; }

// The same pair operations for either working precision: Pk<float> = packed FFMA2 lanes, Pk<double> = two scalars.  Used by
// code that is written once for both instantiations (the row-serial energy pass of the batched kernel).
template <typename R> struct Pk;
template <> struct Pk<float> {
  using T = F2;
  static __device__ __forceinline__ T make(float2 a) { return pk2(a); }
  static __device__ __forceinline__ T zero() { return pk2(0.0f, 0.0f); }
  static __device__ __forceinline__ T fma(T a, T b, T c) { return fma2(a, b, c); }
  static __device__ __forceinline__ T add(T a, T b) { return add2(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return sub2(a, b); }
  static __device__ __forceinline__ T rot(T a) { const float2 t = up2(a); return pk2(-t.y, t.x); }     // i a
  static __device__ __forceinline__ float hsum(T a) { return hsum2(a); }
  static __device__ __forceinline__ float2 get(T a) { return up2(a); }
};
template <> struct Pk<double> {
  struct T { double x, y; };
  static __device__ __forceinline__ T make(double2 a) { return T{a.x, a.y}; }
  static __device__ __forceinline__ T zero() { return T{0.0, 0.0}; }
  static __device__ __forceinline__ T fma(T a, T b, T c) { return T{::fma(a.x, b.x, c.x), ::fma(a.y, b.y, c.y)}; }
  static __device__ __forceinline__ T add(T a, T b) { return T{a.x + b.x, a.y + b.y}; }
  static __device__ __forceinline__ T sub(T a, T b) { return T{a.x - b.x, a.y - b.y}; }
  static __device__ __forceinline__ T rot(T a) { return T{-a.y, a.x}; }
  static __device__ __forceinline__ double hsum(T a) { return a.x + a.y; }
  static __device__ __forceinline__ double2 get(T a) { return make_double2(a.x, a.y); }
};

// dE / K of the move p -> p + d, packed (same polynomial as site_delta_over_k, regrouped so that the (re, im) lanes stay
// together until two horizontal sums):
//   quad = Re(conj(d) (csum d + 2 g)) + Re(conj(-i d) v),   g = csum p - (clz L + crz Rz + clt W + crt E),  v = xl W - xr E
//   s    = Re(conj(d) (d + 2 p))
// (Re(conj(d) i v) = Re(conj(-i d) v): the rotation sits on d, where it is a lane swap with one negation -- an operand modifier.)
__device__ __forceinline__ float site_delta_over_k_packed(const RowCoef<float>& c, float2 p, float2 L, float2 Rz, float2 W, float2 E, F2 d) {
  const F2 pp = pk2(p);
  F2 n = mul2s(c.clz, pk2(L));
  n = fma2s(c.crz, pk2(Rz), n);
  n = fma2s(c.clt, pk2(W), n);
  n = fma2s(c.crt, pk2(E), n);
  const F2 g = fma2s(c.csum, pp, neg2(n));
  F2 v = mul2s(c.xl, pk2(W));
  v = fma2s(-c.xr, pk2(E), v);
  const F2 two = pk2(2.0f, 2.0f);
  const float2 dd = up2(d);
  F2 m = mul2(d, fma2(g, two, mul2s(c.csum, d)));
  m = fma2(pk2(-dd.y, dd.x), neg2(v), m);                   // (d.y, -d.x) . v
  const float quad = hsum2(m);
  const float s = hsum2(mul2(d, fma2(pp, two, d)));
  const float p2 = __fmaf_rn(p.x, p.x, p.y * p.y);
  return __fmaf_rn(c.mass, s, __fmaf_rn(c.hu * s, __fmaf_rn(2.0f, p2, s), quad));
}

// One proposal of Lattice.step_lattice_loc at a site with value pv and neighbours L, Rz (rows above / below), W, E (columns
// left / right): draws from the Philox block x, d = sigma sgmul (z_re, z_im), dE/K, accept rule.  q = the site's value after
// the move.  EVERY sweep kernel goes through this function, so the fp32 paths take bit-identical decisions.
template <typename R, typename C>
__device__ __forceinline__ bool site_propose(const RowCoef<R>& rc, C pv, C L, C Rz, C W, C E, R sigma, uint2 x, C& q) {
  R zre, zim;
  Draw<R>::normals(x.x, x.y, zre, zim);
  const R w = sigma * rc.sgmul;
  const R dre = w * zre, dim = w * zim;
  const R den = site_delta_over_k<R, C>(rc, pv, L, Rz, W, E, dre, dim);
  const bool acc = Draw<R>::accept(x.y, den, rc.kacc);
  q = acc ? Real<R>::make(pv.x + dre, pv.y + dim) : pv;
  return acc;
}
#ifndef CYL_NO_PACKED_F32
template <>
__device__ __forceinline__ bool site_propose<float, float2>(const RowCoef<float>& rc, float2 pv, float2 L, float2 Rz, float2 W, float2 E,
                                                            float sigma, uint2 x, float2& q) {
  // Box-Muller as Draw<float>::normals, with the constant of sqrt(-2 ln u) = sqrt(2 ln 2) sqrt(-lg2 u) folded into the width
  float rad;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-lg2_fast(u01_f32(x.x))));
  float s, c;
  // angle 2 pi (k + 1/2) / 65536 from the [1,2) mantissa trick of u16_f32 with its subtraction folded into the scaling (one FFMA)
  const uint32_t k16 = __byte_perm(x.x, x.y, 0x0004) & 0xFFFFu;
  __sincosf(__fmaf_rn(__uint_as_float(0x3F800000u | (k16 << 7)), 6.283185307179586f, -6.283137370279965f), &s, &c);
  const F2 d = mul2s(rad * ((1.1774100225154747f * sigma) * rc.sgmul), pk2(c, s));
  const float den = site_delta_over_k_packed(rc, pv, L, Rz, W, E, d);
  const bool acc = Draw<float>::accept(x.y, den, rc.kacc);
  q = acc ? up2(add2(pk2(pv), d)) : pv;
  return acc;
}
#endif

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
// bit-identical to each other); the fp64 kernels in fp64 (the oracle restates exactly that).  Defined in cyl_rowtab.cuh
// (rm_logs + rm_apply).
template <typename R> __device__ __forceinline__ double rm_batch_update(double sigma, double step_counter, double n, double k);
