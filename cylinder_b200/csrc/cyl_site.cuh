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
// bit-identical to each other); the fp64 kernels in fp64 (the oracle restates exactly that).  Defined in cyl_rowtab.cuh
// (rm_logs + rm_apply).
template <typename R> __device__ __forceinline__ double rm_batch_update(double sigma, double step_counter, double n, double k);
