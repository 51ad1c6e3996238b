// Whole-lattice field energy in row-moment form.
//
// Lattice.surface_field_energy (On_lattice_simple.py:185-213; 0th-order override
// On_lattice_rescaled_0thorder.py:247-278) is linear in five amplitude-independent sums per z-row,
//   S1 = sum_j |psi|^2   S2 = sum_j |psi|^4   S3 = sum_j |dz|^2   S4 = sum_j |dth|^2   S5 = sum_j Im(conj(psi) dth)
// (left differences divided by the pixel lengths, :170-182), so one O(Z*T) reduction serves any number of
// trial amplitudes at O(Z) each.  S6 = sum_j |psi|, S7 + i S8 = sum_j psi feed Lattice.measure (:121-131).
#pragma once
#include "cyl_common.cuh"

#define CYL_NMOM 8

// one site's contribution to the row moments; L = psi[i-1][j], W = psi[i][j-1]
template <typename C>
__device__ __forceinline__ void moments_accumulate(double (&m)[CYL_NMOM], C p, C L, C W, double inv_lz2, double inv_lth2,
                                                   double inv_lth) {
  const double pr = p.x, pi = p.y;
  const double p2 = pr * pr + pi * pi;
  const double zr = pr - (double)L.x, zi = pi - (double)L.y;
  const double tr = pr - (double)W.x, ti = pi - (double)W.y;
  m[0] += p2;
  m[1] += p2 * p2;
  m[2] += (zr * zr + zi * zi) * inv_lz2;
  m[3] += (tr * tr + ti * ti) * inv_lth2;
  m[4] += (pr * ti - pi * tr) * inv_lth;            // Im(conj(p) * dth)
  m[5] += sqrt(p2);
  m[6] += pr;
  m[7] += pi;
}

// fp32 storage: the per-site terms are formed in fp32 (the inputs carry no more; the differences of neighbouring values are
// exact or nearly so) and only the accumulation runs in fp64 -- a third of the fp64 instructions of the generic form.
__device__ __forceinline__ void moments_accumulate(double (&m)[CYL_NMOM], float2 p, float2 L, float2 W, double inv_lz2, double inv_lth2,
                                                   double inv_lth) {
  const float p2 = p.x * p.x + p.y * p.y;
  const float zr = p.x - L.x, zi = p.y - L.y;
  const float tr = p.x - W.x, ti = p.y - W.y;
  m[0] += (double)p2;
  m[1] += (double)(p2 * p2);
  m[2] += (double)(zr * zr + zi * zi) * inv_lz2;
  m[3] += (double)(tr * tr + ti * ti) * inv_lth2;
  m[4] += (double)(p.x * ti - p.y * tr) * inv_lth;     // Im(conj(p) * dth)
  m[5] += (double)sqrtf(p2);
  m[6] += (double)p.x;
  m[7] += (double)p.y;
}

// Row-moment kernels: a CTA of 8 warps takes 8 / wpr rows with wpr warps per row (wpr = 1, 2, 4 or 8, chosen by the row length:
// a long row walked by one warp is a chain of dependent memory latencies).  Reduction of the row's warps through shared memory.
#define CYL_MOM_WARPS 8
__host__ __device__ __forceinline__ int moments_warps_per_row(int T) { return T >= 2048 ? 8 : T >= 1024 ? 4 : T >= 512 ? 2 : 1; }
__device__ __forceinline__ void moments_store(double (&m)[CYL_NMOM], int wpr, int row_in_cta, int seg, double* __restrict__ out,
                                              double (*red)[CYL_NMOM]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < CYL_NMOM; ++q) m[q] = warp_sum(m[q]);
  if (wpr == 1) {
    if (lane == 0 && out) {
#pragma unroll
      for (int q = 0; q < CYL_NMOM; ++q) out[q] = m[q];
    }
    return;
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < CYL_NMOM; ++q) red[threadIdx.x >> 5][q] = m[q];
  }
  __syncthreads();
  if (seg == 0 && lane < CYL_NMOM && out) {
    double t = 0.0;
    for (int s = 0; s < wpr; ++s) t += red[row_in_cta * wpr + s][lane];
    out[lane] = t;
  }
}

// Row weights of the energy: E_i = W[0] S1 + W[1] S2 + W[2] S3 + W[3] S4 + W[4] S5 + W[5], at trial amplitude a_eval
// given the metric reference amplitude a_ref (only the 0th-order variant distinguishes them).  g0e / gme = metric at
// z_i / z_{i-1/2} for a_eval, g0r = metric at z_i for a_ref.  WITHOUT the global lz*lth factor.
// O(u) fluctuation correction to the thermal background of the 1st-order variant, per cell
// (On_lattice_rescaled_1storder.py:87-108; the reference's ValueError branches are its log arguments going <= 0)
__device__ __forceinline__ double background_correction_1st(double alpha, double C, double u, double temperature, double area, int Z, int T) {
  const double pi = 3.14159265358979323846;
  const double area_factor = 2.0 / sqrt(pi);                    // 1st:34
  const double upper = 1 * area_factor;                         // 1st:89
  const double lower = (1.0 / (double)max(Z, T)) * area_factor; // 1st:90
  const double alpha_cell = alpha * area;                       // 1st:92
  const double hi = C * (upper * upper) + alpha_cell;
  double lo = C * (lower * lower) + alpha_cell;
  double integral = 0.0;                                        // 1st:103-105
  if (hi > 0) {
    if (!(lo > 0)) {                                            // 1st:98-102
      const double l2 = sqrt(fabs(alpha_cell) / C) + .0000001;
      lo = C * (l2 * l2) + alpha_cell;
    }
    integral = pi / C * (log(hi) - log(lo));                    // 1st:96
  }
  double correction = u * (temperature * temperature * temperature) * 4 * (integral * integral);   // 1st:106
  correction /= (double)Z * (double)T;                          // 1st:107
  return correction;
}
__device__ __forceinline__ double background_correction_1st(const CylParams& p, double area, int Z, int T) {
  return background_correction_1st(p.alpha, p.C, p.u, p.temperature, area, Z, T);
}

__device__ __forceinline__ void field_energy_weights(int variant, const CylParams& p, double lz, double lth, int Z, int T,
                                                     const Geo& g0e, const Geo& gme, const Geo& g0r, double (&W)[6]) {
  const double Cn2 = p.C * p.n * p.n;
  double w0, wir, zs, A;
  W[5] = 0.0;
  if (variant == CYL_VARIANT_DECOUPLED) {                      // dec:179-202: only the C n^2 |A psi|^2 coupling
    W[0] = Cn2 * (gme.ath * gme.ath) * (g0e.gz * fast_rcp(gme.gth));            // dec:190-198
    W[1] = W[2] = W[3] = W[4] = 0.0;
    return;
  }
  if (variant == CYL_VARIANT_SIMPLE) {
    zs = gme.gz;                                               // :191
    w0 = zs * g0e.gth;                                         // :193
    wir = g0e.gz * fast_rcp(gme.gth);                          // :194
    A = gme.ath;                                               // :195
  } else {
    zs = g0r.gz;                                               // 0th:253
    w0 = g0r.gz * g0e.gth;                                     // 0th:255
    wir = w0 * fast_rcp(g0r.gth * g0r.gth);                    // 0th:254,256
    A = (variant == CYL_VARIANT_RESCALED_0TH) ? g0e.ath : gme.ath;      // 0th:257 / 1st:275
    const double area = (lz * g0e.gz) * (lth * g0e.gth);       // 0th:261, :68
    double thermal = -p.temperature;                           // 0th:66-92 / 1st:86
    if (variant == CYL_VARIANT_RESCALED_1ST) thermal += background_correction_1st(p, area, Z, T);   // 1st:87-108
    const double bg = thermal * fast_rcp(area) + ((p.alpha < 0) ? .5 * p.alpha * p.alpha * fast_rcp(p.u) : 0.0);
    W[5] = T * bg * w0;                                        // 0th:275 / 1st:293
  }
  W[0] = p.alpha * w0 + Cn2 * (A * A) * wir;                   // :199 (alpha part) + :210
  W[1] = p.u / 2 * w0;                                         // :199
  W[2] = p.C * fast_rcp(zs * zs) * w0;                         // :200-202
  W[3] = p.C * wir;                                            // :206
  W[4] = p.C * p.n * 2 * A * wir;                              // :208
}

__device__ __forceinline__ double field_energy_row(int variant, const CylParams& p, double lz, double lth, int Z, int T,
                                                   const Geo& g0e, const Geo& gme, const Geo& g0r, const double* S) {
  double W[6];
  field_energy_weights(variant, p, lz, lth, Z, T, g0e, gme, g0r, W);
  return W[0] * S[0] + W[1] * S[1] + W[2] * S[2] + W[3] * S[3] + W[4] * S[4] + W[5];
}

// convenience: evaluate the metric with sincos (standalone kernels; the sweep kernels use cached sin/cos tables)
__device__ __forceinline__ double field_energy_row_eval(int variant, const CylParams& p, double lz, double lth, int Z, int T, int i,
                                                        double a_eval, double a_ref, const double* S) {
  const double k = p.wavenumber, r = p.radius;
  const Geo g0e = geo_eval(k, r, a_eval, i * lz), gme = geo_eval(k, r, a_eval, (i - .5) * lz);
  const Geo g0r = geo_eval(k, r, a_ref, i * lz);
  return field_energy_row(variant, p, lz, lth, Z, T, g0e, gme, g0r, S);
}

// ------------------------------------------------------------------------------------------------------
// Energy accumulated on the fly by the sweep kernels (distributed over the sites as they are finalised)
// ------------------------------------------------------------------------------------------------------
// energy weights of a row: current amplitude, and (proposed - current); pixel-length factors folded in
// cst = the row's constant term W[5] (background energy of the rescaled variants): {current, proposed - current}; the batched
// kernel adds it once per row inside its energy pass, the HBM path keeps the constants in a side sum and leaves these 0.
template <typename R>
struct __align__(16) RowW { R cur[5]; R dif[5]; R cst[2]; };

// energy accumulators of one thread: E(a,a), E(a',a) - E(a,a), sum |psi|^2, sum |psi|
template <typename R> struct EAcc { R ec, ed, s2, sa; };
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double sqrt_fast(double x) { return sqrt(x); }

template <typename R, typename C>
__device__ __forceinline__ void energy_own(EAcc<R>& e, const RowW<R>& w, C q, bool measure, bool with_abs = true) {
  const R p2 = q.x * q.x + q.y * q.y, p4 = p2 * p2;
  e.ed += w.dif[0] * p2 + w.dif[1] * p4;
  if (measure) { e.ec += w.cur[0] * p2 + w.cur[1] * p4; e.s2 += p2; if (with_abs) e.sa += sqrt_fast(p2); }
}
// z-bond between the site `hi` of a row and the site `lo` of the row above it; (wc, wd) = that row's |dz|^2 weights
template <typename R, typename C>
__device__ __forceinline__ void energy_zbond(EAcc<R>& e, R wc, R wd, C hi, C lo, bool measure) {
  const R dr = hi.x - lo.x, di = hi.y - lo.y, m = dr * dr + di * di;
  e.ed += wd * m;
  if (measure) e.ec += wc * m;
}
// theta-bond between the site `hi` (column j) and `lo` (column j-1) of one row
template <typename R, typename C>
__device__ __forceinline__ void energy_tbond(EAcc<R>& e, const RowW<R>& w, C hi, C lo, bool measure) {
  const R tr = hi.x - lo.x, ti = hi.y - lo.y;
  const R m4 = tr * tr + ti * ti, m5 = hi.x * ti - hi.y * tr;        // |dth|^2, Im(conj(psi) dth)
  e.ed += w.dif[3] * m4 + w.dif[4] * m5;
  if (measure) e.ec += w.cur[3] * m4 + w.cur[4] * m5;
}

