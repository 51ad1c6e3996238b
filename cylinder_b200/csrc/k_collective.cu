// Whole-lattice ("collective") proposals of the reference's lattice class, evaluated without materialising the proposed field:
//   CYL_MOVE_SHIFT   psi + c                                   Lattice.step_lattice_all        On_lattice_simple.py:247-259
//   CYL_MOVE_WAVE    psi + c exp(i (i qz/Z + j qth/T))         Lattice.step_lattice_wavevector On_lattice_simple.py:267-300
//   CYL_MOVE_TWIST   rows [r0, r0+nr) mod Z:  psi exp(i (2 pi j n_rot/T + phase))   Lattice.step_row_rotate :303-382
// cyl_move_row_moments_* returns the row moments (cyl_energy.cuh) of the PROPOSED field -- the transform is applied to every
// value as it is loaded, so the derived arrays of the reference (psi_squared, dz, dth of the proposal) are never stored -- and
// cyl_field_energy_f64 turns them into the proposal's energy; cyl_move_apply_* commits the transform where accept[b] != 0.
// Semantics adopted where the reference as committed deviates from its own intent: the proposal's
// derivatives are complex (the reference's step_lattice_wavevector stores them in real arrays, dropping the imaginary parts)
// and a twisted block takes every row's own neighbours (the reference copies dth/dz of the first row into all rows).
#include "cyl_common.cuh"
#include "cyl_energy.cuh"

template <typename C>
__device__ __forceinline__ C move_value(const CylMove& m, C v, int i, int j, int Z, int T) {
  using R = decltype(v.x);
  if (m.kind == CYL_MOVE_SHIFT) {
    v.x += (R)m.c_re; v.y += (R)m.c_im;
  } else if (m.kind == CYL_MOVE_WAVE) {
    double s, c;
    sincos((double)i * m.qz / Z + (double)j * m.qth / T, &s, &c);        // :269-271
    v.x += (R)(m.c_re * c - m.c_im * s); v.y += (R)(m.c_re * s + m.c_im * c);
  } else if (m.kind == CYL_MOVE_TWIST) {
    int d = i - m.r0;
    if (d < 0) d += Z;
    if (d < m.nr) {
      double s, c;
      sincos(6.283185307179586476925 / T * (double)j * (double)m.n_rot + m.phase, &s, &c);   // :307,:319
      const double re = (double)v.x * c - (double)v.y * s, im = (double)v.x * s + (double)v.y * c;
      v.x = (R)re; v.y = (R)im;
    }
  }
  return v;
}

struct Cplx { double re, im; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx cexpi(double ph) { Cplx r; sincos(ph, &r.im, &r.re); return r; }

// Rows shared between warps as in row_moments_kernel (k_energy.cu).  The phase of a wave or a twist factorises into a row
// factor and a column factor exp(i j w); a thread walks its columns with a fixed stride, so it takes the column factor from a
// recurrence (one complex product per site; at most T / 32 <= 128 steps, error ~1e-14) instead of a sincos per loaded value.
template <typename R>
__global__ void move_row_moments_kernel(const typename Real<R>::complex_t* __restrict__ psi, int B, const int32_t* __restrict__ Zs,
                                        int Zmax, int T, const CylParams* __restrict__ params, const CylMove* __restrict__ moves,
                                        double* __restrict__ S) {
  using C = typename Real<R>::complex_t;
  __shared__ double red[CYL_MOM_WARPS][CYL_NMOM];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpr = moments_warps_per_row(T), rows_per_cta = CYL_MOM_WARPS / wpr;
  const int row_in_cta = warp / wpr, seg = warp - row_in_cta * wpr;
  const int i = blockIdx.x * rows_per_cta + row_in_cta;
  const int Z = Zs[b];
  double* out = (i < Zmax) ? S + ((size_t)b * Zmax + i) * CYL_NMOM : nullptr;
  double m[CYL_NMOM];
#pragma unroll
  for (int q = 0; q < CYL_NMOM; ++q) m[q] = 0.0;
  if (i < Z) {
    const CylParams p = params[b];
    const CylMove mv = moves[b];
    const double pi = 3.14159265358979323846;
    const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;
    const double inv_lz2 = 1.0 / (lz * lz), inv_lth2 = 1.0 / (lth * lth), inv_lth = 1.0 / lth;
    const int im = i == 0 ? Z - 1 : i - 1;
    const C* row = psi + ((size_t)b * Zmax + i) * T;
    const C* rowm = psi + ((size_t)b * Zmax + im) * T;
    const int j0 = seg * 32 + lane, stride = wpr * 32;
    if (mv.kind == CYL_MOVE_SHIFT) {
#pragma unroll 4
      for (int j = j0; j < T; j += stride) {
        C pv = row[j], L = rowm[j], W = row[j == 0 ? T - 1 : j - 1];
        pv.x += (R)mv.c_re; pv.y += (R)mv.c_im; L.x += (R)mv.c_re; L.y += (R)mv.c_im; W.x += (R)mv.c_re; W.y += (R)mv.c_im;
        moments_accumulate(m, pv, L, W, inv_lz2, inv_lth2, inv_lth);
      }
    } else {
      const bool wave = mv.kind == CYL_MOVE_WAVE;
      const double w = wave ? mv.qth / T : 6.283185307179586476925 / T * (double)mv.n_rot;   // column phase per site
      const double ph0 = wave ? 0.0 : mv.phase;
      Cplx cf = cexpi((double)j0 * w + ph0);                    // column factor of the thread's current column
      const Cplx step = cexpi((double)stride * w), back = cexpi(-w);
      const Cplx wrap = cexpi((double)(T - 1) * w + ph0);       // column T - 1, the left neighbour of column 0
      // wave: value + c exp(i i qz / Z) cf;   twist: value * cf on the rows of the block
      const Cplx c = {mv.c_re, mv.c_im};
      const Cplx ci = cmul(c, cexpi((double)i * mv.qz / Z)), cm = cmul(c, cexpi((double)im * mv.qz / Z));
      int d = i - mv.r0, dm = im - mv.r0;
      if (d < 0) d += Z;
      if (dm < 0) dm += Z;
      const bool tw_i = !wave && d < mv.nr, tw_m = !wave && dm < mv.nr;
#pragma unroll 2
      for (int j = j0; j < T; j += stride) {
        C pv = row[j], L = rowm[j], W = row[j == 0 ? T - 1 : j - 1];
        const Cplx cw = j == 0 ? wrap : cmul(cf, back);
        if (wave) {
          const Cplx a = cmul(ci, cf), bq = cmul(cm, cf), cq = cmul(ci, cw);
          pv.x += (R)a.re; pv.y += (R)a.im; L.x += (R)bq.re; L.y += (R)bq.im; W.x += (R)cq.re; W.y += (R)cq.im;
        } else {
          if (tw_i) {
            const Cplx a = cmul({(double)pv.x, (double)pv.y}, cf), cq = cmul({(double)W.x, (double)W.y}, cw);
            pv.x = (R)a.re; pv.y = (R)a.im; W.x = (R)cq.re; W.y = (R)cq.im;
          }
          if (tw_m) {
            const Cplx bq = cmul({(double)L.x, (double)L.y}, cf);
            L.x = (R)bq.re; L.y = (R)bq.im;
          }
        }
        moments_accumulate(m, pv, L, W, inv_lz2, inv_lth2, inv_lth);
        cf = cmul(cf, step);
      }
    }
  }
  moments_store(m, wpr, row_in_cta, seg, out, red);          // rows in [Z, Zmax) get zeros
}

template <typename R>
__global__ void move_apply_kernel(typename Real<R>::complex_t* __restrict__ psi, int B, const int32_t* __restrict__ Zs, int Zmax, int T,
                                  const CylMove* __restrict__ moves, const uint8_t* __restrict__ accept) {
  const int b = blockIdx.y;
  if (accept && !accept[b]) return;
  const int Z = Zs[b];
  const CylMove mv = moves[b];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < Z * T; s += gridDim.x * blockDim.x) {
    const int i = s / T, j = s - i * T;
    auto* q = psi + ((size_t)b * Zmax + i) * T + j;
    *q = move_value(mv, *q, i, j, Z, T);
  }
}

static int check_moves(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const char* who) {
  CYL_CHECK_ARG(psi && Z && moves && B > 0 && Zmax > 0 && T > 0, "%s: bad arguments", who);
  return CYL_OK;
}

template <typename R>
static int move_row_moments(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params, const CylMove* moves,
                            double* S, void* stream) {
  int rc = check_moves(psi, B, Z, Zmax, T, moves, "cyl_move_row_moments");
  if (rc) return rc;
  CYL_CHECK_ARG(params && S, "cyl_move_row_moments: bad arguments");
  const int rows_per_cta = CYL_MOM_WARPS / moments_warps_per_row(T);
  if (B > 65535) { cyl_set_error("cyl_move_row_moments: at most 65535 replicas per call (got %d); split the batch", B); return CYL_E_UNSUPPORTED; }
  dim3 grid((Zmax + rows_per_cta - 1) / rows_per_cta, B);
  move_row_moments_kernel<R><<<grid, CYL_MOM_WARPS * 32, 0, as_stream(stream)>>>((const typename Real<R>::complex_t*)psi, B, Z, Zmax, T, params,
                                                                         moves, S);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

template <typename R>
static int move_apply(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const uint8_t* accept, void* stream) {
  int rc = check_moves(psi, B, Z, Zmax, T, moves, "cyl_move_apply");
  if (rc) return rc;
  const int per = (Zmax * T + 255) / 256;
  if (B > 65535) { cyl_set_error("cyl_move_apply: at most 65535 replicas per call (got %d); split the batch", B); return CYL_E_UNSUPPORTED; }
  dim3 grid(per < 64 ? per : 64, B);
  move_apply_kernel<R><<<grid, 256, 0, as_stream(stream)>>>((typename Real<R>::complex_t*)psi, B, Z, Zmax, T, moves, accept);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

extern "C" int cyl_move_row_moments_f32(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                        const CylMove* moves, double* S, void* stream) {
  return move_row_moments<float>(psi, B, Z, Zmax, T, params, moves, S, stream);
}
extern "C" int cyl_move_row_moments_f64(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                        const CylMove* moves, double* S, void* stream) {
  return move_row_moments<double>(psi, B, Z, Zmax, T, params, moves, S, stream);
}
extern "C" int cyl_move_apply_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const uint8_t* accept,
                                  void* stream) {
  return move_apply<float>(psi, B, Z, Zmax, T, moves, accept, stream);
}
extern "C" int cyl_move_apply_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const uint8_t* accept,
                                  void* stream) {
  return move_apply<double>(psi, B, Z, Zmax, T, moves, accept, stream);
}
