// Batched-replica checkerboard sweeps, lattice resident in shared memory.
//
// One CTA owns one replica for the whole launch (nsweeps sweeps): psi is read from HBM once, kept in shared
// memory, written back once; the per-row metric coefficients, the amplitude-independent sin/cos tables and the
// per-row energy weights live in shared memory too.  Per sweep:
//   1. site moves  (Lattice.step_lattice_loc, On_lattice_simple.py:598-686) colour by colour, every site once;
//      draws from Philox4x32-10 keyed by seed with counter (site, replica, sweep, tag) -- stateless, so a sweep can
//      be re-run or resumed bit-identically;
//   2. Robbins-Monro update of sigma_site from the sweep's accept count (:688-696, batched; cyl_site.cuh);
//   3. energy pass: every thread forms the five energy densities of its sites and dots them with the
//      row weights of the current amplitude and of the PROPOSED amplitude (weight difference, so the energy change of
//      the amplitude move is accumulated directly, not as a difference of two large sums); one CTA tree reduction.
//      The same pass accumulates the profiles of Lattice.measure (:121-131) by warp shuffles;
//   4. global amplitude move a' = a + sigma_amp N(0,1) on E_surf(a') + E_field(a', a) (Lattice.run :544-550 through
//      engine.step_real_group; accept rule B.2; Robbins-Monro on sigma_amp);
//   5. on accept, rebuild the row coefficients from the sin/cos tables (no trigonometry).
// The replicas are the (alpha, C, n, kappa, k) parameter grid that scripts/taskarray.sh farms out as separate CPU
// jobs in the reference.
#include "cyl_common.cuh"
#include "cyl_energy.cuh"
#include "cyl_site.cuh"
#include "cyl_surface.cuh"

#define SWEEP_NT 256
#define SWEEP_NWARP (SWEEP_NT / 32)

struct SmemScalars {
  double amp, amp_new, sigma_site, sigma_amp, e_field, e_surf, e_surf_new;
  double step_counter, site_proposed, site_accepted, amp_proposed, amp_accepted, amp_counter, flags;
  double obs[CYL_OBS_DOUBLES];      // observable sums of this launch, flushed once at the end
  int amp_accept;
};

// Optional per-phase cycle accounting (build with CYL_EXTRA_NVCC_FLAGS=-DCYL_PHASE_TIMING): thread 0 accumulates
// clock64() deltas into obs slots 10..15 = site moves, accept count, weights/surface, energy pass, decision, rebuild.
#ifdef CYL_PHASE_TIMING
#define PHASE_MARK(k) do { if (tid == 0) { const long long t__ = clock64(); sc->obs[10 + (k)] += (double)(t__ - t_phase); t_phase = t__; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

#define AMP_CHUNK 64                // amplitude-move draws are precomputed for this many sweeps at a time

// sin/cos(k z) at z_i and z_{i-1/2}
struct RowTrig { double s0, c0, sm, cm; };

// energy weights of a row: current amplitude, and (proposed - current); pixel-length factors folded in
template <typename R>
struct __align__(16) RowW { R cur[5]; R dif[5]; R pad[2]; };

template <typename R>
struct SweepSmemLayout {
  using C = typename Real<R>::complex_t;
  size_t off_coef, off_trig, off_w, off_prof, off_amp, off_scr, off_sc, total;
  __host__ __device__ SweepSmemLayout(int Zmax, int T) {
    size_t o = (size_t)Zmax * T * sizeof(C);
    o = (o + 15) & ~(size_t)15;
    off_coef = o; o += (size_t)Zmax * sizeof(RowCoef<R>);
    off_trig = o; o += (size_t)Zmax * sizeof(RowTrig);
    off_w = o;    o += (size_t)Zmax * sizeof(RowW<R>);
    off_prof = o; o += (size_t)Zmax * 3 * sizeof(R);          // per-launch profile sums, flushed once at the end
    o = (o + 15) & ~(size_t)15;
    off_amp = o;  o += 2 * AMP_CHUNK * sizeof(double);          // unit normal + accept uniform per sweep
    off_scr = o;  o += 8 * SWEEP_NWARP * sizeof(double);
    off_sc = o;   o += sizeof(SmemScalars);
    total = (o + 15) & ~(size_t)15;
  }
};

template <typename R>
__device__ __forceinline__ void build_row_coefs(RowCoef<R>* coef, const RowTrig* trig, const CylParams& p, int variant,
                                                double a, double lz, double lth, int Z) {
  const double ra = fast_radius_rescaled(p.radius, a);
  for (int i = threadIdx.x; i < Z; i += blockDim.x) {
    const RowTrig t = trig[i], tp = trig[i + 1 == Z ? 0 : i + 1];
    const Geo g0 = geo_from_sc(p.wavenumber, ra, a, t.s0, t.c0);
    const Geo gm = geo_from_sc(p.wavenumber, ra, a, t.sm, t.cm);
    const Geo gp = geo_from_sc(p.wavenumber, ra, a, tp.sm, tp.cm);      // z_{i+1/2} = z_{(i+1)-1/2}, periodic
    coef[i] = row_coef<R>(p, variant, lz, lth, g0, gm, gp);
  }
}

// Sum N doubles over the CTA (result in every thread).  scratch: N * SWEEP_NWARP doubles.
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < N; ++q) v[q] = warp_sum(v[q]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < N; ++q) scratch[q * SWEEP_NWARP + warp] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < N; ++q) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SWEEP_NWARP; ++w) t += scratch[q * SWEEP_NWARP + w];
    v[q] = t;
  }
}

// one site move; returns 1 when accepted
template <typename R, typename C>
__device__ __forceinline__ int site_move(C* psi, int idx, int im, int ip, int jm, int jp, const RowCoef<R>& rc, R sigma,
                                         uint32_t site, uint32_t c1, uint32_t key) {
  const C pv = psi[idx], L = psi[im], Rz = psi[ip], W = psi[jm], E = psi[jp];
  const uint2 x = philox2x32_10(site, c1, key);
  R zre, zim;
  Draw<R>::normals(x.x, x.y, zre, zim);
  const R w = sigma * rc.sgmul;
  const R dre = w * zre, dim = w * zim;
  const R den = site_delta_over_k<R, C>(rc, pv, L, Rz, W, E, dre, dim);
  const bool acc = Draw<R>::accept(x.y, den, rc.kacc);
  if (acc) psi[idx] = Real<R>::make(pv.x + dre, pv.y + dim);
  return acc ? 1 : 0;
}

template <typename R>
__global__ void __launch_bounds__(SWEEP_NT, (sizeof(R) == 4 ? 4 : 2))
sweep_smem_kernel(typename Real<R>::complex_t* __restrict__ psi_g, int B, const int32_t* __restrict__ Zs, int Zmax, int T,
                  const CylParams* __restrict__ params, double* __restrict__ chain, const __grid_constant__ PhiloxKeys keys,
                  uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, double* __restrict__ obs,
                  double* __restrict__ prof) {
  using C = typename Real<R>::complex_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SweepSmemLayout<R> lay(Zmax, T);
  C* psi = reinterpret_cast<C*>(smem_raw);
  RowCoef<R>* coef = reinterpret_cast<RowCoef<R>*>(smem_raw + lay.off_coef);
  RowTrig* trig = reinterpret_cast<RowTrig*>(smem_raw + lay.off_trig);
  RowW<R>* roww = reinterpret_cast<RowW<R>*>(smem_raw + lay.off_w);
  R* prof_s = reinterpret_cast<R*>(smem_raw + lay.off_prof);
  double* ampz = reinterpret_cast<double*>(smem_raw + lay.off_amp);
  double* ampu = ampz + AMP_CHUNK;
  double* scr = reinterpret_cast<double*>(smem_raw + lay.off_scr);
  SmemScalars* sc = reinterpret_cast<SmemScalars*>(smem_raw + lay.off_sc);

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Z = Zs[b];
  const int nsite = Z * T;
  const CylParams p = params[b];
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z, lth = (2 * pi * p.radius) / T;     // On_lattice_simple.py:47,:49
  const uint32_t rep = (uint32_t)(replica0 + b);
  const SiteStream stream = site_stream(seed, (uint64_t)(replica0 + b));
  double* ch = chain + (size_t)b * CYL_CHAIN_DOUBLES;
  C* psi_b = psi_g + (size_t)b * Zmax * T;
  const bool extras = (flags & (CYL_SWEEP_AMP_MOVES | CYL_SWEEP_MEASURE)) != 0;
  const bool measure = (flags & CYL_SWEEP_MEASURE) != 0;

  // ---- prologue: state, lattice, tables ----
  if (tid == 0) {
    sc->amp = ch[CYL_CH_AMPLITUDE];
    sc->sigma_site = ch[CYL_CH_SIGMA_SITE];
    sc->sigma_amp = ch[CYL_CH_SIGMA_AMP];
    sc->step_counter = ch[CYL_CH_STEP_COUNTER];
    sc->site_proposed = ch[CYL_CH_SITE_PROPOSED];
    sc->site_accepted = ch[CYL_CH_SITE_ACCEPTED];
    sc->amp_proposed = ch[CYL_CH_AMP_PROPOSED];
    sc->amp_accepted = ch[CYL_CH_AMP_ACCEPTED];
    sc->amp_counter = ch[CYL_CH_AMP_COUNTER];
    sc->flags = ch[CYL_CH_FLAGS];
    sc->e_field = ch[CYL_CH_E_FIELD];
    sc->amp_accept = 0;
    for (int q = 0; q < CYL_OBS_DOUBLES; ++q) sc->obs[q] = 0.0;
  }
  for (int s = tid; s < nsite; s += SWEEP_NT) psi[s] = psi_b[s];
  for (int s = tid; s < 3 * Z; s += SWEEP_NT) prof_s[s] = R(0);
  for (int i = tid; i < Z; i += SWEEP_NT) {
    RowTrig t;
    sincos(p.wavenumber * (lz * i), &t.s0, &t.c0);
    sincos(p.wavenumber * (lz * (i - .5)), &t.sm, &t.cm);
    trig[i] = t;
  }
  __syncthreads();
  build_row_coefs<R>(coef, trig, p, variant, sc->amp, lz, lth, Z);
  if (extras && warp == 0) {
    const double es = surface_energy_warp(p, sc->amp, nullptr, nullptr, true);
    if (lane == 0) sc->e_surf = es;
  }
  __syncthreads();

  const double amp_pt = (p.amp_target_acceptance > 0) ? p.amp_target_acceptance : 0.3;
  const double amp_ratio = 1.0 / (amp_pt * (1.0 - amp_pt));
  const int ncol = num_colours(Z, T);
  const int Th = (T + 1) >> 1, Teven = T & ~1;
  const int di = SWEEP_NT / Th, dj = SWEEP_NT % Th;
  const R inv_lz2 = (R)(1.0 / (lz * lz)), inv_lth2 = (R)(1.0 / (lth * lth)), inv_lth = (R)(1.0 / lth);

  for (int sw = 0; sw < nsweeps; ++sw) {
    const uint64_t sweep = sweep0 + (uint64_t)sw;
    if (extras && (sw & (AMP_CHUNK - 1)) == 0) {
      // draws of the next AMP_CHUNK amplitude moves, one sweep per thread (counter-based: any order, any thread)
      if (tid < AMP_CHUNK && sw + tid < nsweeps) {
        const uint64_t s2 = sweep + (uint64_t)tid;
        amp_draw(philox4x32_10(0u, rep, (uint32_t)s2, philox_c3(CYL_TAG_AMP, s2), keys), ampz[tid], ampu[tid]);
      }
      __syncthreads();
    }
    const uint32_t c1 = (uint32_t)sweep + stream.off;
    const R sigma = (R)sc->sigma_site;
    int nacc = 0;
#ifdef CYL_PHASE_TIMING
    long long t_phase = clock64();
#endif
    // ---- 1. site moves ----
    if (ncol == 2) {
      // both lengths even: colour = (i + j) & 1, T/2 sites of each colour in every row
      for (int colour = 0; colour < 2; ++colour) {
        int i = tid / Th, jj = tid - i * Th;
        while (i < Z) {
          const int j = 2 * jj + ((colour + i) & 1);
          const int row = i * T;
          const int im = (i == 0 ? nsite : row) - T, ip = (i == Z - 1 ? 0 : row + T);
          const int jm = (j == 0 ? T : j) - 1, jp = (j == T - 1 ? 0 : j + 1);
          nacc += site_move<R, C>(psi, row + j, im + j, ip + j, row + jm, row + jp, coef[i], sigma, (uint32_t)(row + j), c1,
                                  stream.key);
          jj += dj; i += di;
          if (jj >= Th) { jj -= Th; ++i; }
        }
        __syncthreads();
      }
    } else {
      // an odd ring: 3 colours, colour = (axis_colour(i) + axis_colour(j)) % 3.  With T even (Z odd) the rows that hold
      // sites of a colour are enumerable (pass 0: even rows incl. the last, pass 1: all rows, pass 2: odd rows + the
      // last), so no thread iterates over empty rows; with T odd every row is visited and filtered.
      const bool t_even = (T & 1) == 0;
      for (int colour = 0; colour < 3; ++colour) {
        const int nrow_pass = (t_even && colour != 1) ? (Z + 1) >> 1 : Z;
        int ii = tid / Th, jj = tid - ii * Th;
        while (ii < nrow_pass) {
          const int i = !t_even || colour == 1 ? ii : (colour == 0 ? 2 * ii : min(2 * ii + 1, Z - 1));
          int beta = colour - axis_colour(i, Z);
          if (beta < 0) beta += 3;
          int j;
          bool valid;
          if (beta < 2) { j = beta + 2 * jj; valid = j < Teven; }
          else          { j = T - 1;         valid = (T & 1) && jj == 0; }
          if (valid) {
            const int row = i * T;
            const int im = (i == 0 ? nsite : row) - T, ip = (i == Z - 1 ? 0 : row + T);
            const int jm = (j == 0 ? T : j) - 1, jp = (j == T - 1 ? 0 : j + 1);
            nacc += site_move<R, C>(psi, row + j, im + j, ip + j, row + jm, row + jp, coef[i], sigma, (uint32_t)(row + j), c1,
                                    stream.key);
          }
          jj += dj; ii += di;
          if (jj >= Th) { jj -= Th; ++ii; }
        }
        __syncthreads();
      }
    }
    PHASE_MARK(0);
    // ---- 2. accept count; Robbins-Monro on sigma_site (one thread, overlapped with the weights phase) ----
    const double acc_total = block_sum<double>((double)nacc, scr);
    PHASE_MARK(1);
    if (tid == SWEEP_NT - 1) {
      if (flags & CYL_SWEEP_ADAPT_SIGMA)
        sc->sigma_site = rm_batch_update<R>(sc->sigma_site, sc->step_counter, (double)nsite, acc_total);
      sc->step_counter += nsite;
      sc->site_proposed += nsite;
      sc->site_accepted += acc_total;
    }
    if (!extras) { __syncthreads(); continue; }

    // The proposal a' = a + sigma_amp N(0,1) is known to every thread: the unit normal was drawn ahead of time, a and
    // sigma_amp were fixed by the previous sweep's decision.
    const double a = sc->amp;
    const double a_new = a + sc->sigma_amp * ampz[sw & (AMP_CHUNK - 1)];
    const bool try_move = (flags & CYL_SWEEP_AMP_MOVES) && (fabs(a_new) < p.amp_bound);
    double red[6] = {0, 0, 0, 0, 0, 0};       // e_cur, e_dif, const_cur, const_dif, sum |psi|^2, sum |psi|

    // ---- 3a. row weights on warps 0..6: lane pair (2i, 2i+1) = (current, proposed amplitude) of row i;
    //          warp 7 meanwhile evaluates the surface energy of the proposal ----
    if (warp == SWEEP_NWARP - 1) {
      if (try_move) {
        const double es = surface_energy_warp(p, a_new, nullptr, nullptr, true);
        if (lane == 0) sc->e_surf_new = es;
      }
    } else {
      const double ae = (try_move && (tid & 1)) ? a_new : a;
      const double ra = fast_radius_rescaled(p.radius, a), rae = (tid & 1) ? fast_radius_rescaled(p.radius, ae) : ra;
      for (int base = 0; base < 2 * Z; base += SWEEP_NT - 32) {
        const int t = base + tid;
        const int i = min(t >> 1, Z - 1);
        const RowTrig tr = trig[i];
        const Geo g0e = geo_from_sc(p.wavenumber, rae, ae, tr.s0, tr.c0), gme = geo_from_sc(p.wavenumber, rae, ae, tr.sm, tr.cm);
        Geo g0r = g0e;                                                   // a_ref = current amplitude (:75); 0th order only
        if (variant != CYL_VARIANT_SIMPLE && (tid & 1)) g0r = geo_from_sc(p.wavenumber, ra, a, tr.s0, tr.c0);
        double W[6];
        field_energy_weights(variant, p, lz, lth, T, g0e, gme, g0r, W);
        RowW<R> w;
        double dif5 = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double other = __shfl_xor_sync(0xffffffffu, W[q], 1);    // even lane receives W(a')
          if (q < 5) { w.cur[q] = (R)W[q]; w.dif[q] = (R)(other - W[q]); }
          else dif5 = other - W[q];
        }
        if (t < 2 * Z && !(tid & 1)) {
          roww[i] = w;
          red[2] += W[5];
          red[3] += dif5;
        }
      }
    }
    __syncthreads();
    PHASE_MARK(2);

    // ---- 3b. energy pass: one warp per row ----
    auto site_terms = [&](const C pv, const C L, const C W, const RowW<R>& w, R& ec, R& ed, R& s2, R& sa, R& sre, R& sim) {
      const R p2 = pv.x * pv.x + pv.y * pv.y;
      const R zr = pv.x - L.x, zi = pv.y - L.y, tr = pv.x - W.x, ti = pv.y - W.y;
      const R m2 = p2 * p2;
      const R m3 = (zr * zr + zi * zi) * inv_lz2;
      const R m4 = (tr * tr + ti * ti) * inv_lth2;
      const R m5 = (pv.x * ti - pv.y * tr) * inv_lth;                   // Im(conj(p) dth)
      ed += w.dif[0] * p2 + w.dif[1] * m2 + w.dif[2] * m3 + w.dif[3] * m4 + w.dif[4] * m5;
      if (measure) {
        ec += w.cur[0] * p2 + w.cur[1] * m2 + w.cur[2] * m3 + w.cur[3] * m4 + w.cur[4] * m5;
        s2 += p2; sa += sqrt(p2); sre += pv.x; sim += pv.y;
      }
    };
    struct __align__(16) Pair { C a, b; };
    for (int i = warp; i < Z; i += SWEEP_NWARP) {
      const int row = i * T, rm = (i == 0 ? nsite : row) - T;
      const RowW<R> w = roww[i];
      R ec = 0, ed = 0, s2 = 0, sa = 0, sre = 0, sim = 0;
      if ((T & 1) == 0) {
        for (int pp = lane; 2 * pp < T; pp += 32) {                      // two adjacent sites per lane, 16-byte loads
          const int j = 2 * pp;
          const Pair own = *reinterpret_cast<const Pair*>(psi + row + j);
          const Pair up = *reinterpret_cast<const Pair*>(psi + rm + j);
          const C Wc = psi[row + (j == 0 ? T : j) - 1];
          site_terms(own.a, up.a, Wc, w, ec, ed, s2, sa, sre, sim);
          site_terms(own.b, up.b, own.a, w, ec, ed, s2, sa, sre, sim);
        }
      } else {
        for (int j = lane; j < T; j += 32)
          site_terms(psi[row + j], psi[rm + j], psi[row + (j == 0 ? T : j) - 1], w, ec, ed, s2, sa, sre, sim);
      }
      red[0] += (double)ec;
      red[1] += (double)ed;
      if (measure) {
        red[4] += (double)s2;
        red[5] += (double)sa;
        if (prof) {                                                      // Lattice.measure :121-131
          // three row sums by one transposed butterfly: 6 shuffles instead of 15
          R v0 = sre, v1 = sim, v2 = sa, v3 = R(0);
          {
            const bool hi = lane & 16;                                   // xor 16: upper half keeps (v2, v3)
            const R s0 = hi ? v0 : v2, s1 = hi ? v1 : v3;
            const R r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
            v0 = (hi ? v2 : v0) + r0;
            v1 = (hi ? v3 : v1) + r1;
          }
          {
            const bool hi = lane & 8;                                    // xor 8: keeps one of the two
            const R s0 = hi ? v0 : v1;
            const R r0 = __shfl_xor_sync(0xffffffffu, s0, 8);
            v0 = (hi ? v1 : v0) + r0;
          }
          v0 += __shfl_xor_sync(0xffffffffu, v0, 4);
          v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
          v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
          // lane 0: sum re, lane 8: sum im, lane 16: sum |psi|, (lane 24: 0)
          if ((lane & 7) == 0 && lane < 24) prof_s[3 * i + (lane >> 3)] += v0;
        }
      }
    }
    PHASE_MARK(3);
    block_sum_n<6>(red, scr);
    // ---- 4. measurement + amplitude decision (thread 0) ----
    if (tid == 0) {
      const double e_dif = (red[1] + red[3]) * lz * lth;
      if (measure) {                                                     // measure_avgs :135-156 + engine record
        const double e_cur = (red[0] + red[2]) * lz * lth;               // :213
        sc->e_field = e_cur;
        double* o = sc->obs;
        o[CYL_OB_COUNT] += 1.0;
        o[CYL_OB_ABS_A] += fabs(a);
        o[CYL_OB_A] += a;
        o[CYL_OB_A2] += a * a;
        o[CYL_OB_PSI2] += red[4] / nsite;
        o[CYL_OB_ABSPSI] += red[5] / nsite;
        o[CYL_OB_E_FIELD] += e_cur;
        o[CYL_OB_E_SURF] += sc->e_surf;
        o[CYL_OB_E_FIELD2] += e_cur * e_cur;
        o[CYL_OB_SIGMA] += sc->sigma_site;
      }
      int accept = 0;
      if (flags & CYL_SWEEP_AMP_MOVES) {
        if (try_move) {
          const double dE = (sc->e_surf_new - sc->e_surf) + e_dif;
          if (!isfinite(dE)) sc->flags = (double)((unsigned)sc->flags | 1u);
          if (dE <= 0) accept = 1;                                       // accept rule B.2
          else if (p.temperature == 0) accept = 0;
          else accept = (ampu[sw & (AMP_CHUNK - 1)] <= exp(-dE / p.temperature));
        }
        // Robbins-Monro on sigma_amp (engine, SURVEY.md B.3: m = 1, target p*)
        sc->amp_counter += 1;
        sc->amp_proposed += 1;
        if (flags & CYL_SWEEP_ADAPT_SIGMA) {
          const double cf = sc->sigma_amp * amp_ratio / fmax(sc->amp_counter, 200.0);
          if (accept) sc->sigma_amp += cf * (1 - amp_pt);
          else sc->sigma_amp -= cf * amp_pt;
        }
        if (accept) {
          sc->amp = a_new;
          sc->e_surf = sc->e_surf_new;
          sc->amp_accepted += 1;
        }
      }
      sc->amp_accept = accept;
    }
    __syncthreads();
    PHASE_MARK(4);
    // ---- 5. metric tables follow the amplitude ----
    if (sc->amp_accept) {
      build_row_coefs<R>(coef, trig, p, variant, sc->amp, lz, lth, Z);
      __syncthreads();
    }
    PHASE_MARK(5);
  }

  // ---- epilogue ----
  __syncthreads();
  for (int s = tid; s < nsite; s += SWEEP_NT) psi_b[s] = psi[s];
  if (measure && prof) {
    double* pr = prof + (size_t)b * Zmax * 3;
    for (int s = tid; s < 3 * Z; s += SWEEP_NT) pr[s] += (double)prof_s[s];
  }
  if (measure && obs && tid < CYL_OBS_DOUBLES) obs[(size_t)b * CYL_OBS_DOUBLES + tid] += sc->obs[tid];
  if (tid == 0) {
    ch[CYL_CH_AMPLITUDE] = sc->amp;
    ch[CYL_CH_SIGMA_SITE] = sc->sigma_site;
    ch[CYL_CH_SIGMA_AMP] = sc->sigma_amp;
    ch[CYL_CH_E_FIELD] = sc->e_field;
    if (extras) ch[CYL_CH_E_SURF] = sc->e_surf;
    ch[CYL_CH_STEP_COUNTER] = sc->step_counter;
    ch[CYL_CH_SITE_PROPOSED] = sc->site_proposed;
    ch[CYL_CH_SITE_ACCEPTED] = sc->site_accepted;
    ch[CYL_CH_AMP_PROPOSED] = sc->amp_proposed;
    ch[CYL_CH_AMP_ACCEPTED] = sc->amp_accepted;
    ch[CYL_CH_AMP_COUNTER] = sc->amp_counter;
    ch[CYL_CH_FLAGS] = sc->flags;
  }
}

template <typename R>
static int sweep_smem(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params, double* chain,
                      uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps, int variant, uint32_t flags,
                      double* obs, double* prof, void* stream) {
  CYL_CHECK_ARG(psi && Z && params && chain && B > 0 && Zmax > 0 && T > 1 && nsweeps >= 0, "cyl_sweep_smem: bad arguments");
  CYL_CHECK_ARG(variant == CYL_VARIANT_SIMPLE || variant == CYL_VARIANT_RESCALED_0TH, "cyl_sweep_smem: unknown variant %d", variant);
  CYL_CHECK_ARG(!(flags & CYL_SWEEP_MEASURE) || obs, "cyl_sweep_smem: CYL_SWEEP_MEASURE needs obs");
  if (nsweeps == 0) return CYL_OK;
  const SweepSmemLayout<R> lay(Zmax, T);
  int dev = 0, max_optin = 0;
  CYL_CUDA(cudaGetDevice(&dev));
  CYL_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (lay.total > (size_t)max_optin) {
    cyl_set_error("cyl_sweep_smem: lattice %d x %d needs %zu B of shared memory (> %d); use the HBM path", Zmax, T,
                  lay.total, max_optin);
    return CYL_E_UNSUPPORTED;
  }
  auto kern = sweep_smem_kernel<R>;
  CYL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  const PhiloxKeys keys = philox_expand(seed);
  kern<<<B, SWEEP_NT, lay.total, as_stream(stream)>>>((typename Real<R>::complex_t*)psi, B, Z, Zmax, T, params, chain, keys,
                                                      seed, replica0, sweep0, nsweeps, variant, flags, obs, prof);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

extern "C" int cyl_sweep_smem_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                  double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                  int variant, uint32_t flags, double* obs, double* prof, void* stream) {
  return sweep_smem<float>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, stream);
}
extern "C" int cyl_sweep_smem_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                  double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                  int variant, uint32_t flags, double* obs, double* prof, void* stream) {
  return sweep_smem<double>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, stream);
}
