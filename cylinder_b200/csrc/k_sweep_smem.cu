// Batched-replica checkerboard sweeps, lattice resident in shared memory.
//
// One CTA owns one replica for the whole launch (nsweeps sweeps): psi is read from HBM once, kept in shared
// memory, written back once; the per-row site-move coefficients, the amplitude-independent sin/cos tables and the
// per-row energy weights live in shared memory too.  The replicas are the (alpha, C, n, kappa, k) parameter grid that
// scripts/taskarray.sh farms out as separate CPU jobs in the reference.
//
// A sweep is organised so that only FOUR CTA-wide barriers lie on its critical path (colour 0, colour 1, energy, decision;
// five with three colours, one more when an amplitude move is accepted) and no phase leaves the CTA to one or two threads
// for long:
//   T. proposal tables.  a' = a + sigma_amp N(0,1) is known up front (its draws are precomputed, counter-based).  Warps 0..6
//      build the row weights of the field energy for a' (one row per thread, whole warps counted down from warp 6) and store
//      them beside the current ones.  No barrier follows: nothing here is read before the energy pass, so every warp goes
//      straight on to its sites.  (E_surf(a') and the Robbins-Monro logarithms, which only the decision reads, are evaluated by
//      the last warp during the energy pass, where it holds no rows.)
//   1. site moves (Lattice.step_lattice_loc, On_lattice_simple.py:598-686) colour by colour, every site once; draws from
//      Philox2x32-10 at counter (site, sweep) -- stateless, so a sweep can be re-run or resumed bit-identically.  For even T
//      the loops are column-mapped: a thread keeps its column pair and walks down the rows with an even stride, so colour
//      parity, theta wrap and the neighbour offsets are loop constants and site pointer, coefficient pointer and Philox
//      counter advance by constants (2 and 3 colours).
//   E. energy pass over the swept lattice: per column pair the five moments (|psi|^2, |psi|^4, |dz|^2, |dth|^2,
//      Im(conj(psi) dth)) of the two sites with their bonds upwards and to the left, dotted with the row
//      weights for a and for (a' - a) -- the energy change of the amplitude move is accumulated directly, not as a difference
//      of two large sums; the rows' constant terms (background energy of the rescaled variants) ride in the same sums.  Then,
//      with CYL_SWEEP_MEASURE and a profile buffer, the row sums of Lattice.measure (:121-131): two threads per row, each half
//      a row of 16-byte loads.  Warp-level sums of {accepts, E(a,a), E(a',a) - E(a,a), sum |psi|^2, sum |psi|}.
//   D. decisions on three threads of different warps, each a few dozen instructions: Robbins-Monro on sigma_site
//      (On_lattice_simple.py:688-696, batched; its logarithms were taken in phase T), the measurement record, and the
//      amplitude decision (Lattice.run :544-550 through engine.step_real_group; accept rule B.2 as T ln u <= -dE with T ln u
//      precomputed; Robbins-Monro on sigma_amp).
//   R. only when the move was accepted: coefficient rows and current weights of the new amplitude from the sin/cos tables
//      (cyl_rowtab.cuh, working precision, no trigonometry), all eight warps, one barrier.
#include <cstdlib>
#include <type_traits>
#include "cyl_common.cuh"
#include "cyl_energy.cuh"
#include "cyl_site.cuh"
#include "cyl_rowtab.cuh"
#include "cyl_surface.cuh"

#define SWEEP_NT 256
#define SWEEP_NWARP (SWEEP_NT / 32)

template <typename R>
struct SmemScalars {
  RepConst<R> rc;
  double amp, sigma_site, sigma_amp, e_field, e_surf, e_surf_new;
  double step_counter, site_proposed, site_accepted, amp_proposed, amp_accepted, amp_counter, flags;
  double obs[CYL_OBS_DOUBLES];      // observable sums of this launch, flushed once at the end
  RmLogs<R> rm;                     // logarithms of this sweep's Robbins-Monro factors (phase T -> phase D)
  int amp_accept;
  CylParams p;                      // the replica's parameters: read where they are used (fixed shared-memory address) instead of
                                    // a 96-byte copy per thread that the register allocator has to carry or spill
  alignas(16) uint32_t keys[12];    // the ten Philox round keys of this replica's site stream: three loads per colour pass
};                                  // instead of re-deriving the Weyl sequence (18 instructions) in every warp

// Optional per-phase cycle accounting (build with CYL_EXTRA_NVCC_FLAGS=-DCYL_PHASE_TIMING): thread 0 accumulates
// clock64() deltas into obs slots 10..15 = proposal tables, site moves, energy + profiles + reduction, decision, rebuild.
#ifdef CYL_PHASE_TIMING
#define PHASE_MARK(k) do { if (tid == 0) { const long long t__ = clock64(); sc->obs[10 + (k)] += (double)(t__ - t_phase); t_phase = t__; } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

// one-off ablation switches (scripts/build_variant.sh ... -DSMEM_ABL_NO_ENERGY etc.); the product build has them all on
#ifdef SMEM_ABL_NO_ENERGY
#define SMEM_ENERGY_ON false
#else
#define SMEM_ENERGY_ON true
#endif
#ifdef SMEM_ABL_NO_WEIGHTS
#define SMEM_WEIGHTS_ON false
#else
#define SMEM_WEIGHTS_ON true
#endif
#define SWEEP_FLAG_LONGEST_FIRST 0x80000000u   // set by the launcher, not part of the ABI: see the kernel prologue
#define AMP_CHUNK 32                // amplitude-move draws are precomputed for this many sweeps at a time

// Energy weights of a row in this kernel: for the current amplitude, and for the proposal (their DIFFERENCE is taken where
// it is used, in working precision, so the chain samples exactly the model with the rounded weights); cst = the rows'
// constant terms (background energy of the rescaled variants) for {current, proposal}.
template <typename R>
struct __align__(16) RowWS { R cur[5]; R nxt[5]; R cst[2]; };

// Shared memory of one CTA.  The chain scalars, the reduction scratch and the amplitude draws sit at FIXED offsets at the front
// (their addresses are compile-time constants relative to the window: before, every use re-derived the variable offsets from
// Zmax and T, a dozen instructions each time); the lattice and the row tables follow, at offsets the host computes once and
// passes as a kernel parameter (constant-bank operands).
template <typename R>
struct SweepSmemLayout {
  using C = typename Real<R>::complex_t;
  static constexpr uint32_t off_sc = 0;
  static constexpr uint32_t off_scr = (uint32_t)((sizeof(SmemScalars<R>) + 15) & ~(size_t)15);
  static constexpr uint32_t off_amp = off_scr + 8 * SWEEP_NWARP * (uint32_t)sizeof(double);   // unit normal + T ln(u) per sweep
  static constexpr uint32_t off_psi = off_amp + 2 * AMP_CHUNK * (uint32_t)sizeof(double);
  uint32_t off_coef, off_trig, off_w, off_prof, total;
  __host__ __device__ SweepSmemLayout(int Zmax, int T) {
    size_t o = off_psi + (size_t)Zmax * T * sizeof(C);
    o = (o + 15) & ~(size_t)15;
    off_coef = (uint32_t)o; o += (size_t)Zmax * sizeof(RowCoef<R>);
    off_trig = (uint32_t)o; o += (size_t)Zmax * sizeof(RowTrig<R>);
    off_w = (uint32_t)o;    o += (size_t)Zmax * sizeof(RowWS<R>);
    off_prof = (uint32_t)o; o += (size_t)Zmax * 3 * sizeof(R);          // per-launch profile sums, flushed once at the end
    total = (uint32_t)((o + 15) & ~(size_t)15);
  }
};

template <typename R>
__device__ __forceinline__ bool weights_need_ref(int variant) { return variant == CYL_VARIANT_RESCALED_0TH || variant == CYL_VARIANT_RESCALED_1ST; }

// Tables of an accepted amplitude a (the accepted amplitude becomes the metric reference too): coefficient row i on thread i,
// the current weights of row i on thread NT/2 + i, i.e. on other warps, so the two chains run side by side (this phase has one
// barrier behind it and nothing else to do).  Rows fill whole warps: the kernel is bound by instruction issue, and a warp with
// eight active lanes issues as many instructions as a full one.
// NOT inlined on purpose: the prologue of a launch and the accepted move inside a launch must produce bit-identical tables
// (a run of 40 sweeps equals two runs of 20: tests/test_gpu_fullsize.py), and two inlined copies of a floating-point
// expression are only equal as long as the compiler happens to contract them alike.  The proposal's weights (phase T) feed
// nothing but the energy difference of that proposal, so they are never adopted as the current ones.
template <typename R>
__device__ __noinline__ void rebuild_rows(RowCoef<R>* coef, RowWS<R>* roww, const RowTrig<R>* trig, const RepConst<R>* rcp,
                                          double a, int Z, bool weights) {
  const RepConst<R>& rc = *rcp;
  const AmpGeo<R> ag = amp_geo<R>(rc, a);
  const int tid = threadIdx.x;
  const bool split = 2 * Z <= SWEEP_NT;                                  // weights on the upper half of the CTA
  for (int i = tid; i < Z; i += SWEEP_NT) {
    const RowTrig<R> t = trig[i], tp = trig[i + 1 == Z ? 0 : i + 1];
    const GeoT<R> g0 = geo_sc<R>(ag, t.s0, t.c0), gm = geo_sc<R>(ag, t.sm, t.cm);
    const GeoT<R> gp = geo_sc<R>(ag, tp.sm, tp.cm);                  // z_{i+1/2} = z_{(i+1)-1/2}, periodic
    coef[i] = coef_row<R>(rc, g0, gm, gp);
  }
  if (weights) {
    for (int i = split ? tid - SWEEP_NT / 2 : tid; i >= 0 && i < Z; i += SWEEP_NT) {
      const RowTrig<R> t = trig[i];
      const GeoT<R> g0 = geo_sc<R>(ag, t.s0, t.c0), gm = geo_sc<R>(ag, t.sm, t.cm);
      R wf[6];
      weight_row<R>(rc, g0, gm, g0, wf);
      RowWS<R> w;
#pragma unroll
      for (int q = 0; q < 5; ++q) { w.cur[q] = wf[q]; w.nxt[q] = wf[q]; }
      w.cst[0] = wf[5]; w.cst[1] = wf[5];
      roww[i] = w;
    }
  }
}

__device__ __forceinline__ double surface_energy_of(const CylParams& p, double a) { return surface_energy_warp(p, a, nullptr, nullptr, true); }

// The five moments of the field energy (sum |psi|^2, sum |psi|^4, sum |dz psi|^2, sum |dth psi|^2,
// sum Im(conj(psi) dth psi); differences not divided by the pixel lengths) and the profile sums of Lattice.measure
// (On_lattice_simple.py:121-131: sum psi, sum |psi|) over the column pairs [p0, p1) of ONE row, walking along theta: the left
// neighbour is the previous iteration's site, the row above comes by 16-byte loads like the row itself, and the row weights
// are applied once per half row by the caller instead of once per site.  (re, im) lanes stay packed (FFMA2) until the end.
template <typename R, typename C, bool MEAS>
__device__ __forceinline__ void half_row_moments(const C* __restrict__ row, const C* __restrict__ up, int p0, int p1, int T,
                                                 R (&S)[5], R& sre, R& sim, R& sabs) {
  using P = Pk<R>;
  using V = typename P::T;
  struct __align__(16) Quad { C a, b; };
  const Quad* rq = reinterpret_cast<const Quad*>(row);
  const Quad* uq = reinterpret_cast<const Quad*>(up);
  V wl = P::make(row[(p0 == 0 ? T : 2 * p0) - 1]);
  R s2 = 0, s4 = 0, sa = 0;
  V dz = P::zero(), dt = P::zero(), im = P::zero(), sp = P::zero();
#pragma unroll 2
  for (int k = p0; k < p1; ++k) {
    const Quad q = rq[k], l = uq[k];
    const V A = P::make(q.a), B = P::make(q.b);
    const R p2a = q.a.x * q.a.x + q.a.y * q.a.y, p2b = q.b.x * q.b.x + q.b.y * q.b.y;
    s2 += p2a + p2b;
    s4 += p2a * p2a + p2b * p2b;
    if (MEAS) { sa += sqrt_fast(p2a) + sqrt_fast(p2b); sp = P::add(sp, P::add(A, B)); }
    const V za = P::sub(A, P::make(l.a)), zb = P::sub(B, P::make(l.b));
    dz = P::fma(za, za, dz); dz = P::fma(zb, zb, dz);
    const V ta = P::sub(A, wl), tb = P::sub(B, A);
    dt = P::fma(ta, ta, dt); dt = P::fma(tb, tb, dt);
    im = P::fma(P::rot(A), ta, im); im = P::fma(P::rot(B), tb, im);   // (-a.y, a.x) . t = Im(conj(a) t)
    wl = B;
  }
  S[0] = s2; S[1] = s4; S[2] = P::hsum(dz); S[3] = P::hsum(dt); S[4] = P::hsum(im);
  const C spc = P::get(sp);
  sre = spc.x; sim = spc.y; sabs = sa;
}

// one site move; returns the accept flag, `q` = the site's value after the move
template <typename R, typename C>
__device__ __forceinline__ bool site_move(C* psi, int idx, const C L, const C Rz, const C W, const C E, const RowCoef<R>& rc,
                                          R sigma, uint32_t site, uint32_t c1, uint32_t key, C& q, int& nacc) {
  const C pv = psi[idx];
  const uint2 x = philox2x32_10(site, c1, key);
  const bool acc = site_propose<R, C>(rc, pv, L, Rz, W, E, sigma, x, q);
  if (acc) { psi[idx] = q; ++nacc; }
  return acc;
}

// EXTRAS = amplitude moves and/or measurement requested: a compile-time switch so that the plain-sweep instantiation
// does not carry the registers of the energy accumulation.
#ifndef SMEM_MINB_EXTRAS
#define SMEM_MINB_EXTRAS 4
#endif
// TT = the row length as a compile-time constant (50: the reference's default dims, every lattice of the benchmark grid), or 0
// = any row length (run-time T_arg).  With TT known the neighbour offsets in z, the row-table and Philox-counter strides and the
// thread map are immediates.
template <typename R, bool EXTRAS, bool SERIES, int TT>
__global__ void __launch_bounds__(SWEEP_NT, (sizeof(R) == 4 ? (EXTRAS ? SMEM_MINB_EXTRAS : 4) : 2))
sweep_smem_kernel(typename Real<R>::complex_t* __restrict__ psi_g, int B, const int32_t* __restrict__ Zs, int Zmax, int T_arg,
                  const CylParams* __restrict__ params, double* __restrict__ chain, const __grid_constant__ PhiloxKeys keys,
                  uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, double* __restrict__ obs,
                  double* __restrict__ prof, double* __restrict__ series, const __grid_constant__ SweepSmemLayout<R> lay) {
  using C = typename Real<R>::complex_t;
  const int T = TT ? TT : T_arg;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* psi = reinterpret_cast<C*>(smem_raw + lay.off_psi);
  RowCoef<R>* coef = reinterpret_cast<RowCoef<R>*>(smem_raw + lay.off_coef);
  RowTrig<R>* trig = reinterpret_cast<RowTrig<R>*>(smem_raw + lay.off_trig);
  RowWS<R>* roww = reinterpret_cast<RowWS<R>*>(smem_raw + lay.off_w);
  R* prof_s = reinterpret_cast<R*>(smem_raw + lay.off_prof);
  double* ampz = reinterpret_cast<double*>(smem_raw + lay.off_amp);
  double* amplu = ampz + AMP_CHUNK;
  double* scr = reinterpret_cast<double*>(smem_raw + lay.off_scr);
  SmemScalars<R>* sc = reinterpret_cast<SmemScalars<R>*>(smem_raw + lay.off_sc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int b = blockIdx.x;
  if (flags & SWEEP_FLAG_LONGEST_FIRST) {
    // Blocks are dispatched in index order and a replica's run time grows with its row count, so block j takes the replica
    // of rank j in descending row count (ties in index order): the replicas still running when the grid drains are the
    // short ones.  Every CTA finds its replica by itself -- a histogram of Z, then the k-th replica of its bin -- in scratch
    // borrowed from the lattice region; no buffer, no extra launch, and the results do not depend on the assignment.
    int* hist = reinterpret_cast<int*>(smem_raw + lay.off_psi);            // (Zmax + 1) ints <= Zmax * T * sizeof(C)
    int* res = reinterpret_cast<int*>(scr);                                // [0] bin, [1] rank inside the bin, [2] replica, [8..15] warp sums
    for (int z = tid; z <= Zmax; z += SWEEP_NT) hist[z] = 0;
    __syncthreads();
    for (int r = tid; r < B; r += SWEEP_NT) atomicAdd(&hist[min(max(Zs[r], 0), Zmax)], 1);
    __syncthreads();
    if (tid == 0) {
      int cum = 0, z = Zmax;
      for (; z > 0; --z) {
        if (cum + hist[z] > (int)blockIdx.x) break;
        cum += hist[z];
      }
      res[0] = z; res[1] = (int)blockIdx.x - cum; res[2] = (int)blockIdx.x;
    }
    __syncthreads();
    const int zbin = res[0], k = res[1];
    const int chunk = (B + SWEEP_NT - 1) / SWEEP_NT, r0 = min(B, tid * chunk), r1 = min(B, r0 + chunk);
    int cnt = 0;
    for (int r = r0; r < r1; ++r) cnt += (min(max(Zs[r], 0), Zmax) == zbin);
    int incl = cnt;                                                        // inclusive scan over the CTA
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) res[8 + warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += res[8 + w];
    const int excl = before + incl - cnt;
    if (k >= excl && k < excl + cnt) {
      int c = excl;
      for (int r = r0; r < r1; ++r)
        if (min(max(Zs[r], 0), Zmax) == zbin) {
          if (c == k) { res[2] = r; break; }
          ++c;
        }
    }
    __syncthreads();
    b = res[2];
    __syncthreads();                                                       // the scratch is reused below
  }
  const int Z = Zs[b];
  double* ch = chain + (size_t)b * CYL_CHAIN_DOUBLES;
  if (Z < 1 || Z > Zmax) {
    // a row count the shared-memory layout was not sized for (the host cannot validate a device array): leave the replica
    // untouched and say so in its flags (bit 1)
    if (tid == 0) ch[CYL_CH_FLAGS] = (double)((unsigned)ch[CYL_CH_FLAGS] | 2u);
    return;
  }
  const int nsite = Z * T;
  const CylParams& pg = params[b];                                        // prologue reads; the sweeps use the copy in sc->p
  const uint32_t rep = (uint32_t)(replica0 + b);
  const SiteStream stream = site_stream(seed, (uint64_t)(replica0 + b));
  C* psi_b = psi_g + (size_t)b * Zmax * T;
  constexpr bool extras = EXTRAS;
  const bool measure = (flags & CYL_SWEEP_MEASURE) != 0;
  const bool profiles = measure && prof != nullptr;
  const RepConst<R>& rc = sc->rc;

  // ---- prologue: state, lattice, tables ----
  if (tid == 0) {
    sc->p = pg;
    sc->rc = make_rep_const<R>(pg, Z, T, variant);
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
    sc->e_surf = ch[CYL_CH_E_SURF];
    sc->amp_accept = 0;
    for (int q = 0; q < CYL_OBS_DOUBLES; ++q) sc->obs[q] = 0.0;
  }
  if (tid >= 32 && tid < 42) sc->keys[tid - 32] = stream.key + (uint32_t)(tid - 32) * PHILOX_W0;
  for (int s = tid; s < nsite; s += SWEEP_NT) psi[s] = psi_b[s];
  for (int s = tid; s < 3 * Z; s += SWEEP_NT) prof_s[s] = R(0);
  {
    const double pi = 3.14159265358979323846;
    const double kk = pg.wavenumber;
    const double lz = (2 * pi / kk) / Z;                                   // On_lattice_simple.py:47
    for (int i = tid; i < Z; i += SWEEP_NT) trig[i] = row_trig<R>(kk, lz, i);
  }
  __syncthreads();
  rebuild_rows<R>(coef, roww, trig, &sc->rc, sc->amp, Z, extras);
  if (extras && warp == 0) {
    const double es = surface_energy_of(sc->p, sc->amp);
    if (lane == 0) sc->e_surf = es;
  }
  __syncthreads();

  const bool needs_ref = weights_need_ref<R>(variant);
  const CylParams& p = sc->p;
  const double amp_pt = (p.amp_target_acceptance > 0) ? p.amp_target_acceptance : 0.3;
  const double amp_ratio = 1.0 / (amp_pt * (1.0 - amp_pt));
  const double inv_nsite = 1.0 / nsite;
  const int ncol = num_colours(Z, T);
  const int Th = (T + 1) >> 1, Teven = T & ~1;
  const int di = SWEEP_NT / Th, dj = SWEEP_NT % Th;
  // Column-mapped site loops (T even): thread = (row group rg, column pair jj); a thread keeps its two columns and walks
  // down the rows in steps of the (even) number of row groups, so the parity of its rows, hence the column of its colour and
  // the periodic wrap in theta, are fixed per pass and all addresses advance by constants.  Fallback for T odd or T > 256:
  // the generic loops further down.
  const int RGe = ((T & 1) == 0) ? ((SWEEP_NT / Th) & ~1) : 0;
  const bool mapped = RGe >= 2;
  const int rg = mapped ? tid / Th : 0, cj = mapped ? tid - rg * Th : 0;
  const bool col_active = mapped && rg < RGe;

  for (int sw = 0; sw < nsweeps; ++sw) {
    const uint64_t sweep = sweep0 + (uint64_t)sw;
    if (extras && (sw & (AMP_CHUNK - 1)) == 0) {
      // draws of the next AMP_CHUNK amplitude moves, one sweep per thread (counter-based: any order, any thread); the
      // acceptance uniform is kept as T ln(u), so the decision is a comparison (accept rule B.2: u <= exp(-dE/T))
      if (tid < AMP_CHUNK && sw + tid < nsweeps) {
        const uint64_t s2 = sweep + (uint64_t)tid;
        double z, u;
        amp_draw(philox4x32_10(0u, rep, (uint32_t)s2, philox_c3(CYL_TAG_AMP, s2), keys), z, u);
        ampz[tid] = z;
        amplu[tid] = p.temperature * log(u);
      }
      __syncthreads();
    }
    const uint32_t c1 = (uint32_t)sweep + stream.off;
    const R sigma = (R)sc->sigma_site;
#ifdef CYL_PHASE_TIMING
    long long t_phase = clock64();
#endif
    // Chain scalars are read once here (the previous sweep ended with a barrier): during the decision phase several
    // threads work from these copies while thread 0 updates the shared ones.
    const double a = sc->amp, sig_amp = sc->sigma_amp, amp_cnt = sc->amp_counter, e_surf_cur = sc->e_surf;
    const double a_new = extras ? a + sig_amp * ampz[sw & (AMP_CHUNK - 1)] : a;
    const bool try_move = (flags & CYL_SWEEP_AMP_MOVES) && (fabs(a_new) < p.amp_bound);

    // ---- T. tables of the proposal: row weights for a' on whole warps counted down from warp 6.  No barrier: their first
    //         reader is the energy pass, two barriers from here.  (Running them DURING the first half of the energy pass, on
    //         the warps that hold no rows there, with one more barrier before the weights are applied, measured 0.7 % slower.
    //         E_surf(a') used to sit here on warp 7: that warp was then ~1.5k cycles late for the barrier of the first colour
    //         pass whenever its rows left it no slack; in the energy pass it costs nothing: 33.7 -> 32.9 ms per bench step.) ----
    auto proposal_tables = [&]() {
    if (extras) {
        if (try_move) {
          if (warp != SWEEP_NWARP - 1 && SMEM_WEIGHTS_ON) {
            const AmpGeo<R> ae = amp_geo<R>(rc, a_new), ar = amp_geo<R>(rc, a);
            // row i on thread NT - 33 - i: whole warps, counted down from warp 6 -- the low warps carry the extra row of the
            // site loops when Z is not a multiple of the row groups, so the table chain goes to the others
            for (int i = (SWEEP_NT - 33) - tid; i < Z; i += SWEEP_NT - 32) {
              const RowTrig<R> t = trig[i];
              const GeoT<R> g0e = geo_sc<R>(ae, t.s0, t.c0), gme = geo_sc<R>(ae, t.sm, t.cm);
              const GeoT<R> g0r = needs_ref ? geo_sc<R>(ar, t.s0, t.c0) : g0e;     // a_ref = current amplitude (:75)
              R wf[6];
              weight_row<R>(rc, g0e, gme, g0r, wf);
              RowWS<R>& w = roww[i];
#pragma unroll
              for (int q = 0; q < 5; ++q) w.nxt[q] = wf[q];
              w.cst[1] = wf[5];
            }
          }
        }
      }
    };
    proposal_tables();
    PHASE_MARK(0);

    // ---- 1. site moves ----
    int nacc = 0;
    EAcc<R> e = {R(0), R(0), R(0), R(0)};
    // (mapped loops) `o` points at the site, the neighbours sit at element offsets oL / oR (rows above / below, periodic) and
    // oW / oE (columns left / right, fixed per thread and pass); a thread walks down its rows by advancing o, the coefficient
    // pointer and the Philox site counter by constants
    // (Starting the NEXT row's Philox block before the current row's move -- a second independent chain per warp -- measured
    // 6 % slower: the loops are bound by pipe throughput, not by dependency latency.)
    uint32_t kr[10], kc = 0;                                               // round keys: reloaded at the top of every mapped pass
    auto load_keys = [&]() {
      const uint4 k0 = *reinterpret_cast<const uint4*>(&sc->keys[0]), k1 = *reinterpret_cast<const uint4*>(&sc->keys[4]);
      const uint2 k2 = *reinterpret_cast<const uint2*>(&sc->keys[8]);
      kr[0] = k0.x; kr[1] = k0.y; kr[2] = k0.z; kr[3] = k0.w; kr[4] = k1.x; kr[5] = k1.y; kr[6] = k1.z; kr[7] = k1.w; kr[8] = k2.x; kr[9] = k2.y;
      kc = kr[0] ^ c1;
    };
    auto draw = [&](uint32_t site) { return philox2x32_10(site, c1, kc, kr); };
    auto visit_x = [&](C* o, int oL, int oR, int oW, int oE, const RowCoef<R>* rcf, uint2 x) {
      const C L = o[oL], Rz = o[oR], W = o[oW], E = o[oE], pv = o[0];
      C q;
      if (site_propose<R, C>(*rcf, pv, L, Rz, W, E, sigma, x, q)) { o[0] = q; ++nacc; }
    };
    auto visit = [&](C* o, int oL, int oR, int oW, int oE, const RowCoef<R>* rcf, uint32_t site) {
      visit_x(o, oL, oR, oW, oE, rcf, draw(site));
    };
    if (mapped && ncol == 2) {
      // both lengths even: colour = (i + j) & 1
      const int step = RGe * T;
#pragma unroll 1
      for (int colour = 0; colour < 2; ++colour) {
        load_keys();
        if (col_active) {
          const int j = 2 * cj + ((colour + rg) & 1);
          const int oW = (j == 0 ? T : 0) - 1, oE = (j == T - 1 ? 1 - T : 1);
          C* o = psi + rg * T + j;
          const RowCoef<R>* rcf = coef + rg;
          uint32_t site = (uint32_t)(rg * T + j);
          for (int i = rg; i < Z; i += RGe, o += step, rcf += RGe, site += (uint32_t)step)
            visit(o, (i == 0 ? nsite : 0) - T, (i == Z - 1 ? -nsite : 0) + T, oW, oE, rcf, site);
        }
        __syncthreads();
      }
    } else if (mapped) {
      // T even, Z odd: colour = (a(i) + (j & 1)) % 3 with a(i) = i & 1 except a(Z-1) = 2.  Pass 0: even rows (not the last),
      // even columns + the last row, odd columns; pass 1: rows 0..Z-2, the column parity opposite to the row's; pass 2: odd
      // rows, odd columns + the last row, even columns.
      const int half = (Z - 1) >> 1;
#pragma unroll 1
      for (int colour = 0; colour < 3; ++colour) {
        load_keys();
        if (col_active) {
          if (colour == 1) {
            const int bb = (rg & 1) ? 0 : 1;                               // rows keep their parity: the step is even
            const int j = 2 * cj + bb, oW = (j == 0 ? T : 0) - 1, oE = (j == T - 1 ? 1 - T : 1);
            const int step = RGe * T;
            C* o = psi + rg * T + j;
            const RowCoef<R>* rcf = coef + rg;
            uint32_t site = (uint32_t)(rg * T + j);
            for (int i = rg; i < Z - 1; i += RGe, o += step, rcf += RGe, site += (uint32_t)step)
              visit(o, (i == 0 ? nsite : 0) - T, T, oW, oE, rcf, site);
          } else {
            const int bb = colour ? 1 : 0;
            const int j = 2 * cj + bb, oW = (j == 0 ? T : 0) - 1, oE = (j == T - 1 ? 1 - T : 1);
            const int step = 2 * RGe * T;                                  // even rows in pass 0, odd rows in pass 2
            C* o = psi + (2 * rg + bb) * T + j;
            const RowCoef<R>* rcf = coef + (2 * rg + bb);
            uint32_t site = (uint32_t)((2 * rg + bb) * T + j);
            for (int ii = rg; ii < half; ii += RGe, o += step, rcf += 2 * RGe, site += (uint32_t)step)
              visit(o, (ii == 0 && bb == 0 ? nsite : 0) - T, T, oW, oE, rcf, site);
          }
        }
        if (colour != 1 && tid >= SWEEP_NT - Th) {                         // the last row, from the top of the CTA (those
          const int bb = colour ? 0 : 1;                                   // threads have the fewest rows above)
          const int j = 2 * (tid - (SWEEP_NT - Th)) + bb, oW = (j == 0 ? T : 0) - 1, oE = (j == T - 1 ? 1 - T : 1);
          const int row = (Z - 1) * T;
          visit(psi + row + j, -T, -row, oW, oE, coef + (Z - 1), (uint32_t)(row + j));
        }
        __syncthreads();
      }
    } else if (ncol == 2) {
      // both lengths even: colour = (i + j) & 1, T/2 sites of each colour in every row
      for (int colour = 0; colour < 2; ++colour) {
        int i = tid / Th, jj = tid - i * Th;
        while (i < Z) {
          const int j = 2 * jj + ((colour + i) & 1);
          const int row = i * T;
          const int im = (i == 0 ? nsite : row) - T, ip = (i == Z - 1 ? 0 : row + T);
          const int jm = (j == 0 ? T : j) - 1, jp = (j == T - 1 ? 0 : j + 1);
          const C L = psi[im + j], Rz = psi[ip + j], W = psi[row + jm], E = psi[row + jp];
          C q;
          site_move<R, C>(psi, row + j, L, Rz, W, E, coef[i], sigma, (uint32_t)(row + j), c1, stream.key, q, nacc);
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
          const int ai = axis_colour(i, Z);
          int beta = colour - ai;
          if (beta < 0) beta += 3;
          int j;
          bool valid;
          if (beta < 2) { j = beta + 2 * jj; valid = j < Teven; }
          else          { j = T - 1;         valid = (T & 1) && jj == 0; }
          if (valid) {
            const int row = i * T;
            const int iprev = (i == 0 ? Z - 1 : i - 1), inext = (i == Z - 1 ? 0 : i + 1);
            const int jm = (j == 0 ? T : j) - 1, jp = (j == T - 1 ? 0 : j + 1);
            const C L = psi[iprev * T + j], Rz = psi[inext * T + j], W = psi[row + jm], E = psi[row + jp];
            C q;
            site_move<R, C>(psi, row + j, L, Rz, W, E, coef[i], sigma, (uint32_t)(row + j), c1, stream.key, q, nacc);
          }
          jj += dj; ii += di;
          if (jj >= Th) { jj -= Th; ++ii; }
        }
        __syncthreads();
      }
    }
    PHASE_MARK(1);

    // ---- E. field energy of the swept lattice under the current weights and its change under the proposal's, and the row
    //         sums of Lattice.measure (:121-131), in ONE read-only pass over the final field.  Even T: two adjacent threads per
    //         row, each walking half the row along theta (half_row_moments: every site's own terms, its bond to the row above
    //         and to the column on its left -- each bond once); the weights for a and for (a' - a) are applied once per half
    //         row, the rows' constant terms (background energy of the rescaled variants) ride in the same sums, and the two
    //         halves' profile sums meet in one shuffle.  Rows on consecutive lane pairs: conflict-free 16-byte loads.  A
    //         separate pass costs fewer instructions than closing bonds inside the colour passes, where the bookkeeping of who
    //         closes which bond is per site (and divergent with three colours). ----
    bool has_rows = false;                                                 // warp-uniform: did this warp accumulate anything?
#ifndef SMEM_ABL_NO_ESURF
    // E_surf(a') -- an AGM chain of fp64 square roots, ~1.5k cycles of latency on one warp -- on the last warp HERE: with two threads
    // per row the energy pass leaves that warp without rows up to Z = 112, and its only reader is the decision, one barrier from
    // here.  (At the top of the sweep it made the last warp late for the barrier of the first colour pass whenever the row count
    // left that warp no slack: Z a multiple of the row groups, three colours.)
    if (extras && warp == SWEEP_NWARP - 1 && try_move) {
      const double es = surface_energy_of(p, a_new);
      if (lane == 0) sc->e_surf_new = es;
    }
#endif
    // likewise the logarithms of this sweep's Robbins-Monro factors (read by the decision): one lane of the same warp
    if (extras && tid == SWEEP_NT - 1 && (flags & CYL_SWEEP_ADAPT_SIGMA)) sc->rm = rm_logs<R>(sc->step_counter, (double)nsite);
    if ((extras && SMEM_ENERGY_ON) || profiles) {
      if ((T & 1) == 0) {
        // 2 threads per row; a thread walks its half of the row's column pairs
        constexpr int lg = 1, nsplit = 2;                                  // (4 threads per row for Z <= NT/4 measured no faster)
        const int part = tid & 1;
        const int npair = T >> 1, p0 = (part * npair) >> lg, p1 = ((part + 1) * npair) >> lg;
        const int rows_per_warp = 32 >> lg, rows_per_round = SWEEP_NT >> lg;
        has_rows = ((tid & ~31) >> lg) < Z;
        auto rows = [&](auto meas) {
          constexpr bool MEAS = decltype(meas)::value;
          for (int ib = warp * rows_per_warp; ib < Z; ib += rows_per_round) {   // warp-uniform trip count (shuffles below)
            const int i = ib + (lane >> lg);
            R S[5] = {R(0), R(0), R(0), R(0), R(0)}, sre = R(0), sim = R(0), sabs = R(0);
            if (i < Z) half_row_moments<R, C, MEAS>(psi + i * T, psi + (i == 0 ? nsite : i * T) - T, p0, p1, T, S, sre, sim, sabs);
            if (MEAS) e.sa += sabs;
            if (MEAS && profiles) {
              sre += __shfl_xor_sync(0xffffffffu, sre, 1); sim += __shfl_xor_sync(0xffffffffu, sim, 1); sabs += __shfl_xor_sync(0xffffffffu, sabs, 1);
              if (i < Z && part == 0) { prof_s[3 * i] += sre; prof_s[3 * i + 1] += sim; prof_s[3 * i + 2] += sabs; }
            }
            if (i < Z) {
              const RowWS<R> w = roww[i];
              R ed = (w.nxt[0] - w.cur[0]) * S[0] + (w.nxt[1] - w.cur[1]) * S[1] + (w.nxt[2] - w.cur[2]) * S[2] +
                     (w.nxt[3] - w.cur[3]) * S[3] + (w.nxt[4] - w.cur[4]) * S[4];
              if (part == 0) ed += w.cst[1] - w.cst[0];
              e.ed += ed;
              if (MEAS) {
                R ec = w.cur[0] * S[0] + w.cur[1] * S[1] + w.cur[2] * S[2] + w.cur[3] * S[3] + w.cur[4] * S[4];
                if (part == 0) ec += w.cst[0];
                e.ec += ec; e.s2 += S[0];
              }
            }
          }
        };
        if (measure) rows(std::true_type{}); else rows(std::false_type{});
      } else {
        has_rows = true;
        if (extras && SMEM_ENERGY_ON) {
          for (int i = warp; i < Z; i += SWEEP_NWARP) {                    // T odd: a warp per row
            const int row = i * T, up = (i == 0 ? nsite : row) - T;
            const RowWS<R> ws = roww[i];
            RowW<R> w;
#pragma unroll
            for (int q = 0; q < 5; ++q) { w.cur[q] = ws.cur[q]; w.dif[q] = ws.nxt[q] - ws.cur[q]; }
            w.cst[0] = ws.cst[0]; w.cst[1] = ws.cst[1] - ws.cst[0];
            for (int j = lane; j < T; j += 32) {
              const C q = psi[row + j], l = psi[up + j], wl = psi[row + (j == 0 ? T : j) - 1];
              energy_own<R, C>(e, w, q, measure, !profiles);
              energy_zbond<R, C>(e, w.cur[2], w.dif[2], q, l, measure);
              energy_tbond<R, C>(e, w, q, wl, measure);
            }
            if (lane == 0) {
              e.ed += w.cst[1];
              if (measure) e.ec += w.cst[0];
            }
          }
        }
        if (profiles) {
          for (int i = tid; i < Z; i += SWEEP_NT) {
            R sa = 0, sre = 0, sim = 0;
            for (int j = 0; j < T; ++j) {
              const C v = psi[i * T + j];
              sre += v.x; sim += v.y; sa += sqrt_fast(v.x * v.x + v.y * v.y);
            }
            e.sa += sa;
            prof_s[3 * i] += sre; prof_s[3 * i + 1] += sim; prof_s[3 * i + 2] += sa;
          }
        }
      }
    }

    // ---- reduction: sums within a warp (working precision), the eight partials are added by whoever needs them ----
    {
      const int wacc = __reduce_add_sync(0xffffffffu, nacc);
      R v[4] = {e.ec, e.ed, e.s2, e.sa};
      if (extras && has_rows) {                                            // the other warps hold zeros
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
      }
      if (lane == 0) {
        double* s = scr + warp * 8;
        s[0] = (double)wacc;
        if (extras) { s[1] = (double)v[0]; s[2] = (double)v[1]; s[3] = (double)v[2]; s[4] = (double)v[3]; }
      }
    }
    __syncthreads();
    PHASE_MARK(2);
    auto total = [&](int k) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < SWEEP_NWARP; ++w) t += scr[w * 8 + k];
      return t;
    };

    // ---- D. decisions, on threads of different warps: Robbins-Monro on sigma_site (warp 2), the amplitude move (warp 0) ----
    // optional time series: one record of CYL_SERIES_DOUBLES per measured sweep and replica (the engine's me.df)
    double* rec = (SERIES && extras && measure && series) ? series + ((size_t)b * nsweeps + sw) * CYL_SERIES_DOUBLES : nullptr;
    // The measurement record (measure_avgs :135-156 + the engine's record) is not needed by anybody inside the launch, so it
    // is taken AFTER the decision barrier, off the critical path, by a thread of a warp that carries neither table rows nor
    // the extra site row; the partial sums it reads are next written two barriers from here.
    auto measure_record = [&]() {
      const double e_cur = total(1) * rc.lzlth_d;                        // :213
      sc->e_field = e_cur;
      double* o = sc->obs;
      o[CYL_OB_COUNT] += 1.0;
      o[CYL_OB_ABS_A] += fabs(a);
      o[CYL_OB_A] += a;
      o[CYL_OB_A2] += a * a;
      o[CYL_OB_PSI2] += total(3) * inv_nsite;
      o[CYL_OB_ABSPSI] += total(4) * inv_nsite;
      o[CYL_OB_E_FIELD] += e_cur;
      o[CYL_OB_E_SURF] += e_surf_cur;
      o[CYL_OB_E_FIELD2] += e_cur * e_cur;
      o[CYL_OB_SIGMA] += (double)sigma;
      if (rec) {
        rec[CYL_TS_AMPLITUDE] = a;
        rec[CYL_TS_E_FIELD] = e_cur;
        rec[CYL_TS_E_SURF] = e_surf_cur;
        rec[CYL_TS_PSI2] = total(3) * inv_nsite;
        rec[CYL_TS_ABSPSI] = total(4) * inv_nsite;
        rec[CYL_TS_SIGMA_SITE] = (double)sigma;
      }
    };
#define MEASURE_RECORD(t) if (extras && measure && tid == (t)) measure_record();
    if (tid == 64) {
      const double nac = total(0);
      if (rec) rec[CYL_TS_SITE_ACCEPTANCE] = nac * inv_nsite;
      if (flags & CYL_SWEEP_ADAPT_SIGMA) {
        const RmLogs<R> l = extras ? sc->rm : rm_logs<R>(sc->step_counter, (double)nsite);
        sc->sigma_site = rm_apply(sc->sigma_site, l, (double)nsite, nac);
      }
      sc->step_counter += nsite;
      sc->site_proposed += nsite;
      sc->site_accepted += nac;
    }
    if (extras && tid == 0) {
      int accept = 0;
      if (flags & CYL_SWEEP_AMP_MOVES) {
        if (try_move) {
          const double e_dif = total(2) * rc.lzlth_d;
          const double dE = (sc->e_surf_new - e_surf_cur) + e_dif;
          if (!isfinite(dE)) sc->flags = (double)((unsigned)sc->flags | 1u);
          accept = (amplu[sw & (AMP_CHUNK - 1)] <= -dE);                 // accept rule B.2 (T ln u <= 0: any dE <= 0 passes)
        }
        // Robbins-Monro on sigma_amp (engine, SURVEY.md B.3: m = 1, target p*)
        sc->amp_counter = amp_cnt + 1;
        sc->amp_proposed += 1;
        if (flags & CYL_SWEEP_ADAPT_SIGMA) {
          const double cf = sig_amp * amp_ratio * fast_rcp(fmax(amp_cnt + 1, 200.0));
          sc->sigma_amp = accept ? sig_amp + cf * (1 - amp_pt) : sig_amp - cf * amp_pt;
        }
        if (accept) {
          sc->amp = a_new;
          sc->e_surf = sc->e_surf_new;
          sc->amp_accepted += 1;
        }
      }
      sc->amp_accept = accept;
      if (rec) rec[CYL_TS_AMP_ACCEPTED] = (double)accept;
    }
    __syncthreads();
    PHASE_MARK(3);
    MEASURE_RECORD(96)
    // ---- R. metric tables follow the amplitude: the accepted amplitude becomes the metric reference too (rescaled
    //         variants), so the weights are rebuilt rather than taken from the proposal's ----
    if (extras && sc->amp_accept) {
      rebuild_rows<R>(coef, roww, trig, &sc->rc, sc->amp, Z, true);
      __syncthreads();
    }
    PHASE_MARK(4);
  }

  // ---- epilogue ----
  __syncthreads();
  for (int s = tid; s < nsite; s += SWEEP_NT) psi_b[s] = psi[s];
  if (profiles) {
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
                      double* obs, double* prof, double* series, void* stream) {
  CYL_CHECK_ARG(psi && Z && params && chain && B > 0 && Zmax > 0 && T > 1 && nsweeps >= 0, "cyl_sweep_smem: bad arguments");
  CYL_CHECK_ARG(variant >= 0 && variant < CYL_NUM_VARIANTS, "cyl_sweep_smem: unknown variant %d", variant);
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
  const bool extras = (flags & (CYL_SWEEP_AMP_MOVES | CYL_SWEEP_MEASURE)) != 0;
  auto kern = !extras ? sweep_smem_kernel<R, false, false, 0> : (series ? sweep_smem_kernel<R, true, true, 0> : sweep_smem_kernel<R, true, false, 0>);
  // the compile-time row length pays in the plain-sweep instance (-2 %); with the amplitude machinery compiled in, the 64
  // registers are exhausted either way and the specialised instance measured 2.6 % slower (more spills around the phases)
  if (T == 50 && !extras) kern = sweep_smem_kernel<R, false, false, 50>;
  CYL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  const PhiloxKeys keys = philox_expand(seed);
  // longest-first assignment pays when the grid runs in several waves and a CTA lives long enough to amortise the search;
  // CYL_SWEEP_INDEX_ORDER keeps block b on replica b
  int sms = 0;
  CYL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  uint32_t kflags = flags & ~(SWEEP_FLAG_LONGEST_FIRST | CYL_SWEEP_INDEX_ORDER);
  if (B > 8 * sms && nsweeps >= 32 && !(flags & CYL_SWEEP_INDEX_ORDER)) kflags |= SWEEP_FLAG_LONGEST_FIRST;
  kern<<<B, SWEEP_NT, lay.total, as_stream(stream)>>>((typename Real<R>::complex_t*)psi, B, Z, Zmax, T, params, chain, keys,
                                                      seed, replica0, sweep0, nsweeps, variant, kflags, obs, prof, series, lay);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

/* shared-memory bytes one replica of Zmax x T needs (the caller decides between this path and the HBM path from it) */
extern "C" int64_t cyl_sweep_smem_bytes(int Zmax, int T, int is_f64) {
  if (Zmax <= 0 || T <= 0) return -1;
  return (int64_t)(is_f64 ? SweepSmemLayout<double>(Zmax, T).total : SweepSmemLayout<float>(Zmax, T).total);
}

extern "C" int cyl_sweep_smem_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                  double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                  int variant, uint32_t flags, double* obs, double* prof, void* stream) {
  return sweep_smem<float>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, nullptr, stream);
}
extern "C" int cyl_sweep_smem_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                  double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                  int variant, uint32_t flags, double* obs, double* prof, void* stream) {
  return sweep_smem<double>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, nullptr, stream);
}

// The same sweeps with a per-sweep record: series[B][nsweeps][CYL_SERIES_DOUBLES] (requires CYL_SWEEP_MEASURE).
extern "C" int cyl_sweep_smem_series_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                         double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                         int variant, uint32_t flags, double* obs, double* prof, double* series, void* stream) {
  CYL_CHECK_ARG(series && (flags & CYL_SWEEP_MEASURE), "cyl_sweep_smem_series: needs a series buffer and CYL_SWEEP_MEASURE");
  return sweep_smem<float>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, series, stream);
}
extern "C" int cyl_sweep_smem_series_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                                         double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                                         int variant, uint32_t flags, double* obs, double* prof, double* series, void* stream) {
  CYL_CHECK_ARG(series && (flags & CYL_SWEEP_MEASURE), "cyl_sweep_smem_series: needs a series buffer and CYL_SWEEP_MEASURE");
  return sweep_smem<double>(psi, B, Z, Zmax, T, params, chain, seed, replica0, sweep0, nsweeps, variant, flags, obs, prof, series, stream);
}
