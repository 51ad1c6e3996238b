/* cylinder_b200.h -- C-ABI of the B200-native Metropolis engine for the jklebes/cylinder hot path.
 *
 * The reference (pure Python, /root/reference) has no FFI boundary of its own; its boundary is the
 * Python object API.  Each entry point below replaces the inner loop of one reference
 * function, cited as <file>:<line> relative to /root/reference/surfaces_and_fields/ unless noted.  The
 * Python mirror of the reference API (cylinder_b200/*.py) binds these through ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every buffer is a CALLER-OWNED DEVICE pointer (a torch tensor's data_ptr()); the library allocates
 *     nothing that outlives a call and never frees caller memory;
 *   - work is enqueued on the caller's stream (`stream` = cudaStream_t passed as void*; NULL = legacy
 *     default stream); no host synchronisation inside unless the function says so;
 *   - return 0 on success, negative CYL_E_* on error, text in cyl_last_error() (thread-local);
 *   - no C++ exceptions cross the boundary; no global mutable state besides the per-thread error text;
 *   - complex numbers are interleaved (re, im), row-major [z][theta], theta fastest -- the memory layout
 *     of numpy complex64 / complex128, so `.lattice` round-trips without repacking.
 */
#ifndef CYLINDER_B200_H
#define CYLINDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYL_ABI_VERSION 8

/* error codes */
#define CYL_OK             0
#define CYL_E_INVALID     (-1)   /* bad argument */
#define CYL_E_CUDA        (-2)   /* CUDA runtime / driver error (text has the CUDA string) */
#define CYL_E_UNSUPPORTED (-3)   /* shape or option outside what the kernel was built for */

/* dE variants: which reference class the single-site move restates */
#define CYL_VARIANT_SIMPLE        0   /* On_lattice_simple.py:598-686            */
#define CYL_VARIANT_RESCALED_0TH  1   /* On_lattice_rescaled_0thorder.py:116-209 */
#define CYL_VARIANT_RESCALED_1ST  2   /* On_lattice_rescaled_1storder.py:138-227 */
#define CYL_VARIANT_DECOUPLED     3   /* On_lattice_decoupled.py:259-347         */
#define CYL_NUM_VARIANTS          4

/* flags of the sweep entry points */
#define CYL_SWEEP_ADAPT_SIGMA   1u    /* Robbins-Monro width adaptation (On_lattice_simple.py:688-696) */
#define CYL_SWEEP_AMP_MOVES     2u    /* one global amplitude move per sweep (run_lattice.py:72-90)     */
#define CYL_SWEEP_MEASURE       4u    /* accumulate observables after every sweep (:121-156)            */
/* Execution options (they never change results, only how the work is laid out; the library reads no environment
 * variables and keeps no global option state): */
#define CYL_SWEEP_INDEX_ORDER   8u    /* cyl_sweep_smem_*: CTA b takes replica b (default: CTAs take the replicas in
                                         descending row count when the launch is large enough for that to pay)      */
#define CYL_SWEEP_TILE_ROWS(n)  (((uint32_t)(n) & 0xFFu) << 8)   /* cyl_sweep_hbm_*: tile height 32, 30, 28 or 24
                                         (default 0: chosen per lattice so that the wave count is near-integral)     */

/* Physical parameters of one replica = constructor arguments of the reference's
 * Cylinder (system_cylinder.py:15) and Lattice (On_lattice_simple.py:18-19). 12 doubles. */
typedef struct CylParams {
  double wavenumber;            /* k                                                  */
  double radius;                /* r                                                  */
  double gamma;                 /* surface tension                                    */
  double kappa;                 /* bending rigidity                                   */
  double intrinsic_curvature;   /* H0                                                 */
  double alpha, u, C, n;        /* field couplings                                    */
  double temperature;           /* T of both the site moves and the amplitude move    */
  double amp_bound;             /* reject |a'| >= amp_bound (.8 On_lattice_simple.py:83, .99 run.py:281) */
  double amp_target_acceptance; /* Robbins-Monro target of the amplitude move (engine default .3); 0 = .3 */
} CylParams;

/* Mutable per-replica chain state, CYL_CHAIN_DOUBLES doubles per replica (counters are stored as
 * doubles: exact to 2^53). */
#define CYL_CHAIN_DOUBLES 16
#define CYL_CH_AMPLITUDE     0   /* a                         Lattice.amplitude                         */
#define CYL_CH_SIGMA_SITE    1   /* sampling_width            On_lattice_simple.py:101                  */
#define CYL_CH_SIGMA_AMP     2   /* real_group_sampling_width engine                                    */
#define CYL_CH_E_FIELD       3   /* last surface_field_energy(a, a)                                     */
#define CYL_CH_E_SURF        4   /* last calc_surface_energy(a)                                         */
#define CYL_CH_STEP_COUNTER  5   /* Robbins-Monro step counter of the site moves (:689)                 */
#define CYL_CH_SITE_PROPOSED 6
#define CYL_CH_SITE_ACCEPTED 7   /* acceptance_counter                                                  */
#define CYL_CH_AMP_PROPOSED  8
#define CYL_CH_AMP_ACCEPTED  9
#define CYL_CH_AMP_COUNTER  10   /* Robbins-Monro step counter of the amplitude moves                   */
#define CYL_CH_FLAGS        11   /* bit0: non-finite energy seen; bit1: Z[b] outside [1, Zmax] (replica skipped) */

/* Per-replica observable accumulators (sums over measurements), CYL_OBS_DOUBLES doubles. */
#define CYL_OBS_DOUBLES 16
#define CYL_OB_COUNT     0
#define CYL_OB_ABS_A     1   /* sum |a|         measure_avgs :137-138 */
#define CYL_OB_A         2   /* sum a                                 */
#define CYL_OB_A2        3   /* sum a^2                               */
#define CYL_OB_PSI2      4   /* sum of lattice-mean |psi|^2           */
#define CYL_OB_ABSPSI    5   /* sum of lattice-mean |psi|             */
#define CYL_OB_E_FIELD   6
#define CYL_OB_E_SURF    7
#define CYL_OB_E_FIELD2  8
#define CYL_OB_SIGMA     9   /* sum of sigma_site                     */

/* ---------------------------------------------------------------- library ---------------------------- */
int         cyl_abi_version(void);
const char* cyl_last_error(void);                 /* thread-local; valid until the next call on this thread */
int         cyl_set_device(int device);           /* cudaSetDevice for this thread's calls                  */
int         cyl_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin);

/* ---------------------------------------------------------------- surface geometry ------------------- */
/* system_cylinder.py:25-41 evaluated on the device: out[i] = {radius_rescaled, sqrt_g_theta, g_theta,
 * sqrt_g_z, g_z, A_theta}(a[i], z[i]).  Parity hook for tests/golden/geometry.npz. */
int cyl_geometry_eval_f64(double wavenumber, double radius, const double* a, const double* z, int64_t n,
                          double* out /*[n][6]*/, void* stream);

/* Cylinder.calc_surface_energy (system_cylinder.py:119-124): E(m) by AGM, the two bending integrals by a
 * 256-node periodic trapezoid.  out_terms (may be NULL) = {A_integral_0, bending} per replica. */
int cyl_surface_energy_f64(const CylParams* params /*[B]*/, const double* amp /*[B]*/, int B,
                           double* E_out /*[B]*/, double* out_terms /*[B][2] or NULL*/, void* stream);

/* ---------------------------------------------------------------- counter-based RNG ------------------ */
/* Philox4x32-10 known-answer hook: out[i] = philox(ctr[i], key[i]). */
int cyl_philox4x32_10(const uint32_t* ctr /*[n][4]*/, const uint32_t* key /*[n][2]*/, int64_t n,
                      uint32_t* out /*[n][4]*/, void* stream);

/* Philox2x32-10 (the per-proposal generator of the sweep kernels): out[i] = philox(ctr[i], key[i]). */
int cyl_philox2x32_10(const uint32_t* ctr /*[n][2]*/, const uint32_t* key /*[n]*/, int64_t n,
                      uint32_t* out /*[n][2]*/, void* stream);

/* Lattice.random_initialize (On_lattice_simple.py:159-182): psi = rect(U(0,mag_max), U(0,2pi)), drawn from
 * Philox keyed (seed; site, replica0+b).  psi is [B][Zmax][T]; rows >= Z[b] are zeroed. */
int cyl_lattice_init_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, double mag_max,
                         uint64_t seed, int64_t replica0, void* stream);
int cyl_lattice_init_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, double mag_max,
                         uint64_t seed, int64_t replica0, void* stream);

/* ---------------------------------------------------------------- sequential replay (parity mode) ---- */
/* n proposals of Lattice.step_lattice_loc (On_lattice_simple.py:598-686 / 0th-order :116-209) replayed
 * strictly in order from a recorded random tape: site (iz, ith) in the reference's [-1, len-2] convention,
 * unit normals (zre drawn first), uniform u (NaN where the reference drew none).  fp64, arithmetic in the
 * reference's operation order, no FMA contraction.  rm_state = {sigma, step_counter} is read and written.
 * Outputs (each may be NULL): accept[n], dE[n], sigma_after[n]. */
int cyl_replay_f64(double* psi /*[Z][T] complex128*/, int Z, int T, const CylParams* params /*[1]*/,
                   double amp, int variant, const int32_t* iz, const int32_t* ith,
                   const double* zre, const double* zim, const double* u, int64_t n,
                   double* rm_state /*[2]*/, uint8_t* accept_out, double* dE_out, double* sigma_out,
                   void* stream);

/* One proposal psi' = psi + sigma_eff * (zre + i zim) with supplied unit normals (sigma_eff = sigma, or
 * sigma*sqrt(g) for the 0th-order variant): out = {dE, new_re, new_im, sqrt_g}.  Does not write psi.  Serves
 * Lattice.step_lattice_loc when the host draws from Python's `random` exactly like the reference. */
int cyl_site_delta_f64(const double* psi, int Z, int T, const CylParams* params, double amp, int variant,
                       int iz, int ith, double zre, double zim, double sigma, double* out /*[4]*/, void* stream);

/* ---------------------------------------------------------------- field energy ----------------------- */
/* Row moments S1..S5 of psi[B][Zmax][T]: sum_j |psi|^2, |psi|^4, |dz|^2, |dth|^2,
 * Im(conj(psi) dth), left differences divided by the pixel lengths; plus S6 = sum_j |psi|, S7/S8 = sum_j psi.
 * out is [B][Zmax][8] doubles (fp64 accumulation for both storage types). */
int cyl_row_moments_f32(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                        double* S, void* stream);
int cyl_row_moments_f64(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                        double* S, void* stream);

/* Whole-lattice proposals (On_lattice_simple.py: step_lattice_all :247-259, step_lattice_wavevector :267-300,
 * step_row_rotate :303-382).  A move is described by 8 doubles per replica; cyl_move_row_moments_* returns the row moments
 * of the PROPOSED field (feed them to cyl_field_energy_f64), cyl_move_apply_* commits it where accept[b] != 0 (accept may
 * be NULL: all replicas). */
#define CYL_MOVE_SHIFT 0   /* psi + c                                                                  */
#define CYL_MOVE_WAVE  1   /* psi + c exp(i (i qz/Z + j qth/T))                                        */
#define CYL_MOVE_TWIST 2   /* rows [r0, r0+nr) mod Z: psi exp(i (2 pi j n_rot / T + phase))            */
typedef struct CylMove {
  int32_t kind, r0, nr, n_rot;
  double c_re, c_im, qz, qth, phase, reserved;
} CylMove;
int cyl_move_row_moments_f32(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                             const CylMove* moves, double* S, void* stream);
int cyl_move_row_moments_f64(const void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                             const CylMove* moves, double* S, void* stream);
int cyl_move_apply_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const uint8_t* accept,
                       void* stream);
int cyl_move_apply_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylMove* moves, const uint8_t* accept,
                       void* stream);

/* Lattice.surface_field_energy(a_eval, a_ref, ...) (On_lattice_simple.py:185-213; 0th-order :247-278)
 * from row moments: E_out[b]. */
int cyl_field_energy_f64(const double* S /*[B][Zmax][8]*/, int B, const int32_t* Z, int Zmax, int T,
                         const CylParams* params, const double* amp_eval, const double* amp_ref,
                         int variant, double* E_out, void* stream);

/* ---------------------------------------------------------------- batched replica sweeps ------------- */
/* nsweeps checkerboard sweeps of B independent replicas, one CTA per replica, lattice resident in shared
 * memory for the whole launch.  Per sweep: every site proposed once (2 colours, 3 when Z or T is odd),
 * then, per flags, Robbins-Monro update of sigma_site, one amplitude move a' = a + sigma_amp*N(0,1)
 * accepted on E_surf + E_field(a', a) (On_lattice_simple.py:544-550 via engine.step_real_group), and
 * observable accumulation.  Philox counter = (site, replica0+b, sweep0+s, stream tag); key = seed.
 * psi [B][Zmax][T]; chain [B][CYL_CHAIN_DOUBLES]; obs [B][CYL_OBS_DOUBLES] and prof [B][Zmax][3]
 * (sum_theta Re psi, Im psi, |psi| summed over measurements) may be NULL when CYL_SWEEP_MEASURE is off. */
int cyl_sweep_smem_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                       double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                       int variant, uint32_t flags, double* obs, double* prof, void* stream);
int cyl_sweep_smem_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                       double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                       int variant, uint32_t flags, double* obs, double* prof, void* stream);

/* Shared-memory bytes per replica the batched path needs for a Zmax x T lattice (negative: bad arguments); compare with
 * cyl_device_info's max_smem_optin to choose between this path and the HBM path.  Precondition of cyl_sweep_smem_*:
 * 1 <= Z[b] <= Zmax for every replica (device array; a violating replica is skipped and flagged in CYL_CH_FLAGS bit 1). */
int64_t cyl_sweep_smem_bytes(int Zmax, int T, int is_f64);

/* The same sweeps with a time series (the engine's me.df / save_time_series, run.py:324-326): one record per measured
 * sweep and replica, series [B][nsweeps][CYL_SERIES_DOUBLES]; requires CYL_SWEEP_MEASURE. */
#define CYL_SERIES_DOUBLES 8
#define CYL_TS_AMPLITUDE        0   /* amplitude the sweep ran at (before its amplitude decision) */
#define CYL_TS_E_FIELD          1
#define CYL_TS_E_SURF           2
#define CYL_TS_PSI2             3   /* <|psi|^2> over the lattice */
#define CYL_TS_ABSPSI           4   /* <|psi|> */
#define CYL_TS_SIGMA_SITE       5   /* proposal width used by the sweep */
#define CYL_TS_AMP_ACCEPTED     6   /* 1 if the sweep's amplitude move was accepted */
#define CYL_TS_SITE_ACCEPTANCE  7   /* accepted site moves / sites */
int cyl_sweep_smem_series_f32(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                              double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                              int variant, uint32_t flags, double* obs, double* prof, double* series, void* stream);
int cyl_sweep_smem_series_f64(void* psi, int B, const int32_t* Z, int Zmax, int T, const CylParams* params,
                              double* chain, uint64_t seed, int64_t replica0, uint64_t sweep0, int nsweeps,
                              int variant, uint32_t flags, double* obs, double* prof, double* series, void* stream);

/* ---------------------------------------------------------------- one large lattice ------------------ */
/* Workspace of the HBM path.  The lattice lives in a ping-pong pair of halo-padded buffers, each split into
 * two column-parity planes (even / odd padded column) so that one checkerboard colour of a row is contiguous:
 *   work[buf][plane][Z + 2*halo_rows][pitch/2] complex,   padded (row, col) = (i + halo_rows, j + halo_cols),
 * halo rows / columns hold periodic copies.  cyl_hbm_layout reports the geometry; work must hold
 * 2*buffer_elems complex elements, scratch scratch_bytes bytes (header with the Robbins-Monro record, two row-coefficient
 * tables, row energy weights, per-tile partial sums, per-tile-column profile sums; need not be initialised). */
int cyl_hbm_layout(int Z, int T, int is_f64, int* halo_rows, int* halo_cols, int64_t* pitch_elems,
                   int64_t* buffer_elems, int64_t* scratch_bytes);
/* psi [Z][T] (numpy layout) -> work buffer 0 incl. halos; sets *cur = 0.  `cur` is a HOST integer owned by the
 * caller: the index (0/1) of the buffer that holds the lattice.  unpack reads buffer `cur`. */
int cyl_hbm_pack_f32(const void* psi, int Z, int T, void* work, int32_t* cur, void* stream);
int cyl_hbm_pack_f64(const void* psi, int Z, int T, void* work, int32_t* cur, void* stream);
int cyl_hbm_unpack_f32(const void* work, int cur, int Z, int T, void* psi, void* stream);
int cyl_hbm_unpack_f64(const void* work, int cur, int Z, int T, void* psi, void* stream);

/* nsweeps checkerboard sweeps of one Z x T lattice streamed from HBM.  One CTA per 2-D tile; the tile plus a
 * halo of one site per colour is staged into shared memory by TMA (cp.async.bulk.tensor.2d, one box per
 * parity plane, completion on an mbarrier); ALL colours of the sweep are then updated in shared memory (halo
 * sites are recomputed redundantly -- the counter-based RNG makes that bit-identical to the neighbour tile's
 * own result) and the interior is written to the other buffer, so a sweep moves 8 B read + 8 B written per
 * site (fp32) instead of one read + one write per colour.  `cur` (HOST int32, in/out) is flipped per sweep.
 * chain [CYL_CHAIN_DOUBLES] device.  Accept counts feed the Robbins-Monro sigma update once per sweep (carried from
 * sweep to sweep by the kernel itself: one launch per plain sweep).  With CYL_SWEEP_AMP_MOVES / CYL_SWEEP_MEASURE
 * the energies for the current and the proposed amplitude and the profile sums are accumulated by the same kernel
 * while it updates the sites; a small preparation kernel before and a decision kernel after each sweep draw the
 * proposal, build its row tables and take the amplitude / measurement decisions.  obs [CYL_OBS_DOUBLES], prof [Z][3]
 * may be NULL without CYL_SWEEP_MEASURE.  The kernels of one call are launched with programmatic stream
 * serialisation on `stream`; ordering with respect to other work on the stream is the usual stream order. */
int cyl_sweep_hbm_f32(void* work, int32_t* cur, int Z, int T, const CylParams* params, double* chain,
                      uint64_t seed, int64_t replica, uint64_t sweep0, int nsweeps, int variant,
                      uint32_t flags, void* scratch, double* obs, double* prof, void* stream);
int cyl_sweep_hbm_f64(void* work, int32_t* cur, int Z, int T, const CylParams* params, double* chain,
                      uint64_t seed, int64_t replica, uint64_t sweep0, int nsweeps, int variant,
                      uint32_t flags, void* scratch, double* obs, double* prof, void* stream);
/* Row moments of the lattice held in work[cur] -> S [Z][8] (same columns as cyl_row_moments_*). */
int cyl_hbm_row_moments_f32(const void* work, int cur, int Z, int T, const CylParams* params,
                            double* S, void* stream);
int cyl_hbm_row_moments_f64(const void* work, int cur, int Z, int T, const CylParams* params,
                            double* S, void* stream);

/* ---------------------------------------------------------------- Fourier-mode model ----------------- */
/* Cylinder2D.calc_field_energy / calc_field_energy_amplitude_change (system_cylinder2D.py:151-198) by the
 * real-space quadrature of SURVEY.md A.7: E[b] for coefficient sets c[b][(2*nth+1)][(2*nz+1)] complex128
 * at amplitudes a[b].  One warp per chain. */
int cyl_fourier_energy_f64(const CylParams* params /*[B]*/, int nz, int nth, const double* amp /*[B]*/,
                           const double* coeffs /*[B][ncoef] complex128*/, int B, double* E_out, void* stream);

/* Cylinder1D.calc_surface_energy (system_cylinder1D.py:189-194): gamma_eff*int(G_th G_z) + kappa*bending. */
int cyl_fourier_surface_energy_f64(const CylParams* params, const double* amp, int B, double* E_out,
                                   void* stream);

/* The reference's tensor tables, for callers that read them (Cylinder2D.evaluate_A_integrals, system_cylinder2D.py:103-131: tmp_A_integrals
 * [8 nz + 1] complex, index d + 4 nz; Cylinder2D.evaluate_B_integrals, :134-147: the beta' = beta diagonal of tmp_B_matrix,
 * [2nz+1][2nz+1][2nth+1] complex, indices i + nz, j + nz, beta + nth).  B amplitudes / parameter rows per call; 256-node periodic
 * trapezoid per element in place of two scipy quad calls.  The energy entry points above do not use these tables. */
int cyl_fourier_tensor_tables_f64(const CylParams* params, int nz, int nth, const double* amp, int B, double* out_A, double* out_B,
                                  void* stream);

/* The drivers of run.py:286-317 ("method") */
#define CYL_FOURIER_SEQUENTIAL      0   /* run.py:286-294: amplitude move, then fieldsteps coefficient moves          */
#define CYL_FOURIER_SIMULTANEOUS    1   /* run.py:295-304: one joint move of [|a|, |c_i|] (engine step_all)           */
#define CYL_FOURIER_FIXED_AMPLITUDE 2   /* run.py:305-312: coefficient moves only                                     */
#define CYL_FOURIER_NO_FIELD        3   /* run.py:313-318: amplitude moves on the bare surface energy                 */
/* execution layout, OR-ed into `method` (default 0: one CTA per chain when B <= SM count, else one warp per chain;
 * results do not depend on it) */
#define CYL_FOURIER_LAYOUT_CTA   0x100
#define CYL_FOURIER_LAYOUT_WARP  0x200

/* Batched adaptive Metropolis over the Fourier model: run.py:286-317 with the engine arithmetic
 * of SURVEY.md Appendix B (real-group / complex-group moves, Robbins-Monro widths, recursive mean and covariance, proposals
 * sigma_c L z with L L^T = Sigma_c), one warp per chain.  amp [B], coeffs [B][ncoef] complex128 and engine_state
 * [B][cyl_fourier_am_state_doubles(nz, nth)] are read and written:
 *   engine_state = { sigma_real, sigma_complex, real moves, complex moves, measurements, real accepted, complex accepted,
 *                    E_field, E_surf, sum a^2, sum E, 5 reserved | mean[d] | Sigma[d][d] },  d = 1 + ncoef <= 32.
 * ratio_real / ratio_complex are the Robbins-Monro constants for m = 1 / m = ncoef (On_lattice_simple.py:96-100); for
 * the simultaneous method sigma_real is the joint width and ratio_real the constant for m = d. */
int cyl_fourier_am_state_doubles(int nz, int nth);
int cyl_fourier_am_steps_f64(const CylParams* params /*[B]*/, int nz, int nth, double* amp, double* coeffs,
                             double* engine_state, int B, uint64_t seed, int64_t chain0, uint64_t step0, int n_steps,
                             int measure_every, int fieldsteps_per_ampstep, double target_acceptance, double ratio_real,
                             double ratio_complex, int method, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CYLINDER_B200_H */
