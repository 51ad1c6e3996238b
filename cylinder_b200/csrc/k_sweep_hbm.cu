// One large lattice streamed from HBM: TMA-staged halo tiles, a full checkerboard sweep per pass.
//
// Storage (cyl_hbm_layout): ping-pong pair of buffers; each buffer is the halo-padded lattice
//   padded (Rr, Q) = (i + HR, j + HC),  Rr in [0, Z + 2 HR),  Q in [0, pitch)
// split into two column-parity planes: plane = Q & 1, plane column = Q >> 1.  Halo rows/columns are periodic
// copies.  With both lattice lengths even the site colour is (i + j) & 1, so the sites of one colour in a row
// are CONTIGUOUS in one plane: shared-memory accesses of a warp are unit-stride 8-byte (conflict-free) and the
// global stores are fully coalesced.
//
// Kernel (one CTA per TH x TW tile):
//   1. one elected thread arms an mbarrier with the byte count and issues two cp.async.bulk.tensor.2d loads
//      (one box per plane) of the tile plus a halo of NCOL sites on every side;
//   2. colour c = 0..NCOL-1 is updated in shared memory on the tile grown by (NCOL-1-c) sites: the halo sites
//      are recomputed redundantly -- the Philox counter is the GLOBAL site index, so the redundant result is
//      bit-identical to what the neighbouring tile computes for the same site;
//   3. the interior is written to the other buffer (plus its periodic halo copies).
// A sweep therefore reads each site once (+ halo overhead) and writes it once, whatever the number of colours.
//
// Per sweep, plain (site moves + Robbins-Monro): ONE kernel.  The accept count goes to bucketed no-return atomics, and every
// tile of the next sweep derives the adapted proposal width from the previous sweep's record itself.
// Per sweep with measurement / amplitude move (EXTRAS): hbm_prep_kernel (proposal, row weights for a and a', coefficient
// table for a') -> sweep kernel (energies for a and a'-a and the profiles' theta-sums from one pass over each finished tile;
// 5 partial sums per tile) -> hbm_decide_kernel (E_surf(a') ahead of its wait; fixed-order sum of the tiles, Robbins-Monro,
// observables, amplitude decision).  All of them are launched with programmatic stream serialisation (see pdl_wait / pdl_trigger).
//
// Reference semantics restated: Lattice.step_lattice_loc (On_lattice_simple.py:598-686) per site,
// update_sigma (:688-696) batched per sweep, Lattice.measure (:121-131), amplitude move (:544-550).
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>
#include "cyl_common.cuh"
#include "cyl_energy.cuh"
#include "cyl_site.cuh"
#include "cyl_rowtab.cuh"
#include "cyl_surface.cuh"

#define HBM_NT 256
#define HBM_HC 4          // column halo in the global layout (>= NCOL; 4 keeps plane columns 16-byte aligned)
#define HBM_TW 128        // tile width (sites)

// Tile height is a template parameter chosen per launch (see pick_tile_height): co-resident CTAs run in waves, and the
// last, partially filled wave costs as much as a full one, so the height that makes the wave count of THIS lattice on
// THIS device (almost) integral wins -- 4096 x 4096 on 148 SMs x 4 CTAs: 32 rows -> 6.92 waves, 30 rows -> 7.41.

struct HbmGeom {
  int Z, T, ncol, hr, hc;
  int64_t pitch;          // padded row length in complex elements (even)
  int64_t rows;           // Z + 2 hr
  int64_t plane_elems;    // rows * pitch/2
  int64_t buffer_elems;   // 2 planes
};

static HbmGeom hbm_geom(int Z, int T) {
  HbmGeom g;
  g.Z = Z; g.T = T;
  g.ncol = num_colours(Z, T);
  g.hr = g.ncol;
  g.hc = HBM_HC;
  g.pitch = ((int64_t)T + 2 * g.hc + 31) / 32 * 32;
  g.rows = (int64_t)Z + 2 * g.hr;
  g.plane_elems = g.rows * (g.pitch / 2);
  g.buffer_elems = 2 * g.plane_elems;
  return g;
}

// scratch: header (HbmAmp) | RowCoef[2][Z] | RowW[Z] | per-tile partial sums [max tiles][8] | per-block constants of the
// weights [ceil(Z/128)][2] | per-tile-column row sums of the profiles [ceil(T/128)][3][Z]
struct HbmAmp {                      // header: the amplitude proposal of the current sweep + kernel bookkeeping (device)
  double a, a_new, u_acc, e_surf_new;
  int try_move;
  int coef_sel;                      // which of the two RowCoef tables belongs to the current amplitude
  unsigned int decide_ticket;        // decision kernel: CTAs that have filed their share of the tiles' sums (wraps to 0 by itself)
  int pad;
  // Robbins-Monro record of the site moves when no decision kernel runs (sweeps without measurement / amplitude move):
  // slot s % 3 of sweep s holds the width it used, the step counter it started from and its accept count, the latter in
  // HBM_NBUCKET buckets so that the no-return atomics of the tiles finishing together do not queue on one address.
  double rm_sigma[3], rm_counter[3];
  double acc_total;                  // accepts of all sweeps of this call but the last
  unsigned int nacc[3][32];
  double dpart[8][8];                // decision kernel: per-CTA sums of its share of the tiles (HBM_DECIDE_CTAS x 7 used)
};
#define HBM_NBUCKET 32
#define HBM_HDR_BYTES 2048
static_assert(sizeof(HbmAmp) <= HBM_HDR_BYTES, "scratch header");
#ifndef HBM_MINB_F32
#define HBM_MINB_F32 4             // co-resident CTAs per SM the fp32 sweep kernel is compiled for: with the Philox round keys
                                   // in registers it needs ~60; 4 vs 5 resident CTAs measured no difference (issue-bound)
#endif
#ifdef HBM_ABL_NO_ENERGY
#define HBM_ENERGY_ON false
#else
#define HBM_ENERGY_ON true
#endif
#define HBM_NPART 8                  // doubles per tile: accepts, E(a,a), E(a',a)-E(a,a), sum |psi|^2, sum |psi|, 3 spare
#define HBM_MIN_TH 24
#define HBM_PREP_NT 128
template <typename R>
struct HbmScratch {
  size_t off_coef, off_roww, off_part, off_cst, off_rowp, total;
  int ncst, ntx;
  __host__ __device__ HbmScratch(int Z, int T) {
    ncst = (Z + HBM_PREP_NT - 1) / HBM_PREP_NT;
    ntx = (T + HBM_TW - 1) / HBM_TW;
    off_coef = HBM_HDR_BYTES;
    off_roww = (off_coef + 2 * (size_t)Z * sizeof(RowCoef<R>) + 63) & ~(size_t)63;
    off_part = (off_roww + (size_t)Z * sizeof(RowW<R>) + 63) & ~(size_t)63;
    const size_t tiles = (size_t)ntx * ((Z + HBM_MIN_TH - 1) / HBM_MIN_TH);
    off_cst = off_part + tiles * HBM_NPART * sizeof(double);
    off_rowp = off_cst + (size_t)ncst * 2 * sizeof(double);
    total = off_rowp + (size_t)ntx * 3 * Z * sizeof(R);
  }
};

extern "C" int cyl_hbm_layout(int Z, int T, int is_f64, int* halo_rows, int* halo_cols, int64_t* pitch_elems,
                              int64_t* buffer_elems, int64_t* scratch_bytes) {
  CYL_CHECK_ARG(Z >= 8 && T >= 8, "cyl_hbm_layout: the HBM path needs Z, T >= 8 (got %d x %d); use the shared-memory path", Z, T);
  const HbmGeom g = hbm_geom(Z, T);
  if (halo_rows) *halo_rows = g.hr;
  if (halo_cols) *halo_cols = g.hc;
  if (pitch_elems) *pitch_elems = g.pitch;
  if (buffer_elems) *buffer_elems = g.buffer_elems;
  if (scratch_bytes) *scratch_bytes = (int64_t)(is_f64 ? HbmScratch<double>(Z, T).total : HbmScratch<float>(Z, T).total);
  return CYL_OK;
}

__device__ __forceinline__ int wrapi(int i, int n) {      // i in [-n, 2n)
  if (i < 0) i += n;
  if (i >= n) i -= n;
  return i;
}
__device__ __forceinline__ int modi(int i, int n) {       // any i
  i %= n;
  return i < 0 ? i + n : i;
}

// ----------------------------------------------------------------------------------------------------
// pack / unpack
// ----------------------------------------------------------------------------------------------------
template <typename C>
__global__ void hbm_pack_kernel(const C* __restrict__ psi, HbmGeom g, C* __restrict__ work) {
  const int64_t Q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t Rr = blockIdx.y;
  if (Q >= g.pitch) return;
  C v; v.x = 0; v.y = 0;
  if (Q < g.T + 2 * g.hc) {
    int gi = (int)((Rr - g.hr) % g.Z); if (gi < 0) gi += g.Z;
    int gj = (int)((Q - g.hc) % g.T);  if (gj < 0) gj += g.T;
    v = psi[(size_t)gi * g.T + gj];
  }
  work[(Q & 1) * g.plane_elems + Rr * (g.pitch / 2) + (Q >> 1)] = v;
}

template <typename C>
__global__ void hbm_unpack_kernel(const C* __restrict__ work, int cur, HbmGeom g, C* __restrict__ psi) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= g.T) return;
  const int64_t Q = j + g.hc, Rr = i + g.hr;
  const C* buf = work + (size_t)cur * g.buffer_elems;
  psi[(size_t)i * g.T + j] = buf[(Q & 1) * g.plane_elems + Rr * (g.pitch / 2) + (Q >> 1)];
}

template <typename R>
static int hbm_pack(const void* psi, int Z, int T, void* work, int32_t* cur, void* stream) {
  using C = typename Real<R>::complex_t;
  CYL_CHECK_ARG(psi && work && cur && Z >= 8 && T >= 8, "cyl_hbm_pack: bad arguments");
  const HbmGeom g = hbm_geom(Z, T);
  dim3 grid((unsigned)((g.pitch + 255) / 256), (unsigned)g.rows);
  hbm_pack_kernel<C><<<grid, 256, 0, as_stream(stream)>>>((const C*)psi, g, (C*)work);
  *cur = 0;
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
template <typename R>
static int hbm_unpack(const void* work, int cur, int Z, int T, void* psi, void* stream) {
  using C = typename Real<R>::complex_t;
  CYL_CHECK_ARG(psi && work && (cur == 0 || cur == 1) && Z >= 8 && T >= 8, "cyl_hbm_unpack: bad arguments");
  const HbmGeom g = hbm_geom(Z, T);
  dim3 grid((unsigned)((T + 255) / 256), (unsigned)Z);
  hbm_unpack_kernel<C><<<grid, 256, 0, as_stream(stream)>>>((const C*)work, cur, g, (C*)psi);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
extern "C" int cyl_hbm_pack_f32(const void* psi, int Z, int T, void* work, int32_t* cur, void* s) { return hbm_pack<float>(psi, Z, T, work, cur, s); }
extern "C" int cyl_hbm_pack_f64(const void* psi, int Z, int T, void* work, int32_t* cur, void* s) { return hbm_pack<double>(psi, Z, T, work, cur, s); }
extern "C" int cyl_hbm_unpack_f32(const void* work, int cur, int Z, int T, void* psi, void* s) { return hbm_unpack<float>(work, cur, Z, T, psi, s); }
extern "C" int cyl_hbm_unpack_f64(const void* work, int cur, int Z, int T, void* psi, void* s) { return hbm_unpack<double>(work, cur, Z, T, psi, s); }

// ----------------------------------------------------------------------------------------------------
// per-row coefficients for the current amplitude (chain[CYL_CH_AMPLITUDE]) -> scratch
// ----------------------------------------------------------------------------------------------------
template <typename R>
__global__ void hbm_coef_kernel(const CylParams* __restrict__ params, const double* __restrict__ chain, int Z, int T, int variant,
                                RowCoef<R>* __restrict__ coef, HbmAmp* __restrict__ hdr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { hdr->coef_sel = 0; hdr->try_move = 0; hdr->acc_total = 0.0; hdr->decide_ticket = 0u; }   // scratch arrives uninitialised
  if (i < 3 * HBM_NBUCKET) (&hdr->nacc[0][0])[i] = 0u;
  if (i >= Z) return;
  const CylParams p = *params;
  const double a = chain[CYL_CH_AMPLITUDE];
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z;
  // same functions as the shared-memory kernel's tables (cyl_rowtab.cuh): the two fp32 paths stay bit-identical
  const RepConst<R> rc = make_rep_const<R>(p, Z, T, variant);
  const AmpGeo<R> ag = amp_geo<R>(rc, a);
  const RowTrig<R> t = row_trig<R>(p.wavenumber, lz, i), tp = row_trig<R>(p.wavenumber, lz, (i + 1 == Z) ? 0 : i + 1);
  const GeoT<R> g0 = geo_sc<R>(ag, t.s0, t.c0), gm = geo_sc<R>(ag, t.sm, t.cm), gp = geo_sc<R>(ag, tp.sm, tp.cm);
  coef[i] = coef_row<R>(rc, g0, gm, gp);
}

// ----------------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS)
// ----------------------------------------------------------------------------------------------------
// a load the compiler may not sink to its first use
__device__ __forceinline__ double ld_early_f64(const double* p) {
  double v;
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// Programmatic dependent launch (sm_90+): every kernel of the sweep loop is launched with the stream-serialisation attribute,
// so its CTAs may become resident while its predecessor in the stream is still running.  A kernel must execute pdl_wait()
// (returns once the predecessor has completed and its writes are visible) before touching anything the predecessor writes,
// and lets its own successor start early with pdl_trigger().  Rule used below: a kernel triggers only AFTER its own wait, so
// a successor never overtakes the kernel two places up the stream; before its wait a kernel reads only data that is at least
// two kernels old (the sweep kernel: its input tiles, issued by TMA while the decision kernel is still finishing).  One
// exception, argued where it is made: the preparation kernel triggers at its start -- the kernel two places up from ITS successor
// is the decision kernel, which writes nothing a sweep tile reads before its wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// two adjacent complex values of a plane row (16-byte aligned pair)
__device__ __forceinline__ void load_pair(const float2* p, float2& a, float2& b) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  a = make_float2(t.x, t.y); b = make_float2(t.z, t.w);
}
__device__ __forceinline__ void load_pair(const double2* p, double2& a, double2& b) { a = p[0]; b = p[1]; }
__device__ __forceinline__ int warp_id_of(int tid) { return tid >> 5; }
__device__ __forceinline__ int ld_early_i32(const int* p) {
  int v;
  asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ----------------------------------------------------------------------------------------------------
// the sweep kernel
// ----------------------------------------------------------------------------------------------------
template <typename R, int NCOL, int TH_>
struct HbmSmem {
  using C = typename Real<R>::complex_t;
  static constexpr int TH = TH_, TW = HBM_TW;
  // Column halo in shared memory.  4 (not NCOL) so that every TMA box starts on a 16-byte boundary of its plane
  // row: box start = (tj0 + HBM_HC - HS)/2 plane columns = tj0/2, a multiple of 64 complex elements.
  static constexpr int HS = HBM_HC;
  static constexpr int SH = TH + 2 * NCOL, SW = TW + 2 * HS, SWH = SW / 2;
  static constexpr size_t plane_bytes = ((size_t)SH * SWH * sizeof(C) + 127) & ~(size_t)127;
  static constexpr size_t off_coef = 2 * plane_bytes;
  static constexpr size_t off_roww = off_coef + (size_t)SH * sizeof(RowCoef<R>);
  static constexpr size_t off_red = off_roww + (size_t)SH * sizeof(RowW<R>);
  static constexpr size_t off_bar = off_red + ((HBM_NT / 32) * 8 + 8) * sizeof(double);
  static constexpr size_t off_keys = off_bar + 64;                  // Philox round keys (see the prologue)
  static constexpr size_t total = off_keys + 64;
};

struct HbmMaps { CUtensorMap m[2][2]; int esz; };   // [buffer][plane]; esz = TMA elements per complex value

// everything a sweep launch needs besides the tile maps (device pointers into the caller's chain / scratch / outputs)
template <typename R>
struct HbmArgs {
  const RowCoef<R>* coef2;           // [2][Z], selected by hdr->coef_sel
  const RowW<R>* roww;
  const CylParams* params;
  double* chain;
  HbmAmp* hdr;
  double* partials;                  // [HBM_NPART][tiles of this launch]
  const double* cstpart;             // [ncst][2]
  R* rowpart;                        // [ntx][3][Z] or null
  double* obs;
  uint32_t flags;
  int ncst;
};

// EXTRAS: the field energy for the current amplitude and for the proposed one (weights staged per tile row) and the theta-sums of
// the profiles are taken from the finished tile in shared memory by one read-only pass (complete tiles; with three colours the
// tiles away from the lattice's edges); the other tiles accumulate them while the sites are updated -- own terms when a site is
// finalised, a bond by its later-finalised endpoint.  Either way a sweep with measurement / amplitude move needs no second pass
// over the lattice in HBM.  Each tile leaves its partial sums in scratch for hbm_decide_kernel.
template <typename R, int NCOL, int TH_, bool EXTRAS>
__global__ void __launch_bounds__(HBM_NT, sizeof(R) == 4 ? HBM_MINB_F32 : 2)
sweep_hbm_kernel(const __grid_constant__ HbmMaps maps, typename Real<R>::complex_t* __restrict__ work,
                 int cur, const __grid_constant__ HbmGeom g, const __grid_constant__ HbmArgs<R> A,
                 const __grid_constant__ SiteKeys keys, uint64_t sweep, int rm_slot) {
  using C = typename Real<R>::complex_t;
  using L = HbmSmem<R, NCOL, TH_>;
  constexpr int TH = L::TH, TW = L::TW, SH = L::SH, SW = L::SW, SWH = L::SWH, HS = L::HS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* pl0 = reinterpret_cast<C*>(smem_raw);
  C* pl1 = reinterpret_cast<C*>(smem_raw + L::plane_bytes);
  RowCoef<R>* coef = reinterpret_cast<RowCoef<R>*>(smem_raw + L::off_coef);
  RowW<R>* roww = reinterpret_cast<RowW<R>*>(smem_raw + L::off_roww);
  double* red = reinterpret_cast<double*>(smem_raw + L::off_red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::off_bar);
  const double* chain = A.chain;
  const int measure = (A.flags & CYL_SWEEP_MEASURE) ? 1 : 0;

  const int tid = threadIdx.x;
  // Tile of this CTA.  CTAs are dispatched in linear block order.  With three colours the tiles on the lattice's edge take
  // about twice as long as the others (general colour passes with wrapped coordinates, with EXTRAS the bonds closed inside the
  // passes), so the linear index is mapped to put them FIRST -- first and last tile row, then the first and last tile of every
  // other row, then the rest in row order -- and the grid drains with the short tiles (4095^2: 103 -> 95 us per sweep).  With
  // two colours the edge tiles cost little more than the others and the plain order measured 1 - 2 % faster at 4096^2 and
  // 8192^2.  Results do not depend on the order (per-tile partial sums and profile sums are filed by tile coordinates).
  int bx = blockIdx.x, by = blockIdx.y;
  const int gx = gridDim.x, gy = gridDim.y;
  if (NCOL == 3 && gx >= 3 && gy >= 3) {
    int b = by * gx + bx;
    if (b < 2 * gx) { by = b < gx ? 0 : gy - 1; bx = b < gx ? b : b - gx; }
    else if ((b -= 2 * gx) < 2 * (gy - 2)) { by = 1 + (b >> 1); bx = (b & 1) ? gx - 1 : 0; }
    else { b -= 2 * (gy - 2); by = 1 + b / (gx - 2); bx = 1 + b - (by - 1) * (gx - 2); }
  }
  const int ti0 = by * TH, tj0 = bx * TW;                            // interior origin (lattice coords)
  // smem (r, q) <-> padded (ti0 + r, Q0 + q);  padded row of r = 0 is ti0 - hr + hr
  const int Q0 = tj0 + g.hc - HS;
  const int x0 = (Q0 + ((0 - Q0) & 1)) >> 1, x1 = (Q0 + ((1 - Q0) & 1)) >> 1;    // first plane column of each plane's box

  // Prologue.  Global-memory latency under the bulk traffic of this kernel is several microseconds, so everything that
  // comes from global memory is requested up front and in parallel: the tile (TMA, issued by thread 0 as soon as it has
  // initialised the mbarrier), the row coefficients / weights and the proposal width.
  // Plain sweeps follow each other directly, so their tiles come from the preceding kernel: wait first.  With EXTRAS the
  // decision and preparation kernels sit between two sweeps and the tiles are at least two kernels old: load, then wait.
  if (!EXTRAS) pdl_wait();
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr uint32_t bytes = 2u * SH * SWH * sizeof(C);
    mbar_expect_tx(bar, bytes);
    tma_load_2d(pl0, &maps.m[cur][0], x0 * maps.esz, ti0, bar);    // esz: TMA elements per complex value
    tma_load_2d(pl1, &maps.m[cur][1], x1 * maps.esz, ti0, bar);
  }
  if (EXTRAS) pdl_wait();           // everything below comes from the preceding kernel
  pdl_trigger();
  R sigma;
  if (!EXTRAS && warp_id_of(tid) == 1) {
    // No decision kernel between plain sweeps: every tile derives this sweep's proposal width from the previous sweep's
    // record (rm_slot < 0: first sweep of the call, the width comes from the chain state); tile 0 files this sweep's record.
    const int lane = tid & 31;
    HbmAmp* hdr = A.hdr;
    const int r = rm_slot < 0 ? 0 : rm_slot, pslot = r == 0 ? 2 : r - 1;
    double sg, cnt, prev_acc = 0.0;
    if (rm_slot < 0) {
      sg = chain[CYL_CH_SIGMA_SITE];
      cnt = chain[CYL_CH_STEP_COUNTER];
    } else {
      const unsigned tot = __reduce_add_sync(0xffffffffu, hdr->nacc[pslot][lane]);
      const double nsite = (double)g.Z * g.T;
      sg = hdr->rm_sigma[pslot];
      cnt = hdr->rm_counter[pslot];
      if (A.flags & CYL_SWEEP_ADAPT_SIGMA) sg = rm_batch_update<R>(sg, cnt, nsite, (double)tot);
      cnt += nsite;
      prev_acc = (double)tot;
    }
    if (lane == 0) red[(HBM_NT / 32) * 8 + 7] = sg;
    if (bx == 0 && by == 0) {
      if (lane == 0) { hdr->rm_sigma[r] = sg; hdr->rm_counter[r] = cnt; hdr->acc_total += prev_acc; }
      hdr->nacc[r == 2 ? 0 : r + 1][lane] = 0u;                    // the slot of the next sweep; nobody reads it now
    }
  }
  if (EXTRAS) {
    sigma = (R)ld_early_f64(chain + CYL_CH_SIGMA_SITE);
    // which coefficient table is current is only known on the device: fetch the row from both tables together with the
    // selector instead of chaining two global-memory latencies
    const int sel = ld_early_i32(&A.hdr->coef_sel);
    for (int r = tid; r < SH; r += HBM_NT) {
      const int gr = modi(ti0 - NCOL + r, g.Z);
      const RowCoef<R> c0 = A.coef2[gr], c1 = A.coef2[(size_t)g.Z + gr];
      roww[r] = A.roww[gr];
      coef[r] = sel ? c1 : c0;
    }
  } else {
    for (int r = tid; r < SH; r += HBM_NT) coef[r] = A.coef2[modi(ti0 - NCOL + r, g.Z)];
  }
  // The round keys arrive in the constant bank; used from there the compiler re-materialises one move per key per site.
  // Routed through shared memory they are ordinary loaded values and stay in registers for the whole kernel.
  uint32_t* keys_s = reinterpret_cast<uint32_t*>(smem_raw + L::off_keys);
  if (tid < 10) keys_s[tid] = keys.k[tid];
  const uint32_t c1 = (uint32_t)sweep + keys.off;
  __syncthreads();                                                  // mbarrier initialised, row tables staged
  uint32_t kr[10];
#pragma unroll
  for (int r = 0; r < 10; ++r) kr[r] = keys_s[r];
  const uint32_t kc = kr[0] ^ c1;
  if (!EXTRAS) sigma = (R)red[(HBM_NT / 32) * 8 + 7];
  mbar_wait(bar, 0);

  int nacc = 0;
  EAcc<R> e = {R(0), R(0), R(0), R(0)};
  const bool meas = measure != 0;
  // tiles on the lattice boundary wrap their global coordinates, may be partial, and own periodic halo copies
  // (the last clause: with an odd length the last tile may be thinner than the halo, so its neighbour's halo wraps too)
  const bool border = (bx == 0) | (by == 0) | (bx == gx - 1) | (by == gy - 1) |
                      (ti0 + TH + NCOL > g.Z) | (tj0 + TW + HS > g.T);
  // complete tile: all of its TH x TW sites are on the lattice (every tile unless a lattice length is no multiple of the tile's)
  const bool complete = !border || ((ti0 + TH <= g.Z) && (tj0 + TW <= g.T));
  // EXTRAS: the energy and profile sums of this tile come from one pass over the finished tile (below) instead of being
  // accumulated inside the colour passes: complete tiles with two colours, tiles away from the edges with three
  const bool tile_energy = EXTRAS && (NCOL == 2 ? complete : !border);
  constexpr int PS = (int)(L::plane_bytes / sizeof(C));             // plane stride in shared memory (elements)
  C* out = work + (size_t)(cur ^ 1) * g.buffer_elems;
  const int64_t ph = g.pitch / 2;

  if constexpr (NCOL == 2) {
    // Both lengths even: colour of smem (r, q) is (r + q) & 1, so the colour-c sites of row r sit in ONE parity plane
    // at unit stride.  Thread t owns plane column 2 + (t & 63) and rows (t >> 6) + 4 k: the row parity, hence the plane,
    // is fixed per thread, and every address (tile, coefficient row, Philox counter) advances by a constant.
    static_assert(TW == 128 && HBM_NT % 64 == 0 && ((HBM_NT / 64) & 1) == 0, "thread mapping assumes 64 columns per plane row");
    constexpr int RSTEP = HBM_NT / 64;
    // one site move; `counted` = the site belongs to this tile's interior (not a redundantly recomputed halo site)
    // EN = accumulate the energy terms here (border tiles); interior tiles take them in one pass over the finished tile
    auto move = [&](auto MEAS, auto EN, C* own, const C* oth, int r, uint32_t site, bool counted, int c) {
      constexpr bool meas = decltype(MEAS)::value;                // the measurement terms compiled in or out
      constexpr bool en = decltype(EN)::value;
      const RowCoef<R> rc = coef[r];
      const C pv = own[0], Ln = own[-SWH], Rz = own[SWH], W = oth[0], E = oth[1];
      const uint2 x = philox2x32_10(site, c1, kc, kr);
      C q;
      const bool acc = site_propose<R, C>(rc, pv, Ln, Rz, W, E, sigma, x, q);
      if (acc) own[0] = q;
      nacc += (acc && counted);
      if (EXTRAS && HBM_ENERGY_ON && en && !tile_energy && counted) {
        const RowW<R> rw = roww[r];
        energy_own<R, C>(e, rw, q, meas);
        if (c == 1) {                                             // colour 1 closes every bond: its neighbours are final
          energy_zbond<R, C>(e, rw.cur[2], rw.dif[2], q, Ln, meas);
          energy_zbond<R, C>(e, roww[r + 1].cur[2], roww[r + 1].dif[2], Rz, q, meas);
          energy_tbond<R, C>(e, rw, q, W, meas);
          energy_tbond<R, C>(e, rw, E, q, meas);
        }
      }
    };
    auto site_id = [&](int r, int q) -> uint32_t {             // Philox counter word 0 = global site index
      int gi = ti0 - 2 + r, gj = tj0 - HS + q;
      if (border) { gi = wrapi(gi, g.Z); gj = wrapi(gj, g.T); }  // past Z + halo only in partial tiles: never stored
      return (uint32_t)(gi * g.T + gj);
    };
    auto passes = [&](auto MEAS) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int m = c + 1;                                       // rows [m, SH - m)
      const int col = HS / 2 + (tid & 63);                       // interior plane columns [HS/2, HS/2 + 64)
      int r = m + (RSTEP - 1) - (tid >> 6);      // the longer row groups go to the higher warps (scheduler priority)
      const int par = (r + c) & 1, q = 2 * col + par;
      C* own = pl0 + par * PS + r * SWH + col;
      const C* oth = pl0 + (par ^ 1) * PS + r * SWH + col - 1 + par;
      if (!border) {
        uint32_t site = site_id(r, q);
        for (; r < SH - m; r += RSTEP, own += RSTEP * SWH, oth += RSTEP * SWH, site += RSTEP * g.T)
          move(MEAS, std::false_type{}, own, oth, r, site, c == 1 || (r != 1 && r != SH - 2), c);
      } else {
        // edge tiles: the column (wrapped once) is the thread's own; the row index advances with the loop and wraps at most once
        int gi = wrapi(ti0 - 2 + r, g.Z);
        const bool col_in = tj0 - HS + q < g.T;
        uint32_t site = (uint32_t)(gi * g.T + wrapi(tj0 - HS + q, g.T));
        for (; r < SH - m; r += RSTEP, own += RSTEP * SWH, oth += RSTEP * SWH, gi += RSTEP, site += RSTEP * g.T) {
          if (gi >= g.Z) { gi -= g.Z; site -= (uint32_t)(g.Z * g.T); }
          move(MEAS, std::true_type{}, own, oth, r, site, (c == 1 || (r != 1 && r != SH - 2)) && ti0 - 2 + r < g.Z && col_in, c);
        }
      }
      if (c == 0 && tid >= HBM_NT - (SH - 2)) {
        // the 65th colour-0 site of each row of the grown region: halo column q = HS - 1 (odd rows) or HS + TW (even rows)
        const int r1 = 1 + (tid - (HBM_NT - (SH - 2))), q1 = (r1 & 1) ? HS - 1 : HS + TW;
        const int p1 = q1 & 1;
        move(MEAS, std::false_type{}, pl0 + p1 * PS + r1 * SWH + (q1 >> 1), pl0 + (p1 ^ 1) * PS + r1 * SWH + ((q1 - 1) >> 1), r1, site_id(r1, q1), false, 0);
      }
      if (c == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is final: bulk copies will read it
      __syncthreads();
    }
    };
    if (EXTRAS && meas) passes(std::true_type{}); else passes(std::false_type{});
  } else {
    // odd periodic ring(s): 3 colours, colour = (a(i) + a(j)) % 3.  In pass c a row of axis colour a(i) holds the sites whose
    // column has axis colour (c - a(i)) mod 3: every other column (or, for colour 2, only the odd ring's last column).  A warp
    // works on ONE row at a time (rows whose sites belong to another pass are skipped by the whole warp) and each lane
    // examines a pair of adjacent columns, of which at most one has the wanted colour -- so the lanes that get past the filter
    // are dense.  (Visiting every site of the region in every pass and filtering cost 3.7x the 2-colour sweep.)
    auto col3 = [](int s) { return s >= 3 ? s - 3 : s; };
    // one site of pass c at region coordinates (r, q) = lattice (gi, gj), axis colours (ai, bj)
    auto site3 = [&](auto EN, int c, int r, int q, int gi, int gj, int ai, int bj) {
      constexpr bool en = decltype(EN)::value;           // energy terms here (border tiles) or in the tile pass below
      const int par = (Q0 + q) & 1;
      C* own = par ? pl1 : pl0;
      C* oth = par ? pl0 : pl1;
      const int col = ((Q0 + q) >> 1) - (par ? x1 : x0);
      const int colW = ((Q0 + q - 1) >> 1) - (par ? x0 : x1), colE = ((Q0 + q + 1) >> 1) - (par ? x0 : x1);
      const C pv = own[r * SWH + col], Ln = own[(r - 1) * SWH + col], Rz = own[(r + 1) * SWH + col];
      const C W = oth[r * SWH + colW], E = oth[r * SWH + colE];
      const RowCoef<R> rc = coef[r];
      const uint2 x = philox2x32_10((uint32_t)(gi * g.T + gj), c1, kc, kr);
      C qv;
      const bool acc = site_propose<R, C>(rc, pv, Ln, Rz, W, E, sigma, x, qv);
      if (acc) own[r * SWH + col] = qv;
      const int li = r - NCOL, lj = q - HS;
      const bool counted = (li >= 0 && li < TH && lj >= 0 && lj < TW && ti0 + li < g.Z && tj0 + lj < g.T);
      nacc += (acc && counted);
      if (EXTRAS && HBM_ENERGY_ON && en && counted) {
        const RowW<R> rw = roww[r];
        energy_own<R, C>(e, rw, qv, meas);
        // a bond is added by its later-coloured endpoint: here, when the neighbour's colour is lower than ours
        if (col3(axis_colour(wrapi(gi - 1, g.Z), g.Z) + bj) < c) energy_zbond<R, C>(e, rw.cur[2], rw.dif[2], qv, Ln, meas);
        if (col3(axis_colour(wrapi(gi + 1, g.Z), g.Z) + bj) < c) energy_zbond<R, C>(e, roww[r + 1].cur[2], roww[r + 1].dif[2], Rz, qv, meas);
        if (col3(ai + axis_colour(wrapi(gj - 1, g.T), g.T)) < c) energy_tbond<R, C>(e, rw, qv, W, meas);
        if (col3(ai + axis_colour(wrapi(gj + 1, g.T), g.T)) < c) energy_tbond<R, C>(e, rw, E, qv, meas);
      }
    };
    // axis colour and lattice column of region column qq of this pass (-1: outside the region / past the lattice)
    auto column = [&](int qq, int qend, int& gj) {
      const int uj = tj0 - HS + qq;
      if (qq >= qend || uj >= g.T + NCOL) return -1;
      gj = border ? wrapi(uj, g.T) : uj;                           // interior tiles never leave [0, Z) x [0, T), halo included
      return axis_colour(gj, g.T);
    };
    auto passes3 = [&](auto EN) {
#pragma unroll 1
    for (int c = 0; c < NCOL; ++c) {
      const int m = c + 1, mq = m + (HS - NCOL);
      const int ncl = SW - 2 * mq, npairs = (ncl + 1) >> 1, qend = mq + ncl;
      // this thread's column pair is the same in every row of the pass: work out its two candidates once
      const int q0 = mq + 2 * (tid & 63);
      int gj0 = 0, gj1 = 0;
      const int ac0 = column(q0, qend, gj0), ac1 = column(q0 + 1, qend, gj1);
      for (int r = m + (tid >> 6); r < SH - m; r += HBM_NT / 64) {
        // lattice coordinates; tiles past the edge hold zero-filled padding that is never stored
        const int ui = ti0 - NCOL + r;
        if (ui >= g.Z + NCOL) break;
        const int gi = border ? wrapi(ui, g.Z) : ui;
        const int ai = axis_colour(gi, g.Z);
        int bj = c - ai;
        if (bj < 0) bj += 3;
        if (ac0 == bj) site3(EN, c, r, q0, gi, gj0, ai, bj);
        else if (ac1 == bj) site3(EN, c, r, q0 + 1, gi, gj1, ai, bj);
      }
      // the (at most two) pairs beyond the 64 per row, as one flat list over (row, pair)
      const int nextra = npairs - 64;
      for (int idx = tid; idx < nextra * (SH - 2 * m); idx += HBM_NT) {
        const int rr = nextra == 2 ? idx >> 1 : idx, r = m + rr, qx = mq + 2 * (64 + (nextra == 2 ? idx & 1 : 0));
        const int ui = ti0 - NCOL + r;
        if (ui >= g.Z + NCOL) continue;
        const int gi = border ? wrapi(ui, g.Z) : ui;
        const int ai = axis_colour(gi, g.Z);
        int bj = c - ai;
        if (bj < 0) bj += 3;
        int gja = 0, gjb = 0;
        const int aa = column(qx, qend, gja), ab = column(qx + 1, qend, gjb);
        if (aa == bj) site3(EN, c, r, qx, gi, gja, ai, bj);
        else if (ab == bj) site3(EN, c, r, qx + 1, gi, gjb, ai, bj);
      }
      if (c == NCOL - 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is final: bulk copies will read it
      __syncthreads();
    }
    };
    // Tiles away from the lattice edges (the bulk of a large lattice): no row or column of axis colour 2 within reach, tile
    // origins even, so the colour of region site (r, q) is (parity of r + 1) + (parity of q) and every pass is a regular
    // sub-lattice: pass 0 = even lattice rows x even columns on the tile grown by 2, pass 1 = (even, odd) and (odd, even) grown
    // by 1, pass 2 = (odd, odd) on the tile itself.  Thread = (row group tid >> 6, column tid & 63 of the pass's column list):
    // its plane, its neighbour offsets and the parity of its rows are fixed per pass, and tile address, coefficient row and
    // Philox counter advance by constants, as in the 2-colour loops; passes 0 and 2 deal the row groups over every other row.
    // The one or two columns beyond 64 go to the threads at the top of the CTA (fewest rows).  The visited sites, their Philox
    // counters and the arithmetic are those of the general loops above (which the edge tiles keep): identical results.
    auto move3 = [&](C* own, const C* oth, int r, uint32_t site, bool counted) {
      const RowCoef<R> rc = coef[r];
      const C pv = own[0], Ln = own[-SWH], Rz = own[SWH], W = oth[0], E = oth[1];
      const uint2 x = philox2x32_10(site, c1, kc, kr);
      C qv;
      const bool acc = site_propose<R, C>(rc, pv, Ln, Rz, W, E, sigma, x, qv);
      if (acc) own[0] = qv;
      nacc += (acc && counted);
    };
    auto passes3_interior = [&]() {
      static_assert((TH & 1) == 0 && TW == 128 && HBM_NT == 256 && HS == 4, "parity colouring of an interior tile");
      const int k = tid & 63, g4 = tid >> 6;
      const uint32_t T32 = (uint32_t)g.T;
      // site of region (r, q): plane q & 1, plane column q >> 1; its left / right neighbours sit in the other plane
      auto ptr_own = [&](int r, int q) { return pl0 + (q & 1) * PS + r * SWH + (q >> 1); };
      auto ptr_oth = [&](int r, int q) { return pl0 + ((q & 1) ^ 1) * PS + r * SWH + (q >> 1) - 1 + (q & 1); };
      auto site_of = [&](int r, int q) { return (uint32_t)(ti0 - NCOL + r) * T32 + (uint32_t)(tj0 - HS + q); };
      auto inside = [&](int r, int q) { return r >= NCOL && r < NCOL + TH && q >= HS && q < HS + TW; };
      // pass 0: rows r odd in [1, SH - 1), columns q even in [2, SW - 2) = 66 columns
      {
        const int q = 2 + 2 * k;
        int r = 1 + 2 * g4;
        C* own = ptr_own(r, q);
        const C* oth = ptr_oth(r, q);
        uint32_t site = site_of(r, q);
        const bool qin = q >= HS;                                    // q <= 128 < HS + TW
        for (; r < SH - 1; r += 8, own += 8 * SWH, oth += 8 * SWH, site += 8 * T32) move3(own, oth, r, site, qin && r >= NCOL && r < NCOL + TH);
        constexpr int nr0 = (SH - 2) / 2;
        if (tid >= HBM_NT - 2 * nr0) {
          const int idx = tid - (HBM_NT - 2 * nr0), rx = 1 + 2 * (idx >> 1), qx = 130 + 2 * (idx & 1);
          move3(ptr_own(rx, qx), ptr_oth(rx, qx), rx, site_of(rx, qx), inside(rx, qx));
        }
      }
      __syncthreads();
      // pass 1: rows [2, SH - 2); even r (odd lattice row) takes the even columns 4 .. 132, odd r the odd columns 3 .. 131
      {
        int r = 2 + g4;
        const int q = ((r & 1) ? 3 : 4) + 2 * k;
        C* own = ptr_own(r, q);
        const C* oth = ptr_oth(r, q);
        uint32_t site = site_of(r, q);
        const bool qin = q >= HS;                                    // q <= 130
        for (; r < SH - 2; r += 4, own += 4 * SWH, oth += 4 * SWH, site += 4 * T32) move3(own, oth, r, site, qin && r >= NCOL && r < NCOL + TH);
        constexpr int nr1 = SH - 4;
        if (tid >= HBM_NT - nr1) {
          const int rx = 2 + (tid - (HBM_NT - nr1)), qx = (rx & 1) ? 131 : 132;
          move3(ptr_own(rx, qx), ptr_oth(rx, qx), rx, site_of(rx, qx), inside(rx, qx));
        }
      }
      __syncthreads();
      // pass 2: rows r even in [3, SH - 3), columns q odd in [4, SW - 4) = 64 columns: all of them interior
      {
        const int q = 5 + 2 * k;
        int r = 4 + 2 * g4;
        C* own = ptr_own(r, q);
        const C* oth = ptr_oth(r, q);
        uint32_t site = site_of(r, q);
        for (; r < SH - 3; r += 8, own += 8 * SWH, oth += 8 * SWH, site += 8 * T32) move3(own, oth, r, site, true);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is final: bulk copies will read it
      __syncthreads();
    };
    if (!border) passes3_interior(); else passes3(std::true_type{});
  }

  // ---- write the interior (and its periodic halo copies) to the other buffer ----
  // A complete tile (all of its TH x TW sites on the lattice): every (row, plane) is 64 contiguous complex values on both sides,
  // 16-byte aligned: one bulk copy (cp.async.bulk shared -> global, SASS UBLKCP) per row and plane, issued by 2*TH threads,
  // instead of a load + store per 16 bytes by everyone.  The shared-memory writes of the colour passes were made visible to the
  // async proxy by the fence before the last barrier; the issuing threads wait until their sources have been read before the
  // CTA may retire.  Tiles on the lattice's edge add the periodic halo copies of their outermost rows / columns element by
  // element; tiles cut by the lattice's end store everything that way.
  const bool bulk_out = complete;
  // The 2 TH copies are issued by lanes 0..7 of all eight warps, not by the first 2 TH threads: a bulk copy is a uniform-datapath
  // instruction that the compiler serialises over the active lanes of a warp (SASS: R2UR x3 + UBLKCP in a loop), so 32 issuing
  // lanes of one warp are 32 round trips in a row on the tile's critical path; 8 per warp on 8 warps run side by side.
  const int bulk_idx = (tid >> 5) + (HBM_NT / 32) * (tid & 31);      // lanes 0..7 of warp w: w, w + 8, ..., w + 56
  const bool bulk_issuer = bulk_out && (tid & 31) < 8 && bulk_idx < 2 * TH;
  if (bulk_out) {
    static_assert(2 * TH <= 8 * (HBM_NT / 32), "eight copies per warp");
    if (bulk_issuer) {
      const int li = bulk_idx >> 1, p = bulk_idx & 1;
      const C* src = pl0 + p * PS + (li + NCOL) * SWH + HS / 2;
      C* dst = out + p * g.plane_elems + (int64_t)(ti0 + li + g.hr) * ph + ((tj0 + g.hc) >> 1);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                   "r"((uint32_t)((TW / 2) * sizeof(C)))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (border && bulk_out) {
    // periodic halo copies of a complete edge tile: its rows within hr of the lattice's first / last row (with their corner
    // columns), then its columns within hc of the first / last column -- a few hundred elements, one or two per thread
    auto tile_value = [&](int li, int lj) {
      const int r = li + NCOL, q = lj + HS;                          // Q0 is even: plane q & 1, plane column q >> 1
      return (q & 1 ? pl1 : pl0)[r * SWH + (q >> 1)];
    };
    auto halo_col = [&](int gj) -> int64_t {                         // padded column of the periodic copy of lattice column gj, or -1
      const int64_t Q = gj + g.hc;
      return gj < g.hc ? Q + g.T : (gj >= g.T - g.hc ? Q - g.T : (int64_t)-1);
    };
    const int top1 = max(0, min(TH, g.hr - ti0)), bot0 = min(TH, max(top1, g.Z - g.hr - ti0));
    const int nrow = top1 + (TH - bot0);
    for (int idx = tid; idx < nrow * TW; idx += HBM_NT) {
      const int k = idx / TW, lj = idx - k * TW;
      const int li = k < top1 ? k : bot0 + (k - top1);
      const int gi = ti0 + li, gj = tj0 + lj;
      const int64_t Q = gj + g.hc, R2 = gi + g.hr + (gi < g.hr ? g.Z : -g.Z), Q2 = halo_col(gj);
      const C v = tile_value(li, lj);
      out[(Q & 1) * g.plane_elems + R2 * ph + (Q >> 1)] = v;
      if (Q2 >= 0) out[(Q2 & 1) * g.plane_elems + R2 * ph + (Q2 >> 1)] = v;
    }
    const int left1 = max(0, min(TW, g.hc - tj0)), right0 = min(TW, max(left1, g.T - g.hc - tj0));
    const int ncx = left1 + (TW - right0);
    for (int idx = tid; idx < ncx * TH; idx += HBM_NT) {
      const int li = idx / ncx, k = idx - li * ncx;
      const int lj = k < left1 ? k : right0 + (k - left1);
      const int64_t Rr = ti0 + li + g.hr, Q2 = halo_col(tj0 + lj);
      out[(Q2 & 1) * g.plane_elems + Rr * ph + (Q2 >> 1)] = tile_value(li, lj);
    }
  } else if (border) {
    for (int idx = tid; idx < 2 * TH * (TW / 2); idx += HBM_NT) {
      const int par_slot = idx / (TH * (TW / 2));                    // which smem plane
      const int rem = idx - par_slot * (TH * (TW / 2));
      const int li = rem / (TW / 2), cc = rem - li * (TW / 2);
      // interior columns of smem plane `par_slot`: q = HS + lj with (Q0 + q) & 1 == par_slot
      const int lj = 2 * cc + ((par_slot - (Q0 + HS)) & 1);
      const int gi = ti0 + li, gj = tj0 + lj;
      if (gi >= g.Z || gj >= g.T) continue;
      const int r = li + NCOL, q = lj + HS;
      const C* src = par_slot ? pl1 : pl0;
      const C v = src[r * SWH + (((Q0 + q) >> 1) - (par_slot ? x1 : x0))];
      const int64_t Rr = gi + g.hr, Q = gj + g.hc;
      int64_t R2 = -1, Q2 = -1;
      if (gi < g.hr) R2 = Rr + g.Z; else if (gi >= g.Z - g.hr) R2 = Rr - g.Z;
      if (gj < g.hc) Q2 = Q + g.T; else if (gj >= g.T - g.hc) Q2 = Q - g.T;
      out[(Q & 1) * g.plane_elems + Rr * ph + (Q >> 1)] = v;
      if (R2 >= 0) out[(Q & 1) * g.plane_elems + R2 * ph + (Q >> 1)] = v;
      if (Q2 >= 0) {
        out[(Q2 & 1) * g.plane_elems + Rr * ph + (Q2 >> 1)] = v;
        if (R2 >= 0) out[(Q2 & 1) * g.plane_elems + R2 * ph + (Q2 >> 1)] = v;
      }
    }
  }

  const int warp = tid >> 5, lane = tid & 31;
  // ---- EXTRAS, complete tiles (three colours: tiles away from the lattice's edges): the field energy (current weights, and
  //      proposal - current) and the theta-sums of the profiles in ONE read-only pass over the finished tile, as in the batched
  //      kernel's energy pass -- a third of the instructions of closing the bonds inside the colour passes.  Two colours: what is
  //      final in shared memory is every interior site and the colour-0 sites of the ring around it (recomputed redundantly above,
  //      on edge tiles from the periodic halo the tile was loaded with); a bond joins a colour-0 and a colour-1 site, so the tile in
  //      which the colour-1 endpoint is interior sees both ends final and owns the bond -- no bond is missed or taken twice across
  //      tile boundaries.  Thread = (row, 4 strided column pairs per plane): lane & 7 picks the pair inside a group of 16 plane
  //      columns, so the 16-byte loads of a quarter-warp cover 128 contiguous bytes (conflict-free); rows are dealt warp + 8 k, so
  //      the row parity -- which plane holds the colour-1 sites -- is warp-uniform.  Per step a thread loads two colour-1 sites
  //      with the sites above and below, the two colour-0 sites of the same plane columns and the colour-0 neighbour beyond the
  //      pair: five consecutive columns, four theta-bonds, four z-bonds, four sites' own terms.  The row weights are applied once
  //      per thread (the z-bonds downwards carry the next row's weight).  Three colours: see the branch below. ----
  bool tile_pass_done = false;
  if constexpr (EXTRAS) {
    if (tile_energy) {
      static_assert(TH <= 32 && TW == 128 && HBM_NT == 256, "row = warp + 8 k, eight threads per row, 4 x 16 plane columns");
      tile_pass_done = true;
      auto tile_pass = [&](auto MEAS) {
        constexpr bool M = decltype(MEAS)::value;
        using P = Pk<R>;
        using V = typename P::T;
        const int li = warp + 8 * (lane >> 3), seg = lane & 7;
        R sre = R(0), sim = R(0), sab = R(0);
        if (li < TH) {
          const int r = li + NCOL;
          R s2 = R(0), s4 = R(0), sa = R(0);
          V zu = P::zero(), zd = P::zero(), dt = P::zero(), im = P::zero(), sp = P::zero();
          auto own = [&](const C v) {
            const R p2 = v.x * v.x + v.y * v.y;
            s2 += p2;
            s4 += p2 * p2;
            if (M) { sa += sqrt_fast(p2); sp = P::add(sp, P::make(v)); }
          };
          auto tbond = [&](const V hi, const V lo) {                   // as energy_tbond: |dth|^2 and Im(conj(hi) dth)
            const V t = P::sub(hi, lo);
            dt = P::fma(t, t, dt);
            im = P::fma(P::rot(hi), t, im);
          };
          if constexpr (NCOL == 2) {
            const int p1 = (r + 1) & 1;                                // (r + q) & 1 == 1: the plane of this row's colour-1 sites
            const C* a1 = pl0 + p1 * PS + r * SWH + HS / 2 + 2 * seg;
            const C* a0 = pl0 + (p1 ^ 1) * PS + r * SWH + HS / 2 + 2 * seg;
            auto steps = [&](auto PAR) {
              constexpr int par = decltype(PAR)::value;
#pragma unroll 2
              for (int j = 0; j < 4; ++j) {
                C A0, A1, B0, B1, U0, U1, D0, D1;
                load_pair(a1 + 16 * j, A0, A1);
                load_pair(a0 + 16 * j, B0, B1);
                load_pair(a1 + 16 * j - SWH, U0, U1);
                load_pair(a1 + 16 * j + SWH, D0, D1);
                const C X = a0[16 * j + (par ? 2 : -1)];
                const V vA0 = P::make(A0), vA1 = P::make(A1), vB0 = P::make(B0), vB1 = P::make(B1), vX = P::make(X);
                if (par) { tbond(vA0, vB0); tbond(vB1, vA0); tbond(vA1, vB1); tbond(vX, vA1); }     // columns B0 A0 B1 A1 X
                else     { tbond(vA0, vX); tbond(vB0, vA0); tbond(vA1, vB0); tbond(vB1, vA1); }     // columns X A0 B0 A1 B1
                const V u0 = P::sub(vA0, P::make(U0)), u1 = P::sub(vA1, P::make(U1));
                zu = P::fma(u0, u0, zu); zu = P::fma(u1, u1, zu);
                const V d0 = P::sub(P::make(D0), vA0), d1 = P::sub(P::make(D1), vA1);
                zd = P::fma(d0, d0, zd); zd = P::fma(d1, d1, zd);
                own(A0); own(A1); own(B0); own(B1);
              }
            };
            if (p1) steps(std::integral_constant<int, 1>{}); else steps(std::integral_constant<int, 0>{});
          } else {
            // Three colours, tile away from the lattice edges (so no row or column of axis colour 2 within reach): colour =
            // (row parity) + (column parity).  Final in shared memory: the interior and the ring sites of colours 0 and 1; the
            // endpoint of a bond with the ODD coordinate along the bond has the higher colour, and the tile in which that
            // endpoint is interior owns the bond.  Tiles start on even rows and columns (TH and TW are even), so: every interior
            // site takes its bond downwards and its bond to the right -- the ring sites these reach are even, hence final.
            static_assert((TH & 1) == 0 && (TW & 1) == 0, "tile origins must be even");
            const C* b0 = pl0 + r * SWH + HS / 2 + 2 * seg;           // even columns
            const C* b1 = b0 + PS;                                    // odd columns
#pragma unroll 2
            for (int j = 0; j < 4; ++j) {
              C E0, E1, O0, O1, F0, F1, G0, G1;                       // columns in order: E0 O0 E1 O1 X
              load_pair(b0 + 16 * j, E0, E1);
              load_pair(b1 + 16 * j, O0, O1);
              load_pair(b0 + 16 * j + SWH, F0, F1);
              load_pair(b1 + 16 * j + SWH, G0, G1);
              const C X = b0[16 * j + 2];
              const V vE0 = P::make(E0), vE1 = P::make(E1), vO0 = P::make(O0), vO1 = P::make(O1);
              tbond(vO0, vE0); tbond(vE1, vO0); tbond(vO1, vE1); tbond(P::make(X), vO1);
              const V d0 = P::sub(P::make(F0), vE0), d1 = P::sub(P::make(F1), vE1);
              const V d2 = P::sub(P::make(G0), vO0), d3 = P::sub(P::make(G1), vO1);
              zd = P::fma(d0, d0, zd); zd = P::fma(d1, d1, zd); zd = P::fma(d2, d2, zd); zd = P::fma(d3, d3, zd);
              own(E0); own(E1); own(O0); own(O1);
            }
          }
          if (HBM_ENERGY_ON) {
            const RowW<R> w = roww[r];
            const R wnc = roww[r + 1].cur[2], wnd = roww[r + 1].dif[2];
            const R Szu = P::hsum(zu), Szd = P::hsum(zd), S3 = P::hsum(dt), S4 = P::hsum(im);
            e.ed += w.dif[0] * s2 + w.dif[1] * s4 + w.dif[2] * Szu + wnd * Szd + w.dif[3] * S3 + w.dif[4] * S4;
            if (M) {
              e.ec += w.cur[0] * s2 + w.cur[1] * s4 + w.cur[2] * Szu + wnc * Szd + w.cur[3] * S3 + w.cur[4] * S4;
              e.s2 += s2;
              e.sa += sa;
            }
          }
          if (M) { const C spc = P::get(sp); sre = spc.x; sim = spc.y; sab = sa; }
        }
#ifndef HBM_ABL_NO_PROF
        if (M && A.rowpart) {                                          // warp-uniform: the shuffles are taken by whole warps
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            sre += __shfl_xor_sync(0xffffffffu, sre, o); sim += __shfl_xor_sync(0xffffffffu, sim, o); sab += __shfl_xor_sync(0xffffffffu, sab, o);
          }
          if (li < TH && seg < 3) A.rowpart[((size_t)bx * 3 + seg) * g.Z + ti0 + li] = seg == 0 ? sre : seg == 1 ? sim : sab;
        }
#endif
      };
      if (meas) tile_pass(std::true_type{}); else tile_pass(std::false_type{});
    }
  }
  // ---- theta-sums of the profiles (Lattice.measure :121-131) of the tiles that did not take the pass above (tiles cut by the
  //      lattice's end; with three colours every edge tile): a warp per row, wrapped plane addressing ----
#ifndef HBM_ABL_NO_PROF
  if (EXTRAS && meas && A.rowpart && !tile_pass_done) {
    for (int li = warp; li < TH && ti0 + li < g.Z; li += HBM_NT / 32) {
      R sre = R(0), sim = R(0), sab = R(0);
      for (int lj = lane; lj < TW && tj0 + lj < g.T; lj += 32) {
        const int q = lj + HS, par = (Q0 + q) & 1;
        const C v = (par ? pl1 : pl0)[(li + NCOL) * SWH + (((Q0 + q) >> 1) - (par ? x1 : x0))];
        sre += v.x; sim += v.y; sab += sqrt_fast(v.x * v.x + v.y * v.y);
      }
      sre = warp_sum(sre); sim = warp_sum(sim); sab = warp_sum(sab);
      if (lane < 3) A.rowpart[((size_t)bx * 3 + lane) * g.Z + ti0 + li] = lane == 0 ? sre : lane == 1 ? sim : sab;
    }
  }
#endif

  // ---- per-tile partial sums: within a warp in working precision, across warps and tiles in double ----
  constexpr int NV = EXTRAS ? 5 : 1;
  {
    const int wacc = __reduce_add_sync(0xffffffffu, nacc);
    R v[4] = {e.ec, e.ed, e.s2, e.sa};
    if (EXTRAS) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
    }
    if (lane == 0) {
      red[warp * 8] = (double)wacc;
      if (EXTRAS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[warp * 8 + 1 + k] = (double)v[k];
      }
    }
  }
  __syncthreads();
  // the bulk copies of the write-out must have read their shared-memory sources before the CTA retires
  if (bulk_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (!EXTRAS) {
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < HBM_NT / 32; ++w) t += red[w * 8];
      const int r = rm_slot < 0 ? 0 : rm_slot;
      atomicAdd(&A.hdr->nacc[r][(by * gx + bx) & (HBM_NBUCKET - 1)], (unsigned int)t);   // no return value: RED
    }
    return;
  }
  if (tid < NV) {
    double t = 0.0;
    for (int w = 0; w < HBM_NT / 32; ++w) t += red[w * 8 + tid];
    A.partials[(size_t)tid * (gx * gy) + by * gx + bx] = t;
  }
}

// After the last plain sweep of a call: apply its Robbins-Monro update and hand the totals back to the chain state (one warp).
template <typename R>
__global__ void hbm_rm_finalize_kernel(HbmAmp* __restrict__ hdr, double* __restrict__ chain, int Z, int T, uint32_t flags,
                                       int last_slot, int nsweeps) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const unsigned tot = __reduce_add_sync(0xffffffffu, hdr->nacc[last_slot][lane]);
  if (lane != 0) return;
  const double nsite = (double)Z * T;
  double sg = hdr->rm_sigma[last_slot];
  const double cnt = hdr->rm_counter[last_slot];
  if (flags & CYL_SWEEP_ADAPT_SIGMA) sg = rm_batch_update<R>(sg, cnt, nsite, (double)tot);
  chain[CYL_CH_SIGMA_SITE] = sg;
  chain[CYL_CH_STEP_COUNTER] = cnt + nsite;
  chain[CYL_CH_SITE_PROPOSED] += nsite * nsweeps;
  chain[CYL_CH_SITE_ACCEPTED] += hdr->acc_total + (double)tot;
}

// ---- after a sweep: add up the tiles' partial sums (in a fixed order: the result does not depend on scheduling), then the
// per-sweep decisions.  One small CTA (it becomes resident as soon as one tile of the sweep retires).  Everything the decision
// needs besides the sweep's own sums -- the chain state, the proposal header, the parameters, the accumulators it adds to, the
// Robbins-Monro logarithms -- is at least two kernels old and is fetched BEFORE the dependency wait, while the sweep is still
// running; after the wait the kernel costs one round of loads of the partial sums and the decision arithmetic. ----
#define HBM_DECIDE_NT 256
#define HBM_DECIDE_CTAS 8
template <typename R>
struct HbmDecidePre {
  double ch[12];                     // chain state as the previous decision left it (CYL_CH_*)
  double ob[10];                     // observable sums (CYL_OB_COUNT .. CYL_OB_SIGMA)
  double a, a_new, u_acc, e_surf_new, lz, lth, temperature, pt;
  RmLogs<R> rm;
  int try_move, coef_sel;
};
template <typename R>
__device__ void hbm_decide(const HbmArgs<R>& A, const HbmGeom& g, const double* tot, const HbmDecidePre<R>& pre);

// Up to HBM_DECIDE_CTAS CTAs share the tiles' partial sums (one CTA pulls the 160 KB of a 4096^2 sweep at the ~100 GB/s a single
// SM gets from L2: 3 us of the chain); each files the sums of its contiguous share in the header, and the CTA that files last
// (a ticket) adds the shares in index order and decides -- the result does not depend on which CTA that is.  Every CTA prepares
// the decision's inputs ahead of the wait: which one will decide is not known then, and they are idle until the sweep ends.
// DR = tiles per thread and round: 8 for the single CTA of lattices up to 8192 tiles (two rounds at 4096^2), 2 for the CTAs of a
// shared sum, which are also held to 64 registers so that each fits the slot one retiring tile frees.
template <typename R, int DR>
__global__ void __launch_bounds__(HBM_DECIDE_NT, DR == 2 ? 4 : 1)
hbm_decide_kernel(const __grid_constant__ HbmArgs<R> A, const __grid_constant__ HbmGeom g, int ntiles) {
  __shared__ double red[(HBM_DECIDE_NT / 32) * 8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const bool extras = (A.flags & (CYL_SWEEP_AMP_MOVES | CYL_SWEEP_MEASURE)) != 0;
  const int nv = extras ? 5 : 1;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  __shared__ HbmDecidePre<R> pre;      // written and read by thread 0 only; in shared memory to keep the CTA's registers low
  if (tid == 0) {
    const CylParams p = *A.params;
    const HbmAmp* hdr = A.hdr;
    const double pi = 3.14159265358979323846;
#pragma unroll
    for (int q = 0; q < 12; ++q) pre.ch[q] = A.chain[q];
    const bool have_obs = extras && (A.flags & CYL_SWEEP_MEASURE) && A.obs;
#pragma unroll
    for (int q = 0; q < 10; ++q) pre.ob[q] = have_obs ? A.obs[q] : 0.0;
    pre.a = hdr->a; pre.a_new = hdr->a_new; pre.u_acc = hdr->u_acc;               // e_surf_new: warp 1, below
    pre.try_move = hdr->try_move; pre.coef_sel = hdr->coef_sel;
    pre.lz = (2 * pi / p.wavenumber) / g.Z; pre.lth = (2 * pi * p.radius) / g.T;
    pre.temperature = p.temperature;
    pre.pt = (p.amp_target_acceptance > 0) ? p.amp_target_acceptance : 0.3;
    pre.rm = rm_logs<R>(pre.ch[CYL_CH_STEP_COUNTER], (double)g.Z * g.T);
  }
  if (extras && cta == 0)
    for (int t = tid; t < A.ncst; t += HBM_DECIDE_NT) { v[5] += A.cstpart[2 * t]; v[6] += A.cstpart[2 * t + 1]; }
  if (extras && warp == 1) {
    // the proposal's surface energy (an AGM chain of fp64 square roots, ~3 us on one warp): here, under the tail of the sweep,
    // instead of in the preparation kernel, where the next sweep waited for it
    const CylParams p = *A.params;
    const double es = A.hdr->try_move ? surface_energy_warp(p, A.hdr->a_new, nullptr, nullptr, true) : 0.0;
    if (lane == 0) pre.e_surf_new = es;
  }
  pdl_wait();
  pdl_trigger();
  // This CTA's share of the tiles, a thread's tiles in index order.  The loads of a round are issued before anything is added (a
  // loop that adds as it loads pays the ~0.5 us load latency once per iteration: 9 us for a 4096^2 sweep on one CTA, measured with
  // %globaltimer stamps).
  const int share = (ntiles + ncta - 1) / ncta, t_lo = cta * share, t_hi = min(ntiles, t_lo + share);
  for (int t0 = t_lo + tid; t0 < t_hi; t0 += DR * HBM_DECIDE_NT) {
    double x[5][DR];
#pragma unroll
    for (int j = 0; j < DR; ++j) {
      const int t = t0 + j * HBM_DECIDE_NT;
#pragma unroll
      for (int k = 0; k < 5; ++k) x[k][j] = (t < t_hi && k < nv) ? __ldcg(A.partials + (size_t)k * ntiles + t) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < DR; ++j) {
#pragma unroll
      for (int k = 0; k < 5; ++k) v[k] += x[k][j];
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) red[warp * 8 + k] = v[k];
  }
  __syncthreads();
  if (warp != 0) return;
  double tot[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) tot[k] = warp_sum(lane < HBM_DECIDE_NT / 32 ? red[lane * 8 + k] : 0.0);
  if (ncta > 1) {
    int last = 0;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) A.hdr->dpart[cta][k] = tot[k];
      __threadfence();
      last = atomicInc(&A.hdr->decide_ticket, (unsigned int)(ncta - 1)) == (unsigned int)(ncta - 1);
    }
    if (!__shfl_sync(0xffffffffu, last, 0)) return;
    if (lane == 0) {
      __threadfence();
#pragma unroll
      for (int k = 0; k < 7; ++k) tot[k] = 0.0;
      for (int c = 0; c < ncta; ++c) {
#pragma unroll
        for (int k = 0; k < 7; ++k) tot[k] += __ldcg(&A.hdr->dpart[c][k]);
      }
    }
  }
  if (lane == 0) hbm_decide<R>(A, g, tot, pre);
}

// Per-sweep decisions, one thread.  tot = {accepted site moves, E(a,a), E(a',a) - E(a,a), sum |psi|^2, sum |psi|, constants of
// E(a,a), of the difference}, the energies without the pixel-area factor.
template <typename R>
__device__ void hbm_decide(const HbmArgs<R>& A, const HbmGeom& g, const double* v, const HbmDecidePre<R>& pre) {
  double* chain = A.chain;
  const uint32_t flags = A.flags;
  const bool extras = (flags & (CYL_SWEEP_AMP_MOVES | CYL_SWEEP_MEASURE)) != 0;
  const double nsite = (double)g.Z * g.T;
  const double sigma_used = pre.ch[CYL_CH_SIGMA_SITE];
  if (flags & CYL_SWEEP_ADAPT_SIGMA) chain[CYL_CH_SIGMA_SITE] = rm_apply(sigma_used, pre.rm, nsite, v[0]);
  chain[CYL_CH_STEP_COUNTER] = pre.ch[CYL_CH_STEP_COUNTER] + nsite;
  chain[CYL_CH_SITE_PROPOSED] = pre.ch[CYL_CH_SITE_PROPOSED] + nsite;
  chain[CYL_CH_SITE_ACCEPTED] = pre.ch[CYL_CH_SITE_ACCEPTED] + v[0];
  if (!extras) return;
  const double lz = pre.lz, lth = pre.lth;
  const double a = pre.a;
  const double e_dif = (v[2] + v[6]) * lz * lth;
  if (flags & CYL_SWEEP_MEASURE) {
    const double e_cur = (v[1] + v[5]) * lz * lth;                  // :213
    chain[CYL_CH_E_FIELD] = e_cur;
    double* obs = A.obs;
    if (obs) {
      obs[CYL_OB_COUNT] = pre.ob[CYL_OB_COUNT] + 1.0;
      obs[CYL_OB_ABS_A] = pre.ob[CYL_OB_ABS_A] + fabs(a);
      obs[CYL_OB_A] = pre.ob[CYL_OB_A] + a;
      obs[CYL_OB_A2] = pre.ob[CYL_OB_A2] + a * a;
      obs[CYL_OB_PSI2] = pre.ob[CYL_OB_PSI2] + v[3] / nsite;
      obs[CYL_OB_ABSPSI] = pre.ob[CYL_OB_ABSPSI] + v[4] / nsite;
      obs[CYL_OB_E_FIELD] = pre.ob[CYL_OB_E_FIELD] + e_cur;
      obs[CYL_OB_E_SURF] = pre.ob[CYL_OB_E_SURF] + pre.ch[CYL_CH_E_SURF];
      obs[CYL_OB_E_FIELD2] = pre.ob[CYL_OB_E_FIELD2] + e_cur * e_cur;
      obs[CYL_OB_SIGMA] = pre.ob[CYL_OB_SIGMA] + sigma_used;
    }
  }
  if (!(flags & CYL_SWEEP_AMP_MOVES)) return;
  int accept = 0;
  if (pre.try_move) {
    const double dE = (pre.e_surf_new - pre.ch[CYL_CH_E_SURF]) + e_dif;
    if (!isfinite(dE)) chain[CYL_CH_FLAGS] = (double)((unsigned)pre.ch[CYL_CH_FLAGS] | 1u);
    if (dE <= 0) accept = 1;
    else if (pre.temperature == 0) accept = 0;
    else accept = (pre.u_acc <= exp(-dE / pre.temperature));
  }
  const double pt = pre.pt;
  const double ratio = 1.0 / (pt * (1.0 - pt));
  const double amp_counter = pre.ch[CYL_CH_AMP_COUNTER] + 1;
  chain[CYL_CH_AMP_COUNTER] = amp_counter;
  chain[CYL_CH_AMP_PROPOSED] = pre.ch[CYL_CH_AMP_PROPOSED] + 1;
  const double sigma_amp = pre.ch[CYL_CH_SIGMA_AMP];
  if (flags & CYL_SWEEP_ADAPT_SIGMA) {
    const double cf = sigma_amp * ratio * fast_rcp(fmax(amp_counter, 200.0));
    chain[CYL_CH_SIGMA_AMP] = accept ? sigma_amp + cf * (1 - pt) : sigma_amp - cf * pt;
  }
  if (accept) {
    chain[CYL_CH_AMPLITUDE] = pre.a_new;
    chain[CYL_CH_E_SURF] = pre.e_surf_new;
    chain[CYL_CH_AMP_ACCEPTED] = pre.ch[CYL_CH_AMP_ACCEPTED] + 1;
    A.hdr->coef_sel = pre.coef_sel ^ 1;                             // the table built for a' becomes the current one
  }
}

// Add the per-tile-column theta-sums of one sweep into the profile accumulators: thread i = row i.
template <typename R>
__device__ __forceinline__ void hbm_prof_flush_row(const R* __restrict__ rowpart, int ntx, int Z, int i, double* __restrict__ prof) {
  // the loads of 16 tile columns x 3 sums are issued together (a loop that adds as it loads pays the memory latency per
  // iteration; inside the preparation kernel this flush would then outlast the decision it is meant to hide behind)
  double s[3] = {0.0, 0.0, 0.0};
  constexpr int FC = 16;
  for (int t0 = 0; t0 < ntx; t0 += FC) {
    R x[FC][3];
#pragma unroll
    for (int j = 0; j < FC; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) x[j][k] = (t0 + j < ntx) ? rowpart[((size_t)(t0 + j) * 3 + k) * Z + i] : R(0);
    }
#pragma unroll
    for (int j = 0; j < FC; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) s[k] += (double)x[j][k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) prof[3 * i + k] += s[k];
}
template <typename R>
__global__ void hbm_prof_flush_kernel(const R* __restrict__ rowpart, int ntx, int Z, double* __restrict__ prof) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Z) hbm_prof_flush_row<R>(rowpart, ntx, Z, i, prof);
}

// ---- before a sweep with CYL_SWEEP_AMP_MOVES / CYL_SWEEP_MEASURE ----
// Every thread repeats the (counter-based) amplitude draw.  Blocks [0, nrb): thread i = row i builds the row's energy
// weights for (a, a') and the site-move coefficients for a' into the spare table; block nrb evaluates the surface
// energies and fills the header; blocks past it (present when profiles are recorded) add the previous sweep's theta-sums
// into the profile accumulators.
template <typename R>
__global__ void __launch_bounds__(HBM_PREP_NT)
hbm_prep_kernel(const CylParams* __restrict__ params, double* __restrict__ chain, HbmAmp* __restrict__ hdr, int Z, int T, int variant,
                uint64_t seed, uint32_t rep, uint64_t sweep, uint32_t flags, int init, RowCoef<R>* __restrict__ coef2,
                RowW<R>* __restrict__ roww, double* __restrict__ cstpart, const R* __restrict__ rowpart, int ntx,
                double* __restrict__ prof_flush, int nrb) {
  __shared__ double scr[32];
  // Let the next sweep's tiles become resident (and issue their loads) right away.  This kernel starts only after the decision
  // kernel has passed ITS wait, i.e. after the preceding sweep has completed, and before its own wait a sweep tile touches nothing
  // but its input tile -- so the early trigger exposes no unfinished data, and the first wave's load burst (592 tiles, ~4 us) runs
  // under the decision and this preparation instead of after them.
  pdl_trigger();
  // Before the dependency wait: what does not depend on the preceding decision -- the profile flush (theta-sums of the sweep
  // before it, three kernels back), the parameters, the amplitude draw, the replica constants and this row's sin / cos (the
  // fp64 trigonometry is most of this kernel) -- runs while the decision kernel is still summing.
  if ((int)blockIdx.x > nrb) {
    const int i = ((int)blockIdx.x - nrb - 1) * blockDim.x + threadIdx.x;
    if (i < Z) hbm_prof_flush_row<R>(rowpart, ntx, Z, i, prof_flush);
    return;
  }
  const CylParams p = *params;
  const uint4 x = philox4x32<10>(0u, rep, (uint32_t)sweep, philox_c3(CYL_TAG_AMP, sweep), (uint32_t)seed, (uint32_t)(seed >> 32));
  double zn, u;
  amp_draw(x, zn, u);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool row = (int)blockIdx.x < nrb && i < Z;
  const double pi = 3.14159265358979323846;
  const double lz = (2 * pi / p.wavenumber) / Z;
  RepConst<R> rc;
  RowTrig<R> t, tp;
  if (row) {
    rc = make_rep_const<R>(p, Z, T, variant);
    t = row_trig<R>(p.wavenumber, lz, i);
    tp = row_trig<R>(p.wavenumber, lz, (i + 1 == Z) ? 0 : i + 1);
  }
  pdl_wait();
  const double a = chain[CYL_CH_AMPLITUDE];
  double a_new = a + chain[CYL_CH_SIGMA_AMP] * zn;
  const bool try_move = (flags & CYL_SWEEP_AMP_MOVES) && (fabs(a_new) < p.amp_bound);
  if (!try_move) a_new = a;
  if ((int)blockIdx.x == nrb) {
    if (threadIdx.x >= 32) return;
    double es0 = 0.0;
    if (init) es0 = surface_energy_warp(p, a, nullptr, nullptr, true);
    if (threadIdx.x == 0) {
      if (init) chain[CYL_CH_E_SURF] = es0;
      // (the proposal's surface energy is evaluated by the decision kernel, ahead of its dependency wait)
      hdr->a = a; hdr->a_new = a_new; hdr->u_acc = u; hdr->e_surf_new = 0.0; hdr->try_move = try_move ? 1 : 0;
    }
    return;
  }
  const int sel = hdr->coef_sel;
  double c_cur = 0.0, c_dif = 0.0;
  if (row) {
    // same functions as the shared-memory kernel's tables (cyl_rowtab.cuh, working precision): the current weights are the
    // function of the amplitude the smem kernel stores on accept, the proposal's are evaluated with the current amplitude as
    // metric reference (:75), and dif is the difference of the two -- the chain samples the model with these weights
    const AmpGeo<R> ar = amp_geo<R>(rc, a), ae = amp_geo<R>(rc, a_new);
    const GeoT<R> g0 = geo_sc<R>(ar, t.s0, t.c0), gm = geo_sc<R>(ar, t.sm, t.cm);
    const GeoT<R> h0 = geo_sc<R>(ae, t.s0, t.c0), hm = geo_sc<R>(ae, t.sm, t.cm);
    R wc[6], wn[6];
    weight_row<R>(rc, g0, gm, g0, wc);
    weight_row<R>(rc, h0, hm, g0, wn);
    RowW<R> w;
#pragma unroll
    for (int q = 0; q < 5; ++q) { w.cur[q] = wc[q]; w.dif[q] = wn[q] - wc[q]; }
    w.cst[0] = w.cst[1] = R(0);                       // the constants go through the side sums below
    roww[i] = w;
    c_cur = (double)wc[5];
    c_dif = (double)(wn[5] - wc[5]);
    if (try_move) {
      const GeoT<R> hp = geo_sc<R>(ae, tp.sm, tp.cm);
      coef2[(size_t)(sel ^ 1) * Z + i] = coef_row<R>(rc, h0, hm, hp);
    }
  }
  c_cur = block_sum<double>(c_cur, scr);
  c_dif = block_sum<double>(c_dif, scr);
  if (threadIdx.x == 0) { cstpart[2 * blockIdx.x] = c_cur; cstpart[2 * blockIdx.x + 1] = c_dif; }
}

// ----------------------------------------------------------------------------------------------------
// row moments of the plane-split layout: one warp per row
// ----------------------------------------------------------------------------------------------------
template <typename R>
__global__ void hbm_row_moments_kernel(const typename Real<R>::complex_t* __restrict__ work, int cur,
                                       HbmGeom g, const CylParams* __restrict__ params, double* __restrict__ S) {
  using C = typename Real<R>::complex_t;
  __shared__ double red[CYL_MOM_WARPS][CYL_NMOM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpr = moments_warps_per_row(g.T), rows_per_cta = CYL_MOM_WARPS / wpr;
  const int row_in_cta = warp / wpr, seg = warp - row_in_cta * wpr;
  const int i = blockIdx.x * rows_per_cta + row_in_cta;
  double m[CYL_NMOM];
#pragma unroll
  for (int q = 0; q < CYL_NMOM; ++q) m[q] = 0.0;
  if (i < g.Z) {
    const CylParams p = *params;
    const double pi = 3.14159265358979323846;
    const double lz = (2 * pi / p.wavenumber) / g.Z, lth = (2 * pi * p.radius) / g.T;
    const double inv_lz2 = 1.0 / (lz * lz), inv_lth2 = 1.0 / (lth * lth), inv_lth = 1.0 / lth;
    const C* buf = work + (size_t)cur * g.buffer_elems;
    const int64_t ph = g.pitch / 2;
    const int64_t Rr = i + g.hr;
#pragma unroll 4
    for (int j = seg * 32 + lane; j < g.T; j += wpr * 32) {
      const int64_t Q = j + g.hc;
      const C pv = buf[(Q & 1) * g.plane_elems + Rr * ph + (Q >> 1)];
      const C Ln = buf[(Q & 1) * g.plane_elems + (Rr - 1) * ph + (Q >> 1)];          // halo rows make i-1 valid at i = 0
      const C W = buf[((Q - 1) & 1) * g.plane_elems + Rr * ph + ((Q - 1) >> 1)];
      moments_accumulate(m, pv, Ln, W, inv_lz2, inv_lth2, inv_lth);
    }
  }
  moments_store(m, wpr, row_in_cta, seg, i < g.Z ? S + (size_t)i * CYL_NMOM : nullptr, red);
}

template <typename R>
static int hbm_row_moments(const void* work, int cur, int Z, int T, const CylParams* params, double* S, void* stream) {
  CYL_CHECK_ARG(work && (cur == 0 || cur == 1) && params && S && Z >= 8 && T >= 8, "cyl_hbm_row_moments: bad arguments");
  const HbmGeom g = hbm_geom(Z, T);
  const int rows_per_cta = CYL_MOM_WARPS / moments_warps_per_row(T);
  hbm_row_moments_kernel<R><<<(Z + rows_per_cta - 1) / rows_per_cta, CYL_MOM_WARPS * 32, 0, as_stream(stream)>>>(
      (const typename Real<R>::complex_t*)work, cur, g, params, S);
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}
extern "C" int cyl_hbm_row_moments_f32(const void* work, int cur, int Z, int T, const CylParams* params, double* S, void* s) {
  return hbm_row_moments<float>(work, cur, Z, T, params, S, s);
}
extern "C" int cyl_hbm_row_moments_f64(const void* work, int cur, int Z, int T, const CylParams* params, double* S, void* s) {
  return hbm_row_moments<double>(work, cur, Z, T, params, S, s);
}

// ----------------------------------------------------------------------------------------------------
// host driver
// ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CYL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
    if (!f || qres != cudaDriverEntryPointSuccess) {
      cyl_set_error("cuTensorMapEncodeTiled not available from the driver (query result %d)", (int)qres);
      return CYL_E_CUDA;
    }
    cached = (EncodeTiledFn)f;
  }
  *fn = cached;
  return CYL_OK;
}

template <typename R>
static int make_maps(const HbmGeom& g, void* work, int box_cols, int box_rows, HbmMaps* maps) {
  using C = typename Real<R>::complex_t;
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  // one complex64 = one UINT64 element; one complex128 = two FLOAT64 elements (TMA has no 16-byte type)
  // (the box start along the contiguous dimension must fall on a 16-byte boundary, else the load traps)
  const int esz = sizeof(C) / 8;
  const CUtensorMapDataType dt = (sizeof(C) == 8) ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  for (int b = 0; b < 2; ++b)
    for (int pl = 0; pl < 2; ++pl) {
      void* base = (C*)work + (size_t)b * g.buffer_elems + (size_t)pl * g.plane_elems;
      const cuuint64_t dims[2] = {(cuuint64_t)(g.pitch / 2) * esz, (cuuint64_t)g.rows};
      const cuuint64_t strides[1] = {(cuuint64_t)(g.pitch / 2) * sizeof(C)};
      const cuuint32_t box[2] = {(cuuint32_t)(box_cols * esz), (cuuint32_t)box_rows};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = enc(&maps->m[b][pl], dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        cyl_set_error("cuTensorMapEncodeTiled failed (%d): pitch %lld rows %lld box %u x %u", (int)r, (long long)g.pitch,
                      (long long)g.rows, box[0], box[1]);
        return CYL_E_CUDA;
      }
    }
  maps->esz = esz;
  return CYL_OK;
}

template <typename R, int NCOL, int TH_>
static int sweep_hbm_impl(const HbmGeom& g, void* work, int32_t* cur, const CylParams* params, double* chain, uint64_t seed,
                          int64_t replica, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, void* scratch,
                          double* obs, double* prof, cudaStream_t st) {
  using C = typename Real<R>::complex_t;
  using L = HbmSmem<R, NCOL, TH_>;
  static_assert((L::SWH * sizeof(C)) % 16 == 0, "TMA box rows must be a multiple of 16 bytes");
  static_assert(TH_ >= HBM_MIN_TH, "scratch is sized for tiles of at least HBM_MIN_TH rows");
  HbmMaps maps;
  int rc = make_maps<R>(g, work, L::SWH, L::SH, &maps);
  if (rc) return rc;
  const HbmScratch<R> sl(g.Z, g.T);
  const bool extras = (flags & (CYL_SWEEP_AMP_MOVES | CYL_SWEEP_MEASURE)) != 0;
  const bool profiles = (flags & CYL_SWEEP_MEASURE) && prof;
  char* sb = (char*)scratch;
  HbmArgs<R> A;
  A.hdr = reinterpret_cast<HbmAmp*>(sb);
  RowCoef<R>* coef2 = reinterpret_cast<RowCoef<R>*>(sb + sl.off_coef);
  RowW<R>* roww = reinterpret_cast<RowW<R>*>(sb + sl.off_roww);
  double* cstpart = reinterpret_cast<double*>(sb + sl.off_cst);
  R* rowpart = reinterpret_cast<R*>(sb + sl.off_rowp);
  A.coef2 = coef2; A.roww = roww; A.params = params; A.chain = chain;
  A.partials = reinterpret_cast<double*>(sb + sl.off_part);
  A.cstpart = cstpart; A.rowpart = profiles ? rowpart : nullptr; A.obs = obs; A.flags = flags; A.ncst = sl.ncst;
  auto kern = extras ? sweep_hbm_kernel<R, NCOL, TH_, true> : sweep_hbm_kernel<R, NCOL, TH_, false>;
  CYL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
  const dim3 grid((g.T + L::TW - 1) / L::TW, (g.Z + L::TH - 1) / L::TH);
  const SiteKeys keys = site_keys_expand(seed, (uint64_t)replica);
  const int nrb = (g.Z + HBM_PREP_NT - 1) / HBM_PREP_NT;
  const int ntiles = (int)(grid.x * grid.y);
  // one CTA up to 8192 tiles (at 4096 tiles eight CTAs measured 2 us slower than one: 101.0 against 98.9 us per sweep), eight beyond
  // (8192^2 = 16384 tiles: 351 against 385 us per sweep)
  const int ndecide = ntiles > 8192 ? HBM_DECIDE_CTAS : 1;
  hbm_coef_kernel<R><<<nrb, HBM_PREP_NT, 0, st>>>(params, chain, g.Z, g.T, variant, coef2, A.hdr);
  for (int s = 0; s < nsweeps; ++s) {
    const uint64_t sweep = sweep0 + (uint64_t)s;
    if (extras) {
      const bool flush = profiles && s > 0;
      CYL_CUDA(launch_pdl(hbm_prep_kernel<R>, dim3(nrb + 1 + (flush ? nrb : 0)), dim3(HBM_PREP_NT), 0, st, params, chain, A.hdr, g.Z,
                          g.T, variant, seed, (uint32_t)replica, sweep, flags, (int)(s == 0), coef2, roww, cstpart,
                          (const R*)rowpart, sl.ntx, flush ? prof : (double*)nullptr, nrb));
    }
    CYL_CUDA(launch_pdl(kern, grid, dim3(HBM_NT), L::total, st, maps, (C*)work, (int)*cur, g, A, keys, sweep, s == 0 ? -1 : s % 3));
    *cur ^= 1;                                                      // host-side ping-pong index
    if (extras) {
      if (ndecide > 1) CYL_CUDA(launch_pdl(hbm_decide_kernel<R, 2>, dim3(ndecide), dim3(HBM_DECIDE_NT), 0, st, A, g, ntiles));
      else CYL_CUDA(launch_pdl(hbm_decide_kernel<R, 8>, dim3(1), dim3(HBM_DECIDE_NT), 0, st, A, g, ntiles));
    }
  }
  if (!extras && nsweeps > 0)
    CYL_CUDA(launch_pdl(hbm_rm_finalize_kernel<R>, dim3(1), dim3(32), 0, st, A.hdr, chain, g.Z, g.T, flags, (nsweeps - 1) % 3, nsweeps));
  if (profiles && nsweeps > 0)
    CYL_CUDA(launch_pdl(hbm_prof_flush_kernel<R>, dim3(nrb), dim3(HBM_PREP_NT), 0, st, (const R*)rowpart, sl.ntx, g.Z, prof));
  CYL_LAUNCH_CHECK();
  return CYL_OK;
}

// Relative cost of one sweep with tile height TH: waves of co-resident CTAs x site moves per tile (incl. the redundant halo).
template <typename R, int NCOL, int TH_>
static double tile_cost(const HbmGeom& g, int sms) {
  using L = HbmSmem<R, NCOL, TH_>;
  int occ = 0;
  auto kern = sweep_hbm_kernel<R, NCOL, TH_, true>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HBM_NT, L::total) != cudaSuccess || occ <= 0) {
    cudaGetLastError();
    return 1e300;
  }
  const double tiles = (double)((g.T + L::TW - 1) / L::TW) * ((g.Z + TH_ - 1) / TH_);
  const double waves = ceil(tiles / ((double)sms * occ));
  double moves = 0;
  for (int c = 0; c < NCOL; ++c) moves += (double)(TH_ + 2 * (NCOL - 1 - c)) * (L::TW / NCOL + 2);
  return waves * moves * (1.0 + 600.0 / (moves * 8));        // + fixed per-tile cost (load wait, barriers) ~ 600 moves' worth
}

#define HBM_TRY_TH(TH_)                                                                                                   \
  do {                                                                                                                    \
    const double c__ = tile_cost<R, NCOL, TH_>(g, sms);                                                                   \
    if (c__ < best_cost) { best_cost = c__; best = TH_; }                                                                 \
  } while (0)
#define HBM_RUN_TH(TH_)                                                                                                   \
  case TH_:                                                                                                               \
    return sweep_hbm_impl<R, NCOL, TH_>(g, work, cur, params, chain, seed, replica, sweep0, nsweeps, variant, flags, scratch, obs, prof, st)

template <typename R, int NCOL>
static int sweep_hbm_dispatch(const HbmGeom& g, void* work, int32_t* cur, const CylParams* params, double* chain, uint64_t seed,
                              int64_t replica, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, void* scratch,
                              double* obs, double* prof, cudaStream_t st) {
  int dev = 0, sms = 0;
  CYL_CUDA(cudaGetDevice(&dev));
  CYL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int best = 32;
  double best_cost = 1e300;
  const int force = (int)((flags >> 8) & 0xFFu);                    // CYL_SWEEP_TILE_ROWS(n): explicit tile height
  flags &= ~0xFF00u;
  if (force) best = force;
  else { HBM_TRY_TH(32); HBM_TRY_TH(30); HBM_TRY_TH(28); HBM_TRY_TH(24); }
  switch (best) {
    HBM_RUN_TH(32);
    HBM_RUN_TH(30);
    HBM_RUN_TH(28);
    HBM_RUN_TH(24);
    default:
      cyl_set_error("cyl_sweep_hbm: unsupported tile height %d (32, 30, 28, 24)", best);
      return CYL_E_UNSUPPORTED;
  }
}

template <typename R>
static int sweep_hbm(void* work, int32_t* cur, int Z, int T, const CylParams* params, double* chain, uint64_t seed,
                     int64_t replica, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, void* scratch, double* obs,
                     double* prof, void* stream) {
  CYL_CHECK_ARG(work && cur && (*cur == 0 || *cur == 1) && params && chain && scratch && nsweeps >= 0, "cyl_sweep_hbm: bad arguments");
  CYL_CHECK_ARG(Z >= 8 && T >= 8, "cyl_sweep_hbm: the HBM path needs Z, T >= 8 (got %d x %d)", Z, T);
  CYL_CHECK_ARG(variant >= 0 && variant < CYL_NUM_VARIANTS, "cyl_sweep_hbm: unknown variant %d", variant);
  CYL_CHECK_ARG(!(flags & CYL_SWEEP_MEASURE) || obs, "cyl_sweep_hbm: CYL_SWEEP_MEASURE needs obs");
  CYL_CHECK_ARG(((uintptr_t)work & 127) == 0, "cyl_sweep_hbm: work must be 128-byte aligned");
  const HbmGeom g = hbm_geom(Z, T);
  if (nsweeps == 0) return CYL_OK;
  if (g.ncol == 2)
    return sweep_hbm_dispatch<R, 2>(g, work, cur, params, chain, seed, replica, sweep0, nsweeps, variant, flags, scratch, obs, prof, as_stream(stream));
  return sweep_hbm_dispatch<R, 3>(g, work, cur, params, chain, seed, replica, sweep0, nsweeps, variant, flags, scratch, obs, prof, as_stream(stream));
}

extern "C" int cyl_sweep_hbm_f32(void* work, int32_t* cur, int Z, int T, const CylParams* params, double* chain, uint64_t seed,
                                 int64_t replica, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, void* scratch,
                                 double* obs, double* prof, void* stream) {
  return sweep_hbm<float>(work, cur, Z, T, params, chain, seed, replica, sweep0, nsweeps, variant, flags, scratch, obs, prof, stream);
}
extern "C" int cyl_sweep_hbm_f64(void* work, int32_t* cur, int Z, int T, const CylParams* params, double* chain, uint64_t seed,
                                 int64_t replica, uint64_t sweep0, int nsweeps, int variant, uint32_t flags, void* scratch,
                                 double* obs, double* prof, void* stream) {
  return sweep_hbm<double>(work, cur, Z, T, params, chain, seed, replica, sweep0, nsweeps, variant, flags, scratch, obs, prof, stream);
}
