// Shared device/host helpers: error plumbing, surface geometry (system_cylinder.py:25-41), Philox4x32-10,
// uniform/normal conversion, warp/CTA reductions.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/cylinder_b200.h"

// ----------------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------------
void cyl_set_error(const char* fmt, ...);

#define CYL_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      cyl_set_error(__VA_ARGS__);                                 \
      return CYL_E_INVALID;                                       \
    }                                                             \
  } while (0)

#define CYL_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cyl_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return CYL_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

#define CYL_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      cyl_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));      \
      return CYL_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ----------------------------------------------------------------------------------------------------
// real-type traits
// ----------------------------------------------------------------------------------------------------
template <typename R> struct Real;
template <> struct Real<float> {
  using complex_t = float2;
  static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <> struct Real<double> {
  using complex_t = double2;
  static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
};

// ----------------------------------------------------------------------------------------------------
// surface geometry r(z) = r(a)(1 + a sin kz)       system_cylinder.py:25-41
// Always evaluated in fp64: it is O(Z) work per amplitude change, never per site.
// ----------------------------------------------------------------------------------------------------
struct Geo {
  double gth;   // sqrt_g_theta   :28
  double gz;    // sqrt_g_z       :34
  double ath;   // A_theta        :40
};

__host__ __device__ __forceinline__ double geo_radius_rescaled(double r, double a) {   // :37
  return r / sqrt(1.0 + a * a / 2.0);
}

__host__ __device__ __forceinline__ Geo geo_eval(double k, double r, double a, double z) {
  Geo g;
  const double ra = geo_radius_rescaled(r, a);
  double s, c;
  sincos(k * z, &s, &c);
  g.gth = ra * (1.0 + a * s);                           // :28
  const double t = ra * a * k * c;
  g.gz = sqrt(1.0 + t * t);                             // :34
  g.ath = k * ra * a * c / g.gz;                        // :40
  return g;
}

// fp64 reciprocal / reciprocal square root from the fp32 MUFU seed + two Newton steps (relative error ~2e-16 for
// arguments well inside the fp32 range, which every metric factor is).  The sweep kernels rebuild O(Z) metric
// coefficients per amplitude proposal on a handful of threads: that phase is latency-bound, and the IEEE division
// and square-root sequences are its critical path.  The parity-critical entry points (geometry_eval, replay) keep
// the IEEE operations.
__device__ __forceinline__ double fast_rcp(double x) {
  double r = (double)__frcp_rn((float)x);
  r = r * fma(-x, r, 2.0);
  r = r * fma(-x, r, 2.0);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  r = r * fma(-hx * r, r, 1.5);
  r = r * fma(-hx * r, r, 1.5);
  return r;
}
__device__ __forceinline__ double fast_radius_rescaled(double r, double a) {             // :37
  return r * fast_rsqrt(fma(0.5 * a, a, 1.0));
}

// Same from cached sin(kz), cos(kz) -- these do not depend on the amplitude, so the sweep kernels tabulate them
// once per launch and re-evaluate the metric after every accepted amplitude move without trigonometry.
__device__ __forceinline__ Geo geo_from_sc(double k, double ra, double a, double s, double c) {
  Geo g;
  g.gth = ra * (1.0 + a * s);
  const double t = ra * a * k * c;
  const double x = fma(t, t, 1.0);
  const double rs = fast_rsqrt(x);
  g.gz = x * rs;
  g.ath = k * ra * a * c * rs;
  return g;
}

// ----------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter-based: no state, any (counter, key) in O(1).
// ----------------------------------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t m, uint32_t c, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  uint64_t p;                                   // one IMAD.WIDE.U32; hi/lo are the register pair
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(m), "r"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
#else
  const uint64_t p = (uint64_t)m * c;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
  uint32_t h0, l0, h1, l1;
  philox_mulhilo(PHILOX_M0, c0, h0, l0);
  philox_mulhilo(PHILOX_M1, c2, h1, l1);
  const uint32_t n0 = h1 ^ c1 ^ k0;
  const uint32_t n2 = h0 ^ c3 ^ k1;
  c0 = n0; c1 = l1; c2 = n2; c3 = l0;
}

template <int ROUNDS = 10>
__host__ __device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < ROUNDS; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// The ten round keys of one 64-bit seed, expanded once on the host and passed BY VALUE as a kernel parameter: the hot
// loops then read them as constant-bank operands of the XOR (LOP3 ..., c[0x0][..]) instead of recomputing the Weyl
// sequence for every site.
struct PhiloxKeys { uint32_t k[20]; };
static inline PhiloxKeys philox_expand(uint64_t seed) {
  PhiloxKeys ks;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) { ks.k[2 * i] = k0; ks.k[2 * i + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
  return ks;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& ks) {
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c0, c1, c2, c3, ks.k[2 * i], ks.k[2 * i + 1]);
  return make_uint4(c0, c1, c2, c3);
}

// ----------------------------------------------------------------------------------------------------
// Philox2x32-10 for the site moves.  A proposal needs 64 random bits (23-bit Box-Muller radius, 16-bit angle, 23-bit
// acceptance uniform -- cyl_site.cuh), so the hot loops run the two-word generator: 10 multiplies instead of 20, and
// the 32x32->64 multiply is the instruction that loads the busiest pipe of these kernels.
//   counter = (global site index, sweep + off),  key = key
// where (key, off) identify the replica's stream: two words of Philox4x32-10(replica; seed), so that streams of
// different replicas / seeds differ in 64 pseudo-random bits (a bare 32-bit key would collide among 65536 replicas
// with probability ~1/2).
// ----------------------------------------------------------------------------------------------------
#define PHILOX2_M 0xD256D193u

__host__ __device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi, lo;
    philox_mulhilo(PHILOX2_M, c0, hi, lo);
    c0 = hi ^ key ^ c1;
    c1 = lo;
    key += PHILOX_W0;
  }
  return make_uint2(c0, c1);
}

// round keys expanded on the host for single-replica kernels (constant-bank operands of the XOR)
struct SiteKeys { uint32_t k[10]; uint32_t off; uint32_t pad; };
// a ^ b ^ c as ONE three-input logic op.  Left to itself the compiler combines lo ^ key early and pays a second LOP3 per
// round to shorten the dependency chain; these kernels are bound by instruction issue, not by that latency.
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, const SiteKeys& ks) {
  uint32_t hi, lo;
  philox_mulhilo(PHILOX2_M, c0, hi, lo);
  c0 = hi ^ (ks.k[0] ^ c1);                          // c1 = sweep counter: uniform, so is k[0] ^ c1
  c1 = lo;
#pragma unroll
  for (int i = 1; i < 10; ++i) {
    philox_mulhilo(PHILOX2_M, c0, hi, lo);
    c0 = xor3(hi, ks.k[i], c1);
    c1 = lo;
  }
  return make_uint2(c0, c1);
}

// the same with the round keys in registers; kc = k[0] ^ c1 (c1 = the sweep counter word)
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t kc, const uint32_t (&k)[10]) {
  uint32_t hi, lo;
  philox_mulhilo(PHILOX2_M, c0, hi, lo);
  c0 = hi ^ kc;
  c1 = lo;
#pragma unroll
  for (int i = 1; i < 10; ++i) {
    philox_mulhilo(PHILOX2_M, c0, hi, lo);
    c0 = xor3(hi, k[i], c1);
    c1 = lo;
  }
  return make_uint2(c0, c1);
}

// Stream tags in counter word 3 (upper byte), so that site moves, amplitude moves and initialisation of the
// same (site, replica, sweep) never share a counter.
#define CYL_TAG_SITE 0u
#define CYL_TAG_AMP  1u
#define CYL_TAG_INIT 2u
__host__ __device__ __forceinline__ uint32_t philox_c3(uint32_t tag, uint64_t sweep) {
  return (tag << 24) | (uint32_t)((sweep >> 32) & 0xFFFFFFu);
}

// identity of one replica's site-move stream
struct SiteStream { uint32_t key, off; };
__host__ __device__ __forceinline__ SiteStream site_stream(uint64_t seed, uint64_t replica) {
  const uint4 x = philox4x32<10>((uint32_t)replica, (uint32_t)(replica >> 32), 0u, philox_c3(CYL_TAG_SITE, 0), (uint32_t)seed,
                                 (uint32_t)(seed >> 32));
  SiteStream s;
  s.key = x.x;
  s.off = x.y;
  return s;
}
static inline SiteKeys site_keys_expand(uint64_t seed, uint64_t replica) {
  const SiteStream st = site_stream(seed, replica);
  SiteKeys ks;
  uint32_t k = st.key;
  for (int i = 0; i < 10; ++i) { ks.k[i] = k; k += PHILOX_W0; }
  ks.off = st.off;
  ks.pad = 0;
  return ks;
}

// uniforms strictly inside (0,1)
__host__ __device__ __forceinline__ double u01_f64(uint32_t x) {            // 32 bits, midpoint
  return ((double)x + 0.5) * (1.0 / 4294967296.0);
}
// bit fields of a site draw (x0, x1): x0[31:9] radius, x1[31:9] acceptance, angle = x0[7:0] : x1[7:0]
__host__ __device__ __forceinline__ double u23_f64(uint32_t x) {            // top 23 bits, midpoint
  return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0);
}
__host__ __device__ __forceinline__ uint32_t angle16(uint32_t x0, uint32_t x1) { return ((x0 & 0xFFu) << 8) | (x1 & 0xFFu); }
__host__ __device__ __forceinline__ double u16_f64(uint32_t k) { return ((double)k + 0.5) * (1.0 / 65536.0); }
// fp32: the top 23 bits at the midpoint of their cell, (m + 1/2) 2^-23 in (0,1), built without an int->float
// conversion: [1,2) mantissa trick, then an exact subtraction of 1 - 2^-24.  Tracks u01_f64 to 2^-24.
__device__ __forceinline__ float u01_f32(uint32_t x) {
  return __uint_as_float(0x3F800000u | (x >> 9)) - 0.99999994f;
}
// 16-bit field at the midpoint of its cell, (k + 1/2) 2^-16, same trick (exact)
__device__ __forceinline__ float u16_f32(uint32_t k) {
  return __uint_as_float(0x3F800000u | (k << 7)) - 0.99999237060546875f;     // 1 - 2^-17
}

// ----------------------------------------------------------------------------------------------------
// reductions
// ----------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA; result valid in every thread.  scratch: >= 32 elements of T in shared memory.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();                      // scratch may still be read from a previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T t = (lane < nwarp) ? scratch[lane] : T(0);
  t = warp_sum(t);
  return t;
}
