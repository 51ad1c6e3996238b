// Per-instruction issue cost on one SM sub-partition (B200): each kernel runs NCHAIN independent dependency chains per thread of
// one opcode, 8 warps per sub-partition, and reports warp-instructions per cycle per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes scripts/micro/pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define NCH 8
template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, float a, float b, unsigned m, long long* cyc) {
  float x[NCH], y[NCH];
  unsigned u[NCH], v[NCH];
  unsigned long long p[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { x[i] = a + i + threadIdx.x; y[i] = b - i; u[i] = m + i * 77u + threadIdx.x; v[i] = m * 3u + i; p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(y[i]); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (OP == 0) x[i] = fmaf(x[i], a, b);                                                   // FFMA, 2 of 3 sources uniform-ish regs
      if (OP == 1) x[i] = fmaf(x[i], y[i], x[(i + 1) % NCH]);                                  // FFMA, 3 distinct register sources
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) % NCH]), "l"(p[(i + 2) % NCH]));   // FFMA2
      if (OP == 3) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(u[i]), "r"(0xD256D193u)); u[i] = (unsigned)(w >> 32) ^ (unsigned)w; }  // IMAD.WIDE + LOP3
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(m));                         // LOP3
      if (OP == 5) x[i] = x[i] + y[i];                                                          // FADD
      if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));                     // MUFU
      if (OP == 7) { x[i] = fmaf(x[i], y[i], a); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(m)); }   // FFMA + LOP3 alternating
      if (OP == 8) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(u[i]), "r"(0xD256D193u)); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u[i]) : "r"((unsigned)(w >> 32)), "r"((unsigned)w), "r"(m)); x[i] = fmaf(x[i], y[i], a); }  // Philox round + FFMA
      if (OP == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(v[i]), "r"(m));  // IMAD
      if (OP == 10) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) % NCH]));   // FMUL2
      if (OP == 11) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) % NCH]), "l"(p[(i + 2) % NCH])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(m)); }  // FFMA2 + LOP3
    }
  }
  const long long t1 = clock64();
  float s = 0; unsigned us = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) { s += x[i] + y[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]); us ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + us;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 1024>>>(out, 1.0001f, 0.5f, 12345u, cyc);
  k<OP><<<148, 1024>>>(out, 1.0001f, 0.5f, 12345u, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double inst = (double)ITERS * NCH * per_iter * 8;   // warp-instructions per sub-partition (8 warps)
  printf("%-34s %6.3f warp-inst/cycle/SMSP  (%.2f cycles per warp-instruction)\n", name, inst / c, c / inst);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA (imm/uniform sources)", 1);
  run<1>("FFMA (3 register sources)", 1);
  run<2>("FFMA2", 1);
  run<10>("FMUL2", 1);
  run<5>("FADD", 1);
  run<9>("IMAD", 1);
  run<3>("IMAD.WIDE + LOP3", 2);
  run<4>("LOP3", 1);
  run<6>("MUFU.LG2", 1);
  run<7>("FFMA + LOP3", 2);
  run<11>("FFMA2 + LOP3", 2);
  run<8>("IMAD.WIDE + LOP3 + FFMA", 3);
  return 0;
}
