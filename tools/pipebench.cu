// Instruction-throughput microbenchmark for sm_100a: cycles per warp-instruction per SM sub-partition for the
// opcodes the front-end kernels lean on.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int OP>
__global__ void bench(long long* out, int iters, float seed) {
  constexpr int K = 8;   // independent chains
  float f[K]; uint32_t u[K]; u64 p[K];
  for (int k = 0; k < K; ++k) { f[k] = seed + k + threadIdx.x; u[k] = (uint32_t)(seed * 77) + k * 31 + threadIdx.x; p[k] = ((u64)u[k] << 32) | (u[k] * 3u); }
  const float c = seed * 0.5f; const uint32_t cu = (uint32_t)seed + 5;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (OP == 0) f[k] = fmaf(f[k], c, f[k]);
      if (OP == 1) p[k] = fma2(p[k], p[(k + 1) % K], p[k]);
      if (OP == 2) p[k] = add2(p[k], p[(k + 1) % K]);
      if (OP == 3) u[k] = (u[k] & cu) ^ u[(k + 1) % K];                     // LOP3
      if (OP == 4) u[k] = __byte_perm(u[k], u[(k + 1) % K], 0x7531);         // PRMT
      if (OP == 5) u[k] = __funnelshift_l(u[k], u[(k + 1) % K], 1);          // SHF
      if (OP == 6) u[k] = __shfl_xor_sync(0xffffffffu, u[k], 1);             // SHFL
      if (OP == 7) u[k] = __dp4a((int)u[k], (int)u[(k + 1) % K], (int)u[k]); // IDP.4A
      if (OP == 8) u[k] = __vimin3_s16x2(u[k], u[(k + 1) % K], cu);          // VIMNMX3
      if (OP == 9) u[k] = u[k] + u[(k + 1) % K];                             // IADD
      if (OP == 10) f[k] = f[k] + c;                                         // FADD
      if (OP == 11) u[k] = u[k] * cu + u[(k + 1) % K];                       // IMAD
      if (OP == 12) u[k] = __popc(u[k]) + u[(k+1)%K];                        // POPC (+IADD)
      if (OP == 13) f[k] = (float)(int)u[k] + f[k];                          // I2F (+FADD)
      // mixes: do instructions of different pipes overlap (1 inst/cycle) or share one half-rate dispatch port?
      if (OP == 14) { if (k & 1) u[k] = (u[k] & cu) ^ u[(k + 1) % K]; else u[k] = u[k] * cu + u[(k + 1) % K]; }                 // LOP3 | IMAD
      if (OP == 15) { u[k] = (u[k] & cu) ^ u[(k + 1) % K]; f[k] = fmaf(f[k], c, f[k]); }                                        // LOP3 + FFMA
      if (OP == 16) { if (k & 1) u[k] = __byte_perm(u[k], u[(k + 1) % K], 0x7531); else u[k] = __dp4a((int)u[k], (int)u[(k + 1) % K], (int)u[k]); }  // PRMT | IDP
      if (OP == 17) { u[k] = u[k] * cu + u[(k + 1) % K]; f[k] = fmaf(f[k], c, f[k]); }                                          // IMAD + FFMA
      if (OP == 18) { u[k] = __dp4a((int)u[k], (int)u[(k + 1) % K], (int)u[k]); f[k] = fmaf(f[k], c, f[k]); }                   // IDP + FFMA
      if (OP == 19) { if (k & 1) u[k] = __vimin3_s16x2(u[k], u[(k + 1) % K], cu); else u[k] = __dp4a((int)u[k], (int)u[(k + 1) % K], (int)u[k]); }   // VIMNMX3 | IDP
      if (OP == 20) { u[k] = (u[k] & cu) ^ u[(k + 1) % K]; f[k] = f[k] + c; }                                                   // LOP3 + FADD
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int k = 0; k < K; ++k) s += f[k] + (float)u[k] + (float)(p[k] & 0xffff);
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)s; }
}
template <int OP> void run(const char* name, long long* d, int extra) {
  for (int warps : {1, 4, 8, 16}) {
    long long h[2]; const int iters = 2000;
    bench<OP><<<1, warps * 32>>>(d, iters, 1.5f); cudaDeviceSynchronize();
    bench<OP><<<1, warps * 32>>>(d, iters, 1.5f); cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (iters * 8.0 * (1 + extra));   // cycles per instruction of one warp
    const double wps = warps / 4.0 < 1 ? 1 : warps / 4.0;             // warps per sub-partition
    printf("%-10s warps=%2d  cycles/inst/warp=%6.2f  -> cycles per warp-inst per SMSP=%5.2f\n", name, warps, per, per / wps);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0>("FFMA", d, 0); run<1>("FFMA2", d, 0); run<2>("FADD2", d, 0); run<3>("LOP3", d, 0); run<4>("PRMT", d, 0);
  run<5>("SHF", d, 0); run<6>("SHFL", d, 0); run<7>("IDP.4A", d, 0); run<8>("VIMNMX3", d, 0); run<9>("IADD", d, 0);
  run<10>("FADD", d, 0); run<11>("IMAD", d, 0); run<12>("POPC+IADD", d, 1); run<13>("I2F+FADD", d, 1);
  run<14>("LOP3|IMAD", d, 0); run<15>("LOP3+FFMA", d, 1); run<16>("PRMT|IDP", d, 0); run<17>("IMAD+FFMA", d, 1); run<18>("IDP+FFMA", d, 1);
  run<19>("VIMNMX3|IDP", d, 0); run<20>("LOP3+FADD", d, 1);
  return 0;
}
