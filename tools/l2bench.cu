// Microbenchmark: how fast can one CTA pull an 88 KB utterance into shared memory?
//   mode 0: cp.async.bulk (TMA) chunks + mbarrier     mode 1: LDG.128 -> STS.128 by all threads
//   mode 2: cp.async (LDGSTS) 16 B per thread
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loadbench loadbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 2) k(const unsigned char* src, int64_t n_utts, int bytes, int mode, int chunk,
                                           unsigned* counter, unsigned long long* sink, long long* wait_cycles) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ int s_u;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_u = (int)atomicAdd(counter, 1u);
  }
  __syncthreads();
  uint32_t parity = 0;
  unsigned long long acc = 0;
  long long waited = 0;
  int u = s_u;
  while (u < n_utts) {
    const unsigned char* p = src + (int64_t)(u % 296) * bytes;
    long long t0 = clock64();
    if (mode == 0) {
      if (tid < 32) {
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        __syncwarp();
        for (int o = tid * chunk; o < bytes; o += 32 * chunk) {
          int sz = min(chunk, bytes - o);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + o)), "l"(p + o), "r"(sz), "r"(smem_u32(&bar)) : "memory");
        }
      }
      asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(parity) : "memory");
      parity ^= 1;
    } else if (mode == 1) {
      const int4* s4 = reinterpret_cast<const int4*>(p);
      int4* d4 = reinterpret_cast<int4*>(sm);
      const int nv = bytes / 16;
      int v = tid;
      for (; v + 3 * 256 < nv; v += 4 * 256) {
        int4 a = __ldg(s4 + v), b = __ldg(s4 + v + 256), c = __ldg(s4 + v + 512), d = __ldg(s4 + v + 768);
        d4[v] = a; d4[v + 256] = b; d4[v + 512] = c; d4[v + 768] = d;
      }
      for (; v < nv; v += 256) d4[v] = __ldg(s4 + v);
    } else {
      const int nv = bytes / 16;
      for (int v = tid; v < nv; v += 256) __pipeline_memcpy_async(sm + 16 * v, p + 16 * v, 16);
      __pipeline_commit();
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    waited += clock64() - t0;
    acc += reinterpret_cast<unsigned*>(sm)[tid];   // touch
    if (tid == 0) s_u = (int)atomicAdd(counter, 1u);
    __syncthreads();
    u = s_u;
  }
  if (acc == 0x1234567) *sink = acc;
  if (tid == 0) atomicAdd((unsigned long long*)wait_cycles, (unsigned long long)waited);
}
int main() {
  const int bytes = 88208; const int64_t n_distinct = 296; const int64_t n = 40000;
  unsigned char* src; cudaMalloc(&src, n_distinct * bytes + 256); cudaMemset(src, 1, n_distinct * bytes);
  unsigned* counter; cudaMalloc(&counter, 4); unsigned long long* sink; cudaMalloc(&sink, 8); long long* wc; cudaMalloc(&wc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int mode = 0; mode < 3; ++mode)
      for (int chunk : {4096, 16384}) {
        if (mode != 0 && chunk != 4096) continue;
        float best = 1e9; long long w = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemset(counter, 0, 4); cudaMemset(wc, 0, 8);
          cudaEventRecord(e0);
          k<<<148 * ctas, 256, ctas == 1 ? 100000 : 96000>>>(src, n, bytes, mode, chunk, counter, sink, wc);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
          cudaMemcpy(&w, wc, 8, cudaMemcpyDeviceToHost);
        }
        printf("ctas/SM=%d mode=%d chunk=%5d  %.3f ms  %.1f GB/s  avg load %.0f cycles/utt  err=%s\n", ctas, mode, chunk, best,
               n * (double)bytes / best / 1e6, (double)w / n, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
