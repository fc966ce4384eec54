// Probe: can ONE tensor-map TMA op fetch 2048 contiguous int16 samples starting at an ARBITRARY sample offset, and at
// what rate?  The sample buffer is described as an overlapped 3-D tensor
//     t[z][y][x] = buf[256 z + 8 y + x]      extents (263, 32, nz), strides (2 B, 16 B, 512 B)
// so a box (256, 1, 8) at (s & 7, (s >> 3) & 31, s >> 8) is buf[s .. s + 2048).  Compared with cp.async.bulk (16-byte
// aligned starts only).  nvcc -arch=sm_100a --cudart shared tools/tma_probe.cu -o tools/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int kChunk = 2048, kStages = 16, kThreads = 128;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}

__host__ __device__ inline uint16_t pattern(uint64_t i) { return (uint16_t)((i * 2654435761ull) >> 13); }

// mode 0: cp.async.bulk (start must be 16-byte aligned); mode 1: tensor map
__global__ void __launch_bounds__(kThreads) probe(const __grid_constant__ CUtensorMap tmap, const uint16_t* buf, int64_t chunks_per_cta,
                                                  int64_t stride, int shift, int mode, int verify, unsigned long long* bad) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kChunk * 2);
  uint64_t* empty = full + kStages;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(kThreads / 32 - 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * stride + shift;
  if (tid < 32) {
    if (tid == 0) {
      int st = 0; uint32_t ph = 0;
      for (int64_t c = 0; c < chunks_per_cta; ++c) {
        mbar_wait(&empty[st], ph ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(kChunk * 2) : "memory");
        const int64_t s = base + c * kChunk;
        const uint32_t dst = s32(smem + st * kChunk * 2);
        if (mode == 0) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(buf + s),
                       "r"(kChunk * 2), "r"(s32(&full[st])) : "memory");
        } else {
          const int x = (int)(s & 7), y = (int)((s >> 3) & 31), z = (int)(s >> 8);
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst), "l"(&tmap), "r"(s32(&full[st])), "r"(x), "r"(y), "r"(z) : "memory");
        }
        if (++st == kStages) { st = 0; ph ^= 1u; }
      }
    }
    return;
  }
  int st = 0; uint32_t ph = 0;
  unsigned long long mism = 0;
  const int ctid = tid - 32, nct = kThreads - 32;
  for (int64_t c = 0; c < chunks_per_cta; ++c) {
    mbar_wait(&full[st], ph);
    const uint16_t* sm = reinterpret_cast<const uint16_t*>(smem + st * kChunk * 2);
    if (verify) {
      const int64_t s = base + c * kChunk;
      for (int i = ctid; i < kChunk; i += nct) mism += (sm[i] != pattern((uint64_t)(s + i)));
    } else if (sm[ctid] == 0x1234 && sm[ctid + 1000] == 0x4321) mism++;
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
    if (++st == kStages) { st = 0; ph ^= 1u; }
  }
  if (mism) atomicAdd(bad, mism);
}

__global__ void fill(uint16_t* buf, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] = pattern((uint64_t)i);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int64_t n = (int64_t)1 << 31;            // 2 Gi samples = 4 GB (>> L2)
  uint16_t* buf;
  CK(cudaMalloc(&buf, n * 2 + 4096));
  fill<<<1024, 256>>>(buf, n + 2048);
  unsigned long long* bad;
  CK(cudaMalloc(&bad, 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("cuTensorMapEncodeTiled not found\n"); return 1; }
  CUtensorMap tmap;
  const cuuint64_t dims[3] = {263, 32, (cuuint64_t)(n / 256)};
  const cuuint64_t strides[2] = {16, 512};
  const cuuint32_t box[3] = {256, 1, 8};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode overlapped (263,32,nz) strides (16,512) box (256,1,8): CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 2;
  const size_t smem = kStages * kChunk * 2 + 2 * kStages * 8 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = 148 * 2;
  const int64_t chunks = 3000;                   // per CTA: 12 MB
  const int64_t stride = n / grid / 8 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 8, 4, 1, 3, 7}) {
      if (mode == 0 && (shift & 7)) continue;
      CK(cudaMemset(bad, 0, 8));
      probe<<<grid, kThreads, smem>>>(tmap, buf, 64, stride, shift, mode, 1, bad);
      CK(cudaDeviceSynchronize());
      unsigned long long hb = 0;
      CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<<<grid, kThreads, smem>>>(tmap, buf, chunks, stride, shift, mode, 0, bad);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("mode %s shift %d samples: mismatches %llu, %.3f ms, %.0f GB/s\n", mode ? "tensor-map" : "bulk      ", shift, hb, best,
             (double)grid * chunks * kChunk * 2 / best / 1e6);
    }
  return 0;
}
