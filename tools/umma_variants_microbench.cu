// umma_variants_microbench.cu -- cost of the tcgen05.mma variants the conv kernels could use instead of M=128 / A from shared memory:
//   (a) M = 128, A from shared memory (the kernels' form), N = 16 .. 256  -- reference column, incl. the 9-tap fold's N = 144
//   (b) M = 64,  A from shared memory
//   (c) M = 128, A from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): no 4 KB shared-memory read of the A tile per MMA
// One CTA per SM, one thread issues `iters` back-to-back MMAs over two accumulators, commit, wait; clk per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_variants_microbench tools/umma_variants_microbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) bench(int N, int M, int a_tmem, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  {   // A operand in TMEM: columns 496..511 (two K = 16 tiles of 8 columns), every lane
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = 0x3c003c00u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + 496),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // fp16 in, fp32 accumulate
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFF) | ((uint32_t)(2048 >> 4) << 16);          // no-swizzle planar, LBO 2 KB
    const uint32_t b_lo = (((smem_u32(smem) + 8192) >> 4) & 0x3FFF) | ((uint32_t)(N * 16 >> 4) << 16);
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)(N <= 240 ? N : 0);
    const uint32_t at0 = tmem + 496, at1 = tmem + 504;
    const long long t0 = clock64();
    if (!a_tmem) {
      for (int i = 0; i < iters; i += 2) {
#define MMA_SS(D, ACC)                                                                                                          \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t" \
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(D), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(ACC) : "memory")
        MMA_SS(d0, i); MMA_SS(d1, i);
      }
    } else {
      for (int i = 0; i < iters; i += 2) {
#define MMA_TS(D, A, ACC)                                                                                             \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"            \
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(D), "r"(A), "r"(b_lo), "r"(hi), "r"(idesc), "r"(ACC) : "memory")
        MMA_TS(d0, at0, i); MMA_TS(d1, at1, i);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (spins > (1u << 24)) __trap();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  printf("# clk per tcgen05.mma (kind::f16, K = 16, fp32 accumulate), %d back-to-back MMAs over two accumulators, 148 CTAs\n", iters);
  printf("%5s | %12s %12s %12s | %s\n", "N", "M128 A=smem", "M64 A=smem", "M128 A=tmem", "useful FLOP/clk/SM at M128 A=smem (peak 8192)");
  for (int N : {16, 32, 48, 64, 96, 128, 144, 192, 240}) {
    double r[3] = {0, 0, 0};
    for (int v = 0; v < 3; ++v) {
      const int M = v == 1 ? 64 : 128;
      if (M == 64 && (N % 8)) continue;
      bench<<<148, 128, 56 * 1024>>>(N, M, v == 2, iters, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d variant %d error: %s\n", N, v, cudaGetErrorString(e)); return 1; }
      long long h;
      cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      r[v] = (double)h / iters;
    }
    printf("%5d | %12.1f %12.1f %12.1f | %.0f\n", N, r[0], r[1], r[2], 2.0 * 128 * N * 16 / r[0]);
  }
  return 0;
}
