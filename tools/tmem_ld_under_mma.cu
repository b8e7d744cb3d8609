// tmem_ld_under_mma.cu -- does a tcgen05.ld / tcgen05.st issued by another warp wait behind the tcgen05.mma instructions already queued in the
// tensor pipe?  One thread queues Q back-to-back MMAs (M=128, N=64, K=16 -> ~48 clk each) into TMEM columns [0, 64) and commits; as soon as they are
// issued, warps 4..7 load (and optionally store) 16 OTHER columns [256, 272) -- or the SAME columns -- and stamp clock64 when their wait returns.
// Reported per Q: clocks from "all MMAs issued" to (a) the commit's mbarrier arrival, (b) wait::ld returning, (c) wait::st returning.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_under_mma tools/tmem_ld_under_mma.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) probe(int Q, int same_cols, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile long long t_issued;
  __shared__ volatile int go;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    go = 0;
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
  if (threadIdx.x == 0) {
    const int N = 64;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFF) | ((uint32_t)(2048 >> 4) << 16);
    const uint32_t b_lo = (((smem_u32(smem) + 8192) >> 4) & 0x3FFF) | ((uint32_t)(N * 16 >> 4) << 16);
    for (int i = 0; i < Q; ++i) {
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(i) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t0 = clock64();
    t_issued = t0;
    __threadfence_block();
    go = 1;
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (spins > (1u << 24)) __trap();
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (warp >= 4) {
    while (!go) { }
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (same_cols ? 0u : 256u);
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + 64u), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 128) { out[1] = t1 - t_issued; out[2] = t2 - t_issued; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 32);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("# clocks after the last of Q queued MMAs (N=64, ~48 clk each) was ISSUED until: commit arrives | wait::ld returns | wait::st returns\n");
  printf("%5s %10s | %10s %10s %10s\n", "Q", "columns", "commit", "ld", "ld+st");
  for (int same = 0; same < 2; ++same)
    for (int Q : {0, 4, 16, 32, 64, 128}) {
      long long h[3] = {0, 0, 0};
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<148, 256, 56 * 1024>>>(Q, same, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
      printf("%5d %10s | %10lld %10lld %10lld\n", Q, same ? "same" : "other", h[0], h[1], h[2]);
    }
  return 0;
}
