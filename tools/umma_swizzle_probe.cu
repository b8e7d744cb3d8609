// umma_swizzle_probe.cu -- does tcgen05.mma accept K-major SWIZZLED smem operands whose start address is shifted by an
// arbitrary number of rows (not a multiple of the 8-row swizzle atom)?  Needed to feed the implicit-GEMM conv from
// TMA-written [pixel][channel] tiles (swizzle 32/64/128B) with per-tap row shifts.
// For each swizzle mode, row shift s and descriptor variant, one MMA chain computes D[128x16] = A[s:s+128, :K] * B^T and
// is compared with the CPU.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_swizzle_probe ...
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int rowbytes; int layout; int shift; int variant; int ksteps; };

// a_rows x K bf16 matrix stored row-major [row][rowbytes] with the hardware swizzle: addr ^= ((addr >> 7) & mask) << 4
__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                Cfg c, int a_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = c.rowbytes / 2;
  const uint32_t mask = (c.rowbytes == 128) ? 7u : (c.rowbytes == 64) ? 3u : 1u;
  uint8_t* a_s = smem;                        // 1024-aligned
  uint8_t* b_s = smem + 32 * 1024;            // 1024-aligned
  const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s);
  for (int i = threadIdx.x; i < a_rows * K; i += 128) {
    const int r = i / K, k = i % K;
    uint32_t addr = a_base + r * c.rowbytes + k * 2;
    addr ^= ((addr >> 7) & mask) << 4;
    *reinterpret_cast<__nv_bfloat16*>(a_s + (addr - a_base)) = A[i];
  }
  for (int i = threadIdx.x; i < 16 * K; i += 128) {
    const int r = i / K, k = i % K;
    uint32_t addr = b_base + r * c.rowbytes + k * 2;
    addr ^= ((addr >> 7) & mask) << 4;
    *reinterpret_cast<__nv_bfloat16*>(b_s + (addr - b_base)) = B[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // bf16 x bf16 -> fp32, N=16
    const uint32_t sbo = 8 * c.rowbytes;
    for (int ks = 0; ks < c.ksteps; ++ks) {
      const uint32_t a_addr = a_base + c.shift * c.rowbytes + ks * 32;
      const uint32_t b_addr = b_base + ks * 32;
      auto desc = [&](uint32_t addr, bool with_base) {
        uint64_t d = 0;
        d |= (uint64_t)((addr >> 4) & 0x3FFF);
        d |= (uint64_t)1 << 16;                                 // LBO (ignored for swizzled K-major)
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        if (with_base) d |= (uint64_t)((addr >> 7) & 7) << 49;  // matrix base offset
        d |= (uint64_t)c.layout << 61;
        return d;
      };
      const uint64_t da = desc(a_addr, c.variant == 1), db = desc(b_addr, false);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(ks)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0, spins = 0;
  while (!ok && ++spins < (1u << 24)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
  const int a_rows = 160;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  struct Mode { const char* name; int rowbytes, layout; } modes[3] = {{"SW32", 32, 6}, {"SW64", 64, 4}, {"SW128", 128, 2}};
  for (auto& m : modes) {
    const int K = m.rowbytes / 2;
    std::vector<__nv_bfloat16> A(a_rows * K), B(16 * K);
    std::vector<float> Af(a_rows * K), Bf(16 * K);
    for (int i = 0; i < a_rows * K; ++i) { Af[i] = (float)((i * 7 + (i / K) * 3) % 13 - 6); A[i] = __float2bfloat16(Af[i]); }
    for (int i = 0; i < 16 * K; ++i) { Bf[i] = (float)((i * 5 + 1) % 7 - 3); B[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant)
      for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 9, 13, 16, 25}) {
        Cfg c{m.rowbytes, m.layout, shift, variant, K / 16};
        cudaMemset(dD, 0xff, 128 * 16 * 4);
        probe<<<1, 128, 40 * 1024>>>(dA, dB, dD, c, a_rows);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s shift %d variant %d: CUDA error %s\n", m.name, shift, variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(128 * 16);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < 16; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += Af[(shift + i) * K + k] * Bf[n * K + k];
            const double err = fabs(ref - D[i * 16 + n]);
            if (err > 1e-3) ++bad;
            if (err > maxerr) maxerr = err;
          }
        printf("%-6s shift %2d base_offset=%s : %s (bad %d, max err %.3g)\n", m.name, shift, variant ? "(addr>>7)&7" : "0", bad ? "FAIL" : "ok", bad, maxerr);
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
