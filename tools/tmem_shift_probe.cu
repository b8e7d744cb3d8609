// tmem_shift_probe.cu -- what does tcgen05.shift.cta_group::1.down do on sm_100a, and what does it cost?
// The folded 3x3 convolutions (umma_conv.cuh) need out[q] = D0[q-1] + D1[q] + D2[q+1] over TMEM rows; today that shifted sum is 32 warp
// shuffles + a shared-memory exchange + a named barrier per tile in the epilogue.  If the tensor core can move accumulator rows itself the
// epilogue becomes three TMEM loads and adds.
// Part 1 (semantics): every lane L writes value L*1000 + col to columns 0..31 (tcgen05.st), one thread issues ONE shift at
//   (lane = lane0, column = col0), commits, waits; every lane reads the columns back.  Printed: which (lane, column) cells changed and how.
// Part 2 (cost): one thread issues `iters` shifts (alone, or interleaved with M=128 N=48 K=16 MMAs), commit, wait; clk per shift.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_shift_probe tools/tmem_shift_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait0(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();
  }
}

__global__ void __launch_bounds__(128, 1) probe(int lane0, int col0, int n_shifts, int mode, int iters, uint32_t* out, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const uint32_t my = tmem + ((uint32_t)(warp * 32) << 16);
  // fill columns 0..31 of every lane
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = (uint32_t)(threadIdx.x * 1000 + c0 + i);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(my + c0), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
                 "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(48 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFF) | ((uint32_t)(2048 >> 4) << 16);
    const uint32_t b_lo = (((smem_u32(smem) + 8192) >> 4) & 0x3FFF) | ((uint32_t)(48 * 16 >> 4) << 16);
    const uint32_t taddr = tmem + ((uint32_t)lane0 << 16) + (uint32_t)col0;
    const long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < n_shifts; ++i) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
    } else if (mode == 1) {        // shifts only, cycling over 4 column blocks of 8
      for (int i = 0; i < iters; ++i) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem + 64 + (uint32_t)((i & 3) * 8)) : "memory");
    } else if (mode == 2) {        // MMAs only (N = 48 into columns 64..111)
      for (int i = 0; i < iters; ++i)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem + 64), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(i) : "memory");
    } else {                       // the folded conv's pattern: 6 MMAs (N = 48) then 6 shifts (D0 twice x 2 blocks, D1 once x 2 blocks)
      for (int i = 0; i < iters; i += 6) {
        for (int t = 0; t < 6; ++t)
          asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem + 64), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(t) : "memory");
        for (int t = 0; t < 6; ++t) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem + 64 + (uint32_t)((t % 4) * 8)) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    mbar_wait0(smem_u32(&bar), 0);
    const long long t2 = clock64();
    clk[0] = t2 - t0; clk[1] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(my + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 32 + c0 + i] = r[i];
  }
  (void)lane;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static int run(int lane0, int col0, int n_shifts, int mode, int iters, uint32_t* d_out, long long* d_clk, std::vector<uint32_t>& h, long long* hc) {
  probe<<<1, 128, 40 * 1024>>>(lane0, col0, n_shifts, mode, iters, d_out, d_clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, d_clk, 16, cudaMemcpyDeviceToHost);
  return 0;
}

int main() {
  uint32_t* d_out; long long* d_clk;
  cudaMalloc(&d_out, 128 * 32 * 4); cudaMalloc(&d_clk, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<uint32_t> h(128 * 32);
  long long hc[2];
  const int cases[][3] = {{0, 0, 1}, {0, 8, 1}, {32, 0, 1}, {0, 4, 1}, {0, 0, 2}, {64, 16, 3}};
  for (auto& c : cases) {
    if (run(c[0], c[1], c[2], 0, 0, d_out, d_clk, h, hc)) return 1;
    printf("== %d shift(s) at lane %d, column %d: cells that changed (lane: column -> source lane), total %lld clk\n", c[2], c[0], c[1], hc[0]);
    int changed_cols[32] = {0}, first_l = -1, last_l = -1;
    for (int l = 0; l < 128; ++l)
      for (int col = 0; col < 32; ++col) {
        const uint32_t v = h[l * 32 + col];
        if (v != (uint32_t)(l * 1000 + col)) { changed_cols[col]++; if (first_l < 0) first_l = l; last_l = l; }
      }
    printf("   columns changed:");
    for (int col = 0; col < 32; ++col) if (changed_cols[col]) printf(" %d(x%d)", col, changed_cols[col]);
    printf("\n   lanes %d..%d; samples:", first_l, last_l);
    const int cc = c[1];
    for (int l : {0, 1, 2, 3, 30, 31, 32, 33, 34, 63, 64, 65, 66, 67, 95, 96, 97, 127})
      printf(" L%d<-%d.%d", l, h[l * 32 + cc] / 1000, h[l * 32 + cc] % 1000);
    printf("\n");
  }
  const int iters = 3072;
  for (int mode : {1, 2, 3}) {
    if (run(0, 0, 0, mode, iters, d_out, d_clk, h, hc)) return 1;
    printf("mode %d (%s): %d ops, %.1f clk/op to completion, %.1f clk/op to issue\n", mode,
           mode == 1 ? "shifts only" : mode == 2 ? "MMAs N=48 only" : "6 MMAs + 6 shifts interleaved (per op)", iters * (mode == 3 ? 2 : 1),
           (double)hc[0] / (iters * (mode == 3 ? 2 : 1)), (double)hc[1] / (iters * (mode == 3 ? 2 : 1)));
  }
  return 0;
}
