// tma_microbench.cu -- TMA (cp.async.bulk.tensor) load throughput per SM for the box shapes the conv kernel uses.
// One CTA per SM, one thread issues `iters` box loads into a ring of 4 smem buffers (mbarrier per buffer), nothing else.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int rank; int c_box, x_box, y_box; int n_img; int W, H; int split; int nbuf; };

__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap tm, Cfg c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  const uint32_t box_bytes = (uint32_t)c.c_box * 2u * c.x_box * c.y_box;
  const uint32_t buf_bytes = (box_bytes * c.split + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int b = i % c.nbuf;
      if (i >= c.nbuf) {   // wait for the previous use of this buffer
        uint32_t ok = 0, ph = ((i / c.nbuf) - 1) & 1;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(ph) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(box_bytes * c.split) : "memory");
      const int item = blockIdx.x + i * gridDim.x;
      for (int s = 0; s < c.split; ++s) {
        const uint32_t dst = smem_u32(smem + (size_t)b * buf_bytes + (size_t)s * box_bytes);
        if (c.rank == 4) {
          const int n = item % c.n_img, y = ((item / c.n_img) * c.split + s) * (c.y_box - 2) % c.H - 1;
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(smem_u32(&bar[b])), "r"(0), "r"(-1), "r"(y), "r"(n) : "memory");
        } else {
          const long long row = ((long long)item * c.split + s) * c.x_box % ((long long)c.n_img * c.W * c.H - c.x_box);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(smem_u32(&bar[b])), "r"(0), "r"((int)row) : "memory");
        }
      }
    }
    for (int i = iters - c.nbuf; i < iters; ++i) {
      uint32_t ok = 0, ph = (i / c.nbuf) & 1;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[i % c.nbuf])), "r"(ph) : "memory");
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fn;
  const int B = 512, H = 64, W = 192;
  long long* d_out; cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%-44s %10s %10s\n", "config", "B/clk/SM", "GB/s(148)");
  struct T { const char* name; int pitch_ch, c_box, rank, x_box, y_box, split; } tests[] = {
      {"4D halo box 64B rows  [10][98][32ch] pitch32", 32, 32, 4, 98, 10, 1},
      {"4D halo box 64B rows  [10][98][32ch] pitch288", 288, 32, 4, 98, 10, 1},
      {"4D halo box 128B rows [10][98][64ch] pitch288", 288, 64, 4, 98, 10, 1},
      {"4D halo box 32B rows  [10][98][16ch] pitch288", 288, 16, 4, 98, 10, 1},
      {"4D halo box 32B rows  [10][98][16ch] pitch16 (dense)", 16, 16, 4, 98, 10, 1},
      {"4D halo box 32B rows  [10][98][16ch] pitch32", 32, 16, 4, 98, 10, 1},
      {"4D halo box 32B rows  [10][98][16ch] pitch48", 48, 16, 4, 98, 10, 1},
      {"4D halo box 64B rows  [10][98][32ch] pitch48", 48, 32, 4, 98, 10, 1},
      {"4D halo box 128B rows [10][98][64ch] pitch64 (dense)", 64, 64, 4, 98, 10, 1},
      {"4D halo box 64B rows  [10][98][32ch] pitch64", 64, 32, 4, 98, 10, 1},
      {"4D halo 64B rows split in 5 boxes of 2+2 rows", 32, 32, 4, 98, 4, 5},
      {"2D box 64B rows  [128px][32ch] pitch32 (contig)", 32, 32, 2, 128, 1, 1},
      {"2D box 64B rows  [128px][32ch] pitch32  x4", 32, 32, 2, 128, 1, 4},
      {"2D box 128B rows [128px][64ch] pitch64 (contig)", 64, 64, 2, 128, 1, 1},
      {"2D box 128B rows [256px][64ch] pitch64 (contig)", 64, 64, 2, 256, 1, 1},
      {"2D box 128B rows [128px][64ch] pitch288", 288, 64, 2, 128, 1, 1},
      {"2D box 64B rows  [128px][32ch] pitch288 x4", 288, 32, 2, 128, 1, 4},
  };
  for (auto& t : tests) {
    void* buf; const size_t bytes = (size_t)B * H * W * t.pitch_ch * 2;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    CUtensorMap tm;
    const CUtensorMapSwizzle sw = t.c_box == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : t.c_box == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r;
    if (t.rank == 4) {
      cuuint64_t gd[4] = {(cuuint64_t)t.pitch_ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t gs[3] = {(cuuint64_t)t.pitch_ch * 2, (cuuint64_t)W * t.pitch_ch * 2, (cuuint64_t)H * W * t.pitch_ch * 2};
      cuuint32_t box[4] = {(cuuint32_t)t.c_box, (cuuint32_t)t.x_box, (cuuint32_t)t.y_box, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gd[2] = {(cuuint64_t)t.pitch_ch, (cuuint64_t)B * H * W};
      cuuint64_t gs[1] = {(cuuint64_t)t.pitch_ch * 2};
      cuuint32_t box[2] = {(cuuint32_t)t.c_box, (cuuint32_t)t.x_box}, es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", t.name, (int)r); continue; }
    const size_t bb = ((size_t)t.c_box * 2 * t.x_box * t.y_box * t.split + 1023) & ~(size_t)1023;
    int nbuf = (int)(180 * 1024 / bb); if (nbuf > 4) nbuf = 4; if (nbuf < 1) { printf("%s: too big\n", t.name); continue; }
    Cfg c{t.rank, t.c_box, t.x_box, t.y_box, B, W, H, t.split, nbuf};
    const int iters = 400 / nbuf * nbuf;
    bench<<<148, 128, 190 * 1024>>>(tm, c, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", t.name, cudaGetErrorString(e)); return 1; }
    long long clk; cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
    const double bytes_sm = (double)iters * t.split * t.c_box * 2.0 * t.x_box * t.y_box;
    printf("%-44s %10.1f %10.0f\n", t.name, bytes_sm / clk, bytes_sm / clk * 1.9 * 148);
    cudaFree(buf);
  }
  return 0;
}
