// Micro-benchmark: sustained tcgen05.mma rate of ONE CTA per SM for the instruction shapes the conv kernels issue
// (M = 128, K = 16 per instruction, bf16 -> fp32), operands resident in shared memory, no TMA and no epilogue:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
// Varies N (32 .. 256), the K-major row width of the operands (64-byte rows / SWIZZLE_64B as the kernels use, or
// 128-byte rows / SWIZZLE_128B) and whether consecutive instructions walk through different A tiles (as the
// row-shifted halo views do) or re-read one tile. Prints cycles per instruction and the fraction of the
// 8192 FLOP/clk/SM dense bf16 rate (M128 x N x K16 takes N/8 ... cycles at peak: 2*128*N*16 / 8192 = N/2).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../unet-bssfp_b200/csrc/sm100_ptx.cuh"

using namespace ub;

struct Cfg {
  int n;          // UMMA N
  int row_bytes;  // 64 or 128: bytes of one K-major operand row (swizzle span)
  int walk;       // 1: A start address moves by one row per instruction (shifted views), 0: fixed
  int iters;      // instructions per CTA
};

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // A: (128 + 64 spare rows) x row_bytes, B: 256 rows x row_bytes; zero-filled (values do not matter for timing)
  const int a_bytes = 192 * c.row_bytes, b_bytes = 256 * c.row_bytes;
  for (int i = threadIdx.x * 16; i < a_bytes + b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tmem_slot));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t swz = c.row_bytes == 128 ? SWZ_128B : SWZ_64B;
    const uint32_t sbo = 8 * c.row_bytes;                      // 8-row groups are contiguous
    const uint64_t a0 = make_smem_desc(smem_u32(smem), 16, sbo, swz);
    const uint64_t b0 = make_smem_desc(smem_u32(smem + a_bytes), 16, sbo, swz);
    const uint32_t idesc = make_idesc_bf16(128, c.n, 0, 0);
    const uint32_t a_lo0 = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo0 = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    const uint32_t row16 = (uint32_t)c.row_bytes >> 4;          // one row in descriptor address units
    const uint32_t kmask = (uint32_t)(c.row_bytes / 32) - 1;    // K steps of 32 bytes inside one row
    const uint32_t wstep = c.walk ? row16 : 0u;
    const unsigned long long t0 = clock64();
    uint32_t col = 0;
    for (int plane = 0; plane < c.iters / 18; ++plane) {       // 18 instructions per accumulator, as one marching plane
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        // tap j / 2 is a row-shifted view of the A tile (walk = 1), K step j % ksteps inside the swizzle span
        const uint32_t a_lo = a_lo0 + (uint32_t)(j >> 1) * wstep + (((uint32_t)j & kmask) << 1);
        const uint32_t b_lo = b_lo0 + (((uint32_t)j & kmask) << 1);
        umma_bf16_lohi(tmem + col, a_lo, a_hi, b_lo, b_hi, idesc, j != 0);
      }
      col ^= 256u;
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const unsigned long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 36000;
  printf("%-6s %-9s %-5s %12s %10s\n", "N", "row_bytes", "walk", "clk/instr", "of peak");
  for (int rb : {64, 128})
    for (int walk : {0, 1})
      for (int n : {32, 48, 64, 96, 128, 192, 256}) {
        Cfg c{n, rb, walk, iters};
        const int smem = (192 + 256) * rb + 2048;
        rate_kernel<<<sms, 128, smem>>>(c, d_cycles);   // warm-up
        rate_kernel<<<sms, 128, smem>>>(c, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d rb=%d: %s\n", n, rb, cudaGetErrorString(e)); return 1; }
        unsigned long long h[256];
        cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        const double per = avg / iters;
        printf("%-6d %-9d %-5d %12.2f %9.1f%%\n", n, rb, walk, per, 100.0 * (n / 2.0) / per);
      }
  return 0;
}
