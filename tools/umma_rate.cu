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
#include "../unet_bssfp_b200/csrc/sm100_ptx.cuh"

using namespace ub;

struct Cfg {
  int n;          // UMMA N
  int row_bytes;  // 64 or 128: bytes of one K-major operand row (swizzle span)
  int walk;       // 1: A start address moves by one row per instruction (shifted views), 0: fixed
  int iters;      // instructions per CTA
  int ld_warps;   // warps 1..ld_warps read a TMEM accumulator (tcgen05.ld 32x32b.x16 x 3) in a loop while the MMAs run
  int sbo_rows;   // rows between consecutive 8-row groups of A (8 = dense tile, 10 = the 10-voxel-wide halo plane)
  int commits;    // tcgen05.commit (to a scratch mbarrier nobody waits on) issued per 18 instructions: 0, 1, 2 or 4
  int fences;     // tcgen05.fence::after_thread_sync issued per 18 instructions (the MMA thread's two barrier waits)
  int st_warps;   // warps 5..4+st_warps stream 16-byte global stores (an epilogue's output traffic through L1TEX)
};

__global__ void __launch_bounds__(288) rate_kernel(Cfg c, unsigned long long* cycles, uint4* sink) {
  __shared__ volatile int done;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long scratch_bar;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // A: (128 + 64 spare rows) x row_bytes, B: 256 rows x row_bytes; zero-filled (values do not matter for timing)
  const int a_bytes = 256 * c.row_bytes, b_bytes = 256 * c.row_bytes;
  for (int i = threadIdx.x * 16; i < a_bytes + b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&scratch_bar), 0x7FFFF);
    fence_mbar_init();
    done = 0;
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
    const uint64_t a0 = make_smem_desc(smem_u32(smem), 16, (uint32_t)(c.sbo_rows * c.row_bytes), swz);
    const uint64_t b0 = make_smem_desc(smem_u32(smem + a_bytes), 16, sbo, swz);
    const uint32_t idesc = make_idesc_bf16(128, c.n, 0, 0);
    const uint32_t a_lo0 = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo0 = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    const uint32_t row16 = (uint32_t)c.row_bytes >> 4;          // one row in descriptor address units
    const uint32_t kmask = (uint32_t)(c.row_bytes / 32) - 1;    // K steps of 32 bytes inside one row
    const uint32_t wstep = c.walk ? row16 : 0u;
    const unsigned long long t0 = clock64();
    uint32_t col = 0;
    for (int plane = 0; plane < c.iters / 18; ++plane) {       // 18 instructions per accumulator, as one marching plane
      for (int f = 0; f < c.fences; ++f) tc_fence_after();
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        // tap j / 2 is a row-shifted view of the A tile (walk = 1), K step j % ksteps inside the swizzle span
        // halo-plane walk: tap (kh, kw) starts (kh * sbo_rows + kw) rows into the plane
        const uint32_t tap = (uint32_t)(j >> 1);
        const uint32_t a_lo = a_lo0 + ((tap / 3) * (uint32_t)c.sbo_rows + tap % 3) * wstep + (((uint32_t)j & kmask) << 1);
        const uint32_t b_lo = b_lo0 + (((uint32_t)j & kmask) << 1);
        umma_bf16_lohi(tmem + col, a_lo, a_hi, b_lo, b_hi, idesc, j != 0);
        if (c.commits == 4 && (j == 4 || j == 8 || j == 13)) umma_commit(smem_u32(&scratch_bar));
        if (c.commits == 2 && j == 8) umma_commit(smem_u32(&scratch_bar));
      }
      if (c.commits >= 1) umma_commit(smem_u32(&scratch_bar));
      col ^= 256u;
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const unsigned long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + 32 * c.ld_warps) {
    // epilogue-like TMEM readers: lane quarter = warp % 4, the accumulator that is NOT being written right now does
    // not matter for timing -- read columns [320, 416) of the second accumulator region
    const int w = (threadIdx.x >> 5) & 3;
    const uint32_t ta = tmem + ((uint32_t)(w * 32) << 16) + 320u;
    uint32_t acc = 0;
    while (!done) {
      uint32_t v0[16], v1[16], v2[16];
      tmem_ld_32x32b_x16(ta, v0);
      tmem_ld_32x32b_x16(ta + 32, v1);
      tmem_ld_32x32b_x16(ta + 64, v2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += v0[j] ^ v1[j] ^ v2[j];
    }
    if (acc == 0x12345678u) sink[0] = make_uint4(acc, 0, 0, 0);
  } else if (threadIdx.x >= 160 && threadIdx.x < 160 + 32 * c.st_warps) {
    uint4* dst = sink + 64 + ((size_t)blockIdx.x * 128 + (threadIdx.x - 160)) * 64;
    uint32_t k = 0;
    while (!done) {
      __stcs(dst + (k & 63), make_uint4(k, k, k, k));
      ++k;
    }
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
  uint4* d_sink;
  cudaMalloc(&d_sink, sizeof(uint4) * (64 + (size_t)sms * 128 * 64));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 36000;
  printf("%-6s %-9s %-5s %-8s %-8s %-8s %12s %10s\n", "N", "row_bytes", "walk", "ld_warps", "st_warps", "sbo_rows", "clk/instr", "of peak (last column block: commits per 18)");
  struct Variant { int rb, walk, ldw, stw, sbo, commits, fences; };
  const Variant variants[] = {{64, 0, 0, 0, 8, 0, 0}, {128, 1, 0, 0, 8, 0, 0}, {64, 1, 4, 4, 10, 0, 0}, {64, 1, 0, 0, 10, 1, 0}, {64, 1, 0, 0, 10, 2, 0}, {64, 1, 0, 0, 10, 4, 0},
                              {64, 1, 0, 0, 10, 0, 2}, {64, 1, 4, 4, 10, 2, 2}};
  for (const Variant& v : variants)
    {
      const int rb = v.rb, walk = v.walk;
      for (int n : {32, 48, 64, 96, 128, 192, 256}) {
        Cfg c{n, rb, walk, iters, v.ldw, v.sbo, v.commits, v.fences, v.stw};
        const int smem = (256 + 256) * rb + 2048;
        rate_kernel<<<sms, 288, smem>>>(c, d_cycles, d_sink);   // warm-up
        rate_kernel<<<sms, 288, smem>>>(c, d_cycles, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d rb=%d: %s\n", n, rb, cudaGetErrorString(e)); return 1; }
        unsigned long long h[256];
        cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        const double per = avg / iters;
        printf("%-6d %-9d %-5d %-8d %-8d %-8d %12.2f %9.1f%%   commits/18 = %d fences/18 = %d\n", n, rb, walk, v.ldw, v.stw, v.sbo, per, 100.0 * (n / 2.0) / per, v.commits, v.fences);
      }
    }
  return 0;
}
