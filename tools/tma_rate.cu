// TMA fetch rate per SM as a function of the box's inner row length (sm_100a micro-benchmark).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_rate tools/tma_rate.cu && tools/tma_rate
// One CTA per SM; ONE thread issues 5-D tiled TMA loads of activation-style boxes (C channels x BW x BH rows of an
// NDHWC bf16 tensor, swizzle = row bytes) into a ring of shared-memory stages and waits for them -- no consumer, so the
// only limits are the TMA unit, the L2 -> SM path and (for the big source) HBM. The question it answers: do the
// launches of the generic conv kernel that sit "near no roof" (space-to-depth stem, transposed convs: 64-byte rows,
// 16-18 B / cycle / SM measured by ncu) run into a per-row limit of the TMA unit? Compare inner rows of 64 and 128 bytes
// at equal box bytes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(m), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

struct Args {
  int rank;                // 5: (c, w, h, d, n) box (C, 8, 16, 1, 1); 3: (c, w, h*d*n) box (C, 8, 16); 2: (c, w*h*d*n) box (C, 128)
  CUtensorMap map;
  int box_bytes, stages, boxes_per_stage, iters;
  int W, H, D, N;          // tensor dims (voxels); tiles walk w, h, d, n
  int bw, bh;
  unsigned long long* cycles;
};

__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ Args A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + A.stages * A.boxes_per_stage * ((A.box_bytes + 1023) & ~1023);
  if (threadIdx.x == 0) {
    for (int i = 0; i < A.stages; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tiles_w = A.W / A.bw, tiles_h = A.H / A.bh;
  const long long total = (long long)tiles_w * tiles_h * A.D * A.N;
  long long t = (long long)blockIdx.x * 977 % total;    // scatter the CTAs over the tensor
  const int stage_bytes = A.boxes_per_stage * ((A.box_bytes + 1023) & ~1023);
  const unsigned long long c0 = clock64();
  int st = 0;
  uint32_t ph = 0;
  for (int it = 0; it < A.iters + A.stages; ++it) {
    if (it >= A.stages) {            // the stage issued `stages` iterations ago must have landed
      mbar_wait(bars + 8 * st, ph);
    }
    if (it < A.iters) {
      mbar_expect_tx(bars + 8 * st, (uint32_t)(A.boxes_per_stage * A.box_bytes));
      for (int b = 0; b < A.boxes_per_stage; ++b) {
        long long q = t;
        const int w = (int)(q % tiles_w); q /= tiles_w;
        const int h = (int)(q % tiles_h); q /= tiles_h;
        const int d = (int)(q % A.D); q /= A.D;
        const int n = (int)q;
        const uint32_t dst = base + st * stage_bytes + b * ((A.box_bytes + 1023) & ~1023);
        if (A.rank == 5) tma_load_5d(dst, &A.map, bars + 8 * st, 0, w * A.bw, h * A.bh, d, n);
        else if (A.rank == 3) tma_load_3d(dst, &A.map, bars + 8 * st, 0, w * A.bw, (n * A.D + d) * A.H + h * A.bh);
        else tma_load_2d(dst, &A.map, bars + 8 * st, 0, (((n * A.D + d) * A.H + h * A.bh) * A.W + w * A.bw * 16) % (A.W * A.H * A.D * A.N - 128));
        t += gridDim.x;
        if (t >= total) t -= total;
      }
    }
    if (++st == A.stages) { st = 0; if (it >= A.stages) ph ^= 1; }
  }
  A.cycles[blockIdx.x] = clock64() - c0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
  printf("TMA fetch rate, one issuing thread per SM, %d SMs; box = C x 8 x 16 voxel rows (row = C * 2 bytes), ring of stages, no consumer\n", sms);
  printf("%-28s %6s %6s %9s %10s %12s %10s\n", "source", "row B", "boxes", "box B", "B/cyc/SM", "TB/s (chip)", "us");
  // sources: 64 MB (L2 resident after the first pass) and 1 GB (HBM)
  for (int big = 0; big < 2; ++big) {
    for (int C : {32, 64}) {
     for (int rank : {5, 3, 2}) {
      for (int boxes_per_stage : {4}) {
        const int W = 128, H = 128, D = big ? 128 : 16, N = big ? (32 * 8 / C) : (32 * 4 / C);   // 1.07 GB / 67 MB
        const size_t bytes = (size_t)W * H * D * N * C * 2;
        void* src;
        CK(cudaMalloc(&src, bytes));
        CK(cudaMemset(src, 1, bytes));
        Args A;
        A.bw = 8; A.bh = 16;
        cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
        cuuint32_t bx[5] = {(cuuint32_t)(C > 64 ? 64 : C), (cuuint32_t)A.bw, (cuuint32_t)A.bh, 1, 1};
        cuuint32_t st[5] = {1, 1, 1, 1, 1};
        const int row_bytes = (C > 64 ? 64 : C) * 2;
        CUtensorMapSwizzle swz = row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
        A.rank = rank;
        if (rank == 3) { gd[2] = (cuuint64_t)H * D * N; }
        if (rank == 2) { gd[1] = (cuuint64_t)W * H * D * N; bx[1] = 128; }
        CUresult r = encode(&A.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, src, gd, gs, bx, st, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
        A.box_bytes = row_bytes * A.bw * A.bh;
        A.boxes_per_stage = boxes_per_stage;
        A.stages = (160 * 1024) / (boxes_per_stage * ((A.box_bytes + 1023) & ~1023));
        if (A.stages > 16) A.stages = 16;
        A.W = W; A.H = H; A.D = D; A.N = N;
        A.cycles = cyc;
        const size_t total_boxes = bytes / ((size_t)(C > 64 ? 2 : 1) * A.box_bytes);   // boxes covering channel block 0 only
        A.iters = (int)(total_boxes / boxes_per_stage / sms) * (big ? 1 : 8);
        const int smem = A.stages * boxes_per_stage * ((A.box_bytes + 1023) & ~1023) + 8 * 16 + 1024;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 2; ++rep) {     // rep 0 warms L2 / the descriptor
          CK(cudaEventRecord(e0));
          tma_rate_kernel<<<sms, 64, smem>>>(A);
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
        }
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long h[256];
        CK(cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        const double per_sm = (double)A.iters * boxes_per_stage * A.box_bytes;
        printf("%-20s rank %d %6d %6d %9d %10.1f %12.2f %10.1f\n", big ? "1 GB tensor (HBM)" : "64 MB tensor (L2)", rank, row_bytes,
               boxes_per_stage, A.box_bytes, per_sm / avg, per_sm * sms / (ms * 1e-3) / 1e12, ms * 1e3);
        CK(cudaFree(src));
      }
     }
    }
  }
  return 0;
}
