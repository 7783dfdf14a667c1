// In-kernel deferred activation of a halo plane (the "apply on the consumer's operand path" half of the
// conv -> norm -> dropout -> LeakyReLU fusion; ref: monai Convolution -> ADN("NDA"), ref:model.py:22-28).
//
// The producer block stores only its raw conv output y (fp16) and per-(n,c) constants; a consuming convolution
// (forward operand A of igemm_march_kernel, operand X of wgrad_march_kernel) gets the plane of chunk 0 not from TMA
// but from a small group of "operand transform" threads: they read the y halo plane from global memory (coalesced
// 16-byte loads, issued one plane ahead), evaluate  a = LeakyReLU(Dropout(y * scale + shift))  in registers and
// store the result into the stage in exactly the layout TMA would have produced (64-byte swizzle), then
// fence.proxy.async and hand the stage to the MMA thread through an mbarrier. Rows of the halo outside the volume
// are written as zeros: the conv's zero padding applies to the ACTIVATIONS, not to y.
//
// Why not TMA + an in-place rewrite in shared memory (the first version of this file): the marching kernels are
// bound by shared-memory bandwidth (the tensor core reads 7 KB of operands per 48-cycle MMA), so the LDS half of a
// read-modify-write queues behind the operand fetches -- measured 2 600 - 3 700 cycles per plane against the
// ~1 650 the MMAs need (profiles/r02_deferred_microbench.txt). Writing the stage once costs what TMA's own write
// would have cost.
//
// Tile: 180 rows (18 h x 10 w voxels) of 32 channels = 64-byte rows, 64-byte TMA swizzle, stage 1024-byte aligned:
// the 16-byte chunk at physical position p of row r holds the logical channel octet p ^ ((r >> 1) & 3). Thread t of
// NT (64 or 128) owns the chunks t, t + NT, ... : row (t >> 2) + (NT / 4) k, physical position t & 3, so its logical
// octet (t & 3) ^ ((t >> 3) & 3) is the same for all of them and its affine constants stay in registers.
#pragma once
#include "pointwise.cuh"

namespace ub {

constexpr int kTfRows = 180, kTfRowsW = 10;

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// F16: the operand is written as fp16 (deferred_act8_f16; the consumer multiplies it with fp16 weights), else as bf16
// (deferred_act8, the canonical form).
template <int NT, bool F16 = false>
struct HaloTransform {
  static_assert(NT == 64 || NT == 128, "transform group: 2 or 4 warps");
  static constexpr int NK = (kTfRows * 4 + NT - 1) / NT;   // chunks per thread (12 / 6)
  DeferredOctet K;
  __half2 sc2[4], sh2[4], slope2;   // F16: the same folded constants, rounded to fp16
  uint32_t inside;     // bit k: row (t >> 2) + (NT / 4) k is a halo row inside the volume (activation is evaluated)
  uint32_t exists;     // bit k: the row is one of the tile's 180 (zeros are written when it is not inside)
  int goff[NK];        // byte offset of the chunk inside a y plane: (h * W + w) * 64 + octet * 16
  int oct;             // logical channel octet of this thread
  uint32_t byte_off;   // (t >> 2) * 64 + (t & 3) * 16

  // t in [0, NT); (h0, w0) = first OUTPUT voxel of the tile (the halo starts one voxel before)
  __device__ __forceinline__ void setup(int t, const NormActArgs& A, int n, int h0, int w0, int H, int W) {
    oct = (t & 3) ^ ((t >> 3) & 3);
    K.load(A, n, 32, oct * 8);
    if (F16) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sc2[k] = __floats2half2_rn(K.sc[2 * k], K.sc[2 * k + 1]);
        sh2[k] = __floats2half2_rn(K.sh[2 * k], K.sh[2 * k + 1]);
      }
      slope2 = __float2half2_rn(K.slope);
    }
    byte_off = (uint32_t)(t >> 2) * 64u + (uint32_t)(t & 3) * 16u;
    inside = exists = 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int r = (t >> 2) + (NT / 4) * k;
      const int h = h0 - 1 + r / kTfRowsW, w = w0 - 1 + r % kTfRowsW;
      const bool ex = r < kTfRows;
      const bool ok = ex && h >= 0 && h < H && w >= 0 && w < W;
      exists |= (ex ? 1u : 0u) << k;
      inside |= (ok ? 1u : 0u) << k;
      goff[k] = ok ? (h * W + w) * 64 + oct * 16 : 0;
    }
  }
  // issue the loads of one plane; plane = global address of y[n][d][0][0][0]
  __device__ __forceinline__ void load(uint4 (&buf)[NK], const uint8_t* plane) const {
#pragma unroll
    for (int k = 0; k < NK; ++k)
      buf[k] = ((inside >> k) & 1u) ? ld_nc_v4(plane + goff[k]) : make_uint4(0u, 0u, 0u, 0u);
  }
  // evaluate and store into the stage (generic pointer to shared memory); plane_elem0 = element index of
  // y[n][d][0][0][0] (dropout counter)
  __device__ __forceinline__ void store(const uint4 (&buf)[NK], uint8_t* stage, unsigned long long plane_elem0) const {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const unsigned long long e0 = plane_elem0 + (unsigned long long)(goff[k] >> 1);
      const bf16x8 in = *reinterpret_cast<const bf16x8*>(&buf[k]);
      bf16x8 r = F16 ? deferred_act8_f16(in, sc2, sh2, slope2, K.slope_le1, K.has_drop, e0, K.seed, K.thresh)
                     : K.apply(in, e0);
      if (!((inside >> k) & 1u)) *reinterpret_cast<uint4*>(&r) = make_uint4(0u, 0u, 0u, 0u);
      if ((exists >> k) & 1u) *reinterpret_cast<bf16x8*>(stage + byte_off + k * (NT * 16)) = r;
    }
  }
};

}  // namespace ub
