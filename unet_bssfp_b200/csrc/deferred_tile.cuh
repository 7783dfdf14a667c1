// In-shared-memory deferred activation of a TMA-delivered halo plane (the "apply on the consumer's operand path"
// half of the conv -> norm -> dropout -> LeakyReLU fusion; ref: monai Convolution -> ADN("NDA"), ref:model.py:22-28).
//
// The producer block stores only its raw conv output y (fp16) and per-(n,c) constants; a consuming convolution
// (forward operand A of igemm_march_kernel, operand X of wgrad_march_kernel) TMA-loads the y halo plane and a small
// group of "operand transform" threads rewrites it IN PLACE with  a = LeakyReLU(Dropout(y * scale + shift))  between
// the TMA arrival and the first tcgen05.mma that reads it (fence.proxy.async, then an mbarrier hand-over to the MMA
// thread). Rows of the halo outside the volume were zero-filled by TMA and are left untouched: the conv's zero
// padding applies to the ACTIVATIONS, not to y.
//
// Measured (B200, 8 x 128^3, profiles/r02_deferred_microbench.txt): the marching kernels are bound by shared-memory
// bandwidth and by the issue slots around the single MMA thread, and the rewrite competes for both, so on these
// kernels the transform does NOT pay against one materialising pass at HBM speed (32 -> 32 forward: 0.80 ms + 0.32 ms
// norm_act pass vs 1.21 ms fused in the fp16 form, 1.73 ms in the bf16 form). A variant that loads y from global
// memory into registers and writes the stage once was slower still (1.83 ms; commit 'Operand transform variant ...').
// The modules therefore defer activations only in front of memory-bound consumers by default; UB_DEFER_CONV=1 turns
// the conv operand path on.
//
// Tile: 180 rows (18 h x 10 w voxels) of 32 channels = 64-byte rows, 64-byte TMA swizzle, stage 1024-byte aligned:
// the 16-byte chunk at physical position p of row r holds the logical channel octet p ^ ((r >> 1) & 3). Thread t of
// NT (64 or 128) owns the chunks t, t + NT, ... : row (t >> 2) + (NT / 4) k, physical position t & 3, so its logical
// octet (t & 3) ^ ((t >> 3) & 3) is the same for all of them and its affine constants stay in registers.
#pragma once
#include "pointwise.cuh"

namespace ub {

constexpr int kTfRows = 180, kTfRowsW = 10;

// F16: the transformed operand is written as fp16 (deferred_act8_f16; the consumer multiplies it with fp16 weights),
// else as bf16 (deferred_act8, the canonical form).
template <int NT, bool F16 = false>
struct HaloTransform {
  static_assert(NT == 64 || NT == 128, "transform group: 2 or 4 warps");
  static constexpr int NK = (kTfRows * 4 + NT - 1) / NT;   // chunks per thread (12 / 6)
  DeferredOctet K;
  __half2 sc2[4], sh2[4], slope2;   // F16: the same folded constants, rounded to fp16
  uint32_t valid;      // bit k: row (t >> 2) + (NT / 4) k is a halo row inside the volume
  int voff[NK];        // h * W + w of that row (dropout counter)
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
    valid = 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int r = (t >> 2) + (NT / 4) * k;
      const int h = h0 - 1 + r / kTfRowsW, w = w0 - 1 + r % kTfRowsW;
      const bool ok = r < kTfRows && h >= 0 && h < H && w >= 0 && w < W;
      valid |= (ok ? 1u : 0u) << k;
      voff[k] = ok ? h * W + w : 0;
    }
  }
  // stage: generic pointer to the plane in shared memory; plane_vox = voxel index of (n, d, 0, 0) in the y tensor.
  // Straight-line code: every chunk is loaded and evaluated unconditionally (the stage is padded to 192 rows, so
  // the rows past 180 are readable), only the STORE is predicated -- the compiler interleaves the independent
  // chunks instead of serialising one branch region per chunk.
  __device__ __forceinline__ void apply(uint8_t* stage, unsigned long long plane_vox) const {
    const unsigned long long e_base = plane_vox * 32ull + (unsigned long long)(oct * 8);
#pragma unroll
    for (int k0 = 0; k0 < NK; k0 += 6) {
      bf16x8 v[6];
#pragma unroll
      for (int k = k0; k < k0 + 6 && k < NK; ++k)
        v[k - k0] = *reinterpret_cast<const bf16x8*>(stage + byte_off + k * (NT * 16));
#pragma unroll
      for (int k = k0; k < k0 + 6 && k < NK; ++k) {
        const unsigned long long e0 = e_base + (unsigned long long)voff[k] * 32ull;
        const bf16x8 r = F16 ? deferred_act8_f16(v[k - k0], sc2, sh2, slope2, K.slope_le1, K.has_drop, e0, K.seed, K.thresh)
                             : K.apply(v[k - k0], e0);
        if ((valid >> k) & 1u) *reinterpret_cast<bf16x8*>(stage + byte_off + k * (NT * 16)) = r;
      }
    }
  }
};

}  // namespace ub
