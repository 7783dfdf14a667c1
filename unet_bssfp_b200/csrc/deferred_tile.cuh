// In-shared-memory deferred activation of a TMA-delivered halo plane (the "apply on the consumer's operand path"
// half of the conv -> norm -> dropout -> LeakyReLU fusion; ref: monai Convolution -> ADN("NDA"), ref:model.py:22-28).
//
// The producer block stores only its raw conv output y (fp16) and per-(n,c) constants; a consuming convolution
// (forward operand A of igemm_march_kernel, operand X of wgrad_march_kernel) TMA-loads the y halo plane and a group
// of 64 or 128 threads rewrites it IN PLACE with  a = bf16(LeakyReLU(Dropout(y * scale + shift)))  between TMA arrival and
// the first tcgen05.mma that reads it (fence.proxy.async, then an mbarrier hand-over to the MMA thread). Rows of the
// halo that lie outside the volume were zero-filled by TMA and are left untouched: the conv's zero padding applies
// to the ACTIVATIONS, not to y.
//
// Tile: 180 rows (18 h x 10 w voxels) of 32 channels = 64-byte rows, 64-byte TMA swizzle, stage 1024-byte aligned:
// the 16-byte chunk at physical position p of row r holds the logical channel octet p ^ ((r >> 1) & 3). Thread t of
// NT (64 or 128) owns the chunks t, t + NT, ... : row (t >> 2) + (NT / 4) k, physical position t & 3, so its logical
// octet (t & 3) ^ ((t >> 3) & 3) is the same for all of them and its 16 affine constants stay in registers.
#pragma once
#include "pointwise.cuh"

namespace ub {

constexpr int kTfRows = 180, kTfRowsW = 10;

template <int NT>
struct HaloTransform {
  static_assert(NT == 64 || NT == 128, "transform group: 2 or 4 warps");
  static constexpr int NK = (kTfRows * 4 + NT - 1) / NT;   // chunks per thread (12 / 6)
  DeferredOctet K;
  uint32_t valid;      // bit k: row (t >> 2) + (NT / 4) k is a halo row inside the volume
  int voff[NK];        // h * W + w of that row (dropout counter)
  int oct;             // logical channel octet of this thread
  uint32_t byte_off;   // (t >> 2) * 64 + (t & 3) * 16

  // t in [0, NT); (h0, w0) = first OUTPUT voxel of the tile (the halo starts one voxel before)
  __device__ __forceinline__ void setup(int t, const NormActArgs& A, int n, int h0, int w0, int H, int W) {
    oct = (t & 3) ^ ((t >> 3) & 3);
    K.load(A, n, 32, oct * 8);
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
  // stage: generic pointer to the plane in shared memory; plane_vox = voxel index of (n, d, 0, 0) in the y tensor
  __device__ __forceinline__ void apply(uint8_t* stage, unsigned long long plane_vox) const {
#pragma unroll
    for (int k0 = 0; k0 < NK; k0 += 6) {
      bf16x8 v[6];
#pragma unroll
      for (int k = k0; k < k0 + 6 && k < NK; ++k)
        if ((valid >> k) & 1u) v[k - k0] = *reinterpret_cast<const bf16x8*>(stage + byte_off + k * (NT * 16));
#pragma unroll
      for (int k = k0; k < k0 + 6 && k < NK; ++k)
        if ((valid >> k) & 1u)
          *reinterpret_cast<bf16x8*>(stage + byte_off + k * (NT * 16)) =
              K.apply(v[k - k0], (plane_vox + (unsigned long long)voff[k]) * 32ull + (unsigned long long)(oct * 8));
    }
  }
};

}  // namespace ub
