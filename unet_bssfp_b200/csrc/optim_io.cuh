// Kernels either side of the network path:
//   N3  multi-tensor AdamW                       ref:src/model.py:144,164,359-361 (torch.optim.AdamW, lr 1e-3)
//   N4  de-normalisation + NIfTI storage order   ref:src/eval.py:39-47, ref:src/model.py:335-357
// Both are HBM-bound: AdamW moves 28 B / parameter (read p, g, m, v; write p, m, v), the volume writer
// 8 B / value.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ub {

// ------------------------------------------------------------------------------------------------
// N3: AdamW over up to kAdamWMaxTensors parameter tensors per launch. The pointer table travels as a
// kernel argument (no device-side table to keep alive, CUDA-graph friendly); every block owns one
// kAdamWChunk-element chunk of one tensor and finds it by scanning the block-offset prefix.
// Arithmetic in fp32 in the order of torch's _single_tensor_adamw:
//   p *= 1 - lr * wd;  m = lerp(m, g, 1 - b1);  v = b2 * v + (1 - b2) g^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// ------------------------------------------------------------------------------------------------
constexpr int kAdamWMaxTensors = 48;
constexpr int kAdamWChunk = 4096;          // elements per block: 256 threads x 4 float4

struct AdamWBatch {
  float* p[kAdamWMaxTensors];
  const float* g[kAdamWMaxTensors];
  float* m[kAdamWMaxTensors];
  float* v[kAdamWMaxTensors];
  int numel[kAdamWMaxTensors];
  int block_begin[kAdamWMaxTensors + 1];   // prefix of chunks per tensor
  int count;
  // all derived on the host in double and rounded once (what torch does with its Python-float hyper-parameters)
  float beta2, one_minus_beta1, one_minus_beta2, eps, decay_mul /* 1 - lr * wd */, step_size /* lr / bc1 */, bc2_sqrt,
      grad_scale;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWBatch& B) {
  g *= B.grad_scale;
  p *= B.decay_mul;
  m = m + B.one_minus_beta1 * (g - m);                     // lerp_(g, 1 - b1), weight < 0.5 branch
  v = B.beta2 * v;                                         // mul_(b2), rounded, then addcmul_(g, g, 1 - b2)
  v = v + B.one_minus_beta2 * g * g;
  const float denom = sqrtf(v) / B.bc2_sqrt + B.eps;
  p -= B.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ AdamWBatch B) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int t = 0;
    while (t + 1 < B.count && (int)blockIdx.x >= B.block_begin[t + 1]) ++t;
    s_t = t;
  }
  __syncthreads();
  const int t = s_t;
  const int begin = ((int)blockIdx.x - B.block_begin[t]) * kAdamWChunk;
  const int n = B.numel[t];
  int end = begin + kAdamWChunk;
  if (end > n) end = n;
  float* __restrict__ p = B.p[t];
  const float* __restrict__ g = B.g[t];
  float* __restrict__ m = B.m[t];
  float* __restrict__ v = B.v[t];
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int i = begin + threadIdx.x * 4;
  if (aligned) {
    for (; i + 3 < end; i += 256 * 4) {
      float4 P = *reinterpret_cast<float4*>(p + i);
      const float4 G = __ldcs(reinterpret_cast<const float4*>(g + i));   // gradients are dead after this pass
      float4 M = *reinterpret_cast<float4*>(m + i);
      float4 V = *reinterpret_cast<float4*>(v + i);
      adamw_one(P.x, G.x, M.x, V.x, B);
      adamw_one(P.y, G.y, M.y, V.y, B);
      adamw_one(P.z, G.z, M.z, V.z, B);
      adamw_one(P.w, G.w, M.w, V.w, B);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(m + i) = M;
      *reinterpret_cast<float4*>(v + i) = V;
    }
    // at most one partial vector per chunk (the tensor's tail): its owner finishes it element-wise
    if (i < end) {
      for (int k = i; k < end; ++k) {
        float P = p[k], M = m[k], V = v[k];
        adamw_one(P, g[k], M, V, B);
        p[k] = P; m[k] = M; v[k] = V;
      }
    }
  } else {
    for (int k = begin + threadIdx.x; k < end; k += 256) {
      float P = p[k], M = m[k], V = v[k];
      adamw_one(P, g[k], M, V, B);
      p[k] = P; m[k] = M; v[k] = V;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// N4: (C, X, Y, Z) fp32, Z fastest (the module layout of one volume) -> NIfTI storage order of the
// channel-last array (X, Y, Z, C) the reference saves: X fastest, then Y, Z, C -- i.e. [C][Z][Y][X] --
// with the min-max de-normalisation v * scale + offset applied in fp64 like the reference
// (get_fdata() is float64) and rounded once to the float32 the image header declares.
// One 32 x 32 (x, z) tile per block, transposed through shared memory: both sides coalesced.
// grid (ceil(Z/32), ceil(X/32), C*Y), block (32, 8)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) denorm_to_nifti_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int X, int Y, int Z, double scale, double offset) {
  __shared__ float tile[32][33];
  const int c = blockIdx.z / Y, y = blockIdx.z - c * Y;
  const int z0 = blockIdx.x * 32, x0 = blockIdx.y * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int x = x0 + threadIdx.y + 8 * r, z = z0 + threadIdx.x;
    if (x < X && z < Z) tile[threadIdx.y + 8 * r][threadIdx.x] = __ldcs(src + (((size_t)c * X + x) * Y + y) * Z + z);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int z = z0 + threadIdx.y + 8 * r, x = x0 + threadIdx.x;
    if (x < X && z < Z) {
      const double v = (double)tile[threadIdx.x][threadIdx.y + 8 * r] * scale + offset;
      __stcs(dst + (((size_t)c * Z + z) * Y + y) * X + x, (float)v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm running-statistics update from saved batch statistics (ub_norm_finalize run without running
// buffers): running = (1 - momentum) * running + momentum * {mean, unbiased variance}. Lets two forward passes
// of one network run on different streams and still apply their updates in the reference's order.
// ------------------------------------------------------------------------------------------------
__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, int c, double count,
                                         float eps, float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double r = (double)rstd[i];
  double var = 1.0 / (r * r) - (double)eps;            // biased batch variance
  if (var < 0.0) var = 0.0;
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  running_mean[i] = (float)((1.0 - momentum) * running_mean[i] + momentum * (double)mean[i]);
  running_var[i] = (float)((1.0 - momentum) * running_var[i] + momentum * unbiased);
}

}  // namespace ub
