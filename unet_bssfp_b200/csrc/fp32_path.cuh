// fp32 mode of the hot path (north_star: "agree within ... 1e-5 (fp32 mode)").
//
// The bf16 tensor-core path cannot meet 1e-5 (bf16 has 8 mantissa bits; kind::tf32 has 10), so the fp32
// mode is a second set of kernels on the CUDA cores: direct convolutions with fp32 products and an fp64 running
// sum across input channels, fp64 statistics, tensors in the reference's own NCDHW fp32 layout. It exists for
// verification and for small volumes -- it is not the throughput path and makes no roofline claim.
//   Conv3d / ConvTranspose3d(k2,s2) fwd, dgrad, wgrad      ref:src/model.py:19-28,50,72-83 (+ monai BasicUNet)
//   InstanceNorm3d / BatchNorm3d (+ running stats), Dropout, LeakyReLU, MaxPool3d(2): forward and backward
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ub {
namespace f32 {

struct ConvGeom {
  int n, c0, c1, co;        // batch, channels of source 0 / source 1 (concat [src0, src1]), output channels
  int d, h, w;              // input spatial size
  int od, oh, ow;           // output spatial size
  int k, s, p;              // cubic kernel, stride, padding
};

// ------------------------------------------------------------------------------------------------
// Conv3d forward: thread = one output element (n, co, od, oh, ow), ow fastest.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_fwd_kernel(const float* __restrict__ s0, const float* __restrict__ s1,
                                                       const float* __restrict__ wt, const float* __restrict__ bias,
                                                       float* __restrict__ out, ConvGeom G, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int x = (int)(t % G.ow); t /= G.ow;
  const int y = (int)(t % G.oh); t /= G.oh;
  const int z = (int)(t % G.od); t /= G.od;
  const int co = (int)(t % G.co);
  const int n = (int)(t / G.co);
  const int ci_total = G.c0 + G.c1;
  const int k3 = G.k * G.k * G.k;
  const size_t vol = (size_t)G.d * G.h * G.w;
  double acc = bias ? (double)bias[co] : 0.0;
  const int z0 = z * G.s - G.p, y0 = y * G.s - G.p, x0 = x * G.s - G.p;
  for (int ci = 0; ci < ci_total; ++ci) {
    const float* src = ci < G.c0 ? s0 + ((size_t)n * G.c0 + ci) * vol : s1 + ((size_t)n * G.c1 + (ci - G.c0)) * vol;
    const float* wp = wt + ((size_t)co * ci_total + ci) * k3;
    float a = 0.f;
    for (int kd = 0; kd < G.k; ++kd) {
      const int zz = z0 + kd;
      if (zz < 0 || zz >= G.d) continue;
      for (int kh = 0; kh < G.k; ++kh) {
        const int yy = y0 + kh;
        if (yy < 0 || yy >= G.h) continue;
        for (int kw = 0; kw < G.k; ++kw) {
          const int xx = x0 + kw;
          if (xx < 0 || xx >= G.w) continue;
          a = fmaf(__ldg(src + ((size_t)zz * G.h + yy) * G.w + xx), __ldg(wp + (kd * G.k + kh) * G.k + kw), a);
        }
      }
    }
    acc += (double)a;
  }
  out[e] = (float)acc;
}

// Conv3d input gradient: thread = one input element (n, ci, z, y, x); writes into dsrc0 / dsrc1.
__global__ void __launch_bounds__(256) conv_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ wt,
                                                         float* __restrict__ d0, float* __restrict__ d1, ConvGeom G,
                                                         long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int x = (int)(t % G.w); t /= G.w;
  const int y = (int)(t % G.h); t /= G.h;
  const int z = (int)(t % G.d); t /= G.d;
  const int ci_total = G.c0 + G.c1;
  const int ci = (int)(t % ci_total);
  const int n = (int)(t / ci_total);
  const int k3 = G.k * G.k * G.k;
  const size_t ovol = (size_t)G.od * G.oh * G.ow;
  double acc = 0.0;
  for (int co = 0; co < G.co; ++co) {
    const float* dp = dout + ((size_t)n * G.co + co) * ovol;
    const float* wp = wt + ((size_t)co * ci_total + ci) * k3;
    float a = 0.f;
    for (int kd = 0; kd < G.k; ++kd) {
      const int zn = z + G.p - kd;
      if (zn < 0 || zn % G.s) continue;
      const int oz = zn / G.s;
      if (oz >= G.od) continue;
      for (int kh = 0; kh < G.k; ++kh) {
        const int yn = y + G.p - kh;
        if (yn < 0 || yn % G.s) continue;
        const int oy = yn / G.s;
        if (oy >= G.oh) continue;
        for (int kw = 0; kw < G.k; ++kw) {
          const int xn = x + G.p - kw;
          if (xn < 0 || xn % G.s) continue;
          const int ox = xn / G.s;
          if (ox >= G.ow) continue;
          a = fmaf(__ldg(dp + ((size_t)oz * G.oh + oy) * G.ow + ox), __ldg(wp + (kd * G.k + kh) * G.k + kw), a);
        }
      }
    }
    acc += (double)a;
  }
  const size_t vol = (size_t)G.d * G.h * G.w;
  const size_t sp = ((size_t)z * G.h + y) * G.w + x;
  if (ci < G.c0) d0[((size_t)n * G.c0 + ci) * vol + sp] = (float)acc;
  else d1[((size_t)n * G.c1 + (ci - G.c0)) * vol + sp] = (float)acc;
}

// Conv3d weight gradient: block = one (co, ci) pair and one depth tap kd (blockIdx.y), its K^2 in-plane taps in
// fp64 registers; threads stride over (n, output voxels).
template <int K>
__global__ void __launch_bounds__(128) conv_wgrad_kernel(const float* __restrict__ s0, const float* __restrict__ s1,
                                                         const float* __restrict__ dout, float* __restrict__ dw,
                                                         ConvGeom G) {
  constexpr int K2 = K * K;
  const int ci_total = G.c0 + G.c1;
  const int co = blockIdx.x / ci_total, ci = blockIdx.x % ci_total;
  const int kd = blockIdx.y;
  const size_t vol = (size_t)G.d * G.h * G.w;
  const size_t ovol = (size_t)G.od * G.oh * G.ow;
  double acc[K2];
#pragma unroll
  for (int i = 0; i < K2; ++i) acc[i] = 0.0;
  for (int n = 0; n < G.n; ++n) {
    const float* src = ci < G.c0 ? s0 + ((size_t)n * G.c0 + ci) * vol : s1 + ((size_t)n * G.c1 + (ci - G.c0)) * vol;
    const float* dp = dout + ((size_t)n * G.co + co) * ovol;
    for (size_t o = threadIdx.x; o < ovol; o += blockDim.x) {
      const int x = (int)(o % G.ow);
      const int y = (int)((o / G.ow) % G.oh);
      const int z = (int)(o / ((size_t)G.ow * G.oh));
      const int zz = z * G.s - G.p + kd;
      if (zz < 0 || zz >= G.d) continue;
      const float g = dp[o];
      const int y0 = y * G.s - G.p, x0 = x * G.s - G.p;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int yy = y0 + kh;
        if (yy < 0 || yy >= G.h) continue;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int xx = x0 + kw;
          if (xx < 0 || xx >= G.w) continue;
          acc[kh * K + kw] += (double)(g * __ldg(src + ((size_t)zz * G.h + yy) * G.w + xx));
        }
      }
    }
  }
  __shared__ double red[4][K2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K2; ++i) {
    double v = acc[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K2; i += blockDim.x)
    dw[(((size_t)co * ci_total + ci) * K + kd) * K2 + i] = (float)(red[0][i] + red[1][i] + red[2][i] + red[3][i]);
}

// out[c] = sum over n and voxels of x[n][c][:]   (bias gradients; block = one channel)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ x, int n, int c, long long vol,
                                                          float* __restrict__ out) {
  const int ch = blockIdx.x;
  double s = 0.0;
  for (int b = 0; b < n; ++b) {
    const float* p = x + ((size_t)b * c + ch) * vol;
    for (long long i = threadIdx.x; i < vol; i += blockDim.x) s += (double)p[i];
  }
  __shared__ double red[8];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[ch] = (float)t;
  }
}

// ------------------------------------------------------------------------------------------------
// ConvTranspose3d(kernel 2, stride 2): weight layout [ci][co][2][2][2]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deconv2_fwd_kernel(const float* __restrict__ src, const float* __restrict__ wt,
                                                          const float* __restrict__ bias, float* __restrict__ out, int n,
                                                          int ci_n, int co_n, int d, int h, int w, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int ow = 2 * w, oh = 2 * h, od = 2 * d;
  const int x = (int)(t % ow); t /= ow;
  const int y = (int)(t % oh); t /= oh;
  const int z = (int)(t % od); t /= od;
  const int co = (int)(t % co_n);
  const int b = (int)(t / co_n);
  const size_t vol = (size_t)d * h * w;
  const size_t sp = ((size_t)(z >> 1) * h + (y >> 1)) * w + (x >> 1);
  const int tap = ((z & 1) * 2 + (y & 1)) * 2 + (x & 1);
  double acc = bias ? (double)bias[co] : 0.0;
  for (int ci = 0; ci < ci_n; ++ci)
    acc += (double)(__ldg(src + ((size_t)b * ci_n + ci) * vol + sp) * __ldg(wt + ((size_t)ci * co_n + co) * 8 + tap));
  out[e] = (float)acc;
}

__global__ void __launch_bounds__(256) deconv2_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ wt,
                                                            float* __restrict__ dsrc, int n, int ci_n, int co_n, int d,
                                                            int h, int w, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int x = (int)(t % w); t /= w;
  const int y = (int)(t % h); t /= h;
  const int z = (int)(t % d); t /= d;
  const int ci = (int)(t % ci_n);
  const int b = (int)(t / ci_n);
  const int ow = 2 * w, oh = 2 * h;
  const size_t ovol = (size_t)8 * d * h * w;
  double acc = 0.0;
  for (int co = 0; co < co_n; ++co) {
    const float* dp = dout + ((size_t)b * co_n + co) * ovol;
    const float* wp = wt + ((size_t)ci * co_n + co) * 8;
    float a = 0.f;
#pragma unroll
    for (int tap = 0; tap < 8; ++tap) {
      const int zz = 2 * z + (tap >> 2), yy = 2 * y + ((tap >> 1) & 1), xx = 2 * x + (tap & 1);
      a = fmaf(__ldg(dp + ((size_t)zz * oh + yy) * ow + xx), __ldg(wp + tap), a);
    }
    acc += (double)a;
  }
  dsrc[e] = (float)acc;
}

// block = one (ci, co) pair
__global__ void __launch_bounds__(128) deconv2_wgrad_kernel(const float* __restrict__ src, const float* __restrict__ dout,
                                                            float* __restrict__ dw, int n, int ci_n, int co_n, int d, int h,
                                                            int w) {
  const int ci = blockIdx.x / co_n, co = blockIdx.x % co_n;
  const size_t vol = (size_t)d * h * w;
  const int ow = 2 * w, oh = 2 * h;
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  for (int b = 0; b < n; ++b) {
    const float* sp = src + ((size_t)b * ci_n + ci) * vol;
    const float* dp = dout + ((size_t)b * co_n + co) * vol * 8;
    for (size_t i = threadIdx.x; i < vol; i += blockDim.x) {
      const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((size_t)w * h));
      const float v = sp[i];
#pragma unroll
      for (int tap = 0; tap < 8; ++tap) {
        const int zz = 2 * z + (tap >> 2), yy = 2 * y + ((tap >> 1) & 1), xx = 2 * x + (tap & 1);
        acc[tap] += (double)(v * __ldg(dp + ((size_t)zz * oh + yy) * ow + xx));
      }
    }
  }
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double v = acc[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8)
    dw[((size_t)ci * co_n + co) * 8 + threadIdx.x] =
        (float)(red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// Normalisation statistics. mode 0: InstanceNorm (per n, c; biased variance); 1: BatchNorm training (per c over
// n and voxels; running stats updated with momentum and the unbiased variance); 2: BatchNorm eval (running stats).
// block = one channel; two passes (mean, then centred second moment) in fp64.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) norm_stats_kernel(const float* __restrict__ y, int n, int c, long long vol, int mode,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, float momentum, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float* __restrict__ scale,
                                                         float* __restrict__ shift, float* __restrict__ mean_out,
                                                         float* __restrict__ rstd_out) {
  __shared__ double red[8];
  const int ch = blockIdx.x;
  const double g = gamma ? (double)gamma[ch] : 1.0, bt = beta ? (double)beta[ch] : 0.0;
  if (mode == 2) {
    const double mean = running_mean[ch], rstd = 1.0 / sqrt((double)running_var[ch] + (double)eps);
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
      const size_t o = (size_t)b * c + ch;
      scale[o] = (float)(g * rstd); shift[o] = (float)(bt - mean * g * rstd);
      mean_out[o] = (float)mean; rstd_out[o] = (float)rstd;
    }
    return;
  }
  const int groups = mode == 0 ? n : 1;
  for (int gi = 0; gi < groups; ++gi) {
    const int b0 = mode == 0 ? gi : 0, b1 = mode == 0 ? gi + 1 : n;
    const double cnt = (double)(b1 - b0) * (double)vol;
    double s = 0.0;
    for (int b = b0; b < b1; ++b) {
      const float* p = y + ((size_t)b * c + ch) * vol;
      for (long long i = threadIdx.x; i < vol; i += blockDim.x) s += (double)p[i];
    }
    const double mean = block_sum_256(s, red) / cnt;
    double q = 0.0;
    for (int b = b0; b < b1; ++b) {
      const float* p = y + ((size_t)b * c + ch) * vol;
      for (long long i = threadIdx.x; i < vol; i += blockDim.x) { const double dlt = (double)p[i] - mean; q += dlt * dlt; }
    }
    const double var = block_sum_256(q, red) / cnt;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    if (threadIdx.x == 0) {
      for (int b = b0; b < b1; ++b) {
        const size_t o = (size_t)b * c + ch;
        scale[o] = (float)(g * rstd); shift[o] = (float)(bt - mean * g * rstd);
        mean_out[o] = (float)mean; rstd_out[o] = (float)rstd;
      }
      if (mode == 1 && running_mean != nullptr) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mean);
        running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unbiased);
      }
    }
  }
}

// counter-based dropout mask of the fp32 mode (element index in NCDHW order)
__device__ __forceinline__ bool keep_element(unsigned long long e, uint32_t seed, uint32_t thresh) {
  uint32_t x = (uint32_t)e ^ (seed * 0x9E3779B9u) ^ ((uint32_t)(e >> 32) * 0x85EBCA6Bu);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x >= thresh;
}

// a = LeakyReLU(Dropout(y * scale + shift)); scale == nullptr: no normalisation
__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float slope, float drop_p,
                                                           uint32_t seed, uint32_t thresh, long long vol, long long total,
                                                           float* __restrict__ a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long nc = e / vol;
  float v = y[e];
  if (scale) v = fmaf(v, scale[nc], shift[nc]);
  if (drop_p > 0.f) v = keep_element((unsigned long long)e, seed, thresh) ? v / (1.f - drop_p) : 0.f;
  a[e] = v > 0.f ? v : v * slope;
}

// MaxPool3d(2): thread = one pooled element
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ a, float* __restrict__ pooled, int d,
                                                          int h, int w, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int pw = w >> 1, ph = h >> 1, pd = d >> 1;
  long long t = e;
  const int x = (int)(t % pw); t /= pw;
  const int y = (int)(t % ph); t /= ph;
  const int z = (int)(t % pd);
  const long long nc = t / pd;
  const float* p = a + (size_t)nc * d * h * w;
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = p[((size_t)(2 * z + (j >> 2)) * h + 2 * y + ((j >> 1) & 1)) * w + 2 * x + (j & 1)];
    m = v > m ? v : m;
  }
  pooled[e] = m;
}

// dz = (dA + routed pooled gradient) * lrelu'(pre-activation) * dropout factor.
// dA or dP may be nullptr (not both). The pooled gradient goes to the FIRST maximum of a window in (d,h,w) scan
// order (torch's rule). The activation branch is taken from sign(y * scale + shift), or sign(a) without a norm.
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ dP,
                                                      const float* __restrict__ a, const float* __restrict__ y,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      float slope, float drop_p, uint32_t seed, uint32_t thresh, int d, int h,
                                                      int w, long long total, float* __restrict__ dz) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long vol = (long long)d * h * w;
  const long long nc = e / vol;
  float g = dA ? dA[e] : 0.f;
  if (dP) {
    long long t = e - nc * vol;
    const int x = (int)(t % w); t /= w;
    const int yy = (int)(t % h);
    const int z = (int)(t / h);
    const int pw = w >> 1, ph = h >> 1, pd = d >> 1;
    if ((x >> 1) < pw && (yy >> 1) < ph && (z >> 1) < pd) {
      const float* p = a + (size_t)nc * vol;
      const int me = ((z & 1) * 2 + (yy & 1)) * 2 + (x & 1);
      float m = -INFINITY;
      int arg = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = p[((size_t)((z & ~1) + (j >> 2)) * h + (yy & ~1) + ((j >> 1) & 1)) * w + (x & ~1) + (j & 1)];
        if (v > m) { m = v; arg = j; }
      }
      if (arg == me) g += dP[((size_t)nc * pd + (z >> 1)) * ph * pw + (size_t)(yy >> 1) * pw + (x >> 1)];
    }
  }
  if (slope != 1.f) {
    const float pre = scale ? fmaf(y[e], scale[nc], shift[nc]) : a[e];
    if (!(pre > 0.f)) g *= slope;
  }
  if (drop_p > 0.f) g = keep_element((unsigned long long)e, seed, thresh) ? g / (1.f - drop_p) : 0.f;
  dz[e] = g;
}

// block = one channel: S1 = sum dz, S2 = sum dz * xhat per (n, c) [InstanceNorm] or per c [BatchNorm training];
// c1 = S1 / cnt, c2 = S2 / cnt per (n, c); dgamma = sum_n S2, dbeta = sum_n S1. mode 2 (eval BN): c1 = c2 = 0.
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              int n, int c, long long vol, int mode, float* __restrict__ c1,
                                                              float* __restrict__ c2, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
  __shared__ double red[8];
  const int ch = blockIdx.x;
  double t1 = 0.0, t2 = 0.0;
  for (int b = 0; b < n; ++b) {
    const size_t o = (size_t)b * c + ch;
    const double m = mean[o], r = rstd[o];
    const float* dp = dz + o * vol;
    const float* yp = y + o * vol;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = threadIdx.x; i < vol; i += blockDim.x) {
      const double g = dp[i];
      s1 += g;
      s2 += g * (((double)yp[i] - m) * r);
    }
    s1 = block_sum_256(s1, red);
    s2 = block_sum_256(s2, red);
    t1 += s1; t2 += s2;
    if (mode == 0 && threadIdx.x == 0) { c1[o] = (float)(s1 / (double)vol); c2[o] = (float)(s2 / (double)vol); }
  }
  if (threadIdx.x == 0) {
    if (mode != 0) {
      const double cnt = (double)n * (double)vol;
      for (int b = 0; b < n; ++b) {
        c1[(size_t)b * c + ch] = mode == 1 ? (float)(t1 / cnt) : 0.f;
        c2[(size_t)b * c + ch] = mode == 1 ? (float)(t2 / cnt) : 0.f;
      }
    }
    if (dgamma) dgamma[ch] = (float)t2;
    if (dbeta) dbeta[ch] = (float)t1;
  }
}

// dy = gamma * rstd * (dz - c1 - xhat * c2), in place over dz
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(float* __restrict__ dz, const float* __restrict__ y,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ scale, const float* __restrict__ c1,
                                                             const float* __restrict__ c2, long long vol, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long nc = e / vol;
  const float xhat = (y[e] - mean[nc]) * rstd[nc];
  dz[e] = scale[nc] * (dz[e] - c1[nc] - xhat * c2[nc]);
}

}  // namespace f32
}  // namespace ub
