// Marching-plane weight gradient for the 3x3x3 convs, one (32 input channels x 32 output channels)
// block per CTA column: dW[kd,kh,kw][ci][co] = sum_v X[v + (kd,kh,kw) - 1][ci] * dY[v][co].
//
// igemm_wgrad_kernel handles one kd per CTA with N = C_out = 32 (smem-bound UMMA, X planes loaded three
// times). Here the depth taps are folded into N by shifting dY instead of X: for one X plane p,
//
//     D[kh][(kw, ci)][(j, co)] += sum_{h,w} X[p, h+kh-1, w+kw-1, ci] * dY[p - 1 + j, h, w, co],   kd = 2 - j
//
// so ONE UMMA (M = 128 = 4 kw-atoms x 32 ci, N = 96 = 3 dY planes x 32 co, K = 16 voxels) covers nine
// taps. Both operands are MN-major views of the tiles exactly as TMA delivered them:
//   A atoms (kw) are one halo row apart (LBO = 64 B), B atoms (j) are one dY plane apart
//   (LBO = 8 KB): the dY planes live in a ring whose first two slots are mirrored behind the last so
//   that three consecutive planes are always contiguous. Planes -1 and D are out-of-bounds TMA loads
//   (zero fill) -- the conv padding in depth.
// A CTA marches along d through its share of (n, h-tile, w-tile, d-segment) items and keeps the three
// kh accumulators (3 x 96 TMEM columns) for its whole lifetime (split-K over CTAs), then writes one
// fp32 partial record [27][32 ci][32 co]; wgrad_reduce_kernel sums the records in a fixed order.
#pragma once
#include "igemm_fwd.cuh"

namespace ub {

struct WgradMarchParams {
  CUtensorMap tm_x[2];      // box (32, 10, 18, 1, 1)
  CUtensorMap tm_dy;        // box (32, 8, 16, 1, 1)
  int n_chunks_src0, n_chunks_total;
  int Nb, D, H, W;
  int tiles_w, tiles_h, nseg, seg_len;
  int ci_total;             // padded input channels (partial pitch)
  int co_total, n_cotiles;  // padded output channels, co_total / 32
  float* partial;           // [gridDim.x][27][ci_total][co_total]
};

constexpr int kWmXStages = 4, kWmXBytes = 12288;
constexpr int kWmYSlots = 8, kWmYBytes = 8192;   // + 2 mirror slots

__global__ void __launch_bounds__(kIgemmThreads, 1)
wgrad_march_kernel(const __grid_constant__ WgradMarchParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t x_base = base;
  const uint32_t y_base = x_base + kWmXStages * kWmXBytes;
  const uint32_t bar_base = y_base + (kWmYSlots + 2) * kWmYBytes;
  const uint32_t x_full = bar_base, x_empty = x_full + 8 * kWmXStages, y_full = x_empty + 8 * kWmXStages,
                 y_empty = y_full + 8 * kWmYSlots, acc_full = y_empty + 8 * kWmYSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (acc_full + 8 - base));

  const int chunk = blockIdx.y / P.n_cotiles;
  const int cot = blockIdx.y - chunk * P.n_cotiles;
  const bool s1 = chunk >= P.n_chunks_src0;
  const int c0 = (s1 ? chunk - P.n_chunks_src0 : chunk) * 32;
  // contiguous range of items for this CTA
  const int items = P.Nb * P.tiles_h * P.tiles_w * P.nseg;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int i_begin = blockIdx.x * per;
  int i_end = i_begin + per; if (i_end > items) i_end = items;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWmXStages; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 1); }
    for (int i = 0; i < kWmYSlots; ++i) { mbar_init(y_full + 8 * i, 1); mbar_init(y_empty + 8 * i, 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&P.tm_x[s1 ? 1 : 0]);
    tma_prefetch_desc(&P.tm_dy);
  }
  if (warp == 5) tmem_alloc_rt(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const CUtensorMap* tmx = &P.tm_x[s1 ? 1 : 0];
      int xs = 0; uint32_t xp = 0;    // X ring
      int ys = 0; uint32_t yp = 0;    // dY ring
      auto load_dy = [&](int q, int nb, int h0, int w0) {
        mbar_wait(y_empty + 8 * ys, yp ^ 1);
        const bool mirror = ys < 2;
        mbar_expect_tx(y_full + 8 * ys, mirror ? 2 * kWmYBytes : kWmYBytes);
        tma_load_5d(y_base + ys * kWmYBytes, &P.tm_dy, y_full + 8 * ys, cot * 32, w0, h0, q, nb);
        if (mirror) tma_load_5d(y_base + (kWmYSlots + ys) * kWmYBytes, &P.tm_dy, y_full + 8 * ys, cot * 32, w0, h0, q, nb);
        if (++ys == kWmYSlots) { ys = 0; yp ^= 1; }
      };
      for (int it = i_begin; it < i_end; ++it) {
        int t = it;
        const int seg = t % P.nseg; t /= P.nseg;
        const int w0 = (t % P.tiles_w) * 8; t /= P.tiles_w;
        const int h0 = (t % P.tiles_h) * 16; t /= P.tiles_h;
        const int nb = t;
        const int d0 = seg * P.seg_len;
        int d1 = d0 + P.seg_len; if (d1 > P.D) d1 = P.D;
        // the item owns X planes [d0, d1) and pairs each with dY planes p-1, p, p+1; planes -1 and D
        // are out-of-bounds TMA coordinates (zero fill = the conv padding in depth)
        load_dy(d0 - 1, nb, h0, w0);
        load_dy(d0, nb, h0, w0);
        for (int p = d0; p < d1; ++p) {
          load_dy(p + 1, nb, h0, w0);
          mbar_wait(x_empty + 8 * xs, xp ^ 1);
          mbar_expect_tx(x_full + 8 * xs, 180 * 64);
          tma_load_5d(x_base + xs * kWmXBytes, tmx, x_full + 8 * xs, c0, w0 - 1, h0 - 1, p, nb);
          if (++xs == kWmXStages) { xs = 0; xp ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16(128, 96, 1, 1);
    const uint64_t a_desc0 = make_smem_desc(0, /*lbo: next kw atom = next halo row*/ 64, /*sbo: next h row*/ 10 * 64, SWZ_64B);
    const uint64_t b_desc0 = make_smem_desc(0, /*lbo: next dY plane*/ kWmYBytes, /*sbo: next 8 voxel rows*/ 8 * 64, SWZ_64B);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
    const bool leader = elect_one();
    int xs = 0; uint32_t xp = 0;
    int ys = 0; uint32_t yp = 0;       // slot / phase of the OLDEST of the three live dY planes
    uint32_t first = 1;
    for (int it = i_begin; it < i_end; ++it) {
      const int seg = it % P.nseg;
      const int d0 = seg * P.seg_len;
      int d1 = d0 + P.seg_len; if (d1 > P.D) d1 = P.D;
      // the first two dY planes of the item
      {
        int s = ys; uint32_t ph = yp;
        for (int k = 0; k < 2; ++k) {
          mbar_wait(y_full + 8 * s, ph);
          if (++s == kWmYSlots) { s = 0; ph ^= 1; }
        }
      }
      for (int p = d0; p < d1; ++p) {
        // newest plane (oldest + 2)
        {
          int s = ys + 2; uint32_t ph = yp;
          if (s >= kWmYSlots) { s -= kWmYSlots; ph ^= 1; }
          mbar_wait(y_full + 8 * s, ph);
        }
        mbar_wait(x_full + 8 * xs, xp);
        tc_fence_after();
        const bool last = p + 1 == d1;
        if (leader) {
          const uint32_t a_lo = a_lo0 + ((x_base + xs * kWmXBytes) >> 4);
          const uint32_t b_lo = b_lo0 + ((y_base + ys * kWmYBytes) >> 4);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_lohi(tmem + kh * 96, a_lo + (uint32_t)((kh * 10 + ks * 20) * 64 >> 4), a_hi,
                             b_lo + (uint32_t)(ks * 16 * 64 >> 4), b_hi, idesc, (ks == 0 ? (first ^ 1u) : 1u));
          umma_commit(x_empty + 8 * xs);
          umma_commit(y_empty + 8 * ys);           // the oldest plane is dead after this step
          if (last) {                              // ... and so are the other two at the end of an item
            int s = ys + 1; if (s >= kWmYSlots) s -= kWmYSlots;
            umma_commit(y_empty + 8 * s);
            s = ys + 2; if (s >= kWmYSlots) s -= kWmYSlots;
            umma_commit(y_empty + 8 * s);
          }
        }
        __syncwarp();
        first = 0;
        if (++xs == kWmXStages) { xs = 0; xp ^= 1; }
        const int adv = last ? 3 : 1;
        ys += adv;
        if (ys >= kWmYSlots) { ys -= kWmYSlots; yp ^= 1; }
      }
    }
    if (leader) umma_commit(acc_full);
    __syncwarp();
  } else {
    // =========================== epilogue: TMEM -> partial record ===========================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool any = i_end > i_begin;
    const size_t tap_elems = (size_t)P.ci_total * P.co_total;
    float* outb = P.partial + (size_t)blockIdx.x * 27 * tap_elems;
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + kh * 96 + j * 32, rr);
        tmem_ld_wait();
        if (warp < 3) {   // warp = kw atom (the 4th atom is unused), lane = ci
          const int kd = 2 - j;
          float4* d4 = reinterpret_cast<float4*>(outb + (size_t)((kd * 3 + kh) * 3 + warp) * tap_elems +
                                                 (size_t)(chunk * 32 + lane) * P.co_total + cot * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            d4[q] = any ? make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]),
                                      __uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc_rt(tmem, 512);
}

}  // namespace ub
