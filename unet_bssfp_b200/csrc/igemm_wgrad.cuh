// Weight-gradient implicit GEMM for sm_100a: dW[tap][ci][co] = sum_voxels X[v + tap][ci] * dY[v][co].
//
// Both operands are read where the forward pass left them -- NDHWC bf16, channels contiguous -- so
// they are *MN-major* UMMA operands (the reduction index K is the voxel row):
//   A (M side) = a halo tile of X, one 32-channel chunk (64-byte rows, SWIZZLE_64B). One UMMA covers
//                M = 128 = 4 atoms of 32 channels; atom a starts `a` rows later (LBO = one row), i.e.
//                the atoms are the kw = 0,1,2 taps of one (kd,kh) pair (+1 unused atom). K = 16 voxels
//                per instruction = two h-rows of 8 w (SBO = box-width rows).
//   B (N side) = the matching 16x8 dY plane tile, NT channels (one or two swizzle atoms).
// Accumulators: one [128 x NT] fp32 block per kh group in TMEM, kept for the CTA's whole share of
// the voxel tiles (split-K over CTAs), then written once to a per-split partial buffer
// [split][tap][ci][co] (fp32). A small reduce kernel sums the splits in a fixed order
// (deterministic) and converts to the torch weight layout.
//
// Work decomposition: blockIdx.x = voxel split; blockIdx.y enumerates (ci chunk, variant, co tile).
// A "variant" fixes the X-tile origin offset and the dY parity offset: kd for 3x3x3, (parity, jd)
// for the 4x4x4 stride-2 conv, the sub-position for the transposed conv.
#pragma once
#include "igemm_fwd.cuh"

namespace ub {

constexpr int kMaxWgVariants = 16;

struct WgradParams {
  CUtensorMap tm_x[2];
  CUtensorMap tm_dy;
  int n_chunks_src0, n_chunks_total;     // 32-channel chunks of X (both sources)
  int bw, bh;                            // X tile box rows
  int x_stride;                          // X coordinate = tile coord * x_stride + off
  int dy_stride;                         // dY coordinate = tile coord * dy_stride + off
  int n_variants;
  int x_off[kMaxWgVariants][3];          // (w,h,d)
  int x_nmul, x_nadd[kMaxWgVariants];    // X batch coordinate = nb * x_nmul + x_nadd[variant] (parity-planar
                                         // space-to-depth source: x_nmul = 8, x_nadd = parity of the variant)
  int dy_off[kMaxWgVariants][3];
  int ngroups, group_row_step;           // groups per variant (kh), rows between group starts
  int natoms;                            // useful atoms per group (kw), <= 4
  int n_cotiles, nt;                     // dY column tiles of nt channels each
  int ncb;                               // channels per dY TMA box (min(nt,64)); swizzle = ncb*2 bytes
  int Nb, Dt, Ht, Wt;                    // tile space (voxels the reduction runs over)
  int tiles_w, tiles_h;
  int ci_total, co_total;                // padded totals (partial buffer pitch)
  float* partial;                        // [nsplit][n_variants*ngroups*natoms][ci_total][co_total]
  int x_stage_bytes, dy_stage_bytes, nstages, tmem_cols;
  // Transposed conv ("dc mode", chunk_atoms > 0): the four M atoms are up to four 32-channel CHUNKS of X (their tiles
  // sit 8 KB apart in the stage, LBO = 8 KB) instead of kw taps, and N = var_boxes sub-positions x nt channels: one
  // CTA reads its X tiles and dY parity tiles once for chunk_atoms x var_boxes (chunk, sub-position) pairs. Before,
  // every chunk re-read dY and every sub-position re-read X: 6.4 GB through L2 for 2.4 GB of operands on the
  // 64 -> 64 layer, which bound the launch (profiles/r02h_deconv_wgrad_ncu_summary.txt).
  int chunk_atoms, var_boxes;
};

// Warps 0-3 epilogue, 4 TMA producer of the X tiles, 5 MMA issuer, 6 TMA producer of the dY tiles (two producer threads:
// a stage of the transposed-conv form is 2-4 X boxes + 4 dY boxes for 8 UMMAs, one thread's issue rate was the bound).
constexpr int kWgradThreads = kIgemmThreads + 32;
__global__ void __launch_bounds__(kWgradThreads, 1)
igemm_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = P.x_stage_bytes + P.dy_stage_bytes;
  const uint32_t bar_base = base + P.nstages * stage_bytes;
  const uint32_t full = bar_base, empty = full + 8 * P.nstages, acc_full = empty + 8 * P.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (acc_full + 8 - base));

  // ---- work item
  const bool dc = P.chunk_atoms > 0;
  const int vboxes = dc ? P.var_boxes : 1;
  int y = blockIdx.y;
  const int cot = y % P.n_cotiles; y /= P.n_cotiles;
  const int nvg = P.n_variants / vboxes;             // variant groups (dc mode: var_boxes sub-positions each)
  const int var = (y % nvg) * vboxes; y /= nvg;      // first variant of the group
  const int chunk = dc ? y * P.chunk_atoms : y;      // first chunk of the group
  int ca_eff = dc ? P.n_chunks_total - chunk : 1;    // chunks of this group
  if (dc && ca_eff > P.chunk_atoms) ca_eff = P.chunk_atoms;
  const int n0 = cot * P.nt;
  const bool s1 = chunk >= P.n_chunks_src0;
  const int c0 = (s1 ? chunk - P.n_chunks_src0 : chunk) * 32;
  const int nn = vboxes * P.nt;                      // UMMA N
  const int ntc = (nn + 31) & ~31;

  // ---- this CTA's share of the plane tiles
  const long long total = (long long)P.Nb * P.Dt * P.tiles_h * P.tiles_w;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per;
  long long t_end = t_begin + per; if (t_end > total) t_end = total;
  const int niter = t_end > t_begin ? (int)(t_end - t_begin) : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nstages; ++i) { mbar_init(full + 8 * i, 2); mbar_init(empty + 8 * i, 1); }   // full: both producers arrive
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&P.tm_x[s1 ? 1 : 0]);
  if (warp == 6 && lane == 0) tma_prefetch_desc(&P.tm_dy);
  if (warp == 5) tmem_alloc_rt(smem_u32(tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nboxes = P.nt / P.ncb;
  const int dy_pitch = P.ncb * 2;

  if (warp == 4 || warp == 6) {
    if (lane == 0) {
      const bool xprod = warp == 4;     // this thread loads the X tiles, the other one the dY tiles
      const uint32_t bytes = xprod ? (uint32_t)(ca_eff * P.bh * P.bw * 64) : (uint32_t)(vboxes * 128 * P.nt * 2);
      const CUtensorMap* tmx = &P.tm_x[s1 ? 1 : 0];
      long long tt = t_begin;
      int tw_i = (int)(tt % P.tiles_w); tt /= P.tiles_w;
      int th_i = (int)(tt % P.tiles_h); tt /= P.tiles_h;
      int d = (int)(tt % P.Dt); tt /= P.Dt;
      int nb = (int)tt;
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < niter; ++it) {
        mbar_wait(empty + 8 * st, ph ^ 1);
        mbar_expect_tx(full + 8 * st, bytes);
        const uint32_t xs = base + st * stage_bytes;
        if (xprod) {
          for (int a = 0; a < ca_eff; ++a)     // dc mode: the chunks of the group, 8 KB apart (bw x bh = 8 x 16 rows of 64 B)
            tma_load_5d(xs + a * (P.bh * P.bw * 64), tmx, full + 8 * st, c0 + a * 32, tw_i * 8 * P.x_stride + P.x_off[var][0],
                        th_i * 16 * P.x_stride + P.x_off[var][1], d * P.x_stride + P.x_off[var][2],
                        nb * P.x_nmul + P.x_nadd[var]);
        } else {
          for (int vb = 0; vb < vboxes; ++vb)
            for (int b = 0; b < nboxes; ++b)
              tma_load_5d(xs + P.x_stage_bytes + (vb * nboxes + b) * 128 * dy_pitch, &P.tm_dy, full + 8 * st, n0 + b * P.ncb,
                          tw_i * 8 * P.dy_stride + P.dy_off[var + vb][0], th_i * 16 * P.dy_stride + P.dy_off[var + vb][1],
                          d * P.dy_stride + P.dy_off[var + vb][2], nb);
        }
        if (++st == P.nstages) { st = 0; ph ^= 1; }
        if (++tw_i == P.tiles_w) {
          tw_i = 0;
          if (++th_i == P.tiles_h) {
            th_i = 0;
            if (++d == P.Dt) { d = 0; ++nb; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    const uint32_t swz_b = dy_pitch == 128 ? SWZ_128B : (dy_pitch == 64 ? SWZ_64B : SWZ_32B);
    const uint32_t idesc = make_idesc_bf16(128, nn, 1, 1);
    const uint64_t a_desc0 = make_smem_desc(0, /*lbo: next atom = next row (dc mode: next chunk tile)*/ dc ? P.bh * P.bw * 64 : 64,
                                            /*sbo*/ P.bw * 64, SWZ_64B);
    const uint64_t b_desc0 = make_smem_desc(0, /*lbo: next 64-channel box*/ 128 * dy_pitch, 8 * dy_pitch, swz_b);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;   // LBO fields, address 0
    const uint32_t a_kstep = (uint32_t)(2 * P.bw * 64) >> 4, b_kstep = (uint32_t)(16 * dy_pitch) >> 4;
    const uint32_t g_step = (uint32_t)(P.group_row_step * 64) >> 4;
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < niter; ++it) {
      mbar_wait(full + 8 * st, ph);
      tc_fence_after();
      if (leader) {
        const uint32_t xs = base + st * stage_bytes;
        const uint32_t a_lo = a_lo0 + (xs >> 4);
        const uint32_t b_lo = b_lo0 + ((xs + P.x_stage_bytes) >> 4);
        for (int g = 0; g < P.ngroups; ++g) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_lohi(tmem + g * ntc, a_lo + g * g_step + ks * a_kstep, a_hi, b_lo + ks * b_kstep, b_hi, idesc,
                           (uint32_t)((it | ks) != 0));
        }
        umma_commit(empty + 8 * st);
      }
      __syncwarp();
      if (++st == P.nstages) { st = 0; ph ^= 1; }
    }
    if (leader) umma_commit(acc_full);
    __syncwarp();
  } else {
    // epilogue: row = atom * 32 + ci  ->  warp index is the atom (kw), lane is the channel
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int ntaps_v = P.ngroups * P.natoms;
    const size_t tap_elems = (size_t)P.ci_total * P.co_total;
    float* outb = P.partial + (size_t)blockIdx.x * ((size_t)P.n_variants * ntaps_v) * tap_elems;
    for (int g = 0; g < P.ngroups; ++g) {
      // legacy: warp = kw atom of chunk `chunk`; dc mode: warp = chunk of the group, the columns run over var_boxes taps
      const int tap = dc ? var : (var * P.ngroups + g) * P.natoms + warp;
      float* dst = outb + (size_t)tap * tap_elems + (size_t)((dc ? chunk + warp : chunk) * 32 + lane) * P.co_total + n0;
      for (int cc = 0; cc * 32 < nn; ++cc) {
        const int ncol = (nn - cc * 32) >= 32 ? 32 : 16;
        uint32_t rr[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + g * ntc + cc * 32;
        if (ncol == 32) {
          tmem_ld_32x32b_x32(taddr, rr);
        } else {
          uint32_t r16[16];
          tmem_ld_32x32b_x16(taddr, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { rr[j] = r16[j]; rr[j + 16] = 0u; }
        }
        tmem_ld_wait();
        if (warp < (dc ? ca_eff : P.natoms)) {
          // dc mode: column cc * 32 belongs to sub-position var + (cc * 32) / nt, channel n0 + (cc * 32) % nt
          float4* d4 = reinterpret_cast<float4*>(dc ? dst + (size_t)((cc * 32) / P.nt) * tap_elems + (cc * 32) % P.nt : dst + cc * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j * 4 < ncol) {
              float4 o;
              if (niter > 0) {
                o.x = __uint_as_float(rr[j * 4 + 0]); o.y = __uint_as_float(rr[j * 4 + 1]);
                o.z = __uint_as_float(rr[j * 4 + 2]); o.w = __uint_as_float(rr[j * 4 + 3]);
              } else {
                o = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              d4[j] = o;
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc_rt(tmem, P.tmem_cols);
}

}  // namespace ub
