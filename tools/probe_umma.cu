// Bring-up probe for the sm_100a building blocks the conv kernels rely on. Standalone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_umma tools/probe_umma.cu
// Each case builds TMA tensor maps + a host-generated list of tcgen05.mma instructions
// (descriptors relative to the smem base), runs them in one CTA and compares the TMEM
// accumulator with a CPU result computed from the same bf16 integers (exact in fp32).
//   K1..K3 : K-major operands, SWIZZLE_128B/64B/32B                (plain GEMM)
//   S1..S3 : row-shifted views into one halo tile (3x3 taps)       (implicit-GEMM fwd/dgrad)
//   W1..W3 : MN-major operands with LBO/SBO striding over the halo (implicit-GEMM wgrad)
//   T1     : TMA elementStrides=2 gather                            (stride-2 conv / deconv dgrad)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../unet_bssfp_b200/csrc/sm100_ptx.cuh"

using namespace ub;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

struct MmaOp {
  uint64_t adesc, bdesc;
  uint32_t idesc, accumulate, tmem_col, pad;
};
struct TmaOp {
  int map, rank;
  uint32_t smem_off, bytes;
  int c[5];
};
struct Maps {
  CUtensorMap m[4];
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ Maps maps, const TmaOp* tmas, int ntma, const MmaOp* mmas, int nmma,
             float* out, int ncols, uint8_t* dump, int dump_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  uint32_t base = smem_u32(smem_raw);
  base = (base + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int i = 0; i < ntma; ++i) total += tmas[i].bytes;
    mbar_expect_tx(bar_tma, total);
    for (int i = 0; i < ntma; ++i) {
      const TmaOp t = tmas[i];
      const CUtensorMap* m = &maps.m[t.map];
      uint32_t dst = base + t.smem_off;
      if (t.rank == 2) tma_load_2d(dst, m, bar_tma, t.c[0], t.c[1]);
      else if (t.rank == 3) tma_load_3d(dst, m, bar_tma, t.c[0], t.c[1], t.c[2]);
      else if (t.rank == 4) tma_load_4d(dst, m, bar_tma, t.c[0], t.c[1], t.c[2], t.c[3]);
      else tma_load_5d(dst, m, bar_tma, t.c[0], t.c[1], t.c[2], t.c[3], t.c[4]);
    }
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    for (int i = 0; i < nmma; ++i) {
      const MmaOp o = mmas[i];
      umma_bf16(tmem + o.tmem_col, o.adesc + (uint64_t)(base >> 4), o.bdesc + (uint64_t)(base >> 4),
                o.idesc, o.accumulate);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * ncols + c0 + j] = __uint_as_float(r[j]);
  }
  for (int i = threadIdx.x; i < dump_bytes; i += blockDim.x) dump[i] = sm[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static CUtensorMapSwizzle swz_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
       : bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
       : bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                      : CU_TENSOR_MAP_SWIZZLE_NONE;
}
static uint32_t swz_desc(int bytes) { return bytes == 128 ? SWZ_128B : bytes == 64 ? SWZ_64B : bytes == 32 ? SWZ_32B : SWZ_NONE; }

// dims[0] is the contiguous dimension. strides given in elements for dims 1..rank-1.
static CUtensorMap make_map(void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                            const uint32_t* box, const uint32_t* estr, int swizzle_bytes) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr ? estr[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_elems[i] * 2;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, gptr, gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(swizzle_bytes),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(3); }
  return m;
}

static uint32_t g_seed = 12345;
static float rnd_int(int lo, int hi) {
  g_seed = g_seed * 1664525u + 1013904223u;
  return (float)(lo + (int)((g_seed >> 8) % (uint32_t)(hi - lo + 1)));
}
static std::vector<__nv_bfloat16> rand_bf16(size_t n, int lo = -3, int hi = 3) {
  std::vector<__nv_bfloat16> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = __float2bfloat16(rnd_int(lo, hi));
  return v;
}
static float bf(const __nv_bfloat16& x) { return __bfloat162float(x); }

struct Case {
  Maps maps;
  std::vector<TmaOp> tmas;
  std::vector<MmaOp> mmas;
  int ncols = 32;
  int dump_bytes = 0;
};

static void run_case(const char* name, Case& cs, std::vector<float>& out, std::vector<uint8_t>* dump = nullptr) {
  TmaOp* d_t; MmaOp* d_m; float* d_o; uint8_t* d_d;
  CK(cudaMalloc(&d_t, sizeof(TmaOp) * (cs.tmas.size() + 1)));
  CK(cudaMalloc(&d_m, sizeof(MmaOp) * (cs.mmas.size() + 1)));
  CK(cudaMalloc(&d_o, sizeof(float) * 128 * 512));
  CK(cudaMalloc(&d_d, 200 * 1024));
  CK(cudaMemcpy(d_t, cs.tmas.data(), sizeof(TmaOp) * cs.tmas.size(), cudaMemcpyHostToDevice));
  if (!cs.mmas.empty()) CK(cudaMemcpy(d_m, cs.mmas.data(), sizeof(MmaOp) * cs.mmas.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_o, 0, sizeof(float) * 128 * 512));
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(cs.maps, d_t, (int)cs.tmas.size(), d_m, (int)cs.mmas.size(), d_o, cs.ncols,
                                 d_d, cs.dump_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] KERNEL FAILED: %s\n", name, cudaGetErrorString(e)); exit(4); }
  out.resize((size_t)128 * cs.ncols);
  CK(cudaMemcpy(out.data(), d_o, sizeof(float) * 128 * cs.ncols, cudaMemcpyDeviceToHost));
  if (dump) { dump->resize(cs.dump_bytes); CK(cudaMemcpy(dump->data(), d_d, cs.dump_bytes, cudaMemcpyDeviceToHost)); }
  cudaFree(d_t); cudaFree(d_m); cudaFree(d_o); cudaFree(d_d);
}

static int report(const char* name, const std::vector<float>& got, const std::vector<float>& ref, int rows,
                  int cols, int ld) {
  double maxerr = 0; int bad = 0;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      double d = fabs((double)got[(size_t)r * ld + c] - (double)ref[(size_t)r * cols + c]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-3) ++bad;
    }
  printf("[%s] %s  max_abs_err=%g  mismatches=%d/%d\n", name, bad ? "FAIL" : "PASS", maxerr, bad, rows * cols);
  fflush(stdout);
  return bad ? 1 : 0;
}

template <class T>
static T* to_dev(const std::vector<T>& v) {
  T* d; CK(cudaMalloc(&d, v.size() * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// ----- K-major GEMM: D[128 x N] = A[128 x K] * B[N x K]^T, operands loaded in chunks of `kc` elements
static int case_kmajor(const char* name, int kc, int N, int K) {
  const int sw = kc * 2;
  auto A = rand_bf16((size_t)128 * K), B = rand_bf16((size_t)N * K);
  auto dA = to_dev(A); auto dB = to_dev(B);
  Case cs; cs.ncols = (N + 31) / 32 * 32;
  uint64_t dimsA[2] = {(uint64_t)K, 128}, strA[1] = {(uint64_t)K};
  uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)K};
  uint32_t boxA[2] = {(uint32_t)kc, 128}, boxB[2] = {(uint32_t)kc, (uint32_t)N};
  cs.maps.m[0] = make_map(dA, 2, dimsA, strA, boxA, nullptr, sw);
  cs.maps.m[1] = make_map(dB, 2, dimsB, strB, boxB, nullptr, sw);
  const int nchunk = K / kc;
  const uint32_t a_bytes = 128 * sw, b_bytes = N * sw;
  const uint32_t b_base = ((nchunk * a_bytes) + 1023) & ~1023u;
  for (int c = 0; c < nchunk; ++c) {
    cs.tmas.push_back({0, 2, c * a_bytes, a_bytes, {c * kc, 0, 0, 0, 0}});
    cs.tmas.push_back({1, 2, b_base + c * b_bytes, b_bytes, {c * kc, 0, 0, 0, 0}});
  }
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  int first = 1;
  for (int c = 0; c < nchunk; ++c)
    for (int k = 0; k < kc / 16; ++k) {
      MmaOp o{};
      o.adesc = make_smem_desc(c * a_bytes + k * 32, 16, 8 * sw, swz_desc(sw));
      o.bdesc = make_smem_desc(b_base + c * b_bytes + k * 32, 16, 8 * sw, swz_desc(sw));
      o.idesc = idesc; o.accumulate = first ? 0 : 1; first = 0; o.tmem_col = 0;
      cs.mmas.push_back(o);
    }
  std::vector<float> ref((size_t)128 * N, 0.f), out;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += bf(A[(size_t)m * K + k]) * bf(B[(size_t)n * K + k]);
      ref[(size_t)m * N + n] = s;
    }
  run_case(name, cs, out);
  int r = report(name, out, ref, 128, N, cs.ncols);
  cudaFree(dA); cudaFree(dB);
  return r;
}

// ----- shifted views: 3x3 (h,w) taps over one halo tile [18][10][C]; M tile = 16(h) x 8(w)
static int case_shift(const char* name, int C, int N, int h0, int w0) {
  const int sw = C * 2, Hg = 32, Wg = 24;
  const uint32_t pitch = sw;
  auto X = rand_bf16((size_t)Hg * Wg * C), W = rand_bf16((size_t)9 * N * C);
  auto dX = to_dev(X); auto dW = to_dev(W);
  Case cs; cs.ncols = (N + 31) / 32 * 32;
  uint64_t dimsX[3] = {(uint64_t)C, (uint64_t)Wg, (uint64_t)Hg}, strX[2] = {(uint64_t)C, (uint64_t)Wg * C};
  uint32_t boxX[3] = {(uint32_t)C, 10, 18};
  uint64_t dimsW[2] = {(uint64_t)C, (uint64_t)9 * N}, strW[1] = {(uint64_t)C};
  uint32_t boxW[2] = {(uint32_t)C, (uint32_t)N};
  cs.maps.m[0] = make_map(dX, 3, dimsX, strX, boxX, nullptr, sw);
  cs.maps.m[1] = make_map(dW, 2, dimsW, strW, boxW, nullptr, sw);
  const uint32_t a_bytes = 180 * pitch, b_base = (a_bytes + 1023) & ~1023u, b_bytes = N * pitch;
  cs.tmas.push_back({0, 3, 0, a_bytes, {0, w0 - 1, h0 - 1, 0, 0}});
  for (int t = 0; t < 9; ++t) cs.tmas.push_back({1, 2, b_base + t * b_bytes, b_bytes, {0, t * N, 0, 0, 0}});
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  int first = 1;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw)
      for (int k = 0; k < C / 16; ++k) {
        MmaOp o{};
        o.adesc = make_smem_desc((kh * 10 + kw) * pitch + k * 32, 16, 10 * pitch, swz_desc(sw));
        o.bdesc = make_smem_desc(b_base + (kh * 3 + kw) * b_bytes + k * 32, 16, 8 * pitch, swz_desc(sw));
        o.idesc = idesc; o.accumulate = first ? 0 : 1; first = 0;
        cs.mmas.push_back(o);
      }
  std::vector<float> ref((size_t)128 * N, 0.f), out;
  for (int h = 0; h < 16; ++h)
    for (int w = 0; w < 8; ++w)
      for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            int hh = h0 + h + kh - 1, ww = w0 + w + kw - 1;
            if (hh < 0 || hh >= Hg || ww < 0 || ww >= Wg) continue;
            for (int c = 0; c < C; ++c)
              s += bf(X[((size_t)hh * Wg + ww) * C + c]) * bf(W[((size_t)(kh * 3 + kw) * N + n) * C + c]);
          }
        ref[(size_t)(h * 8 + w) * N + n] = s;
      }
  run_case(name, cs, out);
  int r = report(name, out, ref, 128, N, cs.ncols);
  cudaFree(dX); cudaFree(dW);
  return r;
}

// ----- MN-major (wgrad): D[(atom, c), n] = sum_{h,w} X[h+kh][w+kw0+atom][c] * dY[h][w][n]
//       A atoms (one per kw tap) are `lbo_rows` rows apart in the halo tile; K groups = h rows.
static int case_wgrad(const char* name, int C, int N, int kh, int h0, int w0) {
  const int swA = C * 2, swB = N * 2, Hg = 32, Wg = 24;
  const int natom = 128 / C;  // atoms along M
  auto X = rand_bf16((size_t)Hg * Wg * C), dY = rand_bf16((size_t)128 * N);
  auto dX = to_dev(X); auto dD = to_dev(dY);
  Case cs; cs.ncols = (N + 31) / 32 * 32;
  uint64_t dimsX[3] = {(uint64_t)C, (uint64_t)Wg, (uint64_t)Hg}, strX[2] = {(uint64_t)C, (uint64_t)Wg * C};
  uint32_t boxX[3] = {(uint32_t)C, 10, 18};
  uint64_t dimsY[2] = {(uint64_t)N, 128}, strY[1] = {(uint64_t)N};
  uint32_t boxY[2] = {(uint32_t)N, 128};
  cs.maps.m[0] = make_map(dX, 3, dimsX, strX, boxX, nullptr, swA);
  cs.maps.m[1] = make_map(dD, 2, dimsY, strY, boxY, nullptr, swB);
  const uint32_t a_bytes = 180 * swA, b_base = (a_bytes + 1023) & ~1023u, b_bytes = 128 * swB;
  cs.tmas.push_back({0, 3, 0, a_bytes, {0, w0 - 1, h0 - 1, 0, 0}});
  cs.tmas.push_back({1, 2, b_base, b_bytes, {0, 0, 0, 0, 0}});
  const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
  for (int ks = 0; ks < 8; ++ks) {  // 16 voxels per MMA = two h rows of 8 w
    MmaOp o{};
    o.adesc = make_smem_desc((kh * 10 + 0) * swA + ks * 2 * 10 * swA, /*lbo=*/swA, /*sbo=*/10 * swA, swz_desc(swA));
    o.bdesc = make_smem_desc(b_base + ks * 2 * 8 * swB, /*lbo=*/swB, /*sbo=*/8 * swB, swz_desc(swB));
    o.idesc = idesc; o.accumulate = ks ? 1 : 0;
    cs.mmas.push_back(o);
  }
  std::vector<float> ref((size_t)128 * N, 0.f), out;
  for (int a = 0; a < natom; ++a)
    for (int c = 0; c < C; ++c)
      for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int h = 0; h < 16; ++h)
          for (int w = 0; w < 8; ++w) {
            int hh = h0 + h + kh - 1, ww = w0 + w + a - 1;
            // atom index a plays the role of kw; a==3 is "garbage" (beyond the 3 taps) but still
            // well-defined data of the halo tile as long as it stays in range; compare anyway.
            float xv = 0;
            int r_h = h + kh, r_w = w + a;           // position inside halo tile
            int lin = r_h * 10 + r_w;                // row index in smem
            int th = lin / 10, tw = lin % 10;        // (wraps into the next h row when r_w >= 10)
            hh = h0 - 1 + th; ww = w0 - 1 + tw;
            if (th < 18 && hh >= 0 && hh < Hg && ww >= 0 && ww < Wg) xv = bf(X[((size_t)hh * Wg + ww) * C + c]);
            s += xv * bf(dY[(size_t)(h * 8 + w) * N + n]);
          }
        ref[(size_t)(a * C + c) * N + n] = s;
      }
  run_case(name, cs, out);
  int r = report(name, out, ref, 128, N, cs.ncols);
  cudaFree(dX); cudaFree(dD);
  return r;
}

// ----- TMA elementStrides = 2 gather: load X[2*h+1][2*w+1][c] for an 8x8 tile
static int case_tma_stride(const char* name) {
  const int C = 32, Hg = 20, Wg = 20;
  auto X = rand_bf16((size_t)Hg * Wg * C, -100, 100);
  auto dX = to_dev(X);
  Case cs; cs.ncols = 32;
  uint64_t dimsX[3] = {(uint64_t)C, (uint64_t)Wg, (uint64_t)Hg}, strX[2] = {(uint64_t)C, (uint64_t)Wg * C};
  uint32_t boxX[3] = {(uint32_t)C, 16, 16}, es[3] = {1, 2, 2};
  cs.maps.m[0] = make_map(dX, 3, dimsX, strX, boxX, es, 0);
  cs.maps.m[1] = cs.maps.m[0];
  cs.tmas.push_back({0, 3, 0, 8 * 8 * C * 2, {0, -1, -1, 0, 0}});  // starts at (-1,-1): rows -1,1,3,...
  cs.dump_bytes = 8 * 8 * C * 2;
  std::vector<float> out; std::vector<uint8_t> dump;
  run_case(name, cs, out, &dump);
  const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(dump.data());
  int bad = 0;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j)
      for (int c = 0; c < C; ++c) {
        int hh = -1 + 2 * i, ww = -1 + 2 * j;
        float e = (hh < 0 || ww < 0) ? 0.f : bf(X[((size_t)hh * Wg + ww) * C + c]);
        if (bf(d[((size_t)i * 8 + j) * C + c]) != e) ++bad;
      }
  printf("[%s] %s  mismatches=%d/%d\n", name, bad ? "FAIL" : "PASS", bad, 8 * 8 * C);
  cudaFree(dX);
  return bad ? 1 : 0;
}

// ----- mixed operand formats (kind::f16 with a_format != b_format): A fp16 x B bf16 and the other combinations.
//       fmt: 0 = fp16, 1 = bf16 (instruction-descriptor encoding). Values are multiples of 0.25 in [-3, 3]:
//       exact in both formats, and a wrong interpretation of the bits changes them by orders of magnitude.
static std::vector<uint16_t> rand_16(size_t n, int fmt) {
  std::vector<uint16_t> v(n);
  for (size_t i = 0; i < n; ++i) {
    const float f = 0.25f * rnd_int(-12, 12);
    if (fmt == 0) { __half h = __float2half(f); v[i] = *reinterpret_cast<uint16_t*>(&h); }
    else { __nv_bfloat16 b = __float2bfloat16(f); v[i] = *reinterpret_cast<uint16_t*>(&b); }
  }
  return v;
}
static float dec16(uint16_t x, int fmt) {
  if (fmt == 0) { __half h = *reinterpret_cast<__half*>(&x); return __half2float(h); }
  __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&x); return __bfloat162float(b);
}
static uint32_t idesc_fmt(uint32_t M, uint32_t N, int afmt, int bfmt, uint32_t a_mn, uint32_t b_mn) {
  uint32_t d = make_idesc_bf16(M, N, a_mn, b_mn);
  d &= ~((7u << 7) | (7u << 10));
  d |= (uint32_t)afmt << 7;
  d |= (uint32_t)bfmt << 10;
  return d;
}
static int case_kmajor_mixed(const char* name, int kc, int N, int K, int afmt, int bfmt) {
  const int sw = kc * 2;
  auto A = rand_16((size_t)128 * K, afmt), B = rand_16((size_t)N * K, bfmt);
  auto dA = to_dev(A); auto dB = to_dev(B);
  Case cs; cs.ncols = (N + 31) / 32 * 32;
  uint64_t dimsA[2] = {(uint64_t)K, 128}, strA[1] = {(uint64_t)K};
  uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)K};
  uint32_t boxA[2] = {(uint32_t)kc, 128}, boxB[2] = {(uint32_t)kc, (uint32_t)N};
  cs.maps.m[0] = make_map(dA, 2, dimsA, strA, boxA, nullptr, sw);
  cs.maps.m[1] = make_map(dB, 2, dimsB, strB, boxB, nullptr, sw);
  const int nchunk = K / kc;
  const uint32_t a_bytes = 128 * sw, b_bytes = N * sw;
  const uint32_t b_base = ((nchunk * a_bytes) + 1023) & ~1023u;
  for (int c = 0; c < nchunk; ++c) {
    cs.tmas.push_back({0, 2, c * a_bytes, a_bytes, {c * kc, 0, 0, 0, 0}});
    cs.tmas.push_back({1, 2, b_base + c * b_bytes, b_bytes, {c * kc, 0, 0, 0, 0}});
  }
  const uint32_t idesc = idesc_fmt(128, N, afmt, bfmt, 0, 0);
  int first = 1;
  for (int c = 0; c < nchunk; ++c)
    for (int k = 0; k < kc / 16; ++k) {
      MmaOp o{};
      o.adesc = make_smem_desc(c * a_bytes + k * 32, 16, 8 * sw, swz_desc(sw));
      o.bdesc = make_smem_desc(b_base + c * b_bytes + k * 32, 16, 8 * sw, swz_desc(sw));
      o.idesc = idesc; o.accumulate = first ? 0 : 1; first = 0; o.tmem_col = 0;
      cs.mmas.push_back(o);
    }
  std::vector<float> ref((size_t)128 * N, 0.f), out;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += dec16(A[(size_t)m * K + k], afmt) * dec16(B[(size_t)n * K + k], bfmt);
      ref[(size_t)m * N + n] = s;
    }
  run_case(name, cs, out);
  int r = report(name, out, ref, 128, N, cs.ncols);
  cudaFree(dA); cudaFree(dB);
  return r;
}
// MN-major mixed (the wgrad operand arrangement: A = activations fp16, B = gradients bf16), no halo: plain
// D[c][n] = sum_v X[v][c] * dY[v][n] over 128 voxels; rows 0..63 of D (atom 0) are compared
static int case_mnmajor_mixed(const char* name, int afmt, int bfmt) {
  const int C = 64, N = 32, V = 128;
  const int swA = C * 2, swB = N * 2;
  auto X = rand_16((size_t)V * C, afmt), dY = rand_16((size_t)V * N, bfmt);
  auto dX = to_dev(X); auto dD = to_dev(dY);
  Case cs; cs.ncols = 32;
  uint64_t dimsX[2] = {(uint64_t)C, (uint64_t)V}, strX[1] = {(uint64_t)C};
  uint64_t dimsY[2] = {(uint64_t)N, (uint64_t)V}, strY[1] = {(uint64_t)N};
  uint32_t boxX[2] = {(uint32_t)C, (uint32_t)V}, boxY[2] = {(uint32_t)N, (uint32_t)V};
  cs.maps.m[0] = make_map(dX, 2, dimsX, strX, boxX, nullptr, swA);
  cs.maps.m[1] = make_map(dD, 2, dimsY, strY, boxY, nullptr, swB);
  const uint32_t a_bytes = V * swA, b_base = (a_bytes + 1023) & ~1023u;
  cs.tmas.push_back({0, 2, 0, a_bytes, {0, 0, 0, 0, 0}});
  cs.tmas.push_back({1, 2, b_base, (uint32_t)(V * swB), {0, 0, 0, 0, 0}});
  const uint32_t idesc = idesc_fmt(128, N, afmt, bfmt, 1, 1);   // atom 0 = the 64 channels; atom 1 (one voxel row later) is not compared
  for (int ks = 0; ks < V / 16; ++ks) {
    MmaOp o{};
    o.adesc = make_smem_desc(ks * 16 * swA, /*lbo=*/swA, /*sbo=*/8 * swA, swz_desc(swA));
    o.bdesc = make_smem_desc(b_base + ks * 16 * swB, /*lbo=*/swB, /*sbo=*/8 * swB, swz_desc(swB));
    o.idesc = idesc; o.accumulate = ks ? 1 : 0;
    cs.mmas.push_back(o);
  }
  std::vector<float> ref((size_t)64 * N, 0.f), out;
  for (int c = 0; c < C; ++c)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int v = 0; v < V; ++v) s += dec16(X[(size_t)v * C + c], afmt) * dec16(dY[(size_t)v * N + n], bfmt);
      ref[(size_t)c * N + n] = s;
    }
  run_case(name, cs, out);
  int r = report(name, out, ref, 64, N, cs.ncols);
  cudaFree(dX); cudaFree(dD);
  return r;
}

int main(int argc, char** argv) {
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  g_encode = (EncodeTiledFn)fn;
  if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
  const char* only = argc > 1 ? argv[1] : nullptr;
  int fails = 0;
#define RUN(tag, call) if (!only || !strcmp(only, tag)) fails += (call)
  RUN("K1", case_kmajor("K1 kmajor sw128 N64 K128", 64, 64, 128));
  RUN("K2", case_kmajor("K2 kmajor sw64  N32 K64", 32, 32, 64));
  RUN("K3", case_kmajor("K3 kmajor sw32  N96 K48", 16, 96, 48));
  RUN("K4", case_kmajor("K4 kmajor sw128 N256 K64", 64, 256, 64));
  RUN("S1", case_shift("S1 shift sw128 C64 N32 interior", 64, 32, 8, 8));
  RUN("S2", case_shift("S2 shift sw128 C64 N32 corner0", 64, 32, 0, 0));
  RUN("S3", case_shift("S3 shift sw64  C32 N32 far edge", 32, 32, 16, 16));
  RUN("S4", case_shift("S4 shift sw32  C16 N64 interior", 16, 64, 8, 8));
  RUN("W1", case_wgrad("W1 wgrad sw128 C64 N32 kh1", 64, 32, 1, 8, 8));
  RUN("W2", case_wgrad("W2 wgrad sw64  C32 N32 kh0 (4 atoms)", 32, 32, 0, 8, 8));
  RUN("W3", case_wgrad("W3 wgrad sw128 C64 N64 kh2 edge", 64, 64, 2, 0, 0));
  RUN("T1", case_tma_stride("T1 tma elementStrides=2"));
  RUN("M0", case_kmajor_mixed("M0 kmajor A bf16 x B bf16 (control)", 32, 96, 64, 1, 1));
  RUN("M1", case_kmajor_mixed("M1 kmajor A fp16 x B bf16", 32, 96, 64, 0, 1));
  RUN("M2", case_kmajor_mixed("M2 kmajor A bf16 x B fp16", 32, 96, 64, 1, 0));
  RUN("M3", case_kmajor_mixed("M3 kmajor A fp16 x B fp16", 32, 96, 64, 0, 0));
  RUN("M4", case_mnmajor_mixed("M4 mn-major A fp16 x B bf16", 0, 1));
  printf("probe done: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
