// stem_igemm.cu — model.0: Conv(3 -> Cout, k3, s2, p1) + folded BN + SiLU on the tensor cores, fused with the predictor's
// uint8/fp32 NCHW -> bf16 NHWC conversion (ultralytics/nn/modules/conv.py:49-55; engine/predictor.py:127-135).
//
// The CUDA-core stem (aux_kernels.cu) was FP32-issue bound at 5-6x its HBM roofline.  Here the 27-deep contraction is
// padded to K = 32 and runs as an M128 x N(Cout) x K32 tcgen05 GEMM per 32x4-pixel output tile:
//   * warps 0 and 2 (one thread each, alternate tiles): TMA brings the [3 ch][9 rows][65+ cols] input patch of the tile (NCHW planes, element type u8 or
//     fp32; out-of-image coordinates are zero-filled = the conv padding);
//   * warps 4..7 / 8..11 (two groups of 128 "builder" threads on alternate tiles, one output pixel each): im2col row of 27
//     values from the patch -> bf16 -> one
//     64-byte row of the A tile in the 64B-swizzled K-major layout (uint8 0..255 is exact in bf16; 1/255 is applied in fp32
//     in the epilogue); fence.proxy.async; arrive on the stage's barrier;
//   * warps 1 and 3 (one thread each, alternate tiles): two K16 tcgen05.mma per tile into a TMEM accumulator stage;
//   * warps 12..15 / 16..19 / 20..23: three epilogue groups on every third tile: tcgen05.ld -> scale, bias, SiLU -> bf16 -> 64B-swizzled
//     staging tile -> TMA store into the NHWC output slice.
// Persistent, one CTA per SM.
#include "dy_common.cuh"
#include "dy_ptx.cuh"
#include "conv_igemm.h"
#include <cstdlib>

namespace dy {

using namespace ptx;

static constexpr int kStemTW = 32, kStemTH = 4;                 // output pixels per tile (128 = UMMA M)
static constexpr int kStemPH = 2 * kStemTH + 1;                 // input rows per patch
// Input columns per patch row.  TMA needs the box to START on a 16-byte boundary in global memory, so the patch begins
// 16 bytes (16 uint8 / 4 fp32 columns) left of the tile instead of 1 column: 16 + 64 + 1 (4 + 64 + 1) columns, padded to 16 B.
static constexpr int kStemPWu8 = 96, kStemPWf32 = 72;
static constexpr int kStemLeftU8 = 16, kStemLeftF32 = 4;
#ifndef DY_STEM_EG
#define DY_STEM_EG 3
#endif
static constexpr int kStemEG = DY_STEM_EG;                                // epilogue groups (the in-kernel timeline shows the epilogue chain, ~2100 cycles per tile and group, as the limiter)
static constexpr int kStemMaxNP = 16, kStemNA = 4, kStemNAcc = 2 * kStemEG;   // patch stages (HBM latency x bandwidth: >= 12 tiles in flight per SM), A stages, accumulator stages (even: two MMA issuers; multiple of kStemEG)
static constexpr int kStemThreads = 128 + 256 + 128 * kStemEG;  // 4 control warps + 2 x 4 builder warps + kStemEG x 4 epilogue warps
static constexpr int kStemABytes = 128 * 64;                    // A tile: 128 rows x 32 bf16
static constexpr int kStemMaxN = ((512 - 16) / kStemNAcc) / 16 * 16;                            // kStemNAcc * N + 16 columns of ragged read must fit 512

struct StemParams {
  CUtensorMap tmIn;        // input planes [B*3][H][W]
  CUtensorMap tmO;         // bf16 NHWC output slice, box {32 ch, 32, 4, 1}
  const float* weight;     // fp32 [Cout][27]
  const float* bias;       // fp32 [Cout]
  float in_scale;          // 1/255 for uint8 input, 1 for fp32
  int Cout, N;             // N = Cout padded to 16
  int tiles_w, tiles_h, B, total_tiles;
  int patch_bytes, np;     // bytes per patch, patch stages
  int dbg;                 // DY_STEM_DBG knock-outs (bring-up): 1 no TMA store, 2 no MMA, 4 no TMA load, 8 no TMEM load
  unsigned long long* trace;   // debug builds only (DY_CONV_TRACE): clock64 stamps of CTA 0, [role 8][iteration 96][event 8]
};

#ifdef DY_CONV_DEBUG
#define ST_TR(role, it, ev) do { if (p.trace && blockIdx.x == 0 && (it) < 96) p.trace[((role) * 96 + (it)) * 8 + (ev)] = clock64(); } while (0)
#else
#define ST_TR(role, it, ev) do { } while (0)
#endif

template <bool U8>
__global__ void __launch_bounds__(kStemThreads, 1) stem_igemm_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t pfull[kStemMaxNP], pempty[kStemMaxNP], afull[kStemNA], aempty[kStemNA], tfull[kStemNAcc], tempty[kStemNAcc];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[kStemMaxN + 32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  constexpr int PW = U8 ? kStemPWu8 : kStemPWf32;
  constexpr int PROW = U8 ? kStemPWu8 : kStemPWf32 * 4;         // bytes per patch row
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t b_tile = smem_base;                                            // weights: N rows x 64 B (<= 8 KB)
  const uint32_t a_tile0 = smem_base + 8192;                                    // kStemNA x 8 KB
  const uint32_t stg0 = a_tile0 + kStemNA * kStemABytes;                        // kStemEG groups x 2 x 8 KB
  const uint32_t patch0 = stg0 + 2 * kStemEG * 8192;                                      // np x patch_bytes (1 KB multiples)
  const uint32_t patch_stride = static_cast<uint32_t>((p.patch_bytes + 1023) & ~1023);

  if (warp == 0 && elect_one()) { prefetch_tmap(&p.tmIn); prefetch_tmap(&p.tmO); }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < p.np; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], 128); }
    for (int i = 0; i < kStemNA; ++i) { mbar_init(&afull[i], 128); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < kStemNAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  // weights -> bf16 [N][32] in the 64B-swizzled K-major layout (rows >= Cout and k >= 27 are zero)
  for (int i = threadIdx.x; i < N * 4; i += kStemThreads) {
    const int n = i >> 2, j = i & 3;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = 8 * j + 2 * e;
      const float a = (n < p.Cout && k0 < 27) ? __ldg(p.weight + n * 27 + k0) : 0.f;
      const float b = (n < p.Cout && k0 + 1 < 27) ? __ldg(p.weight + n * 27 + k0 + 1) : 0.f;
      w[e] = pack_bf16(a, b);
    }
    sts128(b_tile + n * 64 + ((j ^ ((n >> 1) & 3)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
  }
  for (int i = threadIdx.x; i < kStemMaxN + 32; i += kStemThreads) s_bias[i] = i < p.Cout ? 0.5f * __ldg(p.bias + i) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  grid_dep_wait();
  grid_dep_launch();

  const int per = p.total_tiles / static_cast<int>(gridDim.x), rem = p.total_tiles % static_cast<int>(gridDim.x);
  const int tile_begin = static_cast<int>(blockIdx.x) * per + min(static_cast<int>(blockIdx.x), rem);
  const int my_tiles = per + (static_cast<int>(blockIdx.x) < rem ? 1 : 0);

  if (warp == 0 || warp == 2) {
    // ===================== patch producers: two threads (warps 0 and 2) on alternate tiles.  One thread's per-tile chain
    // (wait for the free slot, expect_tx, TMA issue, index arithmetic) is ~450 cycles, more than a tile is allowed to cost;
    // np is even, so every slot always belongs to the same producer =====================
    if (elect_one()) {
      const int pi = warp == 0 ? 0 : 1;
      const int t0 = tile_begin + pi;
      int tw = t0 % p.tiles_w, th = (t0 / p.tiles_w) % p.tiles_h, tb = t0 / (p.tiles_w * p.tiles_h);
      int s = pi; uint32_t ph = 0;
      const uint32_t pf0 = smem_u32(&pfull[0]), pe0 = smem_u32(&pempty[0]);
      const uint32_t tx = (p.dbg & 4) ? 0u : static_cast<uint32_t>(p.patch_bytes);
      const int np = p.np, tiles_w = p.tiles_w, tiles_h = p.tiles_h;
      bool ok = mbar_try_wait_a(pe0 + s * 8, 1u);
      for (int i = pi; i < my_tiles; i += 2) {
        ST_TR(pi, i >> 1, 0);
        if (!ok) mbar_wait_a(pe0 + s * 8, ph ^ 1u);
        ST_TR(pi, i >> 1, 1);
        mbar_arrive_expect_tx_a(pf0 + s * 8, tx);
        int ns = s + 2; uint32_t nph = ph;
        if (ns >= np) { ns -= np; nph ^= 1u; }
        ok = mbar_try_wait_a(pe0 + ns * 8, nph ^ 1u);          // wait-ahead: resolved while the TMA issues
        if (!(p.dbg & 4)) asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(patch0 + s * patch_stride), "l"(reinterpret_cast<uint64_t>(&p.tmIn)), "r"(pf0 + s * 8),
              "r"(2 * tw * kStemTW - (U8 ? kStemLeftU8 : kStemLeftF32)), "r"(2 * th * kStemTH - 1), "r"(3 * tb) : "memory");
        ST_TR(pi, i >> 1, 2);
        for (int k = 0; k < 2; ++k)                             // advance two tiles
          if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; ++tb; } }
        s = ns; ph = nph;
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: two threads on alternate tiles (a tile is only two MMAs: the issuing thread's
    // barrier waits and commits, ~700 cycles per tile, are what has to be hidden) =====================
    if (elect_one()) {
      const int mi = warp == 1 ? 0 : 1;
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const uint64_t hi = umma_desc_kmajor(0, 512, 4u) & 0xffffffff00000000ull;       // 64B swizzle, 8-row groups 512 B apart
      const uint32_t lo_const = static_cast<uint32_t>(umma_desc_kmajor(0, 512, 4u) & 0xffffffffull);
      const uint32_t b_lo = lo_const | ((b_tile & 0x3ffffu) >> 4);
      for (int i = mi; i < my_tiles; i += 2) {
        const int s = i % kStemNA, a = i % kStemNAcc;
        ST_TR(2 + mi, i >> 1, 0);
        mbar_wait(&tempty[a], ((i / kStemNAcc) & 1) ^ 1);
        ST_TR(2 + mi, i >> 1, 1);
        mbar_wait(&afull[s], (i / kStemNA) & 1);
        ST_TR(2 + mi, i >> 1, 2);
        tc_fence_after();
        const uint32_t a_lo = lo_const | (((a_tile0 + s * kStemABytes) & 0x3ffffu) >> 4);
        const uint32_t d = tmem_base + static_cast<uint32_t>(a * N);
        if (!(p.dbg & 2)) {
          umma_bf16_ss(d, hi | a_lo, hi | b_lo, idesc, 0u);
          umma_bf16_ss(d, hi | (a_lo + 2), hi | (b_lo + 2), idesc, 1u);
        }
        umma_commit(&aempty[s]);
        umma_commit(&tfull[a]);
        ST_TR(2 + mi, i >> 1, 3);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== builders: one output pixel (A row) per thread, two groups on alternate tiles =====================
    const int bg = (warp - 4) >> 2;
    const int m = (threadIdx.x - 128) & 127;
    const int py = m >> 5, px = m & 31;
    const uint32_t row_off = static_cast<uint32_t>(m) * 64, sw = static_cast<uint32_t>((m >> 1) & 3);
    for (int i = bg; i < my_tiles; i += 2) {
      const int ps = i % p.np, as = i % kStemNA;
      if (m == 0) ST_TR(4 + bg, i >> 1, 0);
      mbar_wait(&pfull[ps], (i / p.np) & 1);
      if (m == 0) ST_TR(4 + bg, i >> 1, 1);
      // first needed column of this pixel = left margin - 1 + 2*px; the uint8 path reads aligned 16-bit pairs from one byte earlier
      const uint32_t pb = patch0 + ps * patch_stride + static_cast<uint32_t>(2 * py) * PROW +
                          (U8 ? static_cast<uint32_t>(kStemLeftU8 - 2 + 2 * px) : static_cast<uint32_t>(kStemLeftF32 - 1 + 2 * px) * 4u);
      uint32_t hv[27];                      // bf16 bits (upper half of the fp32 pattern) of the 27 patch values, k = (c*3 + ky)*3 + kx
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t ra = pb + static_cast<uint32_t>(c * kStemPH + ky) * PROW;
          const int k = (c * 3 + ky) * 3;
          if constexpr (U8) {
            uint32_t lo16, hi16;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo16) : "r"(ra));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(hi16) : "r"(ra + 2));
            // integer 0..255 -> fp32 via the 2^23 magic number (exact), whose upper 16 bits are the exact bf16
            hv[k + 0] = __float_as_uint(__uint_as_float(0x4b000000u | (lo16 >> 8)) - 8388608.f);      // bytes: [x-2, x-1] [x, x+1]
            hv[k + 1] = __float_as_uint(__uint_as_float(0x4b000000u | (hi16 & 0xffu)) - 8388608.f);
            hv[k + 2] = __float_as_uint(__uint_as_float(0x4b000000u | (hi16 >> 8)) - 8388608.f);
          } else {
            float f0, f1, f2;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f0) : "r"(ra));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f1) : "r"(ra + 4));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f2) : "r"(ra + 8));
            hv[k + 0] = __float_as_uint(f0); hv[k + 1] = __float_as_uint(f1); hv[k + 2] = __float_as_uint(f2);
          }
        }
      if (m == 0) ST_TR(4 + bg, i >> 1, 2);
      mbar_wait(&aempty[as], ((i / kStemNA) & 1) ^ 1);
      if (m == 0) ST_TR(4 + bg, i >> 1, 3);
      const uint32_t arow = a_tile0 + as * kStemABytes + row_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = 8 * j + 2 * e;
          if constexpr (U8) {
            const uint32_t a = k0 < 27 ? hv[k0 < 27 ? k0 : 0] : 0u, b = k0 + 1 < 27 ? hv[k0 + 1 < 27 ? k0 + 1 : 0] : 0u;
            w[e] = __byte_perm(a, b, 0x7632);                                   // upper halves: a -> low 16 bits, b -> high
          } else {
            const float a = k0 < 27 ? __uint_as_float(hv[k0 < 27 ? k0 : 0]) : 0.f, b = k0 + 1 < 27 ? __uint_as_float(hv[k0 + 1 < 27 ? k0 + 1 : 0]) : 0.f;
            w[e] = pack_bf16(a, b);
          }
        }
        sts128(arow + ((static_cast<uint32_t>(j) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
      }
      fence_proxy_async_smem();
      mbar_arrive(&afull[as]);
      mbar_arrive(&pempty[ps]);
      if (m == 0) ST_TR(4 + bg, i >> 1, 4);
    }
  } else if (warp >= 12) {
    // ===================== epilogue groups =====================
    const int ew = warp - 12, g = ew >> 2, q = ew & 3;
    const int row = q * 32 + lane;
    const bool leader = (threadIdx.x == 384 + g * 128);
    const uint32_t stg = stg0 + static_cast<uint32_t>(g) * 2u * 8192u;
    const float hscale = 0.5f * p.in_scale;               // SiLU: h = 0.5*(acc*scale + bias)
    const int nchunks = (N + 31) >> 5;
    uint32_t sctr = 0;
    // Flattened (tile, 32-column chunk) items.  The TMEM load of the NEXT item is issued as soon as this item's values are
    // packed, so its latency (~600 cycles with eight epilogue warps contending) hides behind the fence / barrier / store
    // hand-off instead of heading every item; tile coordinates advance by carry (three divisions per tile were ~150 cycles).
    int tw, th, tb;
    { const int t0 = min(tile_begin + g, p.total_tiles - 1); tw = t0 % p.tiles_w; th = (t0 / p.tiles_w) % p.tiles_h; tb = t0 / (p.tiles_w * p.tiles_h); }
    const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t r[32];
    int i = g, c = 0;
    bool valid = i < my_tiles;
    if (valid) {
      mbar_wait(&tfull[i % kStemNAcc], (i / kStemNAcc) & 1);
      tc_fence_after();
      if (!(p.dbg & 8)) tmem_ld_32x32b_x32(tlane + static_cast<uint32_t>((i % kStemNAcc) * N), r);
    }
    while (valid) {
      const int a = i % kStemNAcc;
      const uint32_t st = stg + (sctr & 1) * 8192u;
      if (leader) ST_TR(6 + g, i / kStemEG, 1);
      tmem_ld_wait();
      if (leader) ST_TR(6 + g, i / kStemEG, 2);
      if (c == nchunks - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
      }
      const uint32_t rowp = st + row * 64;
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        float v[8];
#pragma unroll
        const float4 hb0 = *reinterpret_cast<const float4*>(s_bias + 32 * c + 8 * gi);   // 80-register budget: bias per 8 columns
        const float4 hb1 = *reinterpret_cast<const float4*>(s_bias + 32 * c + 8 * gi + 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 b4 = (e >> 2) ? hb1 : hb0;
          const float bb = (e & 3) == 0 ? b4.x : (e & 3) == 1 ? b4.y : (e & 3) == 2 ? b4.z : b4.w;
          const float h = fmaf(__uint_as_float(r[8 * gi + e]), hscale, bb);
          v[e] = fmaf(h, tanh_fast(h), h);
        }
        sts128(rowp + ((static_cast<uint32_t>(gi) ^ static_cast<uint32_t>((row >> 1) & 3)) << 4),
               make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])));
      }
      if (leader) ST_TR(6 + g, i / kStemEG, 3);
      int ni = i, nc = c + 1;
      if (nc == nchunks) { nc = 0; ni = i + kStemEG; }
      const bool nvalid = ni < my_tiles;
      if (nvalid) {
        const int na = ni % kStemNAcc;
        if (nc == 0) { mbar_wait(&tfull[na], (ni / kStemNAcc) & 1); tc_fence_after(); }
        if (!(p.dbg & 8)) tmem_ld_32x32b_x32(tlane + static_cast<uint32_t>(na * N + 32 * nc), r);
      }
      if (leader) ST_TR(6 + g, i / kStemEG, 4);
      fence_proxy_async_smem();
      if (leader) bulk_wait_group_read<0>();
      named_bar_sync(1 + g, 128);
      if (leader) ST_TR(6 + g, i / kStemEG, 5);
      if (leader && !(p.dbg & 1)) {
        tma_store_4d_a(&p.tmO, st, 32 * c, tw * kStemTW, th * kStemTH, tb);
        bulk_commit_group();
      }
      if (leader) ST_TR(6 + g, i / kStemEG, 6);
      ++sctr;
      if (ni != i)
        for (int k = 0; k < kStemEG; ++k)
          if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; ++tb; } }
      i = ni; c = nc; valid = nvalid;
    }
    if (leader) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Tensor-core stem.  Returns DY_ERR_UNSUPPORTED for shapes it does not cover (the caller falls back to the CUDA-core kernel).
int stem_tc_launch(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout, void* out,
                   int out_ld, cudaStream_t stream) {
  const bool u8 = in_dtype == DY_U8;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if (Cout % 8 != 0 || round_up(Cout, 16) > kStemMaxN || (u8 ? W % 16 : W % 4) != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0)
    return DY_ERR_UNSUPPORTED;
  StemParams p{};
  p.weight = weight; p.bias = bias; p.in_scale = u8 ? 1.f / 255.f : 1.f;
  p.Cout = Cout; p.N = round_up(Cout, 16);
  p.tiles_w = ceil_div(Wo, kStemTW); p.tiles_h = ceil_div(Ho, kStemTH); p.B = B;
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * B;
  DY_CHECK_ARG(total < (1ll << 31), "stem: too many pixels");
  p.total_tiles = static_cast<int>(total);
  const int pw = u8 ? kStemPWu8 : kStemPWf32, esz = u8 ? 1 : 4;
  p.patch_bytes = 3 * kStemPH * pw * esz;
  { const char* e = getenv("DY_STEM_DBG"); p.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("DY_CONV_TRACE"); p.trace = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  {
    const uint64_t dims[3] = {uint64_t(W), uint64_t(H), uint64_t(B) * 3};
    const uint64_t strides[2] = {uint64_t(W) * esz, uint64_t(W) * H * esz};
    const uint32_t box[3] = {uint32_t(pw), uint32_t(kStemPH), 3};
    int rc = encode_map(&p.tmIn, in, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                        u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {uint64_t(Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(B)};
    const uint64_t strides[3] = {uint64_t(out_ld) * 2, uint64_t(Wo) * out_ld * 2, uint64_t(Ho) * Wo * out_ld * 2};
    const uint32_t box[4] = {32, kStemTW, kStemTH, 1};
    int rc = encode_map(&p.tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    if (rc) return rc;
  }
  p.np = u8 ? 16 : 12;                                       // even: the two patch producers own alternate slots
  const int smem = 1024 + 8192 + kStemNA * kStemABytes + 2 * kStemEG * 8192 + p.np * ((p.patch_bytes + 1023) & ~1023);
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  static unsigned long long seen[2] = {0, 0};                // the opt-in is per device
  if (first_use_on_device(&seen[u8 ? 1 : 0])) {
    if (u8) DY_CUDA(cudaFuncSetAttribute(stem_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    else DY_CUDA(cudaFuncSetAttribute(stem_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kStemThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = getenv("DY_NO_PDL") ? 0 : 1;
  if (u8) DY_CUDA(cudaLaunchKernelEx(&cfg, stem_igemm_kernel<true>, p));
  else DY_CUDA(cudaLaunchKernelEx(&cfg, stem_igemm_kernel<false>, p));
  return launch_status("stem_igemm_kernel");
}

}  // namespace dy
