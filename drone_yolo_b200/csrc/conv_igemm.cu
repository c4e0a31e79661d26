// conv_igemm.cu — Conv(k in {1,3}, stride in {1,2}) + bias + SiLU (+ residual) as an implicit GEMM on the
// sm_100a tensor cores.  Replaces the aten/cuDNN conv2d + BatchNorm + SiLU (+ add, + cat) kernel chain the
// reference launches for every Conv / RepVGGBlock / Bottleneck (ultralytics/nn/modules/conv.py:49-55,
// block.py:337-350,1480-1490).
//
// GEMM view: M = output pixels (B*Ho*Wo), N = Cout, K = k*k*Cin.  D[M,N] = A[M,K] * W[N,K]^T.
//   * A is never materialised: for every filter tap (r,s) and every 64-channel block the TMA engine loads a
//     [pixel tile x 64 ch] box of the NHWC activation, shifted by the tap offset; out-of-image coordinates are
//     zero-filled by TMA, which IS the conv padding.  Stride-2 convs read four "parity" views of the input
//     (even/odd rows x even/odd columns), so that every tap is again a dense box.
//   * W is packed [tap][Cout_pad][Cin_pad] (K contiguous) and loaded by a 3-D TMA box.
//   * Both operands land in shared memory in the 128B-swizzled K-major layout tcgen05.mma expects; one elected
//     thread issues M=128 x N=BN x K=16 MMAs accumulating fp32 in TMEM (double-buffered accumulator).
//   * 4 epilogue warps read the accumulator with tcgen05.ld, add the folded-BN bias, apply SiLU, add the optional
//     residual, and store bf16 (or fp32) straight into a channel slice of the destination buffer (concat-write).
//   * Persistent: grid = min(#tiles, #SMs); warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//     warps 4..7 = epilogue.
#include "dy_common.cuh"
#include "dy_ptx.cuh"
#include "conv_igemm.h"
#include <cstring>
#include <cstdlib>

namespace dy {

using namespace ptx;

static constexpr int kBlockM = 128;           // UMMA M
static constexpr int kBlockK = 64;            // bf16 per 128B swizzle row
static constexpr int kABytes = kBlockM * 128; // one A stage
static constexpr int kMaxStages = 8;
static constexpr int kThreads = 384;             // 4 control warps + 8 epilogue warps
static constexpr int kTmemCols = 512;
static constexpr int kMaxDynSmem = 227 * 1024 - 2048;
static constexpr int kHaloW = 16, kHaloH = 18;   // MODE 3 halo tile: 8x16 output pixels + 1-pixel border, row pitch padded to 16 pixels
static constexpr int kHaloBytes = kHaloW * kHaloH * 128;
static constexpr int kMaxAcc = 8;              // accumulator stages in TMEM: min(8, 512 / BN), BN columns apart

// Contiguous tile range per CTA: coordinates advance by carry instead of by integer division (the single-thread
// producer / MMA loops are latency-bound, a div/mod chain per tile was ~600 cycles of pure overhead).
struct TileIter {
  int n_tile, tw_i, th_i, tb_i, remaining;
  __device__ __forceinline__ TileIter(const ConvParams& p, int cta, int ncta) {
    const int total = p.m_tiles * p.n_tiles;
    const int base = total / ncta, rem = total % ncta;
    const int begin = cta * base + min(cta, rem);
    remaining = base + (cta < rem ? 1 : 0);
    n_tile = begin % p.n_tiles;
    int m = begin / p.n_tiles;
    tw_i = m % p.tiles_w; m /= p.tiles_w;
    th_i = m % p.tiles_h;
    tb_i = m / p.tiles_h;
  }
  __device__ __forceinline__ bool valid() const { return remaining > 0; }
  __device__ __forceinline__ void next(int n_tiles, int tiles_w, int tiles_h) {
    --remaining;
    if (++n_tile == n_tiles) {
      n_tile = 0;
      if (++tw_i == tiles_w) {
        tw_i = 0;
        if (++th_i == tiles_h) { th_i = 0; ++tb_i; }
      }
    }
  }
};

// MODE = tap geometry, known at compile time so the single-thread loops carry no table look-ups:
//   0: 1x1 (one tap), 1: 3x3 stride 1 (tap (r,c) shifts the box by (c-1, r-1)), 2: 3x3 stride 2 (four parity views),
//   3: 3x3 stride 1 "halo": Cin <= 64, weights resident; ONE TMA box per tile brings the 8x16-pixel tile plus its border
//      (16 x 18 pixels x 64 ch) and the nine taps are nine shifted UMMA descriptors over that single smem tile.
// CW = chunk width (channels) of the TMA-store epilogue: 64 or 32 bf16 (F32 = false), 32 fp32 (F32 = true): 128-byte or
// 64-byte staging rows.  CW = 0 selects the generic register->global epilogue (odd widths).
//
// Warp roles: 0 = A producer (TMA), 1 = MMA issuer, 2 = TMEM allocator, 3 = B producer (TMA), 4..11 = epilogue.
// Shared memory: [resident weights (b_resident)] [stages x (A 16 KB [+ B BN*128])] [2 x 16 KB output staging].
template <int MODE, int CW, bool F32>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[kMaxAcc];
  __shared__ __align__(8) uint64_t tempty_bar[kMaxAcc];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[256];

  constexpr int NTAPS = MODE == 0 ? 1 : 9;
  constexpr int LOADS = MODE == 3 ? 1 : NTAPS;    // TMA boxes per (tile, 64-channel block)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const bool bres = p.b_resident != 0;
  const uint32_t bres_bytes = bres ? static_cast<uint32_t>(NTAPS * p.kblocks) * b_bytes : 0u;
  const uint32_t stage_bytes = MODE == 3 ? static_cast<uint32_t>(kHaloBytes) : kABytes + (bres ? 0u : b_bytes);
  const int total_tiles = p.m_tiles * p.n_tiles;
#ifdef DY_CONV_DEBUG
  const int dbg = p.dbg;   // DY_CONV_DBG knock-outs for bottleneck hunting: 1 = no epilogue work, 2 = no MMA, 4 = no A loads, 8 = no B loads
#else
  constexpr int dbg = 0;   // knock-outs compile away unless built with -DDY_CONV_DEBUG
#endif
  // `opaque` pins loop invariants in registers: without it the compiler re-derives the shared-window addresses
  // (S2UR SR_CgaCtaId + ULEA) and re-reads kernel parameters inside the single-thread loops, whose cost is pure latency.
  const uint32_t smem_base = opaque(smem_u32(smem));
  const uint32_t stage0 = opaque(smem_base + bres_bytes);  // first pipeline stage (1024-aligned: bres_bytes is a multiple of 2048)
  const uint32_t full0 = opaque(smem_u32(&full_bar[0])), empty0 = opaque(smem_u32(&empty_bar[0]));
  const uint32_t tfull0 = opaque(smem_u32(&tfull_bar[0])), tempty0 = opaque(smem_u32(&tempty_bar[0]));
  const uint32_t bres_b = opaque(smem_u32(&bres_bar));
  const int nstages = opaque(p.stages), nacc = opaque(p.nacc);

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < p.nmaps; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
    if (CW > 0) prefetch_tmap(&p.tmO);
  }
  if (warp == 1 && elect_one()) {
    const uint32_t producers = bres ? 1u : 2u;             // A thread (+ B thread) arrive on every full barrier
    for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], producers); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < nacc; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    mbar_init(&bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on we touch
  // memory it produced.  Let our own dependents start their prologue right away.
  grid_dep_wait();
  grid_dep_launch();

  if (warp == 0) {
    // ===================== A producer: one TMA box per (tap, 64-channel block) =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t a_tx = (dbg & 4) ? 0u : static_cast<uint32_t>(p.TW * p.TH * p.TB) * 128u;
      const int kblocks = opaque(p.kblocks);
      const int n_tiles = opaque(p.n_tiles), tiles_w = opaque(p.tiles_w), tiles_h = opaque(p.tiles_h);
      const int TW = opaque(p.TW), TH = opaque(p.TH), TB = opaque(p.TB);
      for (TileIter it(p, blockIdx.x, gridDim.x); it.valid(); it.next(n_tiles, tiles_w, tiles_h)) {
        const int w0 = it.tw_i * TW, h0 = it.th_i * TH, b0 = it.tb_i * TB;
        if constexpr (MODE == 3) {
          const uint32_t fb = full0 + stage * 8;
          mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
          mbar_arrive_expect_tx_a(fb, (dbg & 4) ? 0u : static_cast<uint32_t>(kHaloBytes));
          if (!(dbg & 4)) tma_load_4d_a(stage0 + stage * stage_bytes, &p.tmA[0], fb, 0, w0 - 1, h0 - 1, b0);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
          continue;
        }
#pragma unroll
        for (int t = 0; t < LOADS; ++t) {
          // compile-time tap geometry
          const int oy = MODE == 0 ? 0 : t / 3 - 1, ox = MODE == 0 ? 0 : t % 3 - 1;
          const int map = MODE == 2 ? ((oy & 1) * 2 + (ox & 1)) : 0;
          const int dx = MODE == 2 ? (ox < 0 ? -1 : 0) : ox, dy = MODE == 2 ? (oy < 0 ? -1 : 0) : oy;
          const CUtensorMap* tm = &p.tmA[map];
          for (int kc = 0; kc < kblocks; ++kc) {
            const uint32_t fb = full0 + stage * 8;
            mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
            mbar_arrive_expect_tx_a(fb, a_tx);
            if (!(dbg & 4)) tma_load_4d_a(stage0 + stage * stage_bytes, tm, fb, kc * kBlockK, w0 + dx, h0 + dy, b0);
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: weights, either once (resident) or per stage =====================
    if (elect_one()) {
      const int kblocks = opaque(p.kblocks);
      if (bres) {
        mbar_arrive_expect_tx_a(bres_b, (dbg & 8) ? 0u : bres_bytes);
        if (!(dbg & 8)) {
          for (int t = 0; t < NTAPS; ++t)
            for (int kc = 0; kc < kblocks; ++kc)
              tma_load_3d_a(smem_base + static_cast<uint32_t>(t * kblocks + kc) * b_bytes, &p.tmB, bres_b, kc * kBlockK, 0, t);
        }
      } else {
        int stage = 0; uint32_t phase = 0;
        const uint32_t b_tx = (dbg & 8) ? 0u : b_bytes;
        const int n_tiles = opaque(p.n_tiles), tiles_w = opaque(p.tiles_w), tiles_h = opaque(p.tiles_h), BN = opaque(p.BN);
        for (TileIter it(p, blockIdx.x, gridDim.x); it.valid(); it.next(n_tiles, tiles_w, tiles_h)) {
          const int n0 = it.n_tile * BN;
#pragma unroll
          for (int t = 0; t < NTAPS; ++t) {
            for (int kc = 0; kc < kblocks; ++kc) {
              const uint32_t fb = full0 + stage * 8;
              mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
              mbar_arrive_expect_tx_a(fb, b_tx);
              if (!(dbg & 8)) tma_load_3d_a(stage0 + stage * stage_bytes + kABytes, &p.tmB, fb, kc * kBlockK, n0, t);
              if (++stage == nstages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(kBlockM, p.BN);
      const int kiters = opaque(NTAPS * p.kblocks);
      const uint32_t sbytes = opaque(stage_bytes), bbytes = opaque(b_bytes);
      // descriptor = constant high word | (address >> 4): only the low word changes
      const uint64_t desc_hi = umma_desc_sw128(0, 1024) & 0xffffffff00000000ull;
      const uint32_t desc_lo_const = static_cast<uint32_t>(umma_desc_sw128(0, 1024) & 0xffffffffull);   // LBO field
      const uint64_t halo_hi = umma_desc_sw128(0, kHaloW * 128) & 0xffffffff00000000ull;               // SBO = one halo row of 16 pixels
      if (bres) mbar_wait_a(bres_b, 0);
      const uint32_t BNu = opaque(static_cast<uint32_t>(p.BN));
      int my_tiles = total_tiles / static_cast<int>(gridDim.x) + (static_cast<int>(blockIdx.x) < total_tiles % static_cast<int>(gridDim.x) ? 1 : 0);
      for (; my_tiles > 0; --my_tiles) {
        mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * BNu;
        if constexpr (MODE == 3) {
          mbar_wait_a(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_lo = desc_lo_const | (((stage0 + stage * sbytes) & 0x3ffffu) >> 4);
          const uint32_t b_lo = desc_lo_const | ((smem_base & 0x3ffffu) >> 4);
          const uint32_t bstep = bbytes >> 4;
          if (!(dbg & 2)) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              // tap (r, c): rows of the halo tile start (r*16 + c) pixels in; 8-pixel row groups are 16 pixels (2048 B) apart.
              // The start is c rows into a 1024-byte swizzle atom.  Measured on B200: tcgen05 applies the 128B-swizzle XOR to
              // ABSOLUTE shared-memory address bits (like TMA), so a row-shifted start needs no base-offset (setting the
              // descriptor's base-offset field to c gives wrong results; DY_HALO_BASEOFF=1 reproduces that).
              const uint32_t r = t / 3, c = t % 3;
              const uint64_t hi = halo_hi | (p.halo_base_offset ? (static_cast<uint64_t>(c) << 49) : 0ull);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_bf16_ss(d_tmem, hi | (a_lo + r * 128 + c * 8 + 2 * k), desc_hi | (b_lo + t * bstep + 2 * k), idesc, (t | k) ? 1u : 0u);
            }
          }
          umma_commit_a(empty0 + stage * 8);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
          umma_commit_a(tfull0 + acc * 8);
          if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
        for (int kb = 0; kb < kiters; ++kb) {
          mbar_wait_a(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_addr = stage0 + stage * sbytes;
          const uint32_t b_addr = bres ? smem_base + static_cast<uint32_t>(kb) * bbytes : a_addr + kABytes;
          const uint32_t a_lo = desc_lo_const | ((a_addr & 0x3ffffu) >> 4), b_lo = desc_lo_const | ((b_addr & 0x3ffffu) >> 4);
          if (!(dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_ss(d_tmem, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit_a(empty0 + stage * 8);          // frees the smem slot when these MMAs retire
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_a(tfull0 + acc * 8);              // accumulator complete -> epilogue
        if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> regs -> bias/SiLU/residual -> (smem -> TMA store | global) ==========
    // 8 warps: warp pair (q, half) owns TMEM lanes [32q, 32q+32) and one half of every CW-column chunk.
    const int ew = warp - 4;
    const int q = ew & 3;                           // TMEM lane quarter this warp may access (== warp % 4)
    const int half = ew >> 2;
    const int row = q * 32 + lane;                  // accumulator row == pixel of the tile
    const int etid = threadIdx.x - 4 * 32;          // 0..255
    const int rows_valid = p.TW * p.TH * p.TB;
    const bool issuer = (etid == 0);                // issues the TMA stores and owns their bulk groups
    uint8_t* stage_base = smem + bres_bytes + static_cast<size_t>(p.stages) * stage_bytes;   // 2 x 16 KB output staging tiles
    const bool silu = (p.act == DY_ACT_SILU);
    const float bscale = silu ? 0.5f : 1.0f;        // SiLU path keeps 0.5*bias: h = 0.5*acc + 0.5*b in one FFMA
    int acc = 0; uint32_t acc_phase = 0;
    int store_ctr = 0;
    const int wl = row % p.TW;                      // this thread's pixel inside the tile never changes
    const int hl = (row / p.TW) % p.TH;
    const int bl = row / (p.TW * p.TH);
    for (TileIter it(p, blockIdx.x, gridDim.x); it.valid(); it.next(p.n_tiles, p.tiles_w, p.tiles_h)) {
      const int w0 = it.tw_i * p.TW, h0 = it.th_i * p.TH, b0 = it.tb_i * p.TB;
      const int w = w0 + wl, h = h0 + hl, b = b0 + bl;
      const bool valid = (row < rows_valid) && (w < p.Wo) && (h < p.Ho) && (b < p.B);
      const size_t pix = (static_cast<size_t>(b) * p.Ho + h) * p.Wo + w;
      const int n0 = it.n_tile * p.BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.BN);

      if constexpr (CW > 0) {
        // ---------------- fast path: every chunk is CW columns wide and leaves through a TMA store ----------------
        constexpr int WC = CW / 2;                  // columns per warp
        s_bias[etid] = (etid < p.BN) ? bscale * __ldg(p.bias + n0 + etid) : 0.f;   // visible after the chunk's first barrier
        const int nchunks = (p.BN + CW - 1) / CW;   // a ragged last chunk only exists when n_tiles == 1: TMA clips columns >= Cout
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * CW + half * WC;
          uint32_t r[WC];
          if constexpr (WC == 32) tmem_ld_32x32b_x32(taddr + col0, r);
          else tmem_ld_32x32b_x16(taddr + col0, r);
          uint8_t* st = stage_base + (store_ctr & 1) * kABytes;
          if (dbg & 1) {                             // knock-out: read the accumulator, release it, do nothing else
            tmem_ld_wait();
            if (c == nchunks - 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&tempty_bar[acc]); }
            continue;
          }
          if (issuer) bulk_wait_group_read<1>();     // the store that last used this staging buffer has read it
          named_bar_sync(1, 256);                    // ... and s_bias of this tile is complete
          tmem_ld_wait();
          if (c == nchunks - 1) {                    // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          float v[WC];
          const float4* bs = reinterpret_cast<const float4*>(s_bias + col0);
          if (silu) {
#pragma unroll
            for (int j = 0; j < WC / 4; ++j) {
              const float4 hb = bs[j];
              const float h0_ = fmaf(__uint_as_float(r[4 * j + 0]), 0.5f, hb.x), h1_ = fmaf(__uint_as_float(r[4 * j + 1]), 0.5f, hb.y);
              const float h2_ = fmaf(__uint_as_float(r[4 * j + 2]), 0.5f, hb.z), h3_ = fmaf(__uint_as_float(r[4 * j + 3]), 0.5f, hb.w);
              v[4 * j + 0] = fmaf(h0_, tanh_fast(h0_), h0_); v[4 * j + 1] = fmaf(h1_, tanh_fast(h1_), h1_);
              v[4 * j + 2] = fmaf(h2_, tanh_fast(h2_), h2_); v[4 * j + 3] = fmaf(h3_, tanh_fast(h3_), h3_);
            }
          } else {
#pragma unroll
            for (int j = 0; j < WC / 4; ++j) {
              const float4 bb = bs[j];
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bb.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
            }
          }
          if (p.res != nullptr && valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_ld + n0 + col0);
#pragma unroll
            for (int g = 0; g < WC / 8; ++g) {
              if (n0 + col0 + 8 * g + 8 > p.Cout) break;     // ragged last chunk: stay inside the residual slice
              const uint4 rr = __ldg(rp + g);
              v[8 * g + 0] += bf16_lo(rr.x); v[8 * g + 1] += bf16_hi(rr.x); v[8 * g + 2] += bf16_lo(rr.y); v[8 * g + 3] += bf16_hi(rr.y);
              v[8 * g + 4] += bf16_lo(rr.z); v[8 * g + 5] += bf16_hi(rr.z); v[8 * g + 6] += bf16_lo(rr.w); v[8 * g + 7] += bf16_hi(rr.w);
            }
          }
          // stage the [128 px x CW ch] chunk: 128-byte rows are 128B-swizzled (CW == 64 bf16, CW == 32 fp32),
          // 64-byte rows (CW == 32 bf16) are linear
          if constexpr (F32) {
            uint8_t* rowp = st + row * 128;
#pragma unroll
            for (int g = 0; g < WC / 4; ++g) {
              const int chunk16 = half * (WC / 4) + g;
              *reinterpret_cast<float4*>(rowp + ((chunk16 ^ (row & 7)) << 4)) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            }
          } else {
            uint8_t* rowp = st + row * (CW * 2);
#pragma unroll
            for (int g = 0; g < WC / 8; ++g) {
              uint4 o;
              o.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]); o.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
              o.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); o.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
              const int chunk16 = half * (WC / 8) + g;
              if constexpr (CW == 64) *reinterpret_cast<uint4*>(rowp + ((chunk16 ^ (row & 7)) << 4)) = o;
              else *reinterpret_cast<uint4*>(rowp + (chunk16 << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 256);
          if (issuer) {
            tma_store_4d(&p.tmO, st, n0 + c * CW, w0, h0, b0);
            bulk_commit_group();
          }
          ++store_ctr;
        }
      } else {
        // ---------------- generic path: fp32 outputs and odd widths, registers -> global ----------------
        const int nchunks = (p.BN + 63) >> 6;
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * 64 + half * 32;
          const int ncols = min(32, p.BN - col0);    // 32, 16 or <= 0 (BN is a multiple of 16)
          uint32_t r[32];
          if (ncols >= 32) tmem_ld_32x32b_x32(taddr + col0, r);
          else if (ncols > 0) tmem_ld_32x32b_x16(taddr + col0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          tmem_ld_wait();
          if (c == nchunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          const int n = n0 + col0;
          if (ncols > 0 && valid && n < p.Cout) {
            for (int j = 0; j < ncols; ++j) {
              if (n + j >= p.Cout) break;
              float x = __uint_as_float(r[j]) + __ldg(p.bias + n + j);
              if (silu) x = silu_fast(x);
              if (p.res != nullptr) x += __bfloat162float(p.res[pix * p.res_ld + n + j]);
              r[j] = __float_as_uint(x);
            }
            if (p.out_f32) {
              float* op = reinterpret_cast<float*>(p.out) + pix * p.out_ld + n;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j * 4 < ncols) {
                  if (n + j * 4 + 4 <= p.Cout)
                    reinterpret_cast<float4*>(op)[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                  else
                    for (int e = 0; e < 4; ++e) if (n + j * 4 + e < p.Cout) op[j * 4 + e] = __uint_as_float(r[j * 4 + e]);
                }
              }
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + n;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (g * 8 < ncols) {
                  if (n + g * 8 + 8 <= p.Cout) {
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
                    o.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
                    o.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
                    o.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                    reinterpret_cast<uint4*>(op)[g] = o;
                  } else {
                    for (int e = 0; e < 8; ++e) if (n + g * 8 + e < p.Cout) op[g * 8 + e] = __float2bfloat16(__uint_as_float(r[g * 8 + e]));
                  }
                }
              }
            }
          }
        }
      }
      if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
    }
    if (CW > 0 && issuer) bulk_wait_group<0>();     // all bulk stores complete before the CTA retires its smem
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------
// Host side: tile-shape choice and tensor-map encoding.
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
                      CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(DY_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, dtype, rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(DY_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] strides [%llu,%llu,%llu] box [%u,%u,%u,%u] base %p",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                (unsigned long long)strides_bytes[0], (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
                rank > 3 ? box[3] : 0, base);
  }
  return DY_OK;
}

int conv_pick_bn(int cout_pad) {
  int best = 16;
  for (int bn = 16; bn <= 256; bn += 16)
    if (cout_pad % bn == 0) best = bn;
  return best;
}

// Pixel-tile shape (TW x TH x TB <= 128 rows) maximising the fraction of MMA rows that are real pixels.
static void pick_tile(int Wo, int Ho, int B, int* TW, int* TH, int* TB) {
  double best = -1.0; int bw = 1, bh = 1, bb = 1;
  for (int tw = 1; tw <= Wo && tw <= 128; ++tw) {
    for (int th = 1; th <= Ho && tw * th <= 128; ++th) {
      const int tb_max = 128 / (tw * th);                                // a tile may span several images
      for (int tb = 1; tb <= tb_max && tb <= B; ++tb) {
        const double tiles = double(ceil_div(Wo, tw)) * ceil_div(Ho, th) * ceil_div(B, tb);
        const double eff = double(Wo) * Ho * B / (tiles * 128.0);
        // prefer wide boxes (longer contiguous runs for TMA) on ties
        const double score = eff + 1e-6 * tw;
        if (score > best) { best = score; bw = tw; bh = th; bb = tb; }
      }
    }
  }
  *TW = bw; *TH = bh; *TB = bb;
}

int conv_build_params(const dy_conv_desc* d, ConvParams* p, ConvLaunch* l) {
  DY_CHECK_ARG(d && p && l, "null conv descriptor");
  DY_CHECK_ARG(d->in && d->weight && d->bias && d->out, "conv: null tensor pointer");
  DY_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize %d unsupported (1 or 3)", d->ksize);
  DY_CHECK_ARG(d->stride == 1 || (d->stride == 2 && d->ksize == 3), "conv: stride %d with k=%d unsupported", d->stride, d->ksize);
  DY_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: bad shape");
  DY_CHECK_ARG(d->Cin % 8 == 0 && d->in_ld % 8 == 0 && d->in_ld >= d->Cin, "conv: Cin/in_ld must be multiples of 8 (16B rows)");
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0,
               "conv: in/weight must be 16B aligned");
  const int out_esz = d->out_dtype == DY_F32 ? 4 : 2;
  DY_CHECK_ARG(d->out_dtype == DY_BF16 || d->out_dtype == DY_F32, "conv: bad out dtype");
  DY_CHECK_ARG((d->out_ld * out_esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "conv: out slice must be 16B aligned");
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "conv: bias must be 16B aligned");
  if (d->residual)
    DY_CHECK_ARG(d->res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "conv: residual slice must be 16B aligned");

  memset(p, 0, sizeof(*p));
  bool halo = false;
  const int k = d->ksize, s = d->stride;
  const int Ho = (d->H + 2 * (k / 2) - k) / s + 1, Wo = (d->W + 2 * (k / 2) - k) / s + 1;
  const int cin_pad = round_up(d->Cin, kBlockK), cout_pad = round_up(d->Cout, 16);
  p->BN = conv_pick_bn(cout_pad);
  p->n_tiles = cout_pad / p->BN;
  p->kblocks = cin_pad / kBlockK;
  p->ntaps = k * k;
  p->Cout = d->Cout;
  p->out = d->out; p->out_ld = d->out_ld; p->out_f32 = d->out_dtype == DY_F32;
  p->res = reinterpret_cast<const __nv_bfloat16*>(d->residual); p->res_ld = d->res_ld;
  p->bias = d->bias; p->act = d->act;

  const uint64_t esz = 2;
  const uint64_t ld = d->in_ld;
  const char* base = reinterpret_cast<const char*>(d->in);
  if (k == 1) {
    // flat GEMM over all pixels: "image" of width M, height 1
    const uint64_t M = uint64_t(d->B) * d->H * d->W;
    DY_CHECK_ARG(M < (1ull << 31), "conv: too many pixels");
    p->Ho = 1; p->Wo = int(M); p->B = 1;
    p->TW = int(M < 128 ? M : 128); p->TH = 1; p->TB = 1;
    const uint64_t dims[4] = {uint64_t(d->Cin), M, 1, 1};
    const uint64_t strides[3] = {ld * esz, M * ld * esz, M * ld * esz};
    const uint32_t box[4] = {kBlockK, uint32_t(p->TW), 1, 1};
    int rc = encode_map(&p->tmA[0], base, 4, dims, strides, box);
    if (rc) return rc;
    p->nmaps = 1;
    p->taps[0] = ConvTap{0, 0, 0, 0};
  } else {
    p->Ho = Ho; p->Wo = Wo; p->B = d->B;
    pick_tile(Wo, Ho, d->B, &p->TW, &p->TH, &p->TB);
    // Halo mode: one activation load per tile instead of nine.  Needs a single 64-channel block, all weights resident and
    // 8x16-pixel tiles that fill the map well.
    {
      const double eff = double(Wo) * Ho / (double(ceil_div(Wo, 8)) * 8 * ceil_div(Ho, 16) * 16);
      const int b_all9 = 9 * p->BN * 128;
      halo = (s == 1 && cin_pad == kBlockK && p->n_tiles == 1 && eff >= 0.8 &&
              b_all9 + 2 * kHaloBytes + 2 * kABytes + 1024 <= kMaxDynSmem && getenv("DY_NO_HALO") == nullptr);
      if (halo) { p->TW = 8; p->TH = 16; p->TB = 1; }
    }
    const uint32_t box[4] = {kBlockK, uint32_t(halo ? kHaloW : p->TW), uint32_t(halo ? kHaloH : p->TH), uint32_t(p->TB)};
    if (s == 1) {
      const uint64_t dims[4] = {uint64_t(d->Cin), uint64_t(d->W), uint64_t(d->H), uint64_t(d->B)};
      const uint64_t strides[3] = {ld * esz, uint64_t(d->W) * ld * esz, uint64_t(d->H) * d->W * ld * esz};
      int rc = encode_map(&p->tmA[0], base, 4, dims, strides, box);
      if (rc) return rc;
      p->nmaps = 1;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p->taps[r * 3 + c] = ConvTap{0, int16_t(c - 1), int16_t(r - 1), 0};
    } else {
      DY_CHECK_ARG(d->H >= 2 && d->W >= 2, "conv: stride-2 needs H,W >= 2");
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          const uint64_t dims[4] = {uint64_t(d->Cin), uint64_t((d->W - px + 1) / 2), uint64_t((d->H - py + 1) / 2), uint64_t(d->B)};
          const uint64_t strides[3] = {2 * ld * esz, 2 * uint64_t(d->W) * ld * esz, uint64_t(d->H) * d->W * ld * esz};
          int rc = encode_map(&p->tmA[py * 2 + px], base + (uint64_t(py) * d->W + px) * ld * esz, 4, dims, strides, box);
          if (rc) return rc;
        }
      p->nmaps = 4;
      // tap offset o in {-1,0,+1}: input coord 2*x+o  ->  parity (o&1), coarse coord x + (o<0 ? -1 : 0)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          const int oy = r - 1, ox = c - 1;
          const int py = oy & 1, px = ox & 1;
          p->taps[r * 3 + c] = ConvTap{int16_t(py * 2 + px), int16_t(ox < 0 ? -1 : 0), int16_t(oy < 0 ? -1 : 0), 0};
        }
    }
  }
  p->tiles_w = ceil_div(p->Wo, p->TW);
  p->tiles_h = ceil_div(p->Ho, p->TH);
  p->m_tiles = p->tiles_w * p->tiles_h * ceil_div(p->B, p->TB);

  {
    const uint64_t dims[3] = {uint64_t(cin_pad), uint64_t(cout_pad), uint64_t(p->ntaps)};
    const uint64_t strides[2] = {uint64_t(cin_pad) * esz, uint64_t(cin_pad) * cout_pad * esz};
    const uint32_t box[3] = {kBlockK, uint32_t(p->BN), 1};
    int rc = encode_map(&p->tmB, d->weight, 3, dims, strides, box);
    if (rc) return rc;
  }

  // output tensor map for the TMA-store epilogue (same pixel box as A, CW channels wide).  A ragged last chunk is only
  // allowed with a single N tile, where every column >= BN is also >= Cout and is clipped by the tensor map.
  p->use_tma_store = 0;
  {
    const bool f32 = d->out_dtype == DY_F32;
    int cw = 0;
    // (TMA clips the innermost dimension at 16-byte granularity: the slice width must be a multiple of 16 bytes)
    if (f32) { if ((p->BN % 32 == 0 || p->n_tiles == 1) && d->Cout % 4 == 0) cw = 32; }
    else if (d->Cout % 8 == 0) {
      if (p->BN % 64 == 0) cw = 64;
      else if (p->BN % 32 == 0) cw = 32;
      else if (p->n_tiles == 1) cw = p->BN > 32 ? 64 : 32;
    }
    if (cw) {
      const uint64_t oes = out_esz;
      const uint64_t old = d->out_ld;
      const uint32_t obox[4] = {uint32_t(cw), uint32_t(p->TW), uint32_t(p->TH), uint32_t(p->TB)};
      const CUtensorMapSwizzle sw = (cw * out_esz == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
      const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      int rc;
      if (k == 1) {
        const uint64_t M = uint64_t(d->B) * d->H * d->W;
        const uint64_t dims[4] = {uint64_t(d->Cout), M, 1, 1};
        const uint64_t strides[3] = {old * oes, M * old * oes, M * old * oes};
        rc = encode_map(&p->tmO, d->out, 4, dims, strides, obox, sw, dt);
      } else {
        const uint64_t dims[4] = {uint64_t(d->Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(d->B)};
        const uint64_t strides[3] = {old * oes, uint64_t(Wo) * old * oes, uint64_t(Ho) * Wo * old * oes};
        rc = encode_map(&p->tmO, d->out, 4, dims, strides, obox, sw, dt);
      }
      if (rc) return rc;
      p->use_tma_store = cw;
    }
  }

  // Shared-memory plan.  Small layers keep ALL their weights resident (loaded once per CTA): the per-stage traffic and the
  // per-stage TMA issue then only cover the activation tile.  Everything else streams B next to A.
  const int b_tile = p->BN * 128;
  const int b_all = p->ntaps * p->kblocks * b_tile;
  const int budget = kMaxDynSmem - 1024 - 2 * kABytes;                  // minus alignment slack and the output staging tiles
  p->b_resident = (p->n_tiles == 1 && b_all <= budget - 4 * kABytes) ? 1 : 0;   // leave room for >= 4 activation stages
  if (getenv("DY_NO_BRES")) p->b_resident = 0;
  if (halo) p->b_resident = 1;
  const int stage_bytes = halo ? kHaloBytes : kABytes + (p->b_resident ? 0 : b_tile);
  int stages = (budget - (p->b_resident ? b_all : 0)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p->stages = stages;
  p->nacc = 512 / p->BN < kMaxAcc ? 512 / p->BN : kMaxAcc;
  l->smem_bytes = (p->b_resident ? b_all : 0) + stages * stage_bytes + 2 * kABytes + 1024;
  p->mode = (k == 1) ? 0 : (halo ? 3 : (s == 1 ? 1 : 2));
  { const char* e = getenv("DY_HALO_BASEOFF"); p->halo_base_offset = e ? atoi(e) : 0; }   // measured on B200: the swizzle XOR uses absolute smem address bits, the field must stay 0
  const int total = p->m_tiles * p->n_tiles;
  const int sms = num_sms();
  l->grid = total < sms ? total : sms;
  { const char* e = getenv("DY_CONV_DBG"); p->dbg = e ? atoi(e) : 0; }
  return DY_OK;
}

template <int MODE, int CW, bool F32>
static int conv_launch_t(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream) {
  static int max_smem_set = 0;
  if (max_smem_set < l->smem_bytes) {
    // 227 KB opt-in limit covers static + dynamic shared memory; the kernel's static part is < 2 KB
    DY_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<MODE, CW, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    max_smem_set = kMaxDynSmem;
  }
  static const bool use_pdl = (getenv("DY_NO_PDL") == nullptr);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(l->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = l->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  DY_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<MODE, CW, F32>, *p));
  return launch_status("conv_igemm_kernel");
}

template <int MODE>
static int conv_launch_m(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream) {
  if (p->use_tma_store == 64) return conv_launch_t<MODE, 64, false>(p, l, stream);
  if (p->use_tma_store == 32 && !p->out_f32) return conv_launch_t<MODE, 32, false>(p, l, stream);
  if (p->use_tma_store == 32 && p->out_f32) return conv_launch_t<MODE, 32, true>(p, l, stream);
  return conv_launch_t<MODE, 0, false>(p, l, stream);
}

int conv_launch(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream) {
  if (p->mode == 0) return conv_launch_m<0>(p, l, stream);
  if (p->mode == 1) return conv_launch_m<1>(p, l, stream);
  if (p->mode == 3) return conv_launch_m<3>(p, l, stream);
  return conv_launch_m<2>(p, l, stream);
}

}  // namespace dy

extern "C" int dy_conv2d(const dy_conv_desc* d, void* stream) {
  dy::ConvParams p; dy::ConvLaunch l;
  int rc = dy::conv_build_params(d, &p, &l);
  if (rc) return rc;
  return dy::conv_launch(&p, &l, static_cast<cudaStream_t>(stream));
}
