// conv_igemm.cu — Conv(k in {1,3}, stride in {1,2}) + bias + SiLU (+ residual) as an implicit GEMM on the
// sm_100a tensor cores.  Replaces the aten/cuDNN conv2d + BatchNorm + SiLU (+ add, + cat) kernel chain the
// reference launches for every Conv / RepVGGBlock / Bottleneck (ultralytics/nn/modules/conv.py:49-55,
// block.py:337-350,1480-1490).
//
// GEMM view: M = output pixels (B*Ho*Wo), N = Cout, K = k*k*Cin.  D[M,N] = A[M,K] * W[N,K]^T.
//   * A is never materialised.  Generic 3x3: for every filter tap (r,s) and 64-channel block the TMA engine loads a
//     [pixel tile x 64 ch] box of the NHWC activation shifted by the tap offset; out-of-image coordinates are zero-filled
//     by TMA, which IS the conv padding.  Stride 2 reads four "parity" views of the input so every tap is a dense box.
//     "Halo" modes (Cin <= 64, stride 1): ONE TMA box per tile brings the 8x16-pixel tile plus its 1-pixel border
//     (10 x 18 pixels) and the nine taps are nine shifted UMMA descriptors over that single shared-memory tile
//     (tcgen05 and TMA both apply the swizzle XOR to absolute shared-memory address bits, so a shifted start is legal).
//     Cin <= 32 uses 64-byte rows (64B swizzle): half the shared-memory traffic and half the MMAs of the 64-channel form.
//   * W is packed [tap][Cout_pad][Cin_pad] (K contiguous), loaded by a 3-D TMA box; small layers keep all taps resident.
//   * One elected thread issues M=128 x N=BN x K=16 tcgen05.mma accumulating fp32 in TMEM (up to 8 accumulator stages).
//   * Two independent epilogue groups (4 warps each, alternate tiles): tcgen05.ld -> bias -> SiLU (1 MUFU) -> + residual
//     (TMA-prefetched into the staging tile) -> bf16/fp32 -> swizzled staging tile -> TMA store into the concat slice.
//   * Persistent: grid = min(#tiles, #SMs); warp 0 = A producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = B producer,
//     4..7 / 8..11 = epilogue groups.  The single-thread loops test the NEXT barrier before issuing the current work
//     ("wait-ahead"): an mbarrier try_wait costs ~200 cycles even when the phase is already complete.
#include "dy_common.cuh"
#include "dy_ptx.cuh"
#include "conv_igemm.h"
#include <cstring>
#include <type_traits>
#include <cstdlib>

namespace dy {

using namespace ptx;

static constexpr int kBlockM = 128;           // UMMA M
static constexpr int kBlockK = 64;            // bf16 per 128B swizzle row
static constexpr int kABytes = kBlockM * 128; // one 64-channel A block
static constexpr int kMaxStages = 8;
static constexpr int kThreads = 384;          // 4 control warps + 2 x 4 epilogue warps (EG = 3: 512, three epilogue groups)
static constexpr int kTmemCols = 512;
static constexpr int kMaxBias = 1024;
static constexpr int kMaxDynSmem = 227 * 1024 - 1024;       // static part: barriers (< 0.5 KB); the bias table lives in the dynamic part
static constexpr int kHaloTW = 8, kHaloTH = 16;             // halo modes: 8x16 output pixels per tile
static constexpr int kHaloRows = kHaloTH + 2;
static constexpr int kMaxAcc = 8;             // accumulator stages in TMEM: min(8, 512 / BN), BN columns apart

// Contiguous tile range per CTA: coordinates advance by carry instead of by integer division.
struct TileIter {
  int n_tile, tw_i, th_i, tb_i, remaining;
  __device__ __forceinline__ TileIter(const ConvParams& p, int cta, int ncta) {
    int m;
    if (p.n_split > 1) {                       // this CTA's n tile is fixed; its group shares the m tiles
      n_tile = cta % p.n_split;
      const int g = cta / p.n_split, ng = ncta / p.n_split;
      const int base = p.m_tiles / ng, rem = p.m_tiles % ng;
      m = g * base + min(g, rem);
      remaining = base + (g < rem ? 1 : 0);
    } else {
      const int total = p.m_tiles * p.n_tiles;
      const int base = total / ncta, rem = total % ncta;
      const int begin = cta * base + min(cta, rem);
      remaining = base + (cta < rem ? 1 : 0);
      n_tile = begin % p.n_tiles;
      m = begin / p.n_tiles;
    }
    tw_i = m % p.tiles_w; m /= p.tiles_w;
    th_i = m % p.tiles_h;
    tb_i = m / p.tiles_h;
  }
  __device__ __forceinline__ bool valid() const { return remaining > 0; }
  // n_iter = number of n tiles iterated inside a CTA (1 when the n tile is fixed)
  __device__ __forceinline__ void next(int n_iter, int tiles_w, int tiles_h) {
    --remaining;
    if (n_iter > 1) {
      if (++n_tile < n_iter) return;
      n_tile = 0;
    }
    if (++tw_i == tiles_w) {
      tw_i = 0;
      if (++th_i == tiles_h) { th_i = 0; ++tb_i; }
    }
  }
};

__device__ __forceinline__ void mbar_wait_unless(uint32_t bar, uint32_t parity, bool already) {
  if (!already) mbar_wait_a(bar, parity);
}

// act(0.5*acc + hb) with hb = 0.5*bias for SiLU (h + h*tanh(h), one MUFU), acc + bias otherwise
__device__ __forceinline__ float act1(uint32_t acc, float hb, bool silu) {
  if (silu) {
    const float h = fmaf(__uint_as_float(acc), 0.5f, hb);
#if defined(DY_CONV_DBG_CONST) && (DY_CONV_DBG_CONST & 16)
    return fmaf(h, 0.75f, h);        // knock-out build: SiLU without its MUFU
#else
    return fmaf(h, tanh_fast(h), h);
#endif
  }
  return __uint_as_float(acc) + hb;
}

// MODE = tap geometry / A staging, known at compile time:
//   0: 1x1 (one tap, flat pixel index), 1: 3x3 stride 1 (tap (r,c) shifts the box by (c-1, r-1)), 2: 3x3 stride 2 (parity views),
//   3: 3x3 stride 1 halo, one 64-channel block, 128-byte rows;  4: 3x3 stride 1 halo, Cin <= 32, 64-byte rows;
//   5: 3x3 stride 2, Cin <= 32, 64-byte rows (a filter row of three taps per pipeline stage).
//   6: 3x3 stride 1 halo, Cin >= 128, Cout <= 128 ("paired halo"): the L2 -> SM path (~56 B/cycle/SM measured) bounds the
//      generic mode 1, which fetches 16 KB of A and 16 KB of B per 4 MMAs (128 B/cycle at the tensor rate).  Here the A
//      operand of a (tile, 64-channel block) is ONE halo tile read by all nine taps (23 KB instead of 144 KB), and every
//      streamed weight tile (tap, block) feeds TWO pixel tiles with their own accumulators: ~41 B/cycle at the tensor rate.
//      Two rings: A (halo tiles) and B (weight tiles), each with its own full / empty barriers.
// CW = chunk width (channels) of the TMA-store epilogue: 64 or 32 bf16 (F32 = false), 32 fp32 (F32 = true).
// CW = 0 selects the generic register->global epilogue (odd widths).
// Shared memory: [resident weights][stages x (A blocks [+ B blocks])][2 groups x 2 staging tiles].
// FUSE2 (MODE 3, CW 32 only): the Detect branch tail.  The SiLU output tile never leaves the SM: its two 32-channel bf16
// staging chunks ARE the K-major A operand of a second GEMM (x W2[N2,64]^T, issued by the epilogue group's leader into a
// private TMEM region); the fp32 result + bias2 goes out through the same staging memory and a TMA store.
// EG = epilogue groups (4 warps each, tiles dealt round-robin).  The 32-channel layers are bound by the per-tile epilogue
// chain (~3000 cycles per group and tile against a ~950-cycle mainloop): they run with three groups.
template <int MODE, int CW, bool F32, bool FUSE2 = false, int EG = 2>
__global__ void __launch_bounds__(128 + 128 * EG, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[kMaxAcc];
  __shared__ __align__(8) uint64_t tempty_bar[kMaxAcc];
  __shared__ __align__(8) uint64_t res_bar[3 * EG];     // [group][ring slot]: residual tile landed
  __shared__ __align__(8) uint64_t pre_bar[4 * EG];     // [group][buffer]: half-resolution addend tile landed
  __shared__ __align__(8) uint64_t d2_bar[3];           // [group]: fused 1x1 tail finished
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ __align__(8) uint64_t afull_bar[kMaxStages];     // mode 6: the halo-tile ring (full_bar / empty_bar then serve the weight ring)
  __shared__ __align__(8) uint64_t aempty_bar[kMaxStages];
  __shared__ uint32_t tmem_base_s;
  constexpr int NTAPS = MODE == 0 ? 1 : 9;
  constexpr bool HALO = (MODE == 3 || MODE == 4);
  constexpr bool PAIRED = (MODE == 6);
  constexpr bool S2 = (MODE == 2 || MODE == 5);         // stride 2: four parity views
  constexpr int ROWB = (MODE == 4 || MODE == 5) ? 64 : 128;   // bytes per pixel / per weight row in the A / B shared-memory tiles
  constexpr int A_BLK = kBlockM * ROWB;                 // one A k-block: 64 channels (16 KB) or 32 channels (8 KB)
  constexpr int KSTEPS = ROWB / 32;                     // K = 16 MMAs per k-block
  constexpr uint32_t LAYOUT = ROWB == 64 ? 4u : 2u;     // UMMA layout type: 64B / 128B swizzle
  constexpr int STG_BYTES = CW == 0 ? 0 : 128 * CW * (F32 ? 4 : 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * ROWB;
  const bool bres = p.b_resident != 0;
  const uint32_t w2_bytes = FUSE2 ? static_cast<uint32_t>(p.N2) * 128u : 0u;      // two 32-channel k-blocks of N2 x 64 B
  const uint32_t bres_bytes = (bres ? static_cast<uint32_t>(NTAPS * p.kblocks) * b_bytes : 0u) + w2_bytes;
#ifdef DY_CONV_DEBUG
  const int dbg = p.dbg;   // DY_CONV_DBG knock-outs for bottleneck hunting: 1 = no epilogue work, 2 = no MMA, 4 = no A loads, 8 = no B loads
#define DY_TR(role, it, ev) do { if (p.trace && blockIdx.x == 0 && (it) < 96) p.trace[((role) * 96 + (it)) * 8 + (ev)] = clock64(); } while (0)
#else
#ifdef DY_CONV_DBG_CONST
  constexpr int dbg = DY_CONV_DBG_CONST;   // compile-time knock-out build (tools/run_knockouts.sh): release-speed loops
#else
  constexpr int dbg = 0;   // knock-outs and the timeline (DY_CONV_TRACE) compile away unless built with -DDY_CONV_DEBUG
#endif
#define DY_TR(role, it, ev) do { } while (0)
#endif
  // `opaque` pins loop invariants in registers: without it the compiler re-derives the shared-window addresses
  // and re-reads kernel parameters inside the single-thread loops, whose cost is pure latency.
  const uint32_t smem_base = opaque(smem_u32(smem));
  float* const s_bias = reinterpret_cast<float*>(smem + p.bias_off);   // [Cout padded + 64] (+ the tail's bias at 512 when FUSE2)
  const uint32_t stage0 = opaque(smem_base + bres_bytes);  // first pipeline stage (1024-aligned: bres_bytes is a multiple of 1024)
  const uint32_t full0 = opaque(smem_u32(&full_bar[0])), empty0 = opaque(smem_u32(&empty_bar[0]));
  const uint32_t tfull0 = opaque(smem_u32(&tfull_bar[0])), tempty0 = opaque(smem_u32(&tempty_bar[0]));
  const uint32_t bres_b = opaque(smem_u32(&bres_bar));
  const int nstages = opaque(p.stages), nacc = opaque(p.nacc);
  const uint32_t sbytes = opaque(static_cast<uint32_t>(p.stage_bytes));
  const int kiters = NTAPS * p.kblocks, kps = HALO ? 1 : p.kps;
  const int n_iter = p.n_split > 1 ? 1 : p.n_tiles;

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < p.nmaps; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
    if (CW > 0 && !FUSE2) prefetch_tmap(&p.tmO);
    if (CW > 0 && p.has_res_tma) prefetch_tmap(&p.tmR);
    if (CW > 0 && p.pre != nullptr) prefetch_tmap(&p.tmP);
    if (CW > 0 && p.has_up) for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmU[i]);
    if (FUSE2) { prefetch_tmap(&p.tmW2); prefetch_tmap(&p.tmO2); }
  }
  if (warp == 1 && elect_one()) {
    const uint32_t producers = (bres || PAIRED) ? 1u : 2u; // A thread (+ B thread) arrive on every full barrier
    for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], producers); mbar_init(&empty_bar[s], PAIRED ? 2 : 1); }
    if constexpr (PAIRED) { for (int s = 0; s < p.a_stages; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aempty_bar[s], 1); } }
    for (int a = 0; a < nacc; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    for (int i = 0; i < 3 * EG; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < 4 * EG; ++i) mbar_init(&pre_bar[i], 1);
    mbar_init(&d2_bar[0], 1); mbar_init(&d2_bar[1], 1); mbar_init(&d2_bar[2], 1);
    mbar_init(&bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  Only the roles that touch
  // memory the previous kernels produced (or still read: the arena reuses buffers) wait for them - the A producer before
  // its first activation load, the epilogue warps before their first residual load / output store.  The weight and
  // bias loads (constants) and the MMA issuer do not wait, so a CTA that lands on a free SM during the previous
  // kernel's tail already has its resident weights on the way.  Our own dependents may start their prologue now.
  grid_dep_launch();

  if (warp == 0) {
    // ===================== A producer =====================
    if (elect_one()) {
      constexpr uint32_t lead = 1u;
      int stage = 0; uint32_t phase = 0;
      [[maybe_unused]] int trk = 0;
      const uint32_t a_tx = (dbg & 4) ? 0u : static_cast<uint32_t>(p.TW * p.TH * p.TB * ROWB);
      const uint32_t halo_tx = (dbg & 4) ? 0u : static_cast<uint32_t>(p.halo_pitch * kHaloRows * ROWB);
      const int kblocks = opaque(p.kblocks);
      const int tiles_w = opaque(p.tiles_w), tiles_h = opaque(p.tiles_h);
      const int TW = opaque(p.TW), TH = opaque(p.TH), TB = opaque(p.TB);
      TileIter it(p, blockIdx.x, gridDim.x);
      bool ok = mbar_try_wait_a(empty0, 1u);
      grid_dep_wait();
      [[maybe_unused]] const int kit = opaque(kiters), kps_r = opaque(kps), kcmax = opaque(kblocks * kBlockK);
      [[maybe_unused]] const uint32_t e_minus_f = empty0 - full0;
      [[maybe_unused]] uint32_t fullb = full0, dst = stage0;
      if constexpr (PAIRED) {
        // one halo tile per (pixel tile, 64-channel block); the two tiles of a pair alternate
        const uint32_t afull0 = smem_u32(&afull_bar[0]), aempty0 = smem_u32(&aempty_bar[0]);
        const int na = opaque(p.a_stages);
        const uint32_t abytes = opaque(static_cast<uint32_t>(p.a_stage_bytes));
        int sa = 0; uint32_t pa = 0;
        while (it.valid()) {
          TileIter t0 = it; it.next(n_iter, tiles_w, tiles_h);
          TileIter t1 = it; const bool two = it.valid();
          if (two) it.next(n_iter, tiles_w, tiles_h);
          for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j == 1 && !two) break;
              const TileIter& t = j ? t1 : t0;
              mbar_wait_a(aempty0 + sa * 8, pa ^ 1u);
              mbar_arrive_expect_tx_a(afull0 + sa * 8, halo_tx);
              if (!(dbg & 4)) tma_load_4d_a(smem_base + static_cast<uint32_t>(sa) * abytes, &p.tmA[0], afull0 + sa * 8, kb * kBlockK,
                                            t.tw_i * TW - 1, t.th_i * TH - 1, t.tb_i * TB);
              if (++sa == na) { sa = 0; pa ^= 1u; }
            }
          }
        }
      } else
      for (; it.valid(); it.next(n_iter, tiles_w, tiles_h)) {
        const int w0 = it.tw_i * TW, h0 = it.th_i * TH, b0 = it.tb_i * TB;
        if constexpr (HALO) {
          const uint32_t fb = full0 + stage * 8;
          DY_TR(0, trk, 0);
          mbar_wait_unless(empty0 + stage * 8, phase ^ 1u, ok);
          DY_TR(0, trk, 1);
          mbar_arrive_expect_tx_p(lead, fb, halo_tx);
          const int ns = stage + 1 == nstages ? 0 : stage + 1;
          const uint32_t nph = stage + 1 == nstages ? phase ^ 1u : phase;
          ok = mbar_try_wait_a(empty0 + ns * 8, nph ^ 1u);                 // wait-ahead: resolved while the TMA issues
          if (!(dbg & 4)) tma_load_4d_p(lead, stage0 + stage * sbytes, &p.tmA[0], fb, 0, w0 - 1, h0 - 1, b0);
          DY_TR(0, trk, 2); ++trk;
          stage = ns; phase = nph;
        } else {
          // lean loop: barrier / destination addresses are carried incrementally, coordinates advance by carry
          int kc = 0, ox = MODE == 0 ? 0 : -1, oy = MODE == 0 ? 0 : -1;
          for (int left = kit; left > 0; left -= kps_r) {
            DY_TR(0, trk, 0);
            mbar_wait_unless(fullb + e_minus_f, phase ^ 1u, ok);
            DY_TR(0, trk, 1);
            mbar_arrive_expect_tx_a(fullb, (left >= kps_r ? static_cast<uint32_t>(kps_r) : static_cast<uint32_t>(left)) * a_tx);
            uint32_t nfullb = fullb + 8, ndst = dst + sbytes, nph = phase;
            if (++stage == nstages) { stage = 0; nfullb = full0; ndst = stage0; nph ^= 1u; }
            ok = mbar_try_wait_a(nfullb + e_minus_f, nph ^ 1u);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (j < kps_r && j < left) {
                // stride 2: input coord 2*x+o -> parity view (o&1), coarse coord x + (o<0 ? -1 : 0)
                const int map = S2 ? ((oy & 1) * 2 + (ox & 1)) : 0;
                const int dx = S2 ? (ox < 0 ? -1 : 0) : ox, dy = S2 ? (oy < 0 ? -1 : 0) : oy;
                if (!(dbg & 4)) tma_load_4d_a(dst + j * A_BLK, &p.tmA[map], fullb, kc, w0 + dx, h0 + dy, b0);
                kc += kBlockK;
                if (kc == kcmax) { kc = 0; if (++ox == 2) { ox = -1; ++oy; } }
              }
            }
            DY_TR(0, trk, 2); ++trk;
            fullb = nfullb; dst = ndst; phase = nph;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: weights, either once (resident) or per stage =====================
    if (elect_one()) {
      constexpr uint32_t lead = 1u;
      const int kblocks = opaque(p.kblocks);
      if constexpr (PAIRED) {
        // weight tiles (tap, block) in the order the MMA issuer consumes them, once per tile PAIR
        TileIter cnt(p, blockIdx.x, gridDim.x);
        const int items = (cnt.remaining + 1) >> 1;
        const uint32_t b_ring0 = smem_base + static_cast<uint32_t>(p.a_stages) * static_cast<uint32_t>(p.a_stage_bytes);
        const uint32_t b_tx = (dbg & 8) ? 0u : b_bytes;
        int sb = 0; uint32_t pb = 0;
        for (int i = 0; i < items; ++i)
          for (int kb = 0; kb < kblocks; ++kb)
            for (int t = 0; t < 9; ++t) {
              mbar_wait_a(empty0 + sb * 8, pb ^ 1u);
              mbar_arrive_expect_tx_a(full0 + sb * 8, b_tx);
              if (!(dbg & 8)) tma_load_3d_a(b_ring0 + static_cast<uint32_t>(sb) * b_bytes, &p.tmB, full0 + sb * 8, kb * kBlockK, 0, t);
              if (++sb == nstages) { sb = 0; pb ^= 1u; }
            }
      } else
      if (bres) {
        const int n0 = (p.n_split > 1 ? static_cast<int>(blockIdx.x) % p.n_split : 0) * p.BN;
        mbar_arrive_expect_tx_p(lead, bres_b, (dbg & 8) ? 0u : bres_bytes);
        if (!(dbg & 8)) {
          for (int t = 0; t < NTAPS; ++t)
            for (int kc = 0; kc < kblocks; ++kc)
              tma_load_3d_p(lead, smem_base + static_cast<uint32_t>(t * kblocks + kc) * b_bytes, &p.tmB, bres_b, kc * kBlockK, n0, t);
          if constexpr (FUSE2) {                      // tail weights behind the nine tap tiles: k-block 0, k-block 1
            const uint32_t w2 = smem_base + static_cast<uint32_t>(NTAPS * kblocks) * b_bytes;
            tma_load_3d_p(lead, w2, &p.tmW2, bres_b, 0, 0, 0);
            tma_load_3d_p(lead, w2 + (w2_bytes >> 1), &p.tmW2, bres_b, 32, 0, 0);
          }
        }
      } else {
        int stage = 0; uint32_t phase = 0;
        const uint32_t b_tx = (dbg & 8) ? 0u : b_bytes;
        const int tiles_w = opaque(p.tiles_w), tiles_h = opaque(p.tiles_h), BN = opaque(p.BN);
        const uint32_t b_off = static_cast<uint32_t>(kps) * A_BLK;
        const int kit = opaque(kiters), kps_r = opaque(kps), kcmax = opaque(kblocks * kBlockK);
        const uint32_t e_minus_f = empty0 - full0;
        const uint32_t bb = opaque(b_bytes);
        uint32_t fullb = full0, dst = stage0 + b_off;
        bool ok = mbar_try_wait_a(empty0, 1u);
        for (TileIter it(p, blockIdx.x, gridDim.x); it.valid(); it.next(n_iter, tiles_w, tiles_h)) {
          const int n0 = it.n_tile * BN;
          int kc = 0, t = 0;
          for (int left = kit; left > 0; left -= kps_r) {
            mbar_wait_unless(fullb + e_minus_f, phase ^ 1u, ok);
            mbar_arrive_expect_tx_a(fullb, (left >= kps_r ? static_cast<uint32_t>(kps_r) : static_cast<uint32_t>(left)) * b_tx);
            uint32_t nfullb = fullb + 8, ndst = dst + sbytes, nph = phase;
            if (++stage == nstages) { stage = 0; nfullb = full0; ndst = stage0 + b_off; nph ^= 1u; }
            ok = mbar_try_wait_a(nfullb + e_minus_f, nph ^ 1u);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (j < kps_r && j < left) {
                if (!(dbg & 8)) tma_load_3d_a(dst + j * bb, &p.tmB, fullb, kc, n0, t);
                kc += kBlockK;
                if (kc == kcmax) { kc = 0; ++t; }
              }
            }
            fullb = nfullb; dst = ndst; phase = nph;
          }
        }
      }
    }
  } else if (warp == 1 || ((HALO || PAIRED) && warp == 2)) {
    // ===================== MMA issuers =====================
    // Halo modes: TWO issuing threads (warps 1 and 2) take alternate tiles, so one thread's barrier waits / commits overlap
    // the other's MMA issue (the tensor pipe executes both streams; the accumulators and smem stages are disjoint).
    // Other modes: one thread; every stage's first k-block carries the try_wait of the next stage (wait-ahead).
    if (elect_one()) {
      constexpr uint32_t lead = 1u;
      const uint32_t idesc = umma_idesc_bf16(kBlockM, p.BN);
      const uint32_t bbytes = opaque(b_bytes);
      // descriptor = constant high word | (address >> 4): only the low word changes
      const uint64_t b_hi = umma_desc_kmajor(0, 8 * ROWB, LAYOUT) & 0xffffffff00000000ull;
      const uint64_t a_hi = (HALO || PAIRED) ? (umma_desc_kmajor(0, static_cast<uint32_t>(p.halo_pitch) * ROWB, LAYOUT) & 0xffffffff00000000ull) : b_hi;
      const uint32_t lo_const = static_cast<uint32_t>(umma_desc_kmajor(0, 1024, LAYOUT) & 0xffffffffull);   // LBO field
      const uint32_t BNu = opaque(static_cast<uint32_t>(p.BN));
      [[maybe_unused]] int trk = 0, trt = 0;
      TileIter count_it(p, blockIdx.x, gridDim.x);
      const int my_total = count_it.remaining;
      if (bres) mbar_wait_a(bres_b, 0);
      if constexpr (PAIRED) {
        // TWO issuing threads (warps 1 and 2): issuer j multiplies every weight tile into the j-th pixel tile of the pair, so
        // each thread issues 4 MMAs per weight stage and the two instruction streams overlap (one thread issuing all 8 next
        // to its barrier traffic ran the tensor pipe at ~97 cycles per MMA instead of 64).  Halo slots: a pair takes two
        // consecutive slots per 64-channel block (issuer j the j-th; the ring is even, so the parity never changes); a final
        // single tile belongs to issuer 0, while issuer 1 just hands the weight stages back.
        const int j = warp - 1;
        const uint32_t afull0 = smem_u32(&afull_bar[0]), aempty0 = smem_u32(&aempty_bar[0]);
        const int na = opaque(p.a_stages), kblocks = opaque(p.kblocks);
        const uint32_t abytes = opaque(static_cast<uint32_t>(p.a_stage_bytes));
        const uint32_t b_ring0 = smem_base + static_cast<uint32_t>(na) * abytes;
        const uint32_t halo_rowstep = static_cast<uint32_t>(p.halo_pitch * ROWB) >> 4;
        const uint32_t BNv = opaque(static_cast<uint32_t>(p.BN));
        const uint32_t bstep = opaque(bbytes >> 4);
        const uint32_t b_lo0 = opaque(lo_const | ((b_ring0 & 0x3ffffu) >> 4));
        const uint32_t a_lo_base = opaque(lo_const | ((smem_base & 0x3ffffu) >> 4));
        const uint32_t astep = opaque(abytes >> 4);
        int sa = j; uint32_t pa = 0;
        int sb = 0; uint32_t pb = 0;
        uint32_t b_lo = b_lo0, fullb = full0;
        const uint32_t e_minus_f = empty0 - full0;
        bool okf = mbar_try_wait_a(full0, 0u);
        for (int t = 0; t < my_total; t += 2) {
          const bool two = t + 1 < my_total;
          const bool mine = (j == 0) || two;                                 // issuer 1 has no tile in a final single item
          const uint32_t slot = (static_cast<uint32_t>(t) & 3u) + static_cast<uint32_t>(j), accp = (static_cast<uint32_t>(t) >> 2) & 1u;
          if (mine) { mbar_wait_a(tempty0 + slot * 8, accp ^ 1u); tc_fence_after(); }
          const uint32_t d = tmem_base + slot * BNv;
          for (int kb = 0; kb < kblocks; ++kb) {
            if (mine) { mbar_wait_a(afull0 + sa * 8, pa); tc_fence_after(); }
            const uint32_t a_lo = a_lo_base + static_cast<uint32_t>(sa) * astep;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait_unless(fullb, pb, okf);
              tc_fence_after();
              uint32_t nfullb = fullb + 8, nb_lo = b_lo + bstep, npb = pb;
              if (++sb == nstages) { sb = 0; nfullb = full0; nb_lo = b_lo0; npb ^= 1u; }
              const uint32_t shift = static_cast<uint32_t>(tap / 3) * halo_rowstep + static_cast<uint32_t>(tap % 3) * (ROWB >> 4);
              if (mine && !(dbg & 2)) {
                okf = umma_bf16_ss_x4_waitahead(d, a_hi | (a_lo + shift), b_hi | b_lo, idesc, (kb | tap) ? 1u : 0u, nfullb, npb);
                umma_commit_a(fullb + e_minus_f);                          // one of the two arrivals that free the weight stage
              } else {
                okf = mbar_try_wait_a(nfullb, npb);
                mbar_arrive_a(fullb + e_minus_f);
              }
              fullb = nfullb; b_lo = nb_lo; pb = npb;
            }
            if (mine) umma_commit_a(aempty0 + sa * 8);
            sa += two ? 2 : 1;
            if (sa >= na) { sa -= na; pa ^= 1u; }
          }
          if (mine) umma_commit_a(tfull0 + slot * 8);
        }
      } else
      if constexpr (HALO) {
        const int m = warp - 1;                                            // this issuer's tile parity
        [[maybe_unused]] const int trole = m == 0 ? 1 : 2;
        int stage = m % nstages; uint32_t phase = static_cast<uint32_t>(m / nstages) & 1u;
        int acc = m % nacc; uint32_t acc_phase = static_cast<uint32_t>(m / nacc) & 1u;
        bool okf = mbar_try_wait_a(full0 + stage * 8, phase);
        bool oke = mbar_try_wait_a(tempty0 + acc * 8, acc_phase ^ 1u);
        const uint32_t halo_rowstep = static_cast<uint32_t>(p.halo_pitch * ROWB) >> 4;   // descriptor units per halo row
        const uint32_t b_lo = lo_const | ((smem_base & 0x3ffffu) >> 4);
        const uint32_t bstep = bbytes >> 4;
        for (int i = m; i < my_total; i += 2) {
          DY_TR(trole, trk, 0);
          mbar_wait_unless(tempty0 + acc * 8, acc_phase ^ 1u, oke);
          mbar_wait_unless(full0 + stage * 8, phase, okf);
          DY_TR(trole, trk, 1);
          tc_fence_after();
          int ns = stage + 2; uint32_t nph = phase;
          if (ns >= nstages) { ns -= nstages; nph ^= 1u; }
          int na = acc + 2; uint32_t nap = acc_phase;
          if (na >= nacc) { na -= nacc; nap ^= 1u; }
          okf = mbar_try_wait_a(full0 + ns * 8, nph);                      // wait-ahead for this issuer's next tile
          oke = mbar_try_wait_a(tempty0 + na * 8, nap ^ 1u);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * BNu;
          const uint32_t a_lo = lo_const | (((stage0 + stage * sbytes) & 0x3ffffu) >> 4);
          if (!(dbg & 2)) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              // tap (r, c): the 8-pixel row groups of the tile start (r*pitch + c) pixels into the halo tile and are one
              // halo row (SBO) apart.  The swizzle XOR is applied to absolute address bits by TMA and tcgen05 alike, so
              // neither the shifted start nor the SBO has to be a multiple of the 8-row swizzle atom.
              const uint32_t r = t / 3, c = t % 3;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16_ss_p(lead, d_tmem, a_hi | (a_lo + r * halo_rowstep + c * (ROWB >> 4) + 2 * k), b_hi | (b_lo + t * bstep + 2 * k), idesc,
                               (t | k) ? 1u : 0u);
            }
          }
          DY_TR(trole, trk, 2);
          umma_commit_p(lead, empty0 + stage * 8);
          umma_commit_p(lead, tfull0 + acc * 8);
          DY_TR(trole, trk, 3); ++trk;
          stage = ns; phase = nph; acc = na; acc_phase = nap;
        }
      } else {
        // Lean single-issuer loop: every address is carried incrementally (no multiplies, no parameter reloads); one asm block
        // per stage issues the next stage's try_wait and all of the stage's MMAs.
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        bool okf = mbar_try_wait_a(full0, 0u);
        bool oke = mbar_try_wait_a(tempty0, 1u);
        const int bres_i = opaque(p.b_resident);
        const int kit = opaque(kiters), kps_r = opaque(kps);
        const uint32_t sstep = opaque(sbytes >> 4), bstep = opaque(bbytes >> 4);
        const uint32_t a_lo0 = opaque(lo_const | ((stage0 & 0x3ffffu) >> 4));
        const uint32_t bres_lo0 = opaque(lo_const | ((smem_base & 0x3ffffu) >> 4));
        const uint32_t b_in_stage = opaque(static_cast<uint32_t>(kps) * (A_BLK >> 4));
        const uint32_t e_minus_f = empty0 - full0;
        uint32_t a_lo = a_lo0, fullb = full0;
        uint32_t tfb = tfull0, teb = tempty0, d_tmem = tmem_base;
        for (int my_tiles = my_total; my_tiles > 0; --my_tiles) {
          DY_TR(2, trt, 0);
          mbar_wait_unless(teb, acc_phase ^ 1u, oke);
          DY_TR(2, trt, 1);
          tc_fence_after();
          uint32_t bres_lo = bres_lo0;
          uint32_t accum = 0;
          for (int left = kit; left > 0; left -= kps_r) {
            DY_TR(1, trk, 0);
            mbar_wait_unless(fullb, phase, okf);
            DY_TR(1, trk, 1);
            tc_fence_after();
            uint32_t nfullb = fullb + 8, na_lo = a_lo + sstep, nph = phase;
            if (++stage == nstages) { stage = 0; nfullb = full0; na_lo = a_lo0; nph ^= 1u; }
            const uint32_t b_lo = bres_i ? bres_lo : a_lo + b_in_stage;
            if constexpr (KSTEPS == 2) {
              // 32-channel k-blocks (64-byte rows): a whole filter row (3 taps, 6 MMAs) per stage when it fits
              if (!(dbg & 2)) {
                if (kps_r == 3 && left >= 3) okf = umma_bf16_ss_k32x3_waitahead(d_tmem, a_hi | a_lo, b_hi | b_lo, bstep, idesc, accum, nfullb, nph);
                else {
                  okf = mbar_try_wait_a(nfullb, nph);
                  for (int j = 0; j < kps_r && j < left; ++j)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                      umma_bf16_ss(d_tmem, a_hi | (a_lo + j * (A_BLK >> 4) + 2 * k), b_hi | (b_lo + j * bstep + 2 * k), idesc, (accum | j | k) ? 1u : 0u);
                }
              } else okf = mbar_try_wait_a(nfullb, nph);
            } else if (!(dbg & 2)) {
              if (kps_r == 2 && left >= 2) okf = umma_bf16_ss_x8_waitahead(d_tmem, a_hi | a_lo, b_hi | b_lo, bstep, idesc, accum, nfullb, nph);
              else okf = umma_bf16_ss_x4_waitahead(d_tmem, a_hi | a_lo, b_hi | b_lo, idesc, accum, nfullb, nph);
            } else {
              okf = mbar_try_wait_a(nfullb, nph);
            }
            accum = 1u;
            bres_lo += bstep * static_cast<uint32_t>(kps_r);
            DY_TR(1, trk, 2);
            umma_commit_a(fullb + e_minus_f);           // frees the smem slot when these MMAs retire
            DY_TR(1, trk, 3); ++trk;
            fullb = nfullb; a_lo = na_lo; phase = nph;
          }
          uint32_t nteb = teb + 8, ntfb = tfb + 8, nd = d_tmem + BNu, nap = acc_phase;
          if (++acc == nacc) { acc = 0; nteb = tempty0; ntfb = tfull0; nd = tmem_base; nap ^= 1u; }
          oke = mbar_try_wait_a(nteb, nap ^ 1u);         // next accumulator stage: resolved behind the queued MMAs
          umma_commit_a(tfb);                            // accumulator complete -> epilogue
          DY_TR(2, trt, 2); ++trt;
          teb = nteb; tfb = ntfb; d_tmem = nd; acc_phase = nap;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> regs -> bias/SiLU/residual -> (smem -> TMA store | global) ==========
    // Two groups of 4 warps take alternate tiles; inside a group warp q owns TMEM lanes [32q, 32q+32) == 32 pixels.
    const int ew = warp - 4;
    const int g = ew >> 2;
    const int q = ew & 3;                           // TMEM lane quarter this warp may access (== warp % 4)
    const int row = q * 32 + lane;                  // accumulator row == pixel of the tile
    const bool leader = (threadIdx.x == 128 + g * 128);   // issues this group's TMA stores / residual loads
    const int barid = 1 + g;
    const bool silu = (p.act == DY_ACT_SILU);
    const float bscale = silu ? 0.5f : 1.0f;        // SiLU path keeps 0.5*bias: h = 0.5*acc + 0.5*b in one FFMA
    {
      const int nb = p.n_tiles * p.BN;              // == Cout padded to 16: the packed bias has that many entries
      for (int i = threadIdx.x - 128; i < nb; i += 128 * EG) s_bias[i] = bscale * __ldg(p.bias + i);
      if constexpr (FUSE2) { for (int i = threadIdx.x - 128; i < 64; i += 128 * EG) s_bias[512 + i] = i < p.N2 ? __ldg(p.bias2 + i) : 0.f; }
      named_bar_sync(1 + EG, 128 * EG);          // ids 1 .. EG are the groups' own barriers
    }
    grid_dep_wait();
    [[maybe_unused]] int trt = 0;
    [[maybe_unused]] const int trole = (lane == 0 && q == 0 && g < 2) ? 3 + g : 99;
#ifdef DY_CONV_DEBUG
#define DY_TRE(ev) do { if (trole < 99) DY_TR(trole, trt, ev); } while (0)
#else
#define DY_TRE(ev) do { } while (0)
#endif
    const int rows_valid = p.TW * p.TH * p.TB;
    const uint32_t stg = smem_base + static_cast<uint32_t>(p.stg_off) + static_cast<uint32_t>(g * p.nbuf) * STG_BYTES;
    const uint32_t resb0 = smem_u32(&res_bar[g * 3]);
    const bool has_res = CW > 0 && !F32 && p.has_res_tma != 0;
    const uint32_t res_tx = static_cast<uint32_t>(rows_valid) * CW * 2u;
    TileIter it(p, blockIdx.x, gridDim.x);
    for (int k = 0; k < g; ++k) it.next(n_iter, p.tiles_w, p.tiles_h);
    int acc = g % nacc; uint32_t acc_phase = static_cast<uint32_t>(g / nacc) & 1u;
    uint32_t sctr = 0;
    const int nbuf = p.nbuf;
    // half-resolution fp32 addend (dy_conv_desc.pre_add): the tile of a chunk - box {32 channels, TW/2, TH/2, TB} = 32 rows x 128 B,
    // 128B swizzle - is brought by TMA THREE chunks ahead into a ring of four 4 KB buffers per group, and every thread reads the
    // row of its pixel (w/2, h/2) from shared memory.  (Per-thread 16-byte global loads of the same rows cost +100 us on the
    // 64 -> 64 @160 layer, a TMA tile requested one chunk ahead +67 us: a chunk lasts ~1000 cycles, a DRAM round trip under load twice that.)
    const bool has_pre = CW == 32 && !F32 && !FUSE2 && p.pre != nullptr;
    const uint32_t preb0 = smem_u32(&pre_bar[g * 4]);
    const uint32_t pre_s = smem_base + static_cast<uint32_t>(p.pre_off) + static_cast<uint32_t>(g) * 16384u;
    const uint32_t pre_tx = static_cast<uint32_t>((p.TW >> 1) * (p.TH >> 1) * p.TB) * 128u;
    uint32_t pre_row = 0;
    TileIter pit = it;                                // the leader's prefetch stream: (tile, chunk) of the next request
    int pc = 0; uint32_t pissued = 0;
    const int pre_chunks = CW > 0 ? (p.BN + CW - 1) / (CW > 0 ? CW : 1) : 1;
    auto issue_pre = [&]() {
      if (!pit.valid()) return;
      const uint32_t bi = pissued & 3u;
      mbar_arrive_expect_tx_a(preb0 + bi * 8, pre_tx);
      tma_load_4d_a(pre_s + bi * 4096u, &p.tmP, preb0 + bi * 8, pit.n_tile * p.BN + pc * (CW > 0 ? CW : 1), (pit.tw_i * p.TW) >> 1, (pit.th_i * p.TH) >> 1,
                    pit.tb_i * p.TB);
      ++pissued;
      if (++pc == pre_chunks) {
        pc = 0;
#pragma unroll
        for (int k = 0; k < EG; ++k) pit.next(n_iter, p.tiles_w, p.tiles_h);
      }
    };
    if (has_pre) {
      const int wl = row % p.TW, hl = (row / p.TW) % p.TH, bl = row / (p.TW * p.TH);
      pre_row = static_cast<uint32_t>((bl * (p.TH >> 1) + (hl >> 1)) * (p.TW >> 1) + (wl >> 1));
      if (leader) { issue_pre(); issue_pre(); issue_pre(); }
    }
    // Residual tiles (same box as the output tile) arrive by TMA.  res_mode 2: TWO chunks ahead in a ring of three staging-sized
    // buffers per group (a chunk lasts ~1000 cycles, a DRAM round trip under load about twice that: one chunk ahead cost the
    // 32 -> 32 @160 layer 57 -> 89 us); where shared memory has no room for the ring: res_mode 1, one chunk ahead into the OTHER
    // staging tile (each thread then overwrites the residual it has read with its result), or res_mode 0, per chunk into the
    // single staging tile (K-heavy tiles: the epilogue has slack).
    const int res_mode = p.res_mode;
    const uint32_t res_slots = res_mode == 2 ? 3u : (res_mode == 1 ? 2u : 1u);
    const uint32_t res_s = res_mode == 2 ? smem_base + static_cast<uint32_t>(p.res_off) + static_cast<uint32_t>(g) * 3u * STG_BYTES : stg;
    TileIter rit = it;                                // the leader's request stream
    int rc_ = 0; uint32_t rslot_i = 0;                // chunk inside rit's tile, slot of the next request
    uint32_t rslot = 0, rphase = 0;                   // slot / phase of the chunk being consumed
    const int res_chunks = CW > 0 ? (p.BN + CW - 1) / (CW > 0 ? CW : 1) : 1;
    auto issue_res = [&]() {
      if (!rit.valid()) return;
      mbar_arrive_expect_tx_a(resb0 + rslot_i * 8, res_tx);
      tma_load_4d_a(res_s + rslot_i * STG_BYTES, &p.tmR, resb0 + rslot_i * 8, rit.n_tile * p.BN + rc_ * (CW > 0 ? CW : 1), rit.tw_i * p.TW, rit.th_i * p.TH,
                    rit.tb_i * p.TB);
      if (++rslot_i == res_slots) rslot_i = 0;
      if (++rc_ == res_chunks) {
        rc_ = 0;
#pragma unroll
        for (int k = 0; k < EG; ++k) rit.next(n_iter, p.tiles_w, p.tiles_h);
      }
    };
    if (has_res && leader && res_mode >= 1) { issue_res(); if (res_mode == 2) issue_res(); }
    while (it.valid()) {
      const int w0 = it.tw_i * p.TW, h0 = it.th_i * p.TH, b0 = it.tb_i * p.TB;
      const int n0 = it.n_tile * p.BN;
      TileIter nx = it;                              // this group's next tile (EG ahead)
#pragma unroll
      for (int k = 0; k < EG; ++k) nx.next(n_iter, p.tiles_w, p.tiles_h);

      DY_TRE(0);
      mbar_wait_a(tfull0 + acc * 8, acc_phase);
      DY_TRE(1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.BN);

      if constexpr (FUSE2) {
        // ---------------- fused Detect tail: SiLU tile (2 x 32 channels, bf16, 64B swizzle) -> second GEMM -> fp32 out ----------------
        static_assert(!FUSE2 || ((MODE == 3 || MODE == 5) && CW == 32 && !F32), "fused tail: 64 output channels, 32-wide chunks");
        const uint32_t d2b = smem_u32(&d2_bar[g]);
        // D2 is written over the tile's own accumulator stage (N2 <= 64 == BN columns; the stage has been read into the staging
        // tiles by then), and the stage goes back to the MMA issuers only after D2 has been read.  Private D2 columns per group
        // left 5 stages for three groups and two issuers, and the parity protocol needs every stage to belong to ONE (issuer,
        // group) pair, i.e. a multiple of lcm(2, EG) stages: with 4 stages a group that ran ahead of the others passed its
        // tfull wait on the completion of an EARLIER tile of the same parity (dead-lock about once per 25 steps of x@640).
        const uint32_t d2_tmem = tmem_base + static_cast<uint32_t>(acc * p.BN);
        if (leader) bulk_wait_group_read<0>();       // the previous tile's output stores have read the staging memory
        named_bar_sync(barid, 128);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float4 hbv[8];
          const float* bs = s_bias + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) hbv[i] = *reinterpret_cast<const float4*>(bs + 4 * i);
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const uint32_t rowp = stg + c * STG_BYTES + row * 64;
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const float4 hb0 = hbv[2 * gi], hb1 = hbv[2 * gi + 1];
            uint4 o;
            o.x = pack_bf16(act1(r[8 * gi + 0], hb0.x, true), act1(r[8 * gi + 1], hb0.y, true));      // (a fused tail always follows a SiLU conv:
            o.y = pack_bf16(act1(r[8 * gi + 2], hb0.z, true), act1(r[8 * gi + 3], hb0.w, true));      //  checked on the host)
            o.z = pack_bf16(act1(r[8 * gi + 4], hb1.x, true), act1(r[8 * gi + 5], hb1.y, true));
            o.w = pack_bf16(act1(r[8 * gi + 6], hb1.z, true), act1(r[8 * gi + 7], hb1.w, true));
            sts128(rowp + ((gi ^ ((row >> 1) & 3)) << 4), o);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(barid, 128);
        if (leader) {
          // D2[128, N2] = tile[128, 64] x W2[N2, 64]^T: two 32-channel k-blocks, two K16 steps each (64-byte rows, 64B swizzle)
          tc_fence_after();
          const uint64_t hi2 = umma_desc_kmajor(0, 512, 4u) & 0xffffffff00000000ull;
          const uint32_t lo2 = static_cast<uint32_t>(umma_desc_kmajor(0, 512, 4u) & 0xffffffffull);
          const uint32_t w2 = smem_base + bres_bytes - w2_bytes;
          const uint32_t idesc2 = umma_idesc_bf16(kBlockM, p.N2);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t a_lo = lo2 | (((stg + c * STG_BYTES) & 0x3ffffu) >> 4), b_lo = lo2 | (((w2 + c * (w2_bytes >> 1)) & 0x3ffffu) >> 4);

            umma_bf16_ss(d2_tmem, hi2 | a_lo, hi2 | b_lo, idesc2, c ? 1u : 0u);
            umma_bf16_ss(d2_tmem, hi2 | (a_lo + 2), hi2 | (b_lo + 2), idesc2, 1u);
          }
          umma_commit_a(d2b);
        }
        mbar_wait_a(d2b, sctr & 1u);                 // one tail GEMM per tile of this group
        tc_fence_after();
        const uint32_t t2addr = d2_tmem + (static_cast<uint32_t>(q * 32) << 16);
        if (p.tail_decode == 1 || p.tail_decode == 2) {
          // ---- fused Detect decode (head.py:100-131): the logits never leave the SM.  Each thread owns one anchor (its
          //      accumulator row) and writes its values straight into the channel-planar prediction tensor (B, 4+nc, A):
          //      a warp covers 4 tile rows x 8 columns = four full 32-byte sectors per plane and store instruction ----
          const float* b2 = s_bias + 512;
          const int w = w0 + (row & 7), h = h0 + (row >> 3);
          const bool inside = (w < p.Wo) && (h < p.Ho);
          float* yp = p.y + (static_cast<size_t>(b0) * (4 + p.y_nc)) * p.y_A + static_cast<size_t>(h) * p.Wo + w;
          if (p.tail_decode == 1) {                  // box branch: 4 sides x 16 bins -> DFL expectation -> dist2bbox(xywh) * stride
            float dist[4];
#pragma unroll
            for (int sd = 0; sd < 4; ++sd) {
              float4 bq[4];                          // this side's 16 bias values: four 16-byte loads in flight with the TMEM load
#pragma unroll
              for (int i = 0; i < 4; ++i) bq[i] = *reinterpret_cast<const float4*>(b2 + sd * 16 + 4 * i);
              uint32_t r[16];
              tmem_ld_32x32b_x16(t2addr + sd * 16, r);
              tmem_ld_wait();
              float x[kRegMax];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                x[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bq[i].x; x[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bq[i].y;
                x[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bq[i].z; x[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bq[i].w;
              }
              dist[sd] = dfl_expect(x);
            }
            tc_fence_before();                       // D2 (and with it the accumulator stage) fully read: hand it back
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);
            float bx[4];
            dist2bbox_xywh(static_cast<float>(w) + 0.5f, static_cast<float>(h) + 0.5f, dist, p.y_stride, bx);
            if (inside) {
#pragma unroll
              for (int c = 0; c < 4; ++c) yp[static_cast<size_t>(c) * p.y_A] = bx[c];
            }
          } else {                                   // class branch: sigmoid -> rows 4 .. 4+nc-1
            float4 bq[4];                            // N2 <= 32 here; the model's class heads have nc <= 16 logits per anchor
#pragma unroll
            for (int i = 0; i < 4; ++i) bq[i] = *reinterpret_cast<const float4*>(b2 + 4 * i);
            uint32_t r[32];
            tmem_ld_32x32b_x16(t2addr, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
            if (p.N2 > 16) tmem_ld_32x32b_x16(t2addr + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
            tmem_ld_wait();
            tc_fence_before();                       // D2 (and with it the accumulator stage) fully read: hand it back
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);
            if (inside) {
              const float bl[16] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w,
                                    bq[2].x, bq[2].y, bq[2].z, bq[2].w, bq[3].x, bq[3].y, bq[3].z, bq[3].w};
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < p.y_nc) yp[static_cast<size_t>(4 + c) * p.y_A] = sigmoid_fast(__uint_as_float(r[c]) + bl[c]);
              if (p.y_nc > 16) {
#pragma unroll
                for (int c = 16; c < 32; ++c)
                  if (c < p.y_nc) yp[static_cast<size_t>(4 + c) * p.y_A] = sigmoid_fast(__uint_as_float(r[c]) + b2[c]);
              }
            }
          }
        } else if (p.tail_decode == 3) {
          // ---- a hidden 1x1 Conv + SiLU as the tail (C2f.cv1 behind the stride-2 conv that feeds it, block.py:227-249):
          //      SiLU(D2 + bias2) -> bf16 -> the two staging tiles (the tail GEMM has consumed them) -> two TMA stores ----
          const int nout = (p.N2 + 31) >> 5;
          for (int oc = 0; oc < nout; ++oc) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t2addr + oc * 32, r);
            const float* b2 = s_bias + 512 + oc * 32;
            float4 hb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) hb[i] = *reinterpret_cast<const float4*>(b2 + 4 * i);
            tmem_ld_wait();
            const uint32_t rowp = stg + oc * STG_BYTES + row * 64;
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              const float4 h0 = hb[2 * gi], h1 = hb[2 * gi + 1];
              uint4 o;
              o.x = pack_bf16(act1(r[8 * gi + 0], 0.5f * h0.x, true), act1(r[8 * gi + 1], 0.5f * h0.y, true));
              o.y = pack_bf16(act1(r[8 * gi + 2], 0.5f * h0.z, true), act1(r[8 * gi + 3], 0.5f * h0.w, true));
              o.z = pack_bf16(act1(r[8 * gi + 4], 0.5f * h1.x, true), act1(r[8 * gi + 5], 0.5f * h1.y, true));
              o.w = pack_bf16(act1(r[8 * gi + 6], 0.5f * h1.z, true), act1(r[8 * gi + 7], 0.5f * h1.w, true));
              sts128(rowp + ((gi ^ ((row >> 1) & 3)) << 4), o);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);   // D2 fully read: the accumulator stage goes back to the MMA issuer
          named_bar_sync(barid, 128);
          if (leader) {
            for (int oc = 0; oc < nout; ++oc) tma_store_4d_a(&p.tmO2, stg + oc * STG_BYTES, oc * 32, w0, h0, b0);
            bulk_commit_group();
          }
        } else {
        const int nout = (p.N2 + 31) >> 5;           // 32-column fp32 output chunks (the last one may be ragged: TMA clips)
        for (int oc = 0; oc < nout; ++oc) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t2addr + oc * 32, r);
          tmem_ld_wait();
          if (oc > 0) {                              // the staging memory is being read by the previous chunk's store
            if (leader) bulk_wait_group_read<0>();
            named_bar_sync(barid, 128);
          }
          const uint32_t rowp = stg + row * 128;     // 128 rows x 32 fp32 = both 8 KB staging tiles, 128B swizzle
          const float* b2 = s_bias + 512 + oc * 32;
#pragma unroll
          for (int gi = 0; gi < 8; ++gi) {
            const float4 bb = *reinterpret_cast<const float4*>(b2 + 4 * gi);
            sts128(rowp + ((gi ^ (row & 7)) << 4),
                   make_uint4(__float_as_uint(__uint_as_float(r[4 * gi + 0]) + bb.x), __float_as_uint(__uint_as_float(r[4 * gi + 1]) + bb.y),
                              __float_as_uint(__uint_as_float(r[4 * gi + 2]) + bb.z), __float_as_uint(__uint_as_float(r[4 * gi + 3]) + bb.w)));
          }
          fence_proxy_async_smem();
          tc_fence_before();
          named_bar_sync(barid, 128);
          if (leader) {
            tma_store_4d_a(&p.tmO2, stg, oc * 32, w0, h0, b0);
            bulk_commit_group();
          }
        }
        __syncwarp();                                // (every chunk's tcgen05 fence has been issued above) D2 fully read
        if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);
        }
        ++sctr;
      } else if constexpr (CW > 0) {
        // ---------------- fast path: every chunk is CW columns wide and leaves through a TMA store ----------------
        constexpr int ROWO = CW * (F32 ? 4 : 2);    // staging row bytes: 128 (128B swizzle) or 64 (64B swizzle)
        const int nchunks = (p.BN + CW - 1) / CW;   // a ragged last chunk only exists when n_tiles == 1: TMA clips columns >= Cout
        for (int c = 0; c < nchunks; ++c) {
          const int buf = nbuf == 2 ? (sctr & 1) : 0;
          const uint32_t st = stg + buf * STG_BYTES;
          uint32_t r[CW];
          [[maybe_unused]] float4 hbv[CW == 32 ? 8 : 1];
          if constexpr (CW == 32) {                  // bias of this chunk -> registers while the TMEM load is in flight (the
            const float* bs = s_bias + n0 + c * CW;  // shared-memory port is saturated by the tensor core: LDS latency is long)
#pragma unroll
            for (int i = 0; i < 8; ++i) hbv[i] = *reinterpret_cast<const float4*>(bs + 4 * i);
          }
          tmem_ld_32x32b_x32(taddr + c * CW, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          if constexpr (CW == 64) tmem_ld_32x32b_x32(taddr + c * CW + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
          if (nbuf != 2) {
            // single staging tile: wait until the previous store has read it (res_mode 0: then fetch this chunk's residual into it)
            if (leader) bulk_wait_group_read<0>();
            named_bar_sync(barid, 128);
            if (has_res && res_mode == 0 && leader) issue_res();
          }
          if (has_res) mbar_wait_a(resb0 + rslot * 8, rphase);             // this chunk's residual tile has landed
          tmem_ld_wait();
          if constexpr (CW == 32 && !F32 && !FUSE2) {
            if (has_pre) {                           // acc += addend (fp32, before bias and activation)
              mbar_wait_a(preb0 + (sctr & 3u) * 8, (sctr >> 2) & 1u);
              const uint32_t prow = pre_s + (sctr & 3u) * 4096u + pre_row * 128u;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 t = lds128(prow + ((static_cast<uint32_t>(i) ^ (pre_row & 7u)) << 4));
                r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + __uint_as_float(t.x)); r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + __uint_as_float(t.y));
                r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + __uint_as_float(t.z)); r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + __uint_as_float(t.w));
              }
            }
          }
          if (c == 0) DY_TRE(2);
          if (c == nchunks - 1) {                    // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);
          }
          if (!(dbg & 1)) {
            const uint32_t rowp = st + row * ROWO;
            [[maybe_unused]] const float* bs = s_bias + n0 + c * CW;
            // The activation and the residual are run-time flags of the layer; as conditions inside the unrolled loops they were
            // if-converted, and every thread issued BOTH variants of every element (64 predicated MUFU per 32 elements: ~550
            // instructions per chunk where ~300 do the work) - in kernels whose P2-level layers are bound by instruction issue.
            // One uniform branch per chunk selects a loop body specialised at compile time instead.
            auto body = [&](auto silu_c, auto res_c) {
              constexpr bool SILU = decltype(silu_c)::value;
              [[maybe_unused]] constexpr bool RES = decltype(res_c)::value;
              if constexpr (F32) {
#pragma unroll
                for (int gi = 0; gi < CW / 4; ++gi) {
                  const float4 hb = hbv[gi];
                  float4 o;
                  o.x = act1(r[4 * gi + 0], hb.x, SILU); o.y = act1(r[4 * gi + 1], hb.y, SILU);
                  o.z = act1(r[4 * gi + 2], hb.z, SILU); o.w = act1(r[4 * gi + 3], hb.w, SILU);
                  sts128(rowp + ((gi ^ (row & 7)) << 4), make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                }
              } else {
#pragma unroll
                for (int gi = 0; gi < CW / 8; ++gi) {
                  const int pos = (ROWO == 128) ? (gi ^ (row & 7)) : (gi ^ ((row >> 1) & 3));
                  const float4 hb0 = CW == 32 ? hbv[(2 * gi) % (CW == 32 ? 8 : 1)] : *reinterpret_cast<const float4*>(bs + 8 * gi);
                  const float4 hb1 = CW == 32 ? hbv[(2 * gi + 1) % (CW == 32 ? 8 : 1)] : *reinterpret_cast<const float4*>(bs + 8 * gi + 4);
                  float v0 = act1(r[8 * gi + 0], hb0.x, SILU), v1 = act1(r[8 * gi + 1], hb0.y, SILU);
                  float v2 = act1(r[8 * gi + 2], hb0.z, SILU), v3 = act1(r[8 * gi + 3], hb0.w, SILU);
                  float v4 = act1(r[8 * gi + 4], hb1.x, SILU), v5 = act1(r[8 * gi + 5], hb1.y, SILU);
                  float v6 = act1(r[8 * gi + 6], hb1.z, SILU), v7 = act1(r[8 * gi + 7], hb1.w, SILU);
                  const uint32_t sp = rowp + (pos << 4);
                  if constexpr (RES) {
                    const uint4 rr = lds128(res_s + rslot * STG_BYTES + row * ROWO + (pos << 4));
                    v0 += bf16_lo(rr.x); v1 += bf16_hi(rr.x); v2 += bf16_lo(rr.y); v3 += bf16_hi(rr.y);
                    v4 += bf16_lo(rr.z); v5 += bf16_hi(rr.z); v6 += bf16_lo(rr.w); v7 += bf16_hi(rr.w);
                  }
                  uint4 o;
                  o.x = pack_bf16(v0, v1); o.y = pack_bf16(v2, v3); o.z = pack_bf16(v4, v5); o.w = pack_bf16(v6, v7);
                  sts128(sp, o);
                }
              }
            };
            using T = std::true_type; using Fa = std::false_type;
            if (silu) { if (has_res) body(T{}, T{}); else body(T{}, Fa{}); }
            else { if (has_res) body(Fa{}, T{}); else body(Fa{}, Fa{}); }
          }
          if (c == 0) DY_TRE(3);
          fence_proxy_async_smem();
          if (nbuf == 2 && leader) bulk_wait_group_read<0>();   // the previous store (other buffer) has read its tile: free after the barrier
          named_bar_sync(barid, 128);
          if (c == 0) DY_TRE(4);
          if (leader) {
            if (!(dbg & 1) && !(dbg & 32)) {
              tma_store_4d_a(&p.tmO, st, n0 + c * CW, w0, h0, b0);
              if (p.has_up) {                        // fused nn.Upsample(2x nearest): the same tile into the four parity views
#pragma unroll
                for (int u = 0; u < 4; ++u) tma_store_4d_a(&p.tmU[u], st, n0 + c * CW, w0, h0, b0);
              }
              bulk_commit_group();
            }
            if (has_pre) issue_pre();                // chunk sctr + 3 into the buffer read one chunk ago (every thread has passed this chunk's barrier since)
            if (has_res && res_mode >= 1) issue_res();   // ring: chunk + 2 into the slot read one chunk ago; res_mode 1: chunk + 1 into the other staging tile
          }
          if (has_res) { if (++rslot == res_slots) { rslot = 0; rphase ^= 1u; } }
          if (c == 0) DY_TRE(5);
          ++sctr;
        }
      } else {
        // ---------------- generic path: odd widths, registers -> global ----------------
        const int wl = row % p.TW, hl = (row / p.TW) % p.TH, bl = row / (p.TW * p.TH);
        const int w = w0 + wl, h = h0 + hl, b = b0 + bl;
        const bool valid = (row < rows_valid) && (w < p.Wo) && (h < p.Ho) && (b < p.B);
        const size_t pix = (static_cast<size_t>(b) * p.Ho + h) * p.Wo + w;
        const int nchunks = (p.BN + 31) >> 5;
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = c * 32;
          const int ncols = min(32, p.BN - col0);    // 32 or 16 (BN is a multiple of 16)
          uint32_t r[32];
          if (ncols >= 32) tmem_ld_32x32b_x32(taddr + col0, r);
          else tmem_ld_32x32b_x16(taddr + col0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          tmem_ld_wait();
          if (c == nchunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty0 + acc * 8);
          }
          const int n = n0 + col0;
          if (valid && n < p.Cout && !(dbg & 1)) {
            for (int j = 0; j < ncols; ++j) {
              if (n + j >= p.Cout) break;
              float x = act1(r[j], s_bias[n + j], silu);
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
              for (int gi = 0; gi < 4; ++gi) {
                if (gi * 8 < ncols) {
                  if (n + gi * 8 + 8 <= p.Cout) {
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(r[8 * gi + 0]), __uint_as_float(r[8 * gi + 1]));
                    o.y = pack_bf16(__uint_as_float(r[8 * gi + 2]), __uint_as_float(r[8 * gi + 3]));
                    o.z = pack_bf16(__uint_as_float(r[8 * gi + 4]), __uint_as_float(r[8 * gi + 5]));
                    o.w = pack_bf16(__uint_as_float(r[8 * gi + 6]), __uint_as_float(r[8 * gi + 7]));
                    reinterpret_cast<uint4*>(op)[gi] = o;
                  } else {
                    for (int e = 0; e < 8; ++e) if (n + gi * 8 + e < p.Cout) op[gi * 8 + e] = __float2bfloat16(__uint_as_float(r[gi * 8 + e]));
                  }
                }
              }
            }
          }
        }
      }
      DY_TRE(7); ++trt;
      acc += EG;
      if (acc >= nacc) { acc -= nacc; acc_phase ^= 1u; }
      it = nx;
    }
    if (CW > 0 && leader) bulk_wait_group_read<0>();   // the bulk stores have READ the staging tiles (the CTA may retire its smem);
                                                       // their global writes complete asynchronously, ordered by the kernel boundary
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

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapSwizzle swizzle, CUtensorMapDataType dtype) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(DY_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  // L2 promotion widens every DRAM fetch of a box row to 64 / 128 / 256 bytes.  256 is right when consecutive pixels are
  // contiguous (the box spans the whole inner dimension of a dense tensor); for a channel SLICE of a wider buffer it drags in
  // the neighbouring channels nobody asked for (measured: a 64-channel slice of a 128-channel buffer read 452 MB for 210,
  // a 32-channel slice of a 96-channel concat buffer 222 MB for 105), so a slice is promoted to its own row length only.
  const unsigned esz = dtype == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4u : (dtype == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1u : 2u);
  const unsigned long long row_bytes = static_cast<unsigned long long>(box[0]) * esz;
  const bool dense_rows = rank > 1 && box[0] == dims[0] && strides_bytes[0] == dims[0] * esz;
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  static const bool narrow = getenv("DY_TMAP_PROMO256") == nullptr;
  if (narrow && !dense_rows && row_bytes < 256) promo = row_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  CUresult r = fn(m, dtype, rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo,
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

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// N tile: a multiple of 16 dividing Cout_pad, <= max_bn.  Cost model per CTA (cycles): waves x (k-iterations x per-k-block
// MMA time + per-tile overhead), where an M128 x N x K64 block costs 4 x max(N/2 [tensor pipe], 32 + N/4 [shared-memory
// operand reads: tcgen05 reads 4 KB of A and N*32 B of B per K=16 step at 128 B/cycle]).
static int pick_bn(int cout_pad, int m_tiles, int kiters, int max_bn) {
  const int sms = num_sms();
  int best = 16; double best_cost = 1e30;
  bool any_tma = false;                                   // widths the TMA-store epilogue can tile: 32-column chunks, or one ragged N tile
  for (int bn = 16; bn <= max_bn && bn <= 256; bn += 16)
    if (cout_pad % bn == 0 && (bn % 32 == 0 || cout_pad == bn)) any_tma = true;
  for (int bn = 16; bn <= max_bn && bn <= 256; bn += 16) {
    if (cout_pad % bn) continue;
    if (any_tma && !(bn % 32 == 0 || cout_pad == bn)) continue;
    const int n_tiles = cout_pad / bn;
    const long long waves = ((long long)m_tiles * n_tiles + sms - 1) / sms;
    const double per_k = 4.0 * (bn / 2.0 > 32.0 + bn / 4.0 ? bn / 2.0 : 32.0 + bn / 4.0);
    const double cost = double(waves) * (kiters * per_k + 600.0) * (1.0 + 0.02 * (n_tiles - 1));   // + A re-reads per extra n tile
    if (cost <= best_cost * 1.001) { if (cost < best_cost) best_cost = cost; best = bn; }          // ties -> wider tile
  }
  if (env_int("DY_CONV_WIDE_BN", 0)) {                      // experiment: always the widest legal N tile
    for (int bn = 16; bn <= max_bn && bn <= 256; bn += 16)
      if (cout_pad % bn == 0 && (!any_tma || bn % 32 == 0 || cout_pad == bn)) best = bn;
  }
  return best;
}

// Pixel-tile shape (TW x TH x TB <= 128 rows) maximising the fraction of MMA rows that are real pixels.
static void pick_tile(int Wo, int Ho, int B, int* TW, int* TH, int* TB, bool even = false) {
  double best = -1.0; int bw = 1, bh = 1, bb = 1;
  const bool flat_ties = env_int("DY_TILE_FLAT", 0) != 0;
  for (int tw = 1; tw <= Wo && tw <= 128; ++tw) {
    for (int th = 1; th <= Ho && tw * th <= 128; ++th) {
      const int tb_max = 128 / (tw * th);                                // a tile may span several images
      for (int tb = 1; tb <= tb_max && tb <= B; ++tb) {
        if (even && ((tw | th) & 1)) continue;                                // half-resolution addend: the tile maps onto whole low-resolution pixels
        const double tiles = double(ceil_div(Wo, tw)) * ceil_div(Ho, th) * ceil_div(B, tb);
        const double eff = double(Wo) * Ho * B / (tiles * 128.0);
        // ties: compact 2-D tiles first (a 16x8 tile re-reads 10 input rows per 8 output rows through its 3x3 taps, a
        // 16x1x8 tile 3 per 1: same L2->SM traffic, but the re-reads miss L2 at batch 64), then wide boxes (long TMA runs)
        const double score = eff + (flat_ties ? 0.0 : 1e-4 * (th < 8 ? th : 8) / 8.0) + 1e-6 * tw;
        if (score > best) { best = score; bw = tw; bh = th; bb = tb; }
      }
    }
  }
  *TW = bw; *TH = bh; *TB = bb;
}

int conv_build_params(const dy_conv_desc* d, ConvParams* p, ConvLaunch* l) {
  DY_CHECK_ARG(d && p && l, "null conv descriptor");
  DY_CHECK_ARG(d->in && d->weight && d->bias && (d->out || d->weight2), "conv: null tensor pointer");
  DY_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize %d unsupported (1 or 3)", d->ksize);
  DY_CHECK_ARG(d->stride == 1 || (d->stride == 2 && d->ksize == 3), "conv: stride %d with k=%d unsupported", d->stride, d->ksize);
  DY_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: bad shape");
  DY_CHECK_ARG(d->Cin % 8 == 0 && d->in_ld % 8 == 0 && d->in_ld >= d->Cin, "conv: Cin/in_ld must be multiples of 8 (16B rows)");
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0,
               "conv: in/weight must be 16B aligned");
  const int out_esz = d->out_dtype == DY_F32 ? 4 : 2;
  DY_CHECK_ARG(d->out_dtype == DY_BF16 || d->out_dtype == DY_F32, "conv: bad out dtype");
  DY_CHECK_ARG(d->weight2 || ((d->out_ld * out_esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0), "conv: out slice must be 16B aligned");
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "conv: bias must be 16B aligned");
  if (d->residual)
    DY_CHECK_ARG(d->res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "conv: residual slice must be 16B aligned");
  const bool fuse2 = d->weight2 != nullptr;
  if (fuse2) {
    DY_CHECK_ARG(d->bias2 && d->Cout2 > 0, "conv: fused tail needs weight2, bias2 and Cout2");
    DY_CHECK_ARG(((d->tail_decode == 1 || d->tail_decode == 2) || ((d->out2_ld * (d->tail_decode == 3 ? 2 : 4)) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out2) & 15) == 0)) &&
                 (reinterpret_cast<uintptr_t>(d->weight2) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->bias2) & 15) == 0,
                 "conv: fused tail tensors must be 16B aligned");
    DY_CHECK_ARG(d->out2 || d->tail_decode == 1 || d->tail_decode == 2, "conv: fused tail needs out2 (or a decoding tail)");
    DY_CHECK_ARG(d->tail_decode >= 0 && d->tail_decode <= 3, "conv: tail_decode must be 0 (raw fp32), 1 (box), 2 (class) or 3 (SiLU, bf16)");
    if (d->tail_decode == 1 || d->tail_decode == 2) {
      DY_CHECK_ARG(d->y && d->y_A > 0 && d->y_nc > 0 && d->y_anchor_off >= 0, "conv: tail_decode needs y, y_A, y_nc, y_anchor_off");
      if (!(d->y_nc <= 32 && (d->tail_decode == 1 ? d->Cout2 == 64 : d->Cout2 >= d->y_nc && d->Cout2 <= 32) && d->W % 4 == 0 && d->y_A % 4 == 0 &&
            d->y_anchor_off % 4 == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0))
        return fail(DY_ERR_UNSUPPORTED, "conv: tail_decode needs nc <= 32, W, A, anchor offset multiples of 4 and a 16B-aligned y");
    }
    const bool geom_halo = d->stride == 1 && d->Cin > 32 && d->Cin <= 64;                      // 64-channel halo mode
    const bool geom_s2 = d->stride == 2 && d->Cin <= 32 && d->tail_decode == 3 && d->Cout2 % 8 == 0;   // 32-channel stride-2 mode
    if (!(d->ksize == 3 && (geom_halo || geom_s2) && d->Cout == 64 && d->Cout2 <= 64 && d->Cout2 % 4 == 0 &&
          d->residual == nullptr && d->up_out == nullptr && d->act == DY_ACT_SILU))
      return fail(DY_ERR_UNSUPPORTED, "conv: fused 1x1 tail needs k3 with (s1, 32 < Cin <= 64) or (s2, Cin <= 32, SiLU bf16 tail), Cout 64, Cout2 <= 64, SiLU, no residual / upsample");
  }

  memset(p, 0, sizeof(*p));
  const int k = d->ksize, s = d->stride;
  const int Ho = (d->H + 2 * (k / 2) - k) / s + 1, Wo = (d->W + 2 * (k / 2) - k) / s + 1;
  const int cin_pad = round_up(d->Cin, kBlockK), cout_pad = round_up(d->Cout, 16);
  DY_CHECK_ARG(cout_pad + 64 <= kMaxBias, "conv: Cout %d too large (max %d)", d->Cout, kMaxBias - 64);
  const bool f32 = d->out_dtype == DY_F32;
  p->ntaps = k * k;
  p->Cout = d->Cout;
  p->out = d->out; p->out_ld = d->out_ld; p->out_f32 = f32;
  p->res = reinterpret_cast<const __nv_bfloat16*>(d->residual); p->res_ld = d->res_ld;
  p->bias = d->bias; p->act = d->act;
  p->n_split = 1; p->kps = 1;
  const int sms = num_sms();

  // ---- mode, pixel tile, N tile ----
  int mode = (k == 1) ? 0 : (s == 1 ? 1 : 2);
  if (mode == 2 && d->Cin <= 32 && !env_int("DY_NO_K32", 0)) mode = 5;          // stride 2 with 64-byte rows: half the MMA / smem / L2 work
  if (k == 3 && s == 1 && !env_int("DY_NO_HALO", 0)) {
    const double eff = double(Wo) * Ho / (double(ceil_div(Wo, kHaloTW)) * kHaloTW * ceil_div(Ho, kHaloTH) * kHaloTH);
    if (eff >= 0.8 || fuse2) {                    // a caller asking for the fused tail accepts ragged tiles (small maps are launch-bound)
      if (d->Cin <= 32 && !env_int("DY_NO_K32", 0)) mode = 4;
      else if (cin_pad == kBlockK) mode = 3;
    }
  }
  if (mode == 1 && !fuse2 && !f32 && d->up_out == nullptr && cin_pad >= 2 * kBlockK && cout_pad <= 128 && cout_pad % 32 == 0 &&
      d->Cout % 8 == 0 && !env_int("DY_NO_PAIRED", 0)) {
    // paired halo (mode 6): needs whole 8x16 tiles to pay (a ragged map wastes MMA rows the generic mode does not)
    const double eff = double(Wo) * Ho / (double(ceil_div(Wo, kHaloTW)) * kHaloTW * ceil_div(Ho, kHaloTH) * kHaloTH);
    if (eff >= 0.8) { mode = 6; p->BN = cout_pad; p->n_tiles = 1; }
  }
  if (mode == 3 || mode == 4) {
    // halo: resident weights of ONE n tile per CTA (9 x BN rows); several n tiles -> static split of the grid
    const int m_tiles = ceil_div(Wo, kHaloTW) * ceil_div(Ho, kHaloTH) * d->B;
    // 64 -> 128 (the merged first Detect convs): one N = 128 tile reads every activation row once for 128 outputs (64
    // cycles per MMA, tensor- and smem-balanced) where two N = 64 halves read it twice (2 x 48): all 147 KB of weights stay
    // resident, at the price of two halo stages and one staging tile per epilogue group.
    const bool wide_halo = (mode == 3 && cout_pad == 128 && !fuse2 && !env_int("DY_HALO_NO_BN128", 0));
    const int bn = fuse2 ? 64 : (wide_halo ? 128 : pick_bn(cout_pad, m_tiles, 9, mode == 4 ? 128 : 64));   // the fused tail contracts over all 64 channels of one tile
    if (cout_pad / bn > 4 || cout_pad / bn > sms) mode = 1;
    else { p->BN = bn; p->n_tiles = cout_pad / bn; p->n_split = p->n_tiles; }
  }
  if (fuse2 && mode != 3 && !(mode == 5 && d->tail_decode == 3))
    return fail(DY_ERR_UNSUPPORTED, "conv: fused 1x1 tail needs the 64-channel halo mode (or the 32-channel stride-2 mode with a SiLU bf16 tail)");
  const bool paired = (mode == 6);
  const bool halo = (mode == 3 || mode == 4 || paired);     // tile geometry / activation box of the halo family
  const int rowb = (mode == 4 || mode == 5) ? 64 : 128;
  const int a_blk = kBlockM * rowb;
  p->mode = mode;
  p->kblocks = ((halo && !paired) || mode == 5) ? 1 : cin_pad / kBlockK;
  const int kiters = p->ntaps * p->kblocks;
  const bool flat = (k == 1 && d->up_out == nullptr && d->pre_add == nullptr);   // 1x1: flat GEMM over all pixels unless a spatial tile is needed
  if (d->up_out) {
    DY_CHECK_ARG(!f32 && d->up_ld % 8 == 0 && d->up_ld >= d->Cout && (reinterpret_cast<uintptr_t>(d->up_out) & 15) == 0,
                 "conv: up_out needs a bf16 primary output and a 16B-aligned slice");
  }
  if (flat) {
    const uint64_t M = uint64_t(d->B) * d->H * d->W;          // "image" of width M, height 1
    DY_CHECK_ARG(M < (1ull << 31), "conv: too many pixels");
    p->Ho = 1; p->Wo = int(M); p->B = 1;
    p->TW = int(M < 128 ? M : 128); p->TH = 1; p->TB = 1;
  } else {
    p->Ho = Ho; p->Wo = Wo; p->B = d->B;
    if (halo) { p->TW = kHaloTW; p->TH = kHaloTH; p->TB = 1; }
    else pick_tile(Wo, Ho, d->B, &p->TW, &p->TH, &p->TB, d->pre_add != nullptr);
  }
  p->tiles_w = ceil_div(p->Wo, p->TW);
  p->tiles_h = ceil_div(p->Ho, p->TH);
  p->m_tiles = p->tiles_w * p->tiles_h * ceil_div(p->B, p->TB);
  if (!halo) {
    p->BN = fuse2 ? 64 : pick_bn(cout_pad, p->m_tiles, kiters, env_int("DY_CONV_MAXBN", 256));   // the fused tail contracts over all 64 channels of one tile
    p->n_tiles = cout_pad / p->BN;
  }
  p->halo_pitch = halo ? (env_int("DY_HALO_PITCH16", 0) ? 16 : kHaloTW + 2) : 0;

  // ---- activation tensor maps ----
  const uint64_t esz = 2;
  const uint64_t ld = d->in_ld;
  const char* base = reinterpret_cast<const char*>(d->in);
  if (flat) {
    const uint64_t M = uint64_t(p->Wo);
    const uint64_t dims[4] = {uint64_t(d->Cin), M, 1, 1};
    const uint64_t strides[3] = {ld * esz, M * ld * esz, M * ld * esz};
    const uint32_t box[4] = {kBlockK, uint32_t(p->TW), 1, 1};
    int rc = encode_map(&p->tmA[0], base, 4, dims, strides, box);
    if (rc) return rc;
    p->nmaps = 1;
  } else if (s == 1) {
    const uint64_t dims[4] = {uint64_t(d->Cin), uint64_t(d->W), uint64_t(d->H), uint64_t(d->B)};
    const uint64_t strides[3] = {ld * esz, uint64_t(d->W) * ld * esz, uint64_t(d->H) * d->W * ld * esz};
    const uint32_t box[4] = {uint32_t(rowb / 2), uint32_t(halo ? p->halo_pitch : p->TW), uint32_t(halo ? kHaloRows : p->TH), uint32_t(p->TB)};
    int rc = encode_map(&p->tmA[0], base, 4, dims, strides, box, mode == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    p->nmaps = 1;
  } else {
    DY_CHECK_ARG(d->H >= 2 && d->W >= 2, "conv: stride-2 needs H,W >= 2");
    const uint32_t box[4] = {uint32_t(rowb / 2), uint32_t(p->TW), uint32_t(p->TH), uint32_t(p->TB)};
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint64_t dims[4] = {uint64_t(d->Cin), uint64_t((d->W - px + 1) / 2), uint64_t((d->H - py + 1) / 2), uint64_t(d->B)};
        const uint64_t strides[3] = {2 * ld * esz, 2 * uint64_t(d->W) * ld * esz, uint64_t(d->H) * d->W * ld * esz};
        int rc = encode_map(&p->tmA[py * 2 + px], base + (uint64_t(py) * d->W + px) * ld * esz, 4, dims, strides, box,
                            rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
      }
    p->nmaps = 4;
  }

  // ---- weight tensor map: [tap][Cout_pad][Cin_pad], box = one k-block of one n tile ----
  {
    const uint64_t dims[3] = {uint64_t(cin_pad), uint64_t(cout_pad), uint64_t(p->ntaps)};
    const uint64_t strides[2] = {uint64_t(cin_pad) * esz, uint64_t(cin_pad) * cout_pad * esz};
    const uint32_t box[3] = {uint32_t(rowb / 2), uint32_t(p->BN), 1};
    int rc = encode_map(&p->tmB, d->weight, 3, dims, strides, box, rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  // ---- output / residual tensor maps for the TMA-store epilogue (same pixel box as the tile, CW channels wide).  A ragged
  // last chunk is only allowed with a single N tile, where every column >= BN is also >= Cout and is clipped by the map.
  p->use_tma_store = 0;
  {
    int cw = 0;
    // (TMA clips the innermost dimension at 16-byte granularity: the slice width must be a multiple of 16 bytes)
    if (f32) { if ((p->BN % 32 == 0 || p->n_tiles == 1) && d->Cout % 4 == 0) cw = 32; }
    else if (d->Cout % 8 == 0) {
      if (p->BN % 64 == 0) cw = 64;
      else if (p->BN % 32 == 0) cw = 32;
      else if (p->n_tiles == 1) cw = p->BN > 32 ? 64 : 32;
      // K-heavy tiles: the epilogue has time to spare, shared memory does not -> narrow staging tiles.  The 64-channel halo
      // mode needs the room for a fourth activation stage (HBM latency x bandwidth ~ 64 KB in flight per SM).
      if (cw == 64 && (mode == 3 || paired || (!halo && (kiters >= 8 || p->BN >= 128))) && p->BN % 32 == 0 &&
          !(mode == 0 && env_int("DY_CONV_CW64_1X1", 0))) cw = 32;
      if (env_int("DY_CONV_CW", 0) == 32 && cw == 64 && p->BN % 32 == 0) cw = 32;
      if (d->pre_add && cw == 64 && p->BN % 32 == 0) cw = 32;   // the addend of a chunk is held in registers next to the accumulator
    }
    if (d->residual && f32) cw = 0;                            // fp32 + residual: generic path (not used by the model)
    if (mode == 4 && env_int("DY_K32_DIRECT", 0)) cw = 0;      // experiment: registers -> global instead of staging + TMA store
    if (fuse2) {
      p->fuse2 = 1; p->N2 = round_up(d->Cout2, 16); p->bias2 = d->bias2;
      {
        const uint64_t dims[3] = {uint64_t(kBlockK), uint64_t(p->N2), 1};
        const uint64_t strides[2] = {uint64_t(kBlockK) * 2, uint64_t(kBlockK) * p->N2 * 2};
        const uint32_t box[3] = {32, uint32_t(p->N2), 1};
        int rc = encode_map(&p->tmW2, d->weight2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
      if (d->tail_decode == 1 || d->tail_decode == 2) {
        // the threads store straight into the level's window of the channel-planar prediction tensor (B, 4+nc, A)
        p->tail_decode = d->tail_decode; p->y_nc = d->y_nc; p->y_stride = d->y_stride;
        p->y = d->y + d->y_anchor_off; p->y_A = d->y_A;
      } else if (d->tail_decode == 3) {
        p->tail_decode = 3;
        const uint64_t dims[4] = {uint64_t(d->Cout2), uint64_t(Wo), uint64_t(Ho), uint64_t(d->B)};
        const uint64_t strides[3] = {uint64_t(d->out2_ld) * 2, uint64_t(Wo) * d->out2_ld * 2, uint64_t(Ho) * Wo * d->out2_ld * 2};
        const uint32_t box[4] = {32, uint32_t(p->TW), uint32_t(p->TH), uint32_t(p->TB)};
        int rc = encode_map(&p->tmO2, d->out2, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (rc) return rc;
      } else {
        const uint64_t dims[4] = {uint64_t(d->Cout2), uint64_t(Wo), uint64_t(Ho), uint64_t(d->B)};
        const uint64_t strides[3] = {uint64_t(d->out2_ld) * 4, uint64_t(Wo) * d->out2_ld * 4, uint64_t(Ho) * Wo * d->out2_ld * 4};
        const uint32_t box[4] = {32, uint32_t(p->TW), uint32_t(p->TH), uint32_t(p->TB)};
        int rc = encode_map(&p->tmO2, d->out2, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        if (rc) return rc;
      }
      p->use_tma_store = 32;                                   // staging: two 8 KB tiles per group (the tail's A operand, then its fp32 output)
    } else if (cw) {
      const uint64_t oes = out_esz;
      const uint32_t obox[4] = {uint32_t(cw), uint32_t(p->TW), uint32_t(p->TH), uint32_t(p->TB)};
      const int rowo = cw * out_esz;
      const CUtensorMapSwizzle sw = rowo == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      uint64_t dims[4], strides[3], rstrides[3];
      if (flat) {
        const uint64_t M = uint64_t(p->Wo);
        dims[0] = uint64_t(d->Cout); dims[1] = M; dims[2] = 1; dims[3] = 1;
        strides[0] = uint64_t(d->out_ld) * oes; strides[1] = strides[2] = M * d->out_ld * oes;
        rstrides[0] = uint64_t(d->res_ld) * 2; rstrides[1] = rstrides[2] = M * d->res_ld * 2;
      } else {
        dims[0] = uint64_t(d->Cout); dims[1] = uint64_t(Wo); dims[2] = uint64_t(Ho); dims[3] = uint64_t(d->B);
        strides[0] = uint64_t(d->out_ld) * oes; strides[1] = uint64_t(Wo) * d->out_ld * oes; strides[2] = uint64_t(Ho) * Wo * d->out_ld * oes;
        rstrides[0] = uint64_t(d->res_ld) * 2; rstrides[1] = uint64_t(Wo) * d->res_ld * 2; rstrides[2] = uint64_t(Ho) * Wo * d->res_ld * 2;
      }
      int rc = encode_map(&p->tmO, d->out, 4, dims, strides, obox, sw, dt);
      if (rc) return rc;
      if (d->residual) {
        rc = encode_map(&p->tmR, d->residual, 4, dims, rstrides, obox, sw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (rc) return rc;
        p->has_res_tma = 1;
      }
      if (d->up_out) {
        // destination pixel (2y+dy, 2x+dx) <- source pixel (y, x): one strided view per (dy, dx)
        const uint64_t ues = 2, W2 = 2 * uint64_t(Wo), H2 = 2 * uint64_t(Ho);
        const uint64_t ustr[3] = {2 * uint64_t(d->up_ld) * ues, 2 * W2 * d->up_ld * ues, H2 * W2 * d->up_ld * ues};
        for (int u = 0; u < 4; ++u) {
          const char* ub = reinterpret_cast<const char*>(d->up_out) + (uint64_t(u >> 1) * W2 + (u & 1)) * d->up_ld * ues;
          rc = encode_map(&p->tmU[u], ub, 4, dims, ustr, obox, sw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
          if (rc) return rc;
        }
        p->has_up = 1;
      }
      p->use_tma_store = cw;
    }
  }
  if (d->up_out && !p->has_up) return fail(DY_ERR_UNSUPPORTED, "conv: up_out needs the TMA-store epilogue (Cout %% 8 == 0)");
  if (d->pre_add) {
    if (!(k == 1 && s == 1 && !fuse2 && !f32 && p->use_tma_store == 32 && d->Cout % 8 == 0 && Ho % 2 == 0 && Wo % 2 == 0 && d->pre_ld % 4 == 0 &&
          d->pre_ld >= d->Cout && (reinterpret_cast<uintptr_t>(d->pre_add) & 15) == 0))
      return fail(DY_ERR_UNSUPPORTED, "conv: pre_add needs ksize 1, stride 1, a bf16 output with Cout %% 8 == 0, even Ho / Wo and a 16B-aligned fp32 slice");
    if (p->TW % 2 || p->TH % 2 || (p->TW / 2) * (p->TH / 2) * p->TB > 32)
      return fail(DY_ERR_UNSUPPORTED, "conv: pre_add needs an even pixel tile (got %dx%dx%d)", p->TW, p->TH, p->TB);
    p->pre = d->pre_add; p->pre_ld = d->pre_ld;
    const uint64_t pdims[4] = {uint64_t(d->Cout), uint64_t(Wo / 2), uint64_t(Ho / 2), uint64_t(d->B)};
    const uint64_t pstr[3] = {uint64_t(d->pre_ld) * 4, uint64_t(Wo / 2) * d->pre_ld * 4, uint64_t(Ho / 2) * (Wo / 2) * d->pre_ld * 4};
    const uint32_t pbox[4] = {32, uint32_t(p->TW / 2), uint32_t(p->TH / 2), uint32_t(p->TB)};
    int rc = encode_map(&p->tmP, d->pre_add, 4, pdims, pstr, pbox, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc) return rc;
  }

  // ---- shared-memory plan ----
  // Small layers keep ALL their weights resident (loaded once per CTA): the per-stage traffic and TMA issue then only
  // cover the activation tile.  Everything else streams B next to A.
  const int cw = p->use_tma_store;
  // K-heavy generic tiles: one staging tile per group (the epilogue has slack), the room goes to fatter pipeline stages
  p->nbuf = (!fuse2 && ((!halo && kiters >= 8 && !env_int("DY_CONV_NBUF2", 0)) || (mode == 3 && p->BN > 64))) ? 1 : 2;   // the fused tail stages two tiles
  if (paired) p->nbuf = env_int("DY_PAIRED_NBUF", 2);
  const int b_tile = p->BN * rowb;
  const int b_all = p->ntaps * p->kblocks * b_tile + (fuse2 ? p->N2 * 128 : 0);
  const int bias_bytes = (fuse2 ? 576 : round_up(p->n_tiles * p->BN + 64, 4)) * 4;
  p->eg = ((mode == 4 || (mode == 3 && !fuse2 && env_int("DY_CONV_EG3_M3", 1))) && cw == 32 && !f32 && p->BN <= 64 && !env_int("DY_CONV_EG2", 0)) ? 3 : 2;   // epilogue groups
  // fused Detect tails: ~6000 cycles of serial epilogue chain per tile and group against a ~1750-cycle mainloop -> a third group
  if (fuse2 && mode == 3 && env_int("DY_TAIL_EG3", 1)) p->eg = 3;
  // the stride-2 conv with a hidden 1x1 + SiLU tail (layer 1 + C2f.cv1): SiLU, staging, tail GEMM, SiLU, two stores per tile and group
  if (fuse2 && mode == 5 && env_int("DY_S2TAIL_EG3", 0)) p->eg = 3;
  // 32-channel halo layers: a fourth group when asked for (640 threads)
  if (mode == 4 && p->eg == 3 && env_int("DY_K32_EG4", 0)) p->eg = 4;
  int staging = p->eg * p->nbuf * 128 * cw * out_esz;                       // groups x nbuf tiles
  // residual staging (see the kernel): the ring where it fits without starving the activation pipeline - the 32-channel halo mode
  // (18 KB of weights) and the paired halo mode; the 64-channel halo mode keeps 72 KB of weights resident and needs its four stages
  p->res_mode = !p->has_res_tma ? 0 : ((mode == 4 || paired) && !env_int("DY_NO_RES_RING", 0) ? 2 : (p->nbuf == 2 ? 1 : 0));
  const int ring_bytes = p->res_mode == 2 ? p->eg * 3 * 128 * cw * 2 : 0;
  const int pre_bytes = (d->pre_add ? p->eg * 16384 : 0) + ring_bytes;      // half-resolution addend: a ring of four 4 KB tiles per group
  if (p->nbuf == 1 && halo && !paired && !env_int("DY_CONV_NBUF1", 0)) {
    // wide halo tile (64 -> 128, 147 KB of resident weights): a single staging tile serialises every chunk behind the
    // previous chunk's store (~1700 cycles per 32-column chunk); take the second one whenever two halo stages still fit
    const int halo_stage = round_up(p->halo_pitch * kHaloRows * rowb, 1024);
    if (kMaxDynSmem - 1024 - 2 * staging - bias_bytes - b_all >= 2 * halo_stage) { p->nbuf = 2; staging *= 2; }
  }
  const int budget = kMaxDynSmem - 1024 - staging - pre_bytes - bias_bytes;
  if (paired) {
    // two rings: halo tiles (one per pixel tile and 64-channel block; a pair holds two while the next two load) and weight
    // tiles (one per tap and block, 8 MMAs each when paired): the weight ring takes what the halo ring leaves
    p->b_resident = 0;
    p->a_stage_bytes = round_up(p->halo_pitch * kHaloRows * rowb, 1024);
    p->a_stages = env_int("DY_PAIRED_ASTAGES", 4) & ~1;                      // even: each of the two MMA issuers keeps its own slot parity
    p->stage_bytes = b_tile;
    int stages = (budget - p->a_stages * p->a_stage_bytes) / p->stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    DY_CHECK_ARG(p->a_stages >= 2 && p->a_stages <= kMaxStages && stages >= 2, "conv: paired halo mode does not fit shared memory (BN %d)", p->BN);
    p->stages = stages;
    p->kps = 1;
  } else if (halo) {
    p->b_resident = 1;
    p->stage_bytes = round_up(p->halo_pitch * kHaloRows * rowb, 1024);
    int stages = (budget - b_all) / p->stage_bytes;
    DY_CHECK_ARG(stages >= 2, "conv: halo mode does not fit shared memory (BN %d)", p->BN);
    p->stages = stages > kMaxStages ? kMaxStages : stages;
    const int cap = env_int("DY_HALO_STAGES", 0);
    if (cap >= 2 && cap < p->stages) p->stages = cap;
    // two MMA issuers take alternate tiles: with an EVEN ring every stage (and accumulator) always belongs to the same
    // issuer, which the mbarrier parity test needs (a waiter may be at most one phase ahead of the barrier)
    p->stages &= ~1;
  } else {
    p->b_resident = (p->n_tiles == 1 && b_all <= budget - env_int("DY_BRES_MIN_STAGES", 3) * kABytes && !env_int("DY_NO_BRES", 0)) ? 1 : 0;   // leave room for >= 3 activation stages (the 64 -> 128 stride-2 layer keeps its 147 KB of weights resident: 91 -> 83 us)
    if (fuse2 && !p->b_resident) return fail(DY_ERR_UNSUPPORTED, "conv: fused 1x1 tail needs resident weights");
    const int per_k = a_blk + (p->b_resident ? 0 : b_tile);
    const int avail = budget - (p->b_resident ? b_all : 0);
    // k-blocks per stage: the issuing thread pays ~300 cycles per stage (barrier wait, commit), a k-block is 4 x max(BN/2,
    // 32 + BN/4) cycles of MMA: fatten the stages of narrow tiles until a stage holds >= ~512 cycles, keeping >= 3 stages.
    int kps = env_int("DY_CONV_KPS", 0);
    if (kps < 1) kps = mode == 5 ? 3 : (p->BN >= 256 ? 1 : 2);
    if (kps > (mode == 5 ? 3 : 2)) kps = mode == 5 ? 3 : 2;
    while (kps > 1 && (kps > kiters || avail / (kps * per_k) < 3)) --kps;
    p->kps = kps;
    p->stage_bytes = kps * per_k;
    int stages = avail / p->stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    DY_CHECK_ARG(stages >= 2, "conv: pipeline does not fit shared memory (BN %d)", p->BN);
    p->stages = stages;
  }
  // accumulator stages: BN columns apart; a ragged last chunk reads up to CW-1 columns past its stage
  {
    const int overrun = cw ? (ceil_div(p->BN, cw) * cw - p->BN) : (ceil_div(p->BN, 32) * 32 - p->BN);
    int nacc = (kTmemCols - overrun) / p->BN;
    if (nacc > kMaxAcc) nacc = kMaxAcc;
    // (the fused tail's second accumulator overwrites the tile's own stage: no columns are set aside for it)
    if (halo) nacc &= ~1;
    if (paired) nacc = nacc >= 4 ? 4 : 0;                                   // the issuer addresses accumulator slots as tile & 3
    // Two MMA issuers on alternate tiles + EG epilogue groups on every EG-th tile: the mbarrier parity protocol (a waiter may be
    // at most ONE phase ahead) holds only if every stage always meets the same issuer and the same group, i.e. the ring is a
    // multiple of lcm(2, EG).  A single issuer completes tiles in order, which bounds every group's lead by itself.
    if (p->eg == 3 && halo) nacc = nacc >= 6 ? 6 : 0;
    if (p->eg == 4) nacc = nacc >= 8 ? 8 : 0;
    DY_CHECK_ARG(nacc >= 2, "conv: BN %d leaves fewer than two accumulator stages", p->BN);
    p->nacc = nacc;
  }
  p->stg_off = (p->b_resident ? b_all : 0) + p->stages * p->stage_bytes + (paired ? p->a_stages * p->a_stage_bytes : 0);
  p->res_off = p->stg_off + staging;
  p->pre_off = p->res_off + ring_bytes;
  p->bias_off = p->res_off + pre_bytes;
  l->smem_bytes = p->bias_off + bias_bytes + 1024;
  const int total = p->m_tiles * p->n_tiles;
  if (p->n_split > 1) {
    const int per = sms / p->n_split < p->m_tiles ? sms / p->n_split : p->m_tiles;
    l->grid = per * p->n_split;
  } else {
    l->grid = total < sms ? total : sms;
  }
  p->dbg = env_int("DY_CONV_DBG", 0);
  { const char* e = getenv("DY_CONV_TRACE"); p->trace = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  if (env_int("DY_CONV_VERBOSE", 0))
    fprintf(stderr, "conv k%d s%d %d->%d %dx%dx%d: mode %d BN %d n_tiles %d n_split %d tile %dx%dx%d m_tiles %d kps %d stages %d x %d B nacc %d bres %d cw %d res_tma %d smem %d grid %d\n",
            k, s, d->Cin, d->Cout, d->B, d->H, d->W, p->mode, p->BN, p->n_tiles, p->n_split, p->TW, p->TH, p->TB, p->m_tiles, p->kps,
            p->stages, p->stage_bytes, p->nacc, p->b_resident, cw * 10 + p->nbuf, p->has_res_tma, l->smem_bytes, l->grid);
  return DY_OK;
}

template <int MODE, int CW, bool F32, bool FUSE2 = false, int EG = 2>
static int conv_launch_t(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream) {
  static unsigned long long seen = 0;                        // one bit per device: the opt-in is a per-device attribute
  if (first_use_on_device(&seen)) {
    // 227 KB opt-in limit covers static + dynamic shared memory; the kernel's static part (barriers) is < 1 KB
    DY_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<MODE, CW, F32, FUSE2, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  }
  static const bool use_pdl = (getenv("DY_NO_PDL") == nullptr);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(l->grid);
  cfg.blockDim = dim3(128 + 128 * EG);
  cfg.dynamicSmemBytes = l->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  DY_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<MODE, CW, F32, FUSE2, EG>, *p));
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
  switch (p->mode) {
    case 0: return conv_launch_m<0>(p, l, stream);
    case 1: return conv_launch_m<1>(p, l, stream);
    case 2: return conv_launch_m<2>(p, l, stream);
    case 3: return p->fuse2 ? (p->eg == 3 ? conv_launch_t<3, 32, false, true, 3>(p, l, stream) : conv_launch_t<3, 32, false, true>(p, l, stream))
                            : (p->eg == 3 ? conv_launch_t<3, 32, false, false, 3>(p, l, stream) : conv_launch_m<3>(p, l, stream));
    case 4: return p->eg == 4 ? conv_launch_t<4, 32, false, false, 4>(p, l, stream)
                              : (p->eg == 3 ? conv_launch_t<4, 32, false, false, 3>(p, l, stream) : conv_launch_m<4>(p, l, stream));
    case 6: return conv_launch_t<6, 32, false>(p, l, stream);
    default: return p->fuse2 ? (p->eg == 3 ? conv_launch_t<5, 32, false, true, 3>(p, l, stream) : conv_launch_t<5, 32, false, true>(p, l, stream))
                             : conv_launch_m<5>(p, l, stream);
  }
}

}  // namespace dy

extern "C" int dy_conv2d(const dy_conv_desc* d, void* stream) {
  dy::ConvParams p; dy::ConvLaunch l;
  int rc = dy::conv_build_params(d, &p, &l);
  if (rc) return rc;
  return dy::conv_launch(&p, &l, static_cast<cudaStream_t>(stream));
}
