// nms.cu — batched, bit-exact NMS for the Detect output.
// Replaces ops.non_max_suppression (ultralytics/utils/ops.py:181-332) and the torchvision.ops.nms call inside it
// (:312): ~12 aten launches + 2 host syncs per image in the reference become two launches per batch.
//
//  kernel 1  nms_filter_kernel   grid (chunks of 2048 anchors, B)
//            confidence test `amax(cls) > conf` (:250) in fp32 like torch's scalar compare, best class with the
//            first-max tie rule (:290) or all (anchor, class) pairs for multi_label (:286-288), optional class
//            filter (:294-295), optional in-place xywh->xyxy of the prediction (:259-260).  Survivors are
//            compacted IN ANCHOR ORDER inside the chunk (warp shuffles + one block scan) as unique 64-bit keys
//                 key = (~score_bits << 32) | (anchor*nc + cls)
//            so ascending key order == descending score, ties broken by the lower candidate index: the order
//            torchvision's stable descending sort produces.
//  kernel 2  nms_select_kernel   one CTA per image
//            repeatedly (a) selects the next <= 4096 smallest keys: one histogram pass over score bins spread over
//            (conf, 1] when the bin holding the wanted rank is small enough to take whole, else exact MSD radix passes,
//            (b) bitonic-sorts them, (c) builds class-offset boxes `xyxy + cls*max_wh` (:305-311) exactly as fp32
//            torch ops do (no FMA contraction), (d) suppresses them against the boxes kept so far, (e) builds the
//            512x512 upper-triangular IoU>thr bitmask in shared memory and (f) lets one warp do the greedy scan.
//            Stops at max_det kept boxes (:313) or after max_nms candidates (:301-302); writes rows
//            (x1,y1,x2,y2,conf,cls), the per-image count and (optionally) the reference's kept indices.
#include "dy_common.cuh"
#include <cstring>

namespace dy {

static constexpr int kFilterThreads = 256;
static constexpr int kChunk = 2048;            // anchors per filter CTA (2 groups x 256 threads x 4 anchors)
static constexpr int kSelThreads = 512;
static constexpr int kK = 512;                 // candidates per greedy round (== kSelThreads)
static constexpr int kWords = kK / 32;
static constexpr int kBins = 2048;
static constexpr int kBig = 4096;               // keys selected + sorted per super-round
static constexpr int kFirst = 1024;             // ... of the first super-round: the bitonic sort costs n log^2 n (2048 keys: 31 k clocks, 1024:
                                                // 13 k) and most images reach max_det within their first few hundred ranks
static constexpr int kMaxPasses = 8;
static constexpr int kMaxClassWords = 32;      // class filter bitmask: nc <= 1024

struct NmsParams {
  const float* pred; float* pred_rw;
  int B, nc, A, nchunks;
  float conf; float iou_f; int iou_inclusive;
  float iou_c;                                 // (1 - 2e-5) * thr / (1 + thr): screening factor, see box_inter / iou_decide (NaN: always decide exactly)
  int max_det, max_nms; float max_wh;
  int agnostic, multi_label, in_place, has_class_filter;
  uint32_t class_mask[kMaxClassWords];
  unsigned long long cap;                      // candidate capacity per image
  unsigned long long* keys;                    // [B][cap]
  int* img_count;                              // [B]
  unsigned int bin_base; int bin_shift;        // score bins of the one-pass selection: bin = (~score_bits - bin_base) >> bin_shift
  int* chunk_base;                             // [B][nchunks]
  int* chunk_cnt;                              // [B][nchunks]
  float* out; int* counts; long long* kept;
  const float* rescale;                        // optional [B][8]: pad_x, pad_y, gain, w0, h0 (scale_boxes + clip_boxes in the output phase)
  int npasses; int pass_shift[kMaxPasses]; int pass_bits[kMaxPasses];
};

__device__ __forceinline__ bool class_allowed(const NmsParams& p, int c) {
  return !p.has_class_filter || ((p.class_mask[c >> 5] >> (c & 31)) & 1u);
}

// ------------------------------------------------------------------------------------------------
// kernel 1: filter + order-preserving compaction per chunk
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFilterThreads) nms_filter_kernel(const __grid_constant__ NmsParams p) {
  __shared__ int warp_tot[kFilterThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a_chunk = chunk * kChunk;
  const size_t A = p.A;
  const float* img = p.pred + static_cast<size_t>(b) * (4 + p.nc) * A;
  const bool vec_ok = (p.A & 3) == 0;

  // pass 1: per-thread candidate counts for its 2 groups of 4 consecutive anchors.  The class planes are read four at a
  // time for both groups before any value is used: eight independent 16-byte loads per thread in flight (one load at a
  // time, as a plain loop over the classes compiles to, left the kernel at 66 % of the HBM bandwidth).
  float best[8]; int bestc[8]; int cnt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bestc[k] = 0; cnt[k] = 0; }
  const int a0g[2] = {a_chunk + tid * 4, a_chunk + (kChunk / 2) + tid * 4};
  for (int c0 = 0; c0 < p.nc; c0 += 4) {
    float v[2][4][4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        const bool on = a0g[g] < p.A && c < p.nc;
        const float* src = img + (4 + (on ? c : 0)) * A + (on ? a0g[g] : 0);
        if (vec_ok) {
          float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          if (on) t = __ldg(reinterpret_cast<const float4*>(src));
          v[g][j][0] = t.x; v[g][j][1] = t.y; v[g][j][2] = t.z; v[g][j][3] = t.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[g][j][k] = (on && a0g[g] + k < p.A) ? __ldg(src + k) : -INFINITY;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        if (c >= p.nc) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.multi_label) {
            if (v[g][j][k] > p.conf && class_allowed(p, c)) cnt[g * 4 + k]++;
          } else if (v[g][j][k] > best[g * 4 + k]) { best[g * 4 + k] = v[g][j][k]; bestc[g * 4 + k] = c; }
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int a0 = a0g[g];
    if (a0 < p.A) {
      if (!p.multi_label) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          cnt[g * 4 + k] = (best[g * 4 + k] > p.conf && class_allowed(p, bestc[g * 4 + k])) ? 1 : 0;
      }
      if (p.in_place) {   // prediction[..., :4] = xywh2xyxy(prediction[..., :4])  (ops.py:259-260, :432-449)
        float* rw = p.pred_rw + static_cast<size_t>(b) * (4 + p.nc) * A;
        if (vec_ok) {
          float4* q = reinterpret_cast<float4*>(rw + a0);
          const size_t A4 = A >> 2;
          const float4 cx = q[0], cy = q[A4], w = q[2 * A4], h = q[3 * A4];
          const float hw[4] = {__fmul_rn(w.x, 0.5f), __fmul_rn(w.y, 0.5f), __fmul_rn(w.z, 0.5f), __fmul_rn(w.w, 0.5f)};
          const float hh[4] = {__fmul_rn(h.x, 0.5f), __fmul_rn(h.y, 0.5f), __fmul_rn(h.z, 0.5f), __fmul_rn(h.w, 0.5f)};
          q[0] = make_float4(__fsub_rn(cx.x, hw[0]), __fsub_rn(cx.y, hw[1]), __fsub_rn(cx.z, hw[2]), __fsub_rn(cx.w, hw[3]));
          q[A4] = make_float4(__fsub_rn(cy.x, hh[0]), __fsub_rn(cy.y, hh[1]), __fsub_rn(cy.z, hh[2]), __fsub_rn(cy.w, hh[3]));
          q[2 * A4] = make_float4(__fadd_rn(cx.x, hw[0]), __fadd_rn(cx.y, hw[1]), __fadd_rn(cx.z, hw[2]), __fadd_rn(cx.w, hw[3]));
          q[3 * A4] = make_float4(__fadd_rn(cy.x, hh[0]), __fadd_rn(cy.y, hh[1]), __fadd_rn(cy.z, hh[2]), __fadd_rn(cy.w, hh[3]));
        } else
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int a = a0 + k;
          if (a < p.A) {
            const float cx = rw[a], cy = rw[A + a], hw = __fmul_rn(rw[2 * A + a], 0.5f), hh = __fmul_rn(rw[3 * A + a], 0.5f);
            rw[a] = __fsub_rn(cx, hw); rw[A + a] = __fsub_rn(cy, hh);
            rw[2 * A + a] = __fadd_rn(cx, hw); rw[3 * A + a] = __fadd_rn(cy, hh);
          }
        }
      }
    }
  }

  // block exclusive scan over (group, thread) order
  int tot[2] = {cnt[0] + cnt[1] + cnt[2] + cnt[3], cnt[4] + cnt[5] + cnt[6] + cnt[7]};
  int offs[2]; int running = 0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    int incl = tot[g];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += n; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kFilterThreads / 32; ++w) { const int t = warp_tot[w]; if (w < warp) wbase += t; all += t; }
    offs[g] = running + wbase + incl - tot[g];
    running += all;
    __syncthreads();
  }
  if (tid == 0) {
    const int base = running ? atomicAdd(&p.img_count[b], running) : 0;
    s_base = base;
    p.chunk_base[b * p.nchunks + chunk] = base;
    p.chunk_cnt[b * p.nchunks + chunk] = running;
  }
  __syncthreads();
  if (running == 0) return;
  unsigned long long* dst = p.keys + static_cast<size_t>(b) * p.cap + s_base;

  // pass 2: write keys in order
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int a0 = a_chunk + g * (kChunk / 2) + tid * 4;
    int o = offs[g];
    if (tot[g] == 0) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (cnt[g * 4 + k] == 0) continue;
      const int a = a0 + k;
      if (p.multi_label) {
        for (int c = 0; c < p.nc; ++c) {
          const float v = __ldg(img + (4 + c) * A + a);
          if (v > p.conf && class_allowed(p, c)) {
            dst[o++] = (static_cast<unsigned long long>(~__float_as_uint(v)) << 32) |
                       static_cast<unsigned long long>(static_cast<unsigned>(a) * static_cast<unsigned>(p.nc) + c);
          }
        }
      } else {
        dst[o++] = (static_cast<unsigned long long>(~__float_as_uint(best[g * 4 + k])) << 32) |
                   static_cast<unsigned long long>(static_cast<unsigned>(a) * static_cast<unsigned>(p.nc) + bestc[g * 4 + k]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: per-image radix select + sort + greedy suppression
// ------------------------------------------------------------------------------------------------
// torchvision's test `inter / (area_i + area_j - inter) > iou_threshold` in fp32 (CPU kernel: the float quotient is
// compared against the double threshold, which is `>= float(thr)` when float(thr) rounds up and `> float(thr)` else).
__device__ __forceinline__ bool iou_suppresses(const float4& a, float area_a, const float4& b, float area_b,
                                               float thr, int inclusive) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  if (!(inter > 0.f)) return false;           // 0/x == 0 and 0/0 == NaN never exceed a threshold in [0,1]
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  // The IEEE division costs ~10 instructions; the approximate quotient (2 ulp) decides every pair that is not within a
  // relative 1e-5 of the threshold, the exact one only the rest (same result as always dividing exactly).
  const float q = __fdividef(inter, uni);
  if (q > thr * 1.00001f) return true;
  if (q < thr * 0.99999f) return false;
  const float ovr = __fdiv_rn(inter, uni);
  return inclusive ? (ovr >= thr) : (ovr > thr);
}

// Division-free screening for the two hot loops (they are instruction-issue-bound: 512 threads share one SM).
//   IoU > thr  <=>  inter > thr * (area_a + area_b - inter)  <=>  inter > c * (area_a + area_b),  c = thr / (1 + thr).
// Every box carries sl = c * (1 - 2e-5) * area (NaN when the area is not a positive finite number), so one add gives the
// lower bound S of the undecided band.  inter < S: the pair certainly does not suppress (11 instructions per pair).
// Otherwise (rare) iou_decide looks again: inter > S * (1 + 4e-5) certainly suppresses; inside the band, or when S is NaN
// (degenerate boxes, thr ~ 0), the exact IEEE division of iou_suppresses decides.  The rounding of S, of c and of the
// reference's own fp32 quotient are all below 1e-6 relative, so outside the band the exact test gives the same answer.
static constexpr float kScreenLo = 0.99998f, kScreenHiOverLo = 1.00004f;

__device__ __forceinline__ float box_inter(const float4& a, const float4& b) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  return __fmul_rn(w, h);
}

__device__ __forceinline__ float screen_area(float area, float c_lo) {
  return (area > 0.f && area < 3.0e38f) ? __fmul_rn(area, c_lo) : __int_as_float(0x7fc00000);
}

// second look at a pair the screen could not rule out
__device__ __forceinline__ bool iou_decide(const float4& a, float area_a, const float4& b, float area_b, float inter, float S,
                                           float thr, int inclusive) {
  if (inter > __fmul_rn(S, kScreenHiOverLo)) return true;
  return iou_suppresses(a, area_a, b, area_b, thr, inclusive);
}

#ifdef DY_CONV_DEBUG
// debug builds only: per-image clock totals of the select kernel's phases (tools/trace_nms.py)
__device__ unsigned long long g_nms_trace[1024 * 16];
#define NMS_T0() long long t_ph = clock64()
#define NMS_TP(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long t_now = clock64(); g_nms_trace[(b & 1023) * 16 + (k)] += t_now - t_ph; t_ph = t_now; } } while (0)
#define NMS_TC(k, v) do { if (threadIdx.x == 0) g_nms_trace[(b & 1023) * 16 + (k)] += (v); } while (0)
#else
#define NMS_T0()
#define NMS_TP(k)
#define NMS_TC(k, v)
#endif

struct SelShared {
  float4 box[kK];
  float area[kK];
  float sarea[kK];                 // iou_c * area (screening form); 16-byte aligned
  unsigned int hist[kBins];
  unsigned int mask[kWords * (kK + 1)];
  unsigned int remv[kWords];
  unsigned int undw[kWords], keptw[kWords];   // greedy fixed point: undecided / kept candidates, one word per warp
  int scan_tmp[kSelThreads / 32];
  int max_tmp[kSelThreads / 32];
  unsigned short kidx[kK];
  int sel_count;
  int kept;
  int found_bin; unsigned int found_below; unsigned int found_count;
};

static constexpr size_t kSelSharedBytes = (sizeof(SelShared) + 15) & ~size_t(15);
__host__ __device__ inline int kept_pad(int max_det) { return (max_det + 4 + 3) & ~3; }
__host__ __device__ inline size_t kept_bytes(int max_det) { return (static_cast<size_t>(kept_pad(max_det)) * (16 + 4 + 4 + 4 + 8) + 15) & ~size_t(15); }

__device__ __forceinline__ void load_offset_box(const NmsParams& p, const float* img, unsigned long long key,
                                                float4* box, float* area) {
  const unsigned id = static_cast<unsigned>(key & 0xffffffffull);
  const unsigned a = id / static_cast<unsigned>(p.nc), cls = id - a * static_cast<unsigned>(p.nc);
  const size_t A = p.A;
  float x1, y1, x2, y2;
  if (p.in_place) { x1 = img[a]; y1 = img[A + a]; x2 = img[2 * A + a]; y2 = img[3 * A + a]; }
  else {
    const float cx = img[a], cy = img[A + a], hw = __fmul_rn(img[2 * A + a], 0.5f), hh = __fmul_rn(img[3 * A + a], 0.5f);
    x1 = __fsub_rn(cx, hw); y1 = __fsub_rn(cy, hh); x2 = __fadd_rn(cx, hw); y2 = __fadd_rn(cy, hh);
  }
  const float c = __fmul_rn(static_cast<float>(cls), p.agnostic ? 0.f : p.max_wh);   // x[:, 5:6] * max_wh
  box->x = __fadd_rn(x1, c); box->y = __fadd_rn(y1, c); box->z = __fadd_rn(x2, c); box->w = __fadd_rn(y2, c);
  *area = __fmul_rn(__fsub_rn(box->z, box->x), __fsub_rn(box->w, box->y));
}

__device__ __forceinline__ unsigned int score_bin(const NmsParams& p, unsigned long long key) {
  const unsigned int u = static_cast<unsigned int>(key >> 32);
  if (u <= p.bin_base) return 0u;
  const unsigned int v = (u - p.bin_base) >> p.bin_shift;
  return v < static_cast<unsigned>(kBins - 1) ? v : static_cast<unsigned>(kBins - 1);
}

__global__ void __launch_bounds__(kSelThreads) nms_select_kernel(const __grid_constant__ NmsParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SelShared& s = *reinterpret_cast<SelShared*>(smem_raw);
  const int b = blockIdx.x;
  // kept-box state lives behind SelShared: boxes, screening areas, areas, ranks, keys (mdp entries each; the entries past
  // the kept count hold a zero box and sarea = +inf, which the screen always rules out: the loop below reads 4 at a time)
  const int mdp = kept_pad(p.max_det);
  float4* kbox = reinterpret_cast<float4*>(smem_raw + kSelSharedBytes);
  float* ksarea = reinterpret_cast<float*>(kbox + mdp);
  float* karea = ksarea + mdp;
  int* krank = reinterpret_cast<int*>(karea + mdp);
  unsigned long long* kkey = reinterpret_cast<unsigned long long*>(krank + mdp);
  unsigned long long* big = reinterpret_cast<unsigned long long*>(smem_raw + kSelSharedBytes + kept_bytes(p.max_det));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.img_count[b];
  const float* img = (p.in_place ? p.pred_rw : p.pred) + static_cast<size_t>(b) * (4 + p.nc) * static_cast<size_t>(p.A);
  const unsigned long long* keys = p.keys + static_cast<size_t>(b) * p.cap;
  const int limit = min(n, p.max_nms);
  if (tid == 0) s.kept = 0;
  for (int i = tid; i < mdp; i += kSelThreads) { kbox[i] = make_float4(0.f, 0.f, 0.f, 0.f); ksarea[i] = __int_as_float(0x7f800000); }
  __syncthreads();

  // Are the scores spread over the bins of the one-pass selection?  A sample of 512 keys: if any warp finds 8 of its 32
  // keys in one bin, the image has near-identical scores by the thousand (e.g. an untrained head) and takes the exact radix
  // passes with warp-aggregated histograms, first super-round of 2048 - the path tuned for clustered, heavily suppressed boxes.
  bool spread;
  {
    unsigned int bin = 0xffffffffu - static_cast<unsigned>(lane);         // distinct dummies for lanes without a key
    if (tid < n) bin = score_bin(p, keys[tid]);
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    spread = !__syncthreads_or(__popc(peers) >= 8);
  }
  NMS_T0();
  NMS_TC(11, n);
  int processed = 0;
  unsigned long long prev_T = 0ull;            // keys taken so far are exactly the keys <= prev_T
  bool first = true;
  bool done = false;
  while (processed < limit && !done) {
    // =============== super-round: the next SK (<= 4096) smallest keys, selected once and sorted in shared memory ===============
    // first super-round: everything when the image has at most 2048 candidates (one sort, rounds of 512: the clustered,
    // heavily suppressed images need all of them anyway), else the top ~1024
    const int cap = first ? ((!spread || n <= kBig / 2) ? kBig / 2 : kFirst) : kBig;
    int SK = min(cap, limit - processed);
    unsigned long long prefix_val = 0ull, prefix_mask = 0ull;
    unsigned int k_rem = static_cast<unsigned>(SK);
    const bool take_all = (n - processed) <= SK;          // everything left fits: no selection needed
    if (take_all) SK = n - processed;
    bool fast = false, bins_ok = false;
    if (spread && (!take_all || SK <= kBig / 2)) {
      // ---- one-pass selection: histogram of kBins score bins spread over (conf, 1]; the bin that holds rank SK is taken
      //      WHOLE when everything up to it fits the sort buffer (a super-round may be any size; only the max_nms limit is
      //      exact).  Identical scores by the thousand (one crowded bin) fall through to the exact radix passes below.
      for (int i = tid; i < kBins; i += kSelThreads) s.hist[i] = 0u;
      __syncthreads();
      for (int i0 = 0; i0 < n; i0 += kSelThreads * 4) {
        unsigned long long key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kSelThreads + tid;
          key[u] = (i < n) ? keys[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kSelThreads + tid;
          // plain shared-memory atomics: only taken for well spread scores (see `spread` above), where the keys of a warp
          // rarely share a bin and a match.any aggregation costs more than the conflicts it saves
          if (i < n && (first || key[u] > prev_T)) atomicAdd(&s.hist[score_bin(p, key[u])], 1u);
        }
      }
      __syncthreads();
      const unsigned h0 = s.hist[tid * 4], h1 = s.hist[tid * 4 + 1], h2 = s.hist[tid * 4 + 2], h3 = s.hist[tid * 4 + 3];
      const int mine = static_cast<int>(h0 + h1 + h2 + h3);
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      unsigned int hmax = max(max(h0, h1), max(h2, h3));                 // most crowded bin (the counting sort below is
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, d));   // quadratic inside a bin)
      if (lane == 31) { s.scan_tmp[warp] = incl; s.max_tmp[warp] = static_cast<int>(hmax); }
      __syncthreads();
      int wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += s.scan_tmp[w];
      int crowd = 0;
      for (int w = 0; w < kSelThreads / 32; ++w) crowd = max(crowd, s.max_tmp[w]);
      bins_ok = crowd <= 192;
      const unsigned int excl = static_cast<unsigned>(wbase + incl - mine);
      if (!take_all && k_rem > excl && k_rem <= excl + static_cast<unsigned>(mine)) {
        unsigned int below = excl; int bin = tid * 4; unsigned int cnt = h0;
        if (k_rem > below + h0) { below += h0; bin++; cnt = h1;
          if (k_rem > below + h1) { below += h1; bin++; cnt = h2;
            if (k_rem > below + h2) { below += h2; bin++; cnt = h3; } } }
        s.found_bin = bin; s.found_below = below; s.found_count = cnt;
      }
      __syncthreads();
      // the rank-SK key sits in bin `found_bin`: take the bins BELOW it (slightly fewer than SK keys, so the sort size
      // stays at the power of two SK was chosen for) unless they hold less than half of what was asked for (a crowded bin:
      // then the bin itself too if everything fits, else the exact passes)
      const unsigned int below = s.found_below, upto = s.found_below + s.found_count;
      const unsigned int room = static_cast<unsigned>(min(cap, limit - processed));
      int last_bin = -1;
      if (take_all) last_bin = kBins - 1;                                  // every live key, T = ~0: the bins only serve the sort
      else if (2u * below >= static_cast<unsigned>(SK)) { last_bin = s.found_bin - 1; SK = static_cast<int>(below); }
      else if (upto <= room) { last_bin = s.found_bin; SK = static_cast<int>(upto); }
      if (last_bin >= 0) {
        fast = true;
        // counting sort by bin (below): start offset of every bin, and a cursor per bin that the gather advances
        unsigned int* start = s.mask;                                      // kBins + 1 words of the (idle) pair matrix
        const unsigned int st[4] = {excl, excl + h0, excl + h0 + h1, excl + h0 + h1 + h2};
#pragma unroll
        for (int k = 0; k < 4; ++k) { start[tid * 4 + k] = st[k]; s.hist[tid * 4 + k] = st[k]; }
        if (tid == kSelThreads - 1) start[kBins] = excl + static_cast<unsigned>(mine);
        if (last_bin >= kBins - 1) prefix_val = ~0ull;
        else {
          const unsigned long long umax = static_cast<unsigned long long>(p.bin_base) +
                                          ((static_cast<unsigned long long>(last_bin) + 1ull) << p.bin_shift) - 1ull;
          prefix_val = umax >= 0xffffffffull ? ~0ull : ((umax << 32) | 0xffffffffull);
        }
      }
      __syncthreads();                                                     // found_* are rewritten by the radix passes
    }
    if (!take_all && !fast) {
      for (int ps = 0; ps < p.npasses; ++ps) {
        const int shift = p.pass_shift[ps];
        const unsigned int dmask = (1u << p.pass_bits[ps]) - 1u;
        for (int i = tid; i < kBins; i += kSelThreads) s.hist[i] = 0u;
        __syncthreads();
        // 4 keys per thread per trip: the loads are independent, so their L2 latency overlaps
        for (int i0 = 0; i0 < n; i0 += kSelThreads * 4) {
          unsigned long long key[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kSelThreads + tid;
            key[u] = (i < n) ? keys[i] : 0ull;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kSelThreads + tid;
            unsigned int digit = 0xffffffffu;
            if (i < n) {
              const bool live = first ? true : (key[u] > prev_T);
              if (live && (key[u] & prefix_mask) == prefix_val) digit = static_cast<unsigned>(key[u] >> shift) & dmask;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit != 0xffffffffu && lane == (__ffs(peers) - 1)) atomicAdd(&s.hist[digit], __popc(peers));
          }
        }
        __syncthreads();
        // locate the bin holding the k_rem-th element: each thread owns 4 consecutive bins
        const unsigned h0 = s.hist[tid * 4], h1 = s.hist[tid * 4 + 1], h2 = s.hist[tid * 4 + 2], h3 = s.hist[tid * 4 + 3];
        const int mine = static_cast<int>(h0 + h1 + h2 + h3);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) s.scan_tmp[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s.scan_tmp[w];
        const unsigned int excl = static_cast<unsigned>(wbase + incl - mine);
        if (k_rem > excl && k_rem <= excl + static_cast<unsigned>(mine)) {
          unsigned int below = excl; int bin = tid * 4; unsigned int cnt = h0;
          if (k_rem > below + h0) { below += h0; bin++; cnt = h1;
            if (k_rem > below + h1) { below += h1; bin++; cnt = h2;
              if (k_rem > below + h2) { below += h2; bin++; cnt = h3; } } }
          s.found_bin = bin; s.found_below = below; s.found_count = cnt;
        }
        __syncthreads();
        prefix_val |= static_cast<unsigned long long>(static_cast<unsigned>(s.found_bin)) << shift;
        prefix_mask |= static_cast<unsigned long long>(dmask) << shift;
        k_rem -= s.found_below;
        const bool takes_whole_bin = (k_rem == s.found_count);
        __syncthreads();
        if (takes_whole_bin) {
          // every key sharing this prefix is wanted: the remaining (lower) digits need no selection.
          // Typical case: after the three score passes the K-th score is unique, so the two id passes are skipped.
          prefix_val |= (shift > 0) ? ((1ull << shift) - 1ull) : 0ull;   // only bits BELOW this digit (bits between the id and score fields are always 0)
          break;
        }
      }
    }
    const unsigned long long T = take_all ? ~0ull : prefix_val;
    NMS_TP(0); NMS_TC(9, 1);

    // ---- gather keys in (prev_T, T] into big[] in ascending order ----
    const bool csort = fast && bins_ok && SK <= kBig / 2;
    if (csort) {
      // Counting sort on the score bins of the one-pass selection: every taken key goes to the next free slot of its bin's
      // segment (upper half of big[] as scratch), then finds its place inside the segment by counting the smaller keys of
      // that segment (a bin holds a key or two; a crowded one costs its length per key, never correctness).  Three barrier
      // intervals instead of the ~30 of a 1024-key bitonic network (21 k clocks per image).
      unsigned long long* tmp = big + kBig / 2;
      const unsigned int* start = s.mask;
      for (int i0 = 0; i0 < n; i0 += kSelThreads * 4) {
        unsigned long long key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kSelThreads + tid;
          key[u] = (i < n) ? keys[i] : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kSelThreads + tid;
          if (i < n && (first || key[u] > prev_T) && key[u] <= T) {
            const unsigned int pos = atomicAdd(&s.hist[score_bin(p, key[u])], 1u);
            tmp[pos] = key[u];
            if (pos < static_cast<unsigned>(kK)) {               // a candidate of the first round: its four box values are ~1.5 us
              const unsigned a = static_cast<unsigned>(key[u] & 0xffffffffull) / static_cast<unsigned>(p.nc);   // of HBM latency away
#pragma unroll
              for (int c = 0; c < 4; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(img + static_cast<size_t>(c) * p.A + a));
            }
          }
        }
      }
      __syncthreads();
      NMS_TP(1);
      for (int q = tid; q < SK; q += kSelThreads) {
        const unsigned long long key = tmp[q];
        const unsigned int bin = score_bin(p, key);
        const unsigned int lo = start[bin], hi = start[bin + 1];
        unsigned int r = lo;
        for (unsigned int j = lo; j < hi; ++j) r += tmp[j] < key ? 1u : 0u;
        big[r] = key;
      }
      __syncthreads();
      NMS_TP(2);
    } else {
    int npow = kK;                                   // sort size: next power of two >= SK (>= 512)
    while (npow < SK) npow <<= 1;
    if (tid == 0) s.sel_count = 0;
    for (int i = tid; i < npow; i += kSelThreads) big[i] = ~0ull;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += kSelThreads * 4) {
      unsigned long long key[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSelThreads + tid;
        key[u] = (i < n) ? keys[i] : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSelThreads + tid;
        const bool live = first ? true : (key[u] > prev_T);
        const bool take = i < n && live && key[u] <= T;
        const unsigned int takers = __ballot_sync(0xffffffffu, take);      // one shared-memory atomic per warp and trip
        if (takers) {
          const int leader = __ffs(takers) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&s.sel_count, __popc(takers));
          base = __shfl_sync(0xffffffffu, base, leader);
          const int slot = base + __popc(takers & ((1u << lane) - 1u));
          if (take && slot < kBig) big[slot] = key[u];
        }
      }
    }
    __syncthreads();
    NMS_TP(1);
    // bitonic network, two strides per barrier: a thread takes the four elements {i, i+h, i+2h, i+3h} through the
    // compare-exchanges of strides 2h and h in registers (36 shared-memory round trips for 2048 keys instead of 66);
    // a stage with an odd number of strides starts with one plain pair step.
    for (int size = 2; size <= npow; size <<= 1) {
      int stride = size >> 1;
      if (__popc(size - 1) & 1) {
        for (int t = tid; t < (npow >> 1); t += kSelThreads) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const unsigned long long x = big[lo], y = big[hi];
          const bool up = (lo & size) == 0;
          if ((x > y) == up) { big[lo] = y; big[hi] = x; }
        }
        __syncthreads();
        stride >>= 1;
      }
      for (; stride >= 2; stride >>= 2) {
        const int h = stride >> 1;
        for (int g = tid; g < (npow >> 2); g += kSelThreads) {
          const int i0 = ((g & ~(h - 1)) << 2) | (g & (h - 1));
          unsigned long long a = big[i0], b = big[i0 + h], c = big[i0 + 2 * h], d = big[i0 + 3 * h];
          const bool up = (i0 & size) == 0;
          unsigned long long t;
          if ((a > c) == up) { t = a; a = c; c = t; }
          if ((b > d) == up) { t = b; b = d; d = t; }
          if ((a > b) == up) { t = a; a = b; b = t; }
          if ((c > d) == up) { t = c; c = d; d = t; }
          big[i0] = a; big[i0 + h] = b; big[i0 + 2 * h] = c; big[i0 + 3 * h] = d;
        }
        __syncthreads();
      }
    }
    NMS_TP(2);
    }
    // =============== greedy rounds of 512 candidates in sorted order ===============
    for (int off = 0, K = 0; off < SK; off += K) {
      // Spread scores: the first round takes what the image can use if nothing were suppressed (max_det plus a quarter; the
      // pair tests grow with K^2), every other round a full block (a round that falls short costs more than it saved).
      K = min(kK, SK - off);
      if (spread && processed + off == 0) K = min(K, (p.max_det + (p.max_det >> 2) + 32 + 31) & ~31);
      NMS_TC(10, 1);
      // ---- (c) class-offset boxes ----
      const bool have = tid < K;
      const unsigned long long mykey = have ? big[off + tid] : ~0ull;
      float4 mybox = make_float4(0.f, 0.f, 0.f, 0.f); float myarea = 0.f;
      float mysa = 0.f;
      if (have) {
        load_offset_box(p, img, mykey, &mybox, &myarea);
        mysa = screen_area(myarea, p.iou_c);
        s.box[tid] = mybox; s.area[tid] = myarea; s.sarea[tid] = mysa;
      }
      NMS_TP(3);
      // ---- (d) suppress against boxes kept in earlier rounds ----
      bool dead = !have;
      const int kept0 = s.kept;
      if (have) {
        // four kept boxes per trip, screened independently (the arrays are padded: no bounds test)
        for (int i = 0; i < kept0 && !dead; i += 4) {
          const float4 sl4 = *reinterpret_cast<const float4*>(&ksarea[i]);
          const float sl[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
          float inter[4], S[4];
          bool maybe = false;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            inter[u] = box_inter(kbox[i + u], mybox);
            S[u] = __fadd_rn(sl[u], mysa);
            maybe |= !(inter[u] < S[u]);
          }
          if (maybe) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (!(inter[u] < S[u]) && i + u < kept0 &&
                  iou_decide(kbox[i + u], karea[i + u], mybox, myarea, inter[u], S[u], p.iou_f, p.iou_inclusive)) dead = true;
          }
        }
      }
      const unsigned dead_bits = __ballot_sync(0xffffffffu, dead);
      if (lane == 0) { s.remv[warp] = dead_bits; s.undw[warp] = ~dead_bits; s.keptw[warp] = 0u; }
      __syncthreads();
      NMS_TP(4);
      // ---- (e) suppression pairs (i suppresses j, i < j) among the candidates that survived (d), stored by COLUMN:
      //      bit (i & 31) of mask[i >> 5][j] says "row i suppresses j" (what the fixed point of (f) needs).  A unit of work is
      //      (column j, 32-row word w) with 32w <= j: the thread tests the rows of word w against box j and writes the word with ONE
      //      plain store - no atomics and no zeroing pass (every word (f) reads, w <= j >> 5, is written by exactly one unit).
      //      Row-major units (row i against a word of 32 columns) needed an atomicOr per suppressed pair: on clustered boxes,
      //      where almost every pair suppresses, that was 60 % of the kernel.  Row word w meets columns 32w .. K-1, i.e. with
      //      v = nwords-1-w there are 32(v+1) units and 16v(v+1) precede them; units are dealt round-robin over the block.
      {
        const int nwords = (K + 31) / 32;
        const int units = 16 * nwords * (nwords + 1);
        int v = 0;
        for (int u = tid; u < units; u += kSelThreads) {
          while (16 * (v + 1) * (v + 2) <= u) ++v;                      // u only grows: v is carried
          const int w = nwords - 1 - v;
          const int j = w * 32 + (u - 16 * v * (v + 1));
          if (j >= K) continue;
          unsigned int bits = 0u;
          const int i0 = w * 32;
          if (!((s.remv[j >> 5] >> (j & 31)) & 1u)) {
            unsigned int todo = ~s.remv[w];                              // rows still alive after (d)
            if (j - i0 < 32) todo &= (1u << (j - i0)) - 1u;              // only i < j (none when j == i0)
            if (todo) {
              const float4 bj = s.box[j]; const float slj = s.sarea[j];
              // screen the 32 rows four at a time (i0 + 31 < kK always), then look again at the few pairs that were not ruled out
              unsigned int maybe = 0u;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                if (!((todo >> (4 * g)) & 0xfu)) continue;
                const float4 sl4 = *reinterpret_cast<const float4*>(&s.sarea[i0 + 4 * g]);
                const float sl[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float inter = box_inter(s.box[i0 + 4 * g + t], bj);
                  if (!(inter < __fadd_rn(sl[t], slj))) maybe |= 1u << (4 * g + t);
                }
              }
              maybe &= todo;
              while (maybe) {
                const int ii = __ffs(maybe) - 1;
                maybe &= maybe - 1u;
                const float4 bi = s.box[i0 + ii];
                if (iou_decide(bi, s.area[i0 + ii], bj, s.area[j], box_inter(bi, bj), __fadd_rn(s.sarea[i0 + ii], slj),
                               p.iou_f, p.iou_inclusive)) bits |= 1u << ii;
              }
            }
          }
          s.mask[w * (kK + 1) + j] = bits;
        }
      }
      __syncthreads();
      NMS_TP(5);
      // ---- (f) greedy choice as a fixed point, all 512 candidates at once (the bit-serial scan by one warp was 20 % of the
      //      kernel).  Candidate j is KEPT once no lower candidate that could still be kept suppresses it, and DEAD once a
      //      kept one does; every step decides at least the lowest undecided candidate, a step costs two barriers, and
      //      the number of steps is the depth of the longest suppression chain (a handful).  Warp w owns word w of the
      //      undecided / kept sets, so no atomics.  The max_det cut is applied to the ranks afterwards: the choice of a
      //      candidate never depends on later ones ----
      {
        unsigned int colr[kWords];                           // rows that suppress me, by word (only words <= mine are non-zero)
#pragma unroll
        for (int w = 0; w < kWords; ++w) colr[w] = (w <= warp) ? s.mask[w * (kK + 1) + tid] : 0u;
        bool und = !dead;
        bool iskept = false;
        while (true) {
          bool nk = false, nd = false;
          if (und) {
            unsigned int anyK = 0u, anyU = 0u;
#pragma unroll
            for (int w = 0; w < kWords; ++w)
              if (colr[w]) { anyK |= colr[w] & s.keptw[w]; anyU |= colr[w] & s.undw[w]; }
            nd = anyK != 0u;
            nk = !nd && anyU == 0u;
          }
          __syncthreads();                                   // every thread has read the sets
          const unsigned int bk = __ballot_sync(0xffffffffu, nk), bd = __ballot_sync(0xffffffffu, nd);
          if (lane == 0 && (bk | bd)) { s.keptw[warp] |= bk; s.undw[warp] &= ~(bk | bd); }
          iskept |= nk;
          und = und && !nk && !nd;
          if (!__syncthreads_or(und)) break;
        }
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWords; ++w) { const int c = __popc(s.keptw[w]); total += c; if (w < warp) before += c; }
        const int rank = before + __popc(s.keptw[warp] & ((1u << lane) - 1u));
        if (iskept && kept0 + rank < p.max_det) s.kidx[rank] = static_cast<unsigned short>(tid);
        if (tid == 0) s.kept = min(p.max_det, kept0 + total);
      }
      __syncthreads();
      NMS_TP(6);
      {
        const int newly = s.kept - kept0;
        if (tid < newly) {
          const int i = s.kidx[tid];
          kbox[kept0 + tid] = s.box[i]; karea[kept0 + tid] = s.area[i]; ksarea[kept0 + tid] = s.sarea[i];
          kkey[kept0 + tid] = big[off + i]; krank[kept0 + tid] = processed + off + i;
        }
      }
      __syncthreads();
      NMS_TP(7);
      if (s.kept >= p.max_det) { done = true; break; }
    }
    processed += SK;
    prev_T = T;
    first = false;
  }

  // ---- output rows x[i] = (x1,y1,x2,y2,conf,cls) (ops.py:327) ----
  __syncthreads();
  const int kept = s.kept;
  if (tid == 0) p.counts[b] = kept;
  const size_t A = p.A;
  for (int t = tid; t < kept; t += kSelThreads) {
    const unsigned long long key = kkey[t];
    const unsigned id = static_cast<unsigned>(key & 0xffffffffull);
    const unsigned a = id / static_cast<unsigned>(p.nc), cls = id - a * static_cast<unsigned>(p.nc);
    float x1, y1, x2, y2;
    if (p.in_place) { x1 = img[a]; y1 = img[A + a]; x2 = img[2 * A + a]; y2 = img[3 * A + a]; }
    else {
      const float cx = img[a], cy = img[A + a], hw = __fmul_rn(img[2 * A + a], 0.5f), hh = __fmul_rn(img[3 * A + a], 0.5f);
      x1 = __fsub_rn(cx, hw); y1 = __fsub_rn(cy, hh); x2 = __fadd_rn(cx, hw); y2 = __fadd_rn(cy, hh);
    }
    if (p.rescale) {
      // ops.scale_boxes (ops.py:92-127): boxes[..., 0/2] -= pad_x, [..., 1/3] -= pad_y, boxes[..., :4] /= gain (fp32, the
      // Python scalar as float32), then ops.clip_boxes (:335-354): clamp to the original image
      const float* rs = p.rescale + static_cast<size_t>(b) * 8;
      const float px = rs[0], py = rs[1], g = rs[2], w0 = rs[3], h0 = rs[4];
      x1 = fminf(fmaxf(__fdiv_rn(__fsub_rn(x1, px), g), 0.f), w0); y1 = fminf(fmaxf(__fdiv_rn(__fsub_rn(y1, py), g), 0.f), h0);
      x2 = fminf(fmaxf(__fdiv_rn(__fsub_rn(x2, px), g), 0.f), w0); y2 = fminf(fmaxf(__fdiv_rn(__fsub_rn(y2, py), g), 0.f), h0);
    }
    float* o = p.out + (static_cast<size_t>(b) * p.max_det + t) * 6;
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
    o[4] = __uint_as_float(~static_cast<unsigned>(key >> 32));
    o[5] = static_cast<float>(cls);
    if (p.kept) {
      long long pos;
      if (n > p.max_nms) pos = krank[t];       // x was re-ordered by score (ops.py:302): index == sorted rank
      else {
        // position in the anchor-ordered candidate list: candidates of earlier chunks + rank inside this chunk
        const int chunk = static_cast<int>(a) / kChunk;
        int before = 0;
        for (int c = 0; c < chunk; ++c) before += p.chunk_cnt[b * p.nchunks + c];
        const unsigned long long* lst = keys + p.chunk_base[b * p.nchunks + chunk];
        int lo = 0, hi = p.chunk_cnt[b * p.nchunks + chunk];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (static_cast<unsigned>(lst[mid] & 0xffffffffull) < id) lo = mid + 1; else hi = mid; }
        pos = before + lo;
      }
      p.kept[static_cast<size_t>(b) * p.max_det + t] = pos;
    }
  }
}

static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

struct NmsWorkspace { size_t keys, img_count, chunk_base, chunk_cnt, total; unsigned long long cap; int nchunks; };

static NmsWorkspace nms_workspace_layout(int B, int nc, int A, int multi_label) {
  NmsWorkspace w{};
  w.nchunks = ceil_div(A, kChunk);
  w.cap = static_cast<unsigned long long>(A) * ((multi_label && nc > 1) ? nc : 1);
  size_t off = 0;
  w.keys = off; off = align16(off + static_cast<size_t>(B) * w.cap * 8);
  w.img_count = off; off = align16(off + static_cast<size_t>(B) * 4);
  w.chunk_base = off; off = align16(off + static_cast<size_t>(B) * w.nchunks * 4);
  w.chunk_cnt = off; off = align16(off + static_cast<size_t>(B) * w.nchunks * 4);
  w.total = off;
  return w;
}

int nms_launch(const dy_nms_desc* d, cudaStream_t stream) {
  DY_CHECK_ARG(d && d->pred && d->out && d->counts && d->workspace, "nms: null pointer");
  DY_CHECK_ARG(d->B > 0 && d->nc > 0 && d->A > 0, "nms: bad shape");
  DY_CHECK_ARG(d->B <= 65535, "nms: B > 65535 unsupported");
  DY_CHECK_ARG(d->nc <= kMaxClassWords * 32, "nms: nc > %d unsupported", kMaxClassWords * 32);
  DY_CHECK_ARG(static_cast<unsigned long long>(d->A) * d->nc < (1ull << 32), "nms: A*nc must fit 32 bits");
  DY_CHECK_ARG(d->conf_thres >= 0.f && d->conf_thres <= 1.f, "nms: conf_thres outside [0,1]");
  DY_CHECK_ARG(d->iou_thres >= 0.0 && d->iou_thres <= 1.0, "nms: iou_thres outside [0,1]");
  DY_CHECK_ARG(d->max_det > 0 && kSelSharedBytes + kept_bytes(d->max_det) + static_cast<size_t>(kBig) * 8 <= 227 * 1024,
               "nms: max_det %d does not fit the select kernel's shared memory (36 B of kept-box state per detection: max ~4000)", d->max_det);
  DY_CHECK_ARG(d->max_nms > 0, "nms: max_nms must be positive");
  const int ml = d->multi_label && d->nc > 1;
  const NmsWorkspace w = nms_workspace_layout(d->B, d->nc, d->A, ml);
  DY_CHECK_ARG(d->workspace_bytes >= w.total, "nms: workspace too small (%zu < %zu)", d->workspace_bytes, w.total);
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0, "nms: workspace must be 16B aligned");

  NmsParams p{};
  p.pred = d->pred; p.pred_rw = const_cast<float*>(d->pred);
  p.B = d->B; p.nc = d->nc; p.A = d->A; p.nchunks = w.nchunks;
  p.conf = d->conf_thres;
  const float f = static_cast<float>(d->iou_thres);
  p.iou_f = f; p.iou_inclusive = static_cast<double>(f) > d->iou_thres ? 1 : 0;
  p.iou_c = f > 1e-6f ? static_cast<float>(static_cast<double>(kScreenLo) * static_cast<double>(f) / (1.0 + static_cast<double>(f))) : nanf("");
  p.max_det = d->max_det; p.max_nms = d->max_nms; p.max_wh = d->max_wh;
  p.agnostic = d->agnostic; p.multi_label = ml; p.in_place = d->xyxy_in_place;
  p.has_class_filter = 0;
  if (d->classes_host && d->n_classes > 0) {
    p.has_class_filter = 1;
    for (int i = 0; i < d->n_classes; ++i) {
      const int c = d->classes_host[i];
      if (c >= 0 && c < d->nc) p.class_mask[c >> 5] |= 1u << (c & 31);
    }
  }
  char* ws = static_cast<char*>(d->workspace);
  p.cap = w.cap;
  p.keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
  p.img_count = reinterpret_cast<int*>(ws + w.img_count);
  p.chunk_base = reinterpret_cast<int*>(ws + w.chunk_base);
  p.chunk_cnt = reinterpret_cast<int*>(ws + w.chunk_cnt);
  p.out = d->out; p.counts = d->counts; p.kept = reinterpret_cast<long long*>(d->kept);
  p.rescale = d->rescale;

  // MSD digit plan over the 64-bit key: 32 score bits, then the significant bits of anchor*nc+cls
  int np = 0;
  p.pass_shift[np] = 53; p.pass_bits[np++] = 11;
  p.pass_shift[np] = 42; p.pass_bits[np++] = 11;
  p.pass_shift[np] = 32; p.pass_bits[np++] = 10;
  int idbits = 1;
  while ((1ull << idbits) < static_cast<unsigned long long>(d->A) * d->nc) ++idbits;
  int hi = idbits;
  while (hi > 0) { const int nb = hi >= 11 ? 11 : hi; p.pass_shift[np] = hi - nb; p.pass_bits[np++] = nb; hi -= nb; }
  p.npasses = np;

  // score bins of the one-pass selection: candidates have conf < score (<= 1 for probabilities), i.e. the high key word
  // u = ~score_bits lies in [~bits(1.0), ~bits(conf)); kBins bins of 2^shift float steps cover that range (scores above 1
  // fall into bin 0, anything beyond the last bin into the last one: the order of bins is the order of keys)
  {
    const float cf = d->conf_thres > 0.f ? d->conf_thres : 0.f;
    unsigned int cbits; memcpy(&cbits, &cf, 4);
    const unsigned int one = 0x3f800000u;
    const unsigned int range = cbits < one ? one - cbits : 0u;
    int sh = 0;
    while ((range >> sh) >= static_cast<unsigned>(kBins)) ++sh;
    p.bin_base = ~one; p.bin_shift = sh;
  }

  DY_CUDA(cudaMemsetAsync(p.img_count, 0, static_cast<size_t>(d->B) * 4, stream));
  const size_t smem = kSelSharedBytes + kept_bytes(d->max_det) + static_cast<size_t>(kBig) * 8;
  {
    // the opt-in is per device: remember the largest request made on each one
    static size_t smem_set[64] = {0};
    int dev = 0;
    DY_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
      DY_CUDA(cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      if (dev >= 0 && dev < 64) smem_set[dev] = smem;
    }
  }
  dim3 fgrid(w.nchunks, d->B);
  nms_filter_kernel<<<fgrid, kFilterThreads, 0, stream>>>(p);
  int rc = launch_status("nms_filter_kernel");
  if (rc) return rc;
  // A plain launch: as a programmatic dependent of the filter kernel the select CTAs were placed while filter CTAs still held
  // most SMs - two per SM on the few that were free - and the issue-bound pair tests of a B = 64 batch ran twice as long.
  nms_select_kernel<<<d->B, kSelThreads, smem, stream>>>(p);
  return launch_status("nms_select_kernel");
}

}  // namespace dy

extern "C" size_t dy_nms_workspace_bytes(int B, int nc, int A, int multi_label) {
  if (B <= 0 || nc <= 0 || A <= 0) return 0;
  return dy::nms_workspace_layout(B, nc, A, multi_label).total;
}

extern "C" int dy_nms(const dy_nms_desc* d, void* stream) { return dy::nms_launch(d, static_cast<cudaStream_t>(stream)); }

#ifdef DY_CONV_DEBUG
// debug builds only: copy out (and clear) the select kernel's per-image phase clocks; 16 counters per image
extern "C" int dy_nms_trace_read(unsigned long long* host, int images) {
  if (images > 1024) images = 1024;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, dy::g_nms_trace, sizeof(unsigned long long) * 16 * images);
  static unsigned long long zeros[1024 * 16];
  cudaMemcpyToSymbol(dy::g_nms_trace, zeros, sizeof(zeros));
  return 0;
}
#endif
