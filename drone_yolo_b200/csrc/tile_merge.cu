// tile_merge.cu — merge step of the tiled-frame dispatch: category-aware greedy NMS over the detections of all tiles of one
// frame, already moved to frame coordinates (double precision, as the reference's caller holds them).
//
// Replaces (called from mix6.py:84-89): supervision.InferenceSlicer's overlap filter — Detections.with_nms ->
// box_non_max_suppression (third-party `supervision`, not under /root/reference: parity unpinned, restated in
// oracle/slicer_np.py).  Rows are ranked by confidence descending (ties: higher row first, the flip of a stable ascending
// argsort); a kept row suppresses every later-ranked row of the same category whose IoU exceeds the threshold (strict >).
// The keep mask comes back in the ORIGINAL row order (the caller's rows stay tile-major, not score-sorted).
//
// Three small launches on one stream: rank (n^2 compares, grid over rows) -> 64x64 IoU bit tiles -> one-CTA scan.
#include "dy_common.cuh"

namespace dy {

namespace {

constexpr int kMaxRows = 16384;            // 256 mask words per row: the scan CTA owns one word per thread

__global__ void __launch_bounds__(256) merge_rank_kernel(const double* __restrict__ rows, int n, int* __restrict__ order) {
  __shared__ double s_score[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const double si = i < n ? rows[static_cast<size_t>(i) * 6 + 4] : 0.0;
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    const int j = base + threadIdx.x;
    s_score[threadIdx.x] = j < n ? rows[static_cast<size_t>(j) * 6 + 4] : 0.0;
    __syncthreads();
    const int lim = min(256, n - base);
    for (int k = 0; k < lim; ++k) {
      const double sj = s_score[k];
      rank += (sj > si || (sj == si && base + k > i)) ? 1 : 0;
    }
    __syncthreads();
  }
  if (i < n) order[rank] = i;
}

struct MergeBox {
  double x1, y1, x2, y2, area, cat;
};

__device__ __forceinline__ MergeBox load_box(const double* __restrict__ rows, int r) {
  const double* p = rows + static_cast<size_t>(r) * 6;
  MergeBox b;
  b.x1 = p[0]; b.y1 = p[1]; b.x2 = p[2]; b.y2 = p[3]; b.cat = p[5];
  b.area = __dmul_rn(__dsub_rn(b.x2, b.x1), __dsub_rn(b.y2, b.y1));
  return b;
}

// grid (words, words): block (c, r) fills word c of the 64 rows of row block r (sorted positions); only c >= r is read.
__global__ void __launch_bounds__(64) merge_mask_kernel(const double* __restrict__ rows, const int* __restrict__ order, int n,
                                                       int words, double iou_thres, int agnostic,
                                                       unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  __shared__ MergeBox s_col[64];
  const int t = threadIdx.x;
  const int jc = cb * 64 + t;
  if (jc < n) s_col[t] = load_box(rows, order[jc]);
  __syncthreads();
  const int i = rb * 64 + t;
  if (i >= n) return;
  const MergeBox a = load_box(rows, order[i]);
  unsigned long long bits = 0ull;
  const int lim = min(64, n - cb * 64);
  for (int k = 0; k < lim; ++k) {
    const int j = cb * 64 + k;
    if (j <= i) continue;
    const MergeBox b = s_col[k];
    if (!agnostic && b.cat != a.cat) continue;
    const double w = fmax(__dsub_rn(fmin(a.x2, b.x2), fmax(a.x1, b.x1)), 0.0);
    const double h = fmax(__dsub_rn(fmin(a.y2, b.y2), fmax(a.y1, b.y1)), 0.0);
    const double inter = __dmul_rn(w, h);
    const double iou = __ddiv_rn(inter, __dsub_rn(__dadd_rn(a.area, b.area), inter));
    if (iou > iou_thres) bits |= 1ull << k;                  // NaN (0/0) compares false: kept, as numpy does
  }
  mask[static_cast<size_t>(i) * words + cb] = bits;
}

// One CTA; the `removed` bit set (one word per 64 ranked rows) lives in shared memory.  Per block of 64 ranked rows: the 64
// diagonal words are fetched in parallel, one thread walks the 64 rows serially on them (shared-memory latency only), then all
// threads OR the mask rows of the kept rows into the later words: (kept row, word) pairs dealt over the CTA, four independent
// loads in flight per thread.  (A first version kept one word per thread in registers and chained a dependent global load per kept
// row: 0.98 ms for 1800 rows, all of it load latency.)
__global__ void __launch_bounds__(256) merge_scan_kernel(const int* __restrict__ order, int n, int words,
                                                        const unsigned long long* __restrict__ mask,
                                                        unsigned char* __restrict__ keep) {
  __shared__ unsigned long long s_removed[256];
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_kept[64];
  const int t = threadIdx.x;
  s_removed[t] = 0ull;
  __syncthreads();
  for (int b = 0; b < words; ++b) {
    const int lim = min(64, n - b * 64);
    if (t < 64) s_diag[t] = t < lim ? mask[static_cast<size_t>(b * 64 + t) * words + b] : 0ull;
    __syncthreads();
    if (t == 0) {
      unsigned long long removed = s_removed[b], kb = 0ull;
      for (int k = 0; k < lim; ++k) {
        if (!((removed >> k) & 1ull)) {
          kb |= 1ull << k;
          removed |= s_diag[k];
        }
      }
      s_keepbits = kb;
    }
    __syncthreads();
    const unsigned long long kb = s_keepbits;
    if (t < 64) {
      if (t < lim) keep[order[b * 64 + t]] = static_cast<unsigned char>((kb >> t) & 1ull);
      if ((kb >> t) & 1ull) s_kept[__popcll(kb & ((1ull << t) - 1ull))] = t;
    }
    __syncthreads();
    const int nw = words - b - 1;
    const int total = __popcll(kb) * nw;
    for (int p0 = t; p0 < total; p0 += 4 * 256) {
      unsigned long long v[4];
      int wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * 256;
        v[u] = 0ull;
        wv[u] = 0;
        if (p < total) {
          const int ki = p / nw;
          wv[u] = b + 1 + (p - ki * nw);
          v[u] = mask[static_cast<size_t>(b * 64 + s_kept[ki]) * words + wv[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v[u]) atomicOr(&s_removed[wv[u]], v[u]);
    }
    __syncthreads();
  }
}

}  // namespace

size_t box_nms_f64_workspace_bytes(int n) {
  if (n <= 0) return 16;
  const size_t words = static_cast<size_t>(ceil_div(n, 64));
  return round_up(n, 4) * sizeof(int) + static_cast<size_t>(n) * words * sizeof(unsigned long long) + 16;
}

int box_nms_f64_launch(const double* rows, int n, double iou_thres, int agnostic, unsigned char* keep, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  DY_CHECK_ARG(n >= 0 && n <= kMaxRows, "box_nms_f64: n = %d outside [0, %d]", n, kMaxRows);
  if (n == 0) return DY_OK;
  DY_CHECK_ARG(rows && keep && workspace, "box_nms_f64: null pointer");
  DY_CHECK_ARG((reinterpret_cast<uintptr_t>(rows) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
               "box_nms_f64: rows and workspace must be 8B aligned");
  DY_CHECK_ARG(workspace_bytes >= box_nms_f64_workspace_bytes(n), "box_nms_f64: workspace of %zu bytes, %zu needed",
               workspace_bytes, box_nms_f64_workspace_bytes(n));
  DY_CHECK_ARG(iou_thres >= 0.0 && iou_thres <= 1.0, "box_nms_f64: iou threshold %f outside [0, 1]", iou_thres);
  const int words = ceil_div(n, 64);
  int* order = static_cast<int*>(workspace);
  auto* mask = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + round_up(n, 4) * sizeof(int));
  merge_rank_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(rows, n, order);
  int rc = launch_status("merge_rank_kernel");
  if (rc != DY_OK) return rc;
  merge_mask_kernel<<<dim3(words, words), 64, 0, stream>>>(rows, order, n, words, iou_thres, agnostic, mask);
  rc = launch_status("merge_mask_kernel");
  if (rc != DY_OK) return rc;
  merge_scan_kernel<<<1, 256, 0, stream>>>(order, n, words, mask, keep);
  return launch_status("merge_scan_kernel");
}

}  // namespace dy

extern "C" size_t dy_box_nms_f64_workspace_bytes(int n) { return dy::box_nms_f64_workspace_bytes(n); }
extern "C" int dy_box_nms_f64(const double* rows, int n, double iou_thres, int class_agnostic, unsigned char* keep,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return dy::box_nms_f64_launch(rows, n, iou_thres, class_agnostic, keep, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}
