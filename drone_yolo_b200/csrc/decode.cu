// decode.cu — fused Detect decode: DFL softmax-expectation + dist2bbox(xywh) + stride scaling + class sigmoid.
// Replaces ~20 aten kernels of Detect._inference (ultralytics/nn/modules/head.py:100-131): the view/cat over
// levels (:104), make_anchors (utils/tal.py:333-345; computed here from the anchor index), DFL.forward
// (nn/modules/block.py:73-76), dist2bbox(xywh=True) (utils/tal.py:348-357), `* strides` (:129) and
// cls.sigmoid() + cat (:131).
//
// Memory-bound: algorithmic bytes per anchor = no*sizeof(raw) read + (4+nc)*4 written (204 B for bf16 raw maps,
// nc=10).  One thread per anchor, 128 anchors per CTA.
//   NHWC input: the CTA's [128 anchors x ld channels] slab is contiguous in memory; it is copied to shared memory
//               by the copy engine, one bulk copy per row (row pitch padded by 16 B -> conflict-free 128-bit row
//               reads), then every thread walks its own row.
//   NCHW input: channel planes are anchor-contiguous, so plain per-channel loads are already coalesced.
// Output (B, 4+nc, A) fp32 is channel-planar: consecutive threads write consecutive anchors of one plane.
#include "dy_common.cuh"
#include "dy_ptx.cuh"

namespace dy {
using namespace ptx;

static constexpr int kDecThreads = 128;

struct DecodeParams {
  const void* lvl[4];
  int ld[4], H[4], W[4], hw[4], tile0[4], aoff[4];
  float stride[4];
  int nl, B, nc, A, ntiles;
  float* out;
};

__device__ __forceinline__ float load_as_float(const float* p) { return *p; }
__device__ __forceinline__ float load_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void write_box(float* out, size_t plane, float ax, float ay, const float (&d)[4], float stride) {
  float b[4];
  dist2bbox_xywh(ax, ay, d, stride, b);
  out[0 * plane] = b[0]; out[1 * plane] = b[1]; out[2 * plane] = b[2]; out[3 * plane] = b[3];
}
template <typename T>
__global__ void __launch_bounds__(kDecThreads) decode_nhwc_kernel(const __grid_constant__ DecodeParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int b = blockIdx.y;
  int l = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nl && static_cast<int>(blockIdx.x) >= p.tile0[i]) l = i;
  const int a0 = (blockIdx.x - p.tile0[l]) * kDecThreads;
  const int cnt = min(kDecThreads, p.hw[l] - a0);
  const int ld = p.ld[l];
  const int row_bytes = ld * static_cast<int>(sizeof(T));          // multiple of 16 (checked on the host)
  const int pitch = row_bytes + 16;
  const uint8_t* src = static_cast<const uint8_t*>(p.lvl[l]) + (static_cast<size_t>(b) * p.hw[l] + a0) * row_bytes;

  // Every thread has the copy engine bring ITS row (row_bytes contiguous bytes) to its padded shared-memory row: one
  // instruction per thread, no register staging, no index arithmetic; completion is counted on one mbarrier.  The
  // register-staged version (10 independent 16-byte loads per thread, then a store with a division per vector) stopped
  // at 62 % of the measured HBM bandwidth; the NCHW kernel below, which needs no staging, runs at the full copy bandwidth.
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, static_cast<uint32_t>(cnt)); fence_mbar_init(); }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= cnt) return;
  mbar_arrive_expect_tx(&bar, static_cast<uint32_t>(row_bytes));
  bulk_load_1d(smem + threadIdx.x * pitch, src + static_cast<size_t>(threadIdx.x) * row_bytes, static_cast<uint32_t>(row_bytes), &bar);
  mbar_wait(&bar, 0);

  const int a = a0 + threadIdx.x;
  const T* row = reinterpret_cast<const T*>(smem + threadIdx.x * pitch);
  const int W = p.W[l];
  const float ax = static_cast<float>(a % W) + 0.5f, ay = static_cast<float>(a / W) + 0.5f;
  float d[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float x[kRegMax];
    if constexpr (sizeof(T) == 2) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(row + s * 16);
      const uint4 v1 = *reinterpret_cast<const uint4*>(row + s * 16 + 8);
      const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) { x[2 * j] = bf16_lo(w[j]); x[2 * j + 1] = bf16_hi(w[j]); }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + s * 16 + j * 4);
        x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
      }
    }
    d[s] = dfl_expect(x);
  }
  const size_t plane = static_cast<size_t>(p.A);
  float* o = p.out + static_cast<size_t>(b) * (4 + p.nc) * plane + p.aoff[l] + a;
  write_box(o, plane, ax, ay, d, p.stride[l]);
  for (int c = 0; c < p.nc; ++c) {
    const float x = load_as_float(row + 4 * kRegMax + c);
    o[(4 + c) * plane] = sigmoid_fast(x);
  }
}

template <typename T>
__global__ void __launch_bounds__(kDecThreads) decode_nchw_kernel(const __grid_constant__ DecodeParams p) {
  const int b = blockIdx.y;
  int l = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nl && static_cast<int>(blockIdx.x) >= p.tile0[i]) l = i;
  const int a = (blockIdx.x - p.tile0[l]) * kDecThreads + threadIdx.x;
  const int hw = p.hw[l];
  if (a >= hw) return;
  const int no = 4 * kRegMax + p.nc;
  const T* base = static_cast<const T*>(p.lvl[l]) + static_cast<size_t>(b) * no * hw + a;
  const int W = p.W[l];
  const float ax = static_cast<float>(a % W) + 0.5f, ay = static_cast<float>(a / W) + 0.5f;
  float d[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float x[kRegMax];
#pragma unroll
    for (int i = 0; i < kRegMax; ++i) x[i] = load_as_float(base + static_cast<size_t>(s * kRegMax + i) * hw);
    d[s] = dfl_expect(x);
  }
  const size_t plane = static_cast<size_t>(p.A);
  float* o = p.out + static_cast<size_t>(b) * (4 + p.nc) * plane + p.aoff[l] + a;
  write_box(o, plane, ax, ay, d, p.stride[l]);
  for (int c = 0; c < p.nc; ++c) {
    const float x = load_as_float(base + static_cast<size_t>(4 * kRegMax + c) * hw);
    o[(4 + c) * plane] = sigmoid_fast(x);
  }
}

int decode_launch(const dy_decode_desc* d, size_t out_offset_bytes, cudaStream_t stream) {
  DY_CHECK_ARG(d && d->out, "decode: null descriptor/out");
  DY_CHECK_ARG(d->nl >= 1 && d->nl <= 4, "decode: nl=%d out of range", d->nl);
  DY_CHECK_ARG(d->B > 0 && d->nc > 0 && d->nc <= 1024, "decode: bad B/nc");
  DY_CHECK_ARG(d->dtype == DY_BF16 || d->dtype == DY_F32, "decode: bad dtype");
  DY_CHECK_ARG(d->B <= 65535, "decode: B > 65535 unsupported");
  DecodeParams p{};
  const int esz = d->dtype == DY_F32 ? 4 : 2;
  int tiles = 0, A = 0, max_ld = 0;
  for (int l = 0; l < d->nl; ++l) {
    DY_CHECK_ARG(d->lvl[l] && d->H[l] > 0 && d->W[l] > 0, "decode: level %d invalid", l);
    p.lvl[l] = d->lvl[l]; p.ld[l] = d->ld[l]; p.H[l] = d->H[l]; p.W[l] = d->W[l];
    p.hw[l] = d->H[l] * d->W[l]; p.stride[l] = d->stride[l];
    p.tile0[l] = tiles; p.aoff[l] = d->A_total ? d->anchor_off[l] : A;
    tiles += ceil_div(p.hw[l], kDecThreads); A += p.hw[l];
    if (d->layout == DY_NHWC) {
      DY_CHECK_ARG(d->ld[l] >= 4 * kRegMax + d->nc, "decode: ld[%d]=%d < no", l, d->ld[l]);
      DY_CHECK_ARG((d->ld[l] * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->lvl[l]) & 15) == 0,
                   "decode: NHWC rows must be 16B aligned");
      if (d->ld[l] > max_ld) max_ld = d->ld[l];
    }
  }
  if (d->A_total) {
    for (int l = 0; l < d->nl; ++l)
      DY_CHECK_ARG(d->anchor_off[l] >= 0 && d->anchor_off[l] + p.hw[l] <= d->A_total, "decode: level %d does not fit A_total", l);
    A = d->A_total;
  }
  p.nl = d->nl; p.B = d->B; p.nc = d->nc; p.A = A; p.ntiles = tiles;
  p.out = reinterpret_cast<float*>(reinterpret_cast<char*>(d->out) + out_offset_bytes);
  dim3 grid(tiles, d->B);
  if (d->layout == DY_NHWC) {
    const int smem = kDecThreads * (max_ld * esz + 16);
    DY_CHECK_ARG(smem <= 200 * 1024, "decode: ld too large");
    if (d->dtype == DY_BF16) {
      if (smem > 48 * 1024) DY_CUDA(cudaFuncSetAttribute(decode_nhwc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      decode_nhwc_kernel<__nv_bfloat16><<<grid, kDecThreads, smem, stream>>>(p);
    } else {
      if (smem > 48 * 1024) DY_CUDA(cudaFuncSetAttribute(decode_nhwc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      decode_nhwc_kernel<float><<<grid, kDecThreads, smem, stream>>>(p);
    }
  } else {
    if (d->dtype == DY_BF16) decode_nchw_kernel<__nv_bfloat16><<<grid, kDecThreads, 0, stream>>>(p);
    else decode_nchw_kernel<float><<<grid, kDecThreads, 0, stream>>>(p);
  }
  return launch_status("decode kernel");
}

}  // namespace dy

extern "C" int dy_detect_decode(const dy_decode_desc* d, void* stream) {
  return dy::decode_launch(d, 0, static_cast<cudaStream_t>(stream));
}
