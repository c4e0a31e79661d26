// aux_kernels.cu — the small memory-bound layers around the tensor-core convs (CUDA cores on purpose: none of
// them is a dense contraction worth reshaping into a GEMM).
//   stem_conv_kernel   model.0 Conv(3->c,3,2)+BN+SiLU fused with the fp32 NCHW -> bf16 NHWC conversion
//                      (ultralytics/nn/modules/conv.py:49-55; engine/predictor.py:132-135)
//   sppf_pool_kernel   SPPF's three chained MaxPool2d(5,1,2) + cat  (nn/modules/block.py:185-191)
//   upsample2x_kernel  nn.Upsample(None,2,'nearest') written straight into a Concat slice (conv.py:331-333)
//   dwconv3x3s2_kernel DWConv(c1,c2,3,2), groups=c2, c1=2*c2, of the "-sf" graph (conv.py:102-107)
#include "dy_common.cuh"
#include <cstdlib>

namespace dy {

// ------------------------------------------------------------------------------------------------
// Stem: one thread per output pixel keeps the 3x3x3 input patch in registers and sweeps the output
// channels 8 at a time (weights broadcast from shared memory as float4), storing 16 B per sweep.
// ------------------------------------------------------------------------------------------------
static constexpr int kStemThreads = 128;
static constexpr int kStemMaxC = 128;
static constexpr int kStemPix = 4;              // output pixels per thread (consecutive along W): weights are fetched once per 4 pixels

template <typename TIn>
__global__ void __launch_bounds__(kStemThreads) stem_conv_kernel(const TIn* __restrict__ in, float in_scale, int B, int H, int W,
                                                                 const float* __restrict__ weight,
                                                                 const float* __restrict__ bias, int Cout,
                                                                 __nv_bfloat16* __restrict__ out, int out_ld) {
  __shared__ __align__(16) float w_s[kStemMaxC * 28];
  __shared__ float b_s[kStemMaxC];
  for (int i = threadIdx.x; i < Cout * 28; i += kStemThreads) {
    const int co = i / 28, k = i - co * 28;
    w_s[i] = k < 27 ? weight[co * 27 + k] : 0.f;
  }
  for (int i = threadIdx.x; i < Cout; i += kStemThreads) b_s[i] = bias[i];
  __syncthreads();

  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int Wq = (Wo + kStemPix - 1) / kStemPix;          // pixel quads per output row
  const long long total = static_cast<long long>(B) * Ho * Wq;
  const long long item = static_cast<long long>(blockIdx.x) * kStemThreads + threadIdx.x;
  if (item >= total) return;
  const int wq = static_cast<int>(item % Wq);
  const int ho = static_cast<int>((item / Wq) % Ho);
  const int b = static_cast<int>(item / (static_cast<long long>(Wq) * Ho));
  const int wo0 = wq * kStemPix;

  // 3 channels x 3 rows x 9 columns of input feed the 4 output pixels
  float x[3][3][2 * kStemPix + 1];
  const size_t plane = static_cast<size_t>(H) * W;
  const TIn* img = in + static_cast<size_t>(b) * 3 * plane;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = 2 * ho - 1 + ky;
      const bool yok = (y >= 0) && (y < H);
#pragma unroll
      for (int j = 0; j < 2 * kStemPix + 1; ++j) {
        const int xx = 2 * wo0 - 1 + j;
        x[c][ky][j] = (yok && xx >= 0 && xx < W) ? static_cast<float>(__ldg(img + c * plane + static_cast<size_t>(y) * W + xx)) * in_scale : 0.f;
      }
    }

  __nv_bfloat16* op = out + (static_cast<size_t>(b) * Ho + ho) * Wo * out_ld + static_cast<size_t>(wo0) * out_ld;
  for (int co = 0; co < Cout; co += 8) {
    float acc[kStemPix][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(w_s + (co + j) * 28);
      float wv[28];
#pragma unroll
      for (int k4 = 0; k4 < 7; ++k4) { const float4 t = wr[k4]; wv[4 * k4] = t.x; wv[4 * k4 + 1] = t.y; wv[4 * k4 + 2] = t.z; wv[4 * k4 + 3] = t.w; }
      const float bj = b_s[co + j];
#pragma unroll
      for (int px = 0; px < kStemPix; ++px) {
        float a = bj;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(x[c][ky][2 * px + kx], wv[c * 9 + ky * 3 + kx], a);
        acc[px][j] = silu_fast(a);
      }
    }
#pragma unroll
    for (int px = 0; px < kStemPix; ++px) {
      if (wo0 + px < Wo) {
        uint4 o;
        o.x = pack_bf16(acc[px][0], acc[px][1]); o.y = pack_bf16(acc[px][2], acc[px][3]);
        o.z = pack_bf16(acc[px][4], acc[px][5]); o.w = pack_bf16(acc[px][6], acc[px][7]);
        *reinterpret_cast<uint4*>(op + static_cast<size_t>(px) * out_ld + co) = o;
      }
    }
  }
}

int stem_tc_launch(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout, void* out,
                   int out_ld, cudaStream_t stream);   // stem_igemm.cu

int stem_launch(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout, void* out,
                int out_ld, cudaStream_t stream) {
  DY_CHECK_ARG(in_dtype == DY_F32 || in_dtype == DY_U8, "stem: input must be fp32 or uint8");
  DY_CHECK_ARG(in && weight && bias && out, "stem: null pointer");
  DY_CHECK_ARG(B > 0 && H > 0 && W > 0, "stem: bad shape");
  DY_CHECK_ARG(Cout % 8 == 0 && Cout > 0 && Cout <= kStemMaxC, "stem: Cout must be a multiple of 8, <= %d", kStemMaxC);
  DY_CHECK_ARG(out_ld % 8 == 0 && out_ld >= Cout && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "stem: out slice must be 16B aligned");
  if (getenv("DY_STEM_CUDA_CORES") == nullptr) {
    // tensor-core stem (stem_igemm.cu); shapes it does not cover (odd widths, Cout > 112) run on the CUDA-core kernel below
    const int rc = stem_tc_launch(in, in_dtype, B, H, W, weight, bias, Cout, out, out_ld, stream);
    if (rc != DY_ERR_UNSUPPORTED) return rc;
  }
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * ((Wo + kStemPix - 1) / kStemPix);
  const long long blocks = (total + kStemThreads - 1) / kStemThreads;
  DY_CHECK_ARG(blocks < (1ll << 31), "stem: too many pixels");
  if (in_dtype == DY_U8)
    stem_conv_kernel<unsigned char><<<static_cast<unsigned>(blocks), kStemThreads, 0, stream>>>(
        static_cast<const unsigned char*>(in), 1.f / 255.f, B, H, W, weight, bias, Cout, static_cast<__nv_bfloat16*>(out), out_ld);
  else
    stem_conv_kernel<float><<<static_cast<unsigned>(blocks), kStemThreads, 0, stream>>>(
        static_cast<const float*>(in), 1.f, B, H, W, weight, bias, Cout, static_cast<__nv_bfloat16*>(out), out_ld);
  return launch_status("stem_conv_kernel");
}

// ------------------------------------------------------------------------------------------------
// SPPF pooling.  One CTA owns one image x one 8-channel group (16 B per pixel) and keeps the whole HxW map in
// shared memory; each mp5 is separable (5-wide row max, then 5-tall column max; window clipped == -inf padding).
// ------------------------------------------------------------------------------------------------
static constexpr int kPoolThreads = 256;

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162 t;
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.x), *reinterpret_cast<__nv_bfloat162*>(&b.x)); r.x = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.y), *reinterpret_cast<__nv_bfloat162*>(&b.y)); r.y = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.z), *reinterpret_cast<__nv_bfloat162*>(&b.z)); r.z = *reinterpret_cast<uint32_t*>(&t);
  t = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a.w), *reinterpret_cast<__nv_bfloat162*>(&b.w)); r.w = *reinterpret_cast<uint32_t*>(&t);
  return r;
}

// One CTA = one image x NV*8 consecutive channels, the whole HxW map in shared memory.  NV = 4 (64 contiguous bytes per
// pixel: full 32-byte sectors on both the strided reads and the three strided writes) when the map fits, else 2 or 1.
template <int NV>
__global__ void __launch_bounds__(kPoolThreads) sppf_pool_kernel(__nv_bfloat16* buf, int H, int W, int C, int ld) {
  extern __shared__ __align__(16) uint8_t pool_smem[];
  const int hw = H * W, n = hw * NV;
  uint4* X = reinterpret_cast<uint4*>(pool_smem);
  uint4* T = X + n;
  const int groups = C / (8 * NV);
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  __nv_bfloat16* base = buf + static_cast<size_t>(b) * hw * ld + g * 8 * NV;
  for (int i = threadIdx.x; i < n; i += kPoolThreads)
    X[i] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(i / NV) * ld + (i % NV) * 8);
  __syncthreads();
  for (int rep = 1; rep <= 3; ++rep) {
    for (int i = threadIdx.x; i < n; i += kPoolThreads) {
      const int px = i / NV, x = px % W;
      uint4 m = X[i];
      for (int dx = -2; dx <= 2; ++dx) { const int xx = x + dx; if (dx != 0 && xx >= 0 && xx < W) m = max_bf16x8(m, X[i + dx * NV]); }
      T[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kPoolThreads) {
      const int px = i / NV, y = px / W;
      uint4 m = T[i];
      for (int dy = -2; dy <= 2; ++dy) { const int yy = y + dy; if (dy != 0 && yy >= 0 && yy < H) m = max_bf16x8(m, T[i + dy * W * NV]); }
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(px) * ld + rep * C + (i % NV) * 8) = m;
      X[i] = m;     // only this thread reads/writes X[i] in this phase; T is the cross-thread source
    }
    __syncthreads();
  }
}

template <int NV>
static int sppf_pool_launch_nv(void* buf, int B, int H, int W, int C, int ld, size_t smem, cudaStream_t stream) {
  static unsigned long long seen = 0;                       // per device: the opt-in up to the 227 KB limit, once
  if (smem > 48 * 1024 && first_use_on_device(&seen))
    DY_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  sppf_pool_kernel<NV><<<B * (C / (8 * NV)), kPoolThreads, smem, stream>>>(static_cast<__nv_bfloat16*>(buf), H, W, C, ld);
  return launch_status("sppf_pool_kernel");
}

int sppf_pool_launch(void* buf, int B, int H, int W, int C, int ld, cudaStream_t stream) {
  DY_CHECK_ARG(buf && B > 0 && H > 0 && W > 0 && C > 0, "sppf_pool: bad argument");
  DY_CHECK_ARG(C % 8 == 0 && ld % 8 == 0 && ld >= 4 * C && (reinterpret_cast<uintptr_t>(buf) & 15) == 0,
               "sppf_pool: C, ld must be multiples of 8 with ld >= 4*C");
  const size_t smem1 = static_cast<size_t>(H) * W * 16 * 2;
  DY_CHECK_ARG(smem1 <= 220 * 1024, "sppf_pool: %dx%d map does not fit shared memory", H, W);
  if (C % 32 == 0 && smem1 * 4 <= 100 * 1024) return sppf_pool_launch_nv<4>(buf, B, H, W, C, ld, smem1 * 4, stream);
  if (C % 16 == 0 && smem1 * 2 <= 100 * 1024) return sppf_pool_launch_nv<2>(buf, B, H, W, C, ld, smem1 * 2, stream);
  return sppf_pool_launch_nv<1>(buf, B, H, W, C, ld, smem1, stream);
}

// ------------------------------------------------------------------------------------------------
// Letterbox: one thread per 4 consecutive output columns of one row (three 4-byte plane stores).  The resize is OpenCV's
// 8-bit INTER_LINEAR: source coordinate (float)((d + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in double, 11-bit
// fixed-point coefficients, horizontal pass in int, vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// ------------------------------------------------------------------------------------------------
struct LbCoef { int i0, i1, a0, a1; };

__device__ __forceinline__ LbCoef lb_coef(int d, double scale, int n_src) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int i = static_cast<int>(floorf(f));
  f -= static_cast<float>(i);
  if (i < 0) { i = 0; f = 0.f; }
  if (i >= n_src - 1) { i = n_src - 1; f = 0.f; }
  LbCoef c;
  c.i0 = i; c.i1 = min(i + 1, n_src - 1);
  c.a1 = __float2int_rn(f * 2048.f);
  c.a0 = __float2int_rn((1.f - f) * 2048.f);
  return c;
}

__global__ void __launch_bounds__(256) letterbox_u8_kernel(const uint8_t* __restrict__ src, int h, int w, int pitch,
                                                           uint8_t* __restrict__ dst, int H, int W, int new_w, int new_h,
                                                           int left, int top, int fill, double scale_x, double scale_y,
                                                           size_t src_image_stride, size_t dst_image_stride) {
  src += blockIdx.y * src_image_stride;          // blockIdx.y = frame of a same-geometry batch
  dst += blockIdx.y * dst_image_stride;
  const int wq = W >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wq * H) return;
  const int y = idx / wq, x0 = (idx - y * wq) * 4;
  uint32_t out[3] = {0u, 0u, 0u};
  const int dy = y - top;
  const bool row_in = dy >= 0 && dy < new_h;
  LbCoef cy{};
  if (row_in) cy = lb_coef(dy, scale_y, h);
  const uint8_t* r0 = src + static_cast<size_t>(cy.i0) * pitch;
  const uint8_t* r1 = src + static_cast<size_t>(cy.i1) * pitch;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dx = x0 + k - left;
    uint32_t v[3] = {static_cast<uint32_t>(fill), static_cast<uint32_t>(fill), static_cast<uint32_t>(fill)};
    if (row_in && dx >= 0 && dx < new_w) {
      const LbCoef cx = lb_coef(dx, scale_x, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int t0 = r0[cx.i0 * 3 + c] * cx.a0 + r0[cx.i1 * 3 + c] * cx.a1;
        const int t1 = r1[cx.i0 * 3 + c] * cx.a0 + r1[cx.i1 * 3 + c] * cx.a1;
        v[2 - c] = static_cast<uint32_t>((((cy.a0 * (t0 >> 4)) >> 16) + ((cy.a1 * (t1 >> 4)) >> 16) + 2) >> 2);   // BGR -> RGB plane
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] |= (v[c] & 0xffu) << (8 * k);
  }
  const size_t plane = static_cast<size_t>(H) * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) *reinterpret_cast<uint32_t*>(dst + c * plane + static_cast<size_t>(y) * W + x0) = out[c];
}

int letterbox_u8_batch_launch(const void* src, int n, size_t src_image_stride, int h, int w, int pitch, void* dst,
                              size_t dst_image_stride, int H, int W, int new_w, int new_h, int left, int top, int fill,
                              cudaStream_t stream) {
  DY_CHECK_ARG(src && dst && h > 0 && w > 0 && H > 0 && W > 0 && pitch >= 3 * w, "letterbox: bad argument");
  DY_CHECK_ARG(n >= 1 && n <= 65535, "letterbox: %d frames outside [1, 65535]", n);
  DY_CHECK_ARG(W % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && dst_image_stride % 4 == 0,
               "letterbox: W and the image stride of dst must be multiples of 4 and dst 4B aligned");
  DY_CHECK_ARG(new_w > 0 && new_h > 0 && left >= 0 && top >= 0 && left + new_w <= W && top + new_h <= H,
               "letterbox: the resized image (%dx%d at %d,%d) does not fit the %dx%d canvas", new_w, new_h, left, top, W, H);
  const double sx = 1.0 / (static_cast<double>(new_w) / w), sy = 1.0 / (static_cast<double>(new_h) / h);
  const int px = (W >> 2) * H;
  letterbox_u8_kernel<<<dim3(ceil_div(px, 256), n), 256, 0, stream>>>(static_cast<const uint8_t*>(src), h, w, pitch,
                                                                     static_cast<uint8_t*>(dst), H, W, new_w, new_h, left, top, fill,
                                                                     sx, sy, src_image_stride, dst_image_stride);
  return launch_status("letterbox_u8_kernel");
}

int letterbox_u8_launch(const void* src, int h, int w, int pitch, void* dst, int H, int W, int new_w, int new_h, int left,
                        int top, int fill, cudaStream_t stream) {
  return letterbox_u8_batch_launch(src, 1, 0, h, w, pitch, dst, 0, H, W, new_w, new_h, left, top, fill, stream);
}

// ------------------------------------------------------------------------------------------------
// 2x nearest upsample: one thread per (output pixel, 8-channel vector).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, int B, int H, int W,
                                                         int C, __nv_bfloat16* __restrict__ out, int out_ld) {
  const int vec = C / 8;
  const long long total = static_cast<long long>(B) * (2 * H) * (2 * W) * vec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long pix = i / vec;
    const int x = static_cast<int>(pix % (2 * W)); pix /= (2 * W);
    const int y = static_cast<int>(pix % (2 * H));
    const int b = static_cast<int>(pix / (2 * H));
    const uint4 val = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + (y >> 1)) * W + (x >> 1)) * in_ld + v * 8);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * 2 * H + y) * 2 * W + x) * out_ld + v * 8) = val;
  }
}

int upsample2x_launch(const void* in, int in_ld, int B, int H, int W, int C, void* out, int out_ld, cudaStream_t stream) {
  DY_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0, "upsample2x: bad argument");
  DY_CHECK_ARG(C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0, "upsample2x: slices must be 16B aligned");
  const long long total = static_cast<long long>(B) * 4 * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), in_ld, B, H, W, C,
                                                                       static_cast<__nv_bfloat16*>(out), out_ld);
  return launch_status("upsample2x_kernel");
}

// ------------------------------------------------------------------------------------------------
// DWConv 3x3 stride 2 pad 1, groups = Cout, 2 input channels per group, + bias + SiLU.
// One thread per (output pixel, 8 output channels = 16 contiguous input channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv3x3s2_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, int B, int H, int W,
                                                          const float* __restrict__ weight, const float* __restrict__ bias,
                                                          int Cout, __nv_bfloat16* __restrict__ out, int out_ld) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int vec = Cout / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * vec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long pix = i / vec;
    const int wo = static_cast<int>(pix % Wo); pix /= Wo;
    const int ho = static_cast<int>(pix % Ho);
    const int b = static_cast<int>(pix / Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + v * 8 + j);
    const float* wp = weight + static_cast<size_t>(v) * 8 * 18;
    for (int ky = 0; ky < 3; ++ky) {
      const int y = 2 * ho - 1 + ky;
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x = 2 * wo - 1 + kx;
        if (x < 0 || x >= W) continue;
        const __nv_bfloat16* ip = in + ((static_cast<size_t>(b) * H + y) * W + x) * in_ld + v * 16;
        const uint4 v0 = *reinterpret_cast<const uint4*>(ip);
        const uint4 v1 = *reinterpret_cast<const uint4*>(ip + 8);
        const uint32_t w32[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // output channel v*8+j reads input channels 2j, 2j+1 of this 16-wide vector
          acc[j] = fmaf(bf16_lo(w32[j]), __ldg(wp + j * 18 + ky * 3 + kx), acc[j]);
          acc[j] = fmaf(bf16_hi(w32[j]), __ldg(wp + j * 18 + 9 + ky * 3 + kx), acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = silu_fast(acc[j]);
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * Ho + ho) * Wo + wo) * out_ld + v * 8) = o;
  }
}

int dwconv_launch(const void* in, int in_ld, int B, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                  void* out, int out_ld, cudaStream_t stream) {
  DY_CHECK_ARG(in && out && weight && bias && B > 0 && H > 0 && W > 0, "dwconv: bad argument");
  DY_CHECK_ARG(Cin == 2 * Cout, "dwconv: only groups == Cout with Cin == 2*Cout is implemented (Cin %d, Cout %d)", Cin, Cout);
  DY_CHECK_ARG(Cout % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0, "dwconv: slices must be 16B aligned, Cout % 8 == 0");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (Cout / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  dwconv3x3s2_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), in_ld, B, H, W, weight,
                                                                        bias, Cout, static_cast<__nv_bfloat16*>(out), out_ld);
  return launch_status("dwconv3x3s2_kernel");
}

}  // namespace dy

extern "C" int dy_stem_conv(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout,
                            void* out, int out_ld, void* stream) {
  return dy::stem_launch(in, in_dtype, B, H, W, weight, bias, Cout, out, out_ld, static_cast<cudaStream_t>(stream));
}
extern "C" int dy_letterbox_u8(const void* src, int h, int w, int src_pitch, void* dst, int H, int W, int new_w, int new_h,
                               int left, int top, int fill, void* stream) {
  return dy::letterbox_u8_launch(src, h, w, src_pitch, dst, H, W, new_w, new_h, left, top, fill, static_cast<cudaStream_t>(stream));
}
extern "C" int dy_letterbox_u8_batch(const void* src, int n, size_t src_image_stride, int h, int w, int src_pitch, void* dst,
                                     size_t dst_image_stride, int H, int W, int new_w, int new_h, int left, int top, int fill,
                                     void* stream) {
  return dy::letterbox_u8_batch_launch(src, n, src_image_stride, h, w, src_pitch, dst, dst_image_stride, H, W, new_w, new_h,
                                       left, top, fill, static_cast<cudaStream_t>(stream));
}
extern "C" int dy_sppf_pool(void* buf, int B, int H, int W, int C, int ld, void* stream) {
  return dy::sppf_pool_launch(buf, B, H, W, C, ld, static_cast<cudaStream_t>(stream));
}
extern "C" int dy_upsample2x(const void* in, int in_ld, int B, int H, int W, int C, void* out, int out_ld, void* stream) {
  return dy::upsample2x_launch(in, in_ld, B, H, W, C, out, out_ld, static_cast<cudaStream_t>(stream));
}
extern "C" int dy_dwconv3x3s2(const void* in, int in_ld, int B, int H, int W, int Cin, const float* weight, const float* bias,
                              int Cout, void* out, int out_ld, void* stream) {
  return dy::dwconv_launch(in, in_ld, B, H, W, Cin, weight, bias, Cout, out, out_ld, static_cast<cudaStream_t>(stream));
}
