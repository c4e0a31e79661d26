// api.cu — C-ABI plumbing: error text, device check, the recorded "program" (layer plan) and the UMMA self-test.
#include "dy_common.cuh"
#include "conv_igemm.h"
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <new>

namespace dy {

// forward declarations of the per-op launchers (decode.cu, nms.cu, aux_kernels.cu)
int decode_launch(const dy_decode_desc* d, size_t out_offset_bytes, cudaStream_t stream);
int nms_launch(const dy_nms_desc* d, cudaStream_t stream);
int stem_launch(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout, void* out,
                int out_ld, cudaStream_t stream);
int sppf_pool_launch(void* buf, int B, int H, int W, int C, int ld, cudaStream_t stream);
int upsample2x_launch(const void* in, int in_ld, int B, int H, int W, int C, void* out, int out_ld, cudaStream_t stream);
int dwconv_launch(const void* in, int in_ld, int B, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                  void* out, int out_ld, cudaStream_t stream);

char* err_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

enum OpKind { OP_CONV, OP_STEM, OP_POOL, OP_UPSAMPLE, OP_DWCONV, OP_DECODE, OP_NMS, OP_SYNC };

struct Op {
  OpKind kind;
  int lane = 0;                       // 0 = the caller's stream, k > 0 = the program's k-th side stream
  int waiter = 0, signaller = 0;      // OP_SYNC: lane `waiter` waits for everything enqueued so far on lane `signaller`
  cudaEvent_t event = nullptr;
  ConvParams conv; ConvLaunch conv_launch;
  // generic scalar arguments for the small ops
  const void* in; void* out; const float* w; const float* b;
  int B, H, W, C, Cin, in_ld, out_ld;
  dy_decode_desc dec;
  dy_nms_desc nms;
  std::vector<int32_t> nms_classes;
};

}  // namespace dy

struct dy_program {
  std::vector<dy::Op*> ops;
  int launches = 0;
  int lane = 0;                           // lane of the ops added from now on (dy_program_set_lane)
  std::vector<cudaStream_t> side;         // side streams, lane k -> side[k - 1]; lower priority than any prioritised caller stream
};

extern "C" {

int dy_version(void) { return 100; }

const char* dy_last_error(void) { return dy::err_buf(); }

int dy_device_check(int device) {
  cudaDeviceProp prop;
  DY_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return dy::fail(DY_ERR_UNSUPPORTED, "device %d is sm_%d%d; libdroneyolo is built for sm_100a only", device, prop.major, prop.minor);
  return DY_OK;
}

int dy_program_create(dy_program** out) {
  DY_CHECK_ARG(out, "program_create: null out");
  *out = new (std::nothrow) dy_program();
  if (!*out) return dy::fail(DY_ERR_NOMEM, "program_create: out of memory");
  return DY_OK;
}

void dy_program_destroy(dy_program* p) {
  if (!p) return;
  for (dy::Op* o : p->ops) {
    if (o->event) cudaEventDestroy(o->event);
    delete o;
  }
  for (cudaStream_t s : p->side) cudaStreamDestroy(s);
  delete p;
}

static int push(dy_program* p, dy::Op* o, int launches) {
  o->lane = p->lane;
  p->ops.push_back(o);
  p->launches += launches;
  return DY_OK;
}

int dy_program_add_conv(dy_program* p, const dy_conv_desc* d) {
  DY_CHECK_ARG(p && d, "program_add_conv: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_CONV;
  int rc = dy::conv_build_params(d, &o->conv, &o->conv_launch);
  if (rc) { delete o; return rc; }
  return push(p, o, 1);
}

int dy_program_add_stem(dy_program* p, const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias,
                        int Cout, void* out, int out_ld) {
  DY_CHECK_ARG(p && in && weight && bias && out, "program_add_stem: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_STEM; o->in = in; o->Cin = in_dtype; o->B = B; o->H = H; o->W = W; o->w = weight; o->b = bias; o->C = Cout; o->out = out; o->out_ld = out_ld;
  return push(p, o, 1);
}

int dy_program_add_sppf_pool(dy_program* p, void* buf, int B, int H, int W, int C, int ld) {
  DY_CHECK_ARG(p && buf, "program_add_sppf_pool: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_POOL; o->out = buf; o->B = B; o->H = H; o->W = W; o->C = C; o->out_ld = ld;
  return push(p, o, 1);
}

int dy_program_add_upsample2x(dy_program* p, const void* in, int in_ld, int B, int H, int W, int C, void* out, int out_ld) {
  DY_CHECK_ARG(p && in && out, "program_add_upsample2x: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_UPSAMPLE; o->in = in; o->in_ld = in_ld; o->B = B; o->H = H; o->W = W; o->C = C; o->out = out; o->out_ld = out_ld;
  return push(p, o, 1);
}

int dy_program_add_dwconv3x3s2(dy_program* p, const void* in, int in_ld, int B, int H, int W, int Cin, const float* weight,
                               const float* bias, int Cout, void* out, int out_ld) {
  DY_CHECK_ARG(p && in && out && weight && bias, "program_add_dwconv: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_DWCONV; o->in = in; o->in_ld = in_ld; o->B = B; o->H = H; o->W = W; o->Cin = Cin; o->w = weight; o->b = bias;
  o->C = Cout; o->out = out; o->out_ld = out_ld;
  return push(p, o, 1);
}

int dy_program_add_decode(dy_program* p, const dy_decode_desc* d) {
  DY_CHECK_ARG(p && d, "program_add_decode: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_DECODE; o->dec = *d;
  return push(p, o, 1);
}

int dy_program_add_nms(dy_program* p, const dy_nms_desc* d) {
  DY_CHECK_ARG(p && d, "program_add_nms: null");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_NMS; o->nms = *d;
  if (d->classes_host && d->n_classes > 0) {
    o->nms_classes.assign(d->classes_host, d->classes_host + d->n_classes);
    o->nms.classes_host = o->nms_classes.data();
  }
  return push(p, o, 2);
}

// Lanes: independent branches of the layer graph (the Detect branches of one pyramid level against the rest of the neck)
// are enqueued on side streams, so that the CTAs of one branch fill the SMs another branch's kernel leaves idle during
// its ramp-up, its last partial wave and its drain.  Side streams are created here (never during a stream capture) with
// the lowest priority: a caller that captures / runs on a prioritised stream keeps the main chain ahead of the branches.
int dy_program_set_lane(dy_program* p, int lane) {
  DY_CHECK_ARG(p && lane >= 0 && lane <= 8, "program_set_lane: lane must be in [0, 8]");
  while (static_cast<int>(p->side.size()) < lane) {
    int least = 0, greatest = 0;
    DY_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    cudaStream_t s = nullptr;
    DY_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, least));
    p->side.push_back(s);
  }
  p->lane = lane;
  return DY_OK;
}

int dy_program_add_sync(dy_program* p, int waiter, int signaller) {
  DY_CHECK_ARG(p && waiter != signaller && waiter >= 0 && signaller >= 0 && waiter <= static_cast<int>(p->side.size()) &&
               signaller <= static_cast<int>(p->side.size()), "program_add_sync: unknown lane (declare it with dy_program_set_lane first)");
  dy::Op* o = new dy::Op();
  o->kind = dy::OP_SYNC; o->waiter = waiter; o->signaller = signaller;
  cudaError_t e = cudaEventCreateWithFlags(&o->event, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete o; return dy::fail(DY_ERR_CUDA, "program_add_sync: %s", cudaGetErrorString(e)); }
  p->ops.push_back(o);
  return DY_OK;
}

static int program_launch_op(dy_program* p, dy::Op* o, size_t in_offset_bytes, size_t out_offset_bytes, cudaStream_t main_stream) {
  cudaStream_t stream = o->lane == 0 ? main_stream : p->side[o->lane - 1];
  switch (o->kind) {
    case dy::OP_SYNC: {
      cudaStream_t sig = o->signaller == 0 ? main_stream : p->side[o->signaller - 1];
      cudaStream_t wai = o->waiter == 0 ? main_stream : p->side[o->waiter - 1];
      DY_CUDA(cudaEventRecord(o->event, sig));
      DY_CUDA(cudaStreamWaitEvent(wai, o->event, 0));
      return DY_OK;
    }
    case dy::OP_CONV: return dy::conv_launch(&o->conv, &o->conv_launch, stream);
    case dy::OP_STEM:
      return dy::stem_launch(static_cast<const char*>(o->in) + in_offset_bytes, o->Cin, o->B, o->H, o->W,
                             o->w, o->b, o->C, o->out, o->out_ld, stream);
    case dy::OP_POOL: return dy::sppf_pool_launch(o->out, o->B, o->H, o->W, o->C, o->out_ld, stream);
    case dy::OP_UPSAMPLE: return dy::upsample2x_launch(o->in, o->in_ld, o->B, o->H, o->W, o->C, o->out, o->out_ld, stream);
    case dy::OP_DWCONV:
      return dy::dwconv_launch(o->in, o->in_ld, o->B, o->H, o->W, o->Cin, o->w, o->b, o->C, o->out, o->out_ld, stream);
    case dy::OP_DECODE: return dy::decode_launch(&o->dec, out_offset_bytes, stream);
    case dy::OP_NMS: return dy::nms_launch(&o->nms, stream);
  }
  return DY_OK;
}

int dy_program_run(dy_program* p, size_t in_offset_bytes, size_t out_offset_bytes, void* stream_) {
  DY_CHECK_ARG(p, "program_run: null program");
  cudaStream_t main_stream = static_cast<cudaStream_t>(stream_);
  // DY_PROGRAM_SYNC=1 (bring-up): wait for every op and name the one that faulted (eager replays only, not under graph capture)
  static const bool sync_each = getenv("DY_PROGRAM_SYNC") != nullptr;
  int op_index = -1;
  for (dy::Op* o : p->ops) {
    ++op_index;
    int rc = program_launch_op(p, o, in_offset_bytes, out_offset_bytes, main_stream);
    if (rc) return rc;
    if (sync_each && o->kind != dy::OP_SYNC) {
      cudaError_t e = cudaStreamSynchronize(o->lane == 0 ? main_stream : p->side[o->lane - 1]);
      if (e != cudaSuccess)
        return dy::fail(DY_ERR_CUDA, "program op %d (kind %d; conv: mode %d, Cout %d, BN %d, map %dx%dx%d, tile %dx%dx%d, m_tiles %d, stages %d, eg %d, fuse2 %d): %s",
                        op_index, int(o->kind), o->conv.mode, o->conv.Cout, o->conv.BN, o->conv.B, o->conv.Ho, o->conv.Wo, o->conv.TW, o->conv.TH,
                        o->conv.TB, o->conv.m_tiles, o->conv.stages, o->conv.eg, o->conv.fuse2, cudaGetErrorString(e));
    }
  }
  return DY_OK;
}

// Per-op device times of one eager replay (the role of the reference's per-layer profile, nn/tasks.py:171-191): every op that
// launches kernels is bracketed by CUDA events on its stream, `reps` back-to-back launches each (the first one is untimed), and
// the mean of the timed ones lands in ms[i]; sync ops get 0.  Not capturable; synchronises the device.
int dy_program_profile(dy_program* p, size_t in_offset_bytes, size_t out_offset_bytes, void* stream_, int reps, float* ms, int n_ms) {
  DY_CHECK_ARG(p && ms && reps >= 1, "program_profile: bad arguments");
  DY_CHECK_ARG(n_ms >= static_cast<int>(p->ops.size()), "program_profile: ms[] holds %d entries, the program has %d ops", n_ms, (int)p->ops.size());
  cudaStream_t main_stream = static_cast<cudaStream_t>(stream_);
  cudaEvent_t e0, e1;
  DY_CUDA(cudaEventCreate(&e0));
  DY_CUDA(cudaEventCreate(&e1));
  int rc = DY_OK, i = 0;
  for (dy::Op* o : p->ops) {
    ms[i] = 0.f;
    cudaStream_t stream = o->lane == 0 ? main_stream : p->side[o->lane - 1];
    if (o->kind == dy::OP_SYNC) { rc = program_launch_op(p, o, in_offset_bytes, out_offset_bytes, main_stream); if (rc) break; ++i; continue; }
    rc = program_launch_op(p, o, in_offset_bytes, out_offset_bytes, main_stream);          // warm (and the one whose result counts)
    if (rc) break;
    if (cudaEventRecord(e0, stream) != cudaSuccess) { rc = dy::fail(DY_ERR_CUDA, "program_profile: event record failed"); break; }
    for (int r = 0; r < reps && rc == DY_OK; ++r) rc = program_launch_op(p, o, in_offset_bytes, out_offset_bytes, main_stream);
    if (rc) break;
    cudaEventRecord(e1, stream);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { rc = dy::fail(DY_ERR_CUDA, "program_profile: op %d: %s", i, cudaGetErrorString(e)); break; }
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    ms[i] = t / static_cast<float>(reps);
    ++i;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int dy_program_num_ops(const dy_program* p) { return p ? static_cast<int>(p->ops.size()) : 0; }

int dy_program_num_launches(const dy_program* p) { return p ? p->launches : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Self-test: a [256 x K] x [N x K]^T GEMM through the tcgen05 conv path (1x1 conv view) against a CUDA-core
// reference, so a wrong descriptor encoding is caught by smoke() in one launch.
// ------------------------------------------------------------------------------------------------
namespace dy {
__global__ void selftest_ref_kernel(const __nv_bfloat16* a, const __nv_bfloat16* w, const float* bias, int M, int N, int K, int Kpad,
                                    float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx % N;
  float acc = bias[n];
  for (int k = 0; k < K; ++k) acc += __bfloat162float(a[m * K + k]) * __bfloat162float(w[n * Kpad + k]);
  out[idx] = acc;
}
__global__ void selftest_fill_kernel(__nv_bfloat16* p, int n, unsigned seed, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (i + 1) * 2654435761u ^ seed;
  x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  p[i] = __float2bfloat16((static_cast<float>(x & 0xffff) / 65536.f - 0.5f) * scale);
}
}  // namespace dy

extern "C" int dy_selftest_umma(int N, int K, float* max_abs_err_host, void* stream_) {
  using namespace dy;
  DY_CHECK_ARG(max_abs_err_host, "selftest: null result pointer");
  DY_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 1024 && K % 8 == 0 && K >= 8 && K <= 4096, "selftest: bad N/K");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int M = 256, Kpad = round_up(K, 64);
  __nv_bfloat16 *a = nullptr, *w = nullptr; float *bias = nullptr, *out = nullptr, *ref = nullptr;
  DY_CUDA(cudaMalloc(&a, sizeof(__nv_bfloat16) * M * K));
  DY_CUDA(cudaMalloc(&w, sizeof(__nv_bfloat16) * N * Kpad));
  DY_CUDA(cudaMalloc(&bias, sizeof(float) * N));
  DY_CUDA(cudaMalloc(&out, sizeof(float) * M * N));
  DY_CUDA(cudaMalloc(&ref, sizeof(float) * M * N));
  DY_CUDA(cudaMemsetAsync(w, 0, sizeof(__nv_bfloat16) * N * Kpad, stream));
  DY_CUDA(cudaMemsetAsync(bias, 0, sizeof(float) * N, stream));
  DY_CUDA(cudaMemsetAsync(out, 0xff, sizeof(float) * M * N, stream));
  selftest_fill_kernel<<<ceil_div(M * K, 256), 256, 0, stream>>>(a, M * K, 1u, 2.f);
  // fill the K valid columns of every weight row (rows are Kpad apart)
  for (int n = 0; n < N; ++n) selftest_fill_kernel<<<ceil_div(K, 256), 256, 0, stream>>>(w + n * Kpad, K, 77u + n, 2.f);
  dy_conv_desc d{};
  d.in = a; d.in_ld = K; d.B = 1; d.H = 1; d.W = M; d.Cin = K; d.weight = w; d.bias = bias; d.Cout = N; d.ksize = 1; d.stride = 1;
  d.out = out; d.out_ld = N; d.out_dtype = DY_F32; d.residual = nullptr; d.res_ld = 0; d.act = DY_ACT_NONE;
  int rc = dy_conv2d(&d, stream);
  if (rc == DY_OK) {
    selftest_ref_kernel<<<ceil_div(M * N, 256), 256, 0, stream>>>(a, w, bias, M, N, K, Kpad, ref);
    rc = launch_status("selftest_ref_kernel");
  }
  float err = NAN;
  if (rc == DY_OK) {
    std::vector<float> ho(M * N), hr(M * N);
    cudaError_t e = cudaMemcpyAsync(ho.data(), out, sizeof(float) * M * N, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hr.data(), ref, sizeof(float) * M * N, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = fail(DY_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
    else {
      err = 0.f;
      for (int i = 0; i < M * N; ++i) {
        const float dlt = std::fabs(ho[i] - hr[i]);
        if (!(dlt <= err)) err = std::isnan(dlt) ? INFINITY : dlt;
      }
    }
  }
  *max_abs_err_host = err;
  cudaFree(a); cudaFree(w); cudaFree(bias); cudaFree(out); cudaFree(ref);
  return rc;
}
