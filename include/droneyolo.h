/*
 * droneyolo.h — C-ABI of libdroneyolo.so: the B200 (sm_100a) kernels behind the Drone-YOLO
 * inference hot path (conv stack -> Detect decode -> NMS).
 *
 * The reference (331658260/Drone-YOLO, a fork of ultralytics 8.3.82) has no FFI of its own: the
 * path sits behind Python modules that dispatch into aten/cuDNN and torchvision.  Each entry point
 * below names the reference interface it replaces (file:line under the reference tree).  Calling
 * convention for all of them:
 *   - plain pointers and sizes; the caller (PyTorch, or any host) owns every buffer;
 *   - device pointers unless the name says "host"; asynchronous on the given cudaStream_t
 *     (passed as void* so this header needs no CUDA include);
 *   - returns DY_OK (0) or a negative dy_status; never throws; dy_last_error() gives the text
 *     of the last failure on the calling thread.
 *   - activations are NHWC ("channels last"); a tensor may be a channel slice of a wider buffer,
 *     described by its pixel stride `ld` (in elements).
 */
#ifndef DRONEYOLO_H_
#define DRONEYOLO_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dy_status {
  DY_OK = 0,
  DY_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer)            */
  DY_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed                       */
  DY_ERR_UNSUPPORTED = -3, /* valid request the kernels do not implement              */
  DY_ERR_NOMEM = -4
} dy_status;

typedef enum dy_dtype { DY_BF16 = 0, DY_F32 = 1, DY_U8 = 2 } dy_dtype;
typedef enum dy_act { DY_ACT_NONE = 0, DY_ACT_SILU = 1 } dy_act;
typedef enum dy_layout { DY_NHWC = 0, DY_NCHW = 1 } dy_layout;

/* Version / capability ------------------------------------------------------------------- */
int dy_version(void);                       /* 10000*major + 100*minor + patch           */
const char* dy_last_error(void);            /* thread-local text of the last error       */
int dy_device_check(int device);            /* DY_OK iff `device` is compute capability 10.x */

/* ----------------------------------------------------------------------------------------
 * Conv + folded-BN bias + SiLU (+ residual) as an implicit GEMM on tcgen05 tensor cores.
 * Replaces: Conv.forward_fuse  ultralytics/nn/modules/conv.py:53-55  (act(conv(x)) after
 *           fuse_conv_and_bn, utils/torch_utils.py:242-269), the deploy form of RepVGGBlock
 *           (nn/modules/block.py:1421-1438,1480-1482), RepConv.forward_fuse (conv.py:198-200),
 *           Bottleneck's residual add (block.py:348-350: x + cv2(cv1(x))), the bare nn.Conv2d
 *           1x1 outputs of Detect (head.py:44,47) and the concat writes of C2f / Concat
 *           (block.py:240-242, conv.py:331-333) via out/out_ld channel slices.
 * in      : bf16 NHWC [B,H,W,Cin] slice, pixel stride in_ld elements (in_ld % 8 == 0, 16B aligned)
 * weight  : bf16 packed [k*k][Cout_pad][Cin_pad] (tap-major, K contiguous), Cin_pad = ceil64(Cin),
 *           Cout_pad = ceil16(Cout); BN already folded in fp32 by the host before the bf16 cast
 * bias    : fp32 [Cout_pad]
 * out     : bf16 or fp32 NHWC [B,Ho,Wo,Cout] slice, pixel stride out_ld; Ho=(H+2p-k)/s+1, p=k/2
 * residual: optional bf16 NHWC [B,Ho,Wo,Cout] slice added AFTER the activation, or NULL
 * up_out  : optional second destination, bf16 NHWC [B,2*Ho,2*Wo,Cout] slice with pixel stride up_ld: the
 *           output is ALSO written 2x nearest-upsampled (nn.Upsample(None, 2, 'nearest') of the model YAMLs,
 *           cfg/models/v8/yolov8-p2-repvgg.yaml:30,34,38, fused into its producer), or NULL
 * weight2 : optional fused 1x1 tail (the Detect branch ends `Conv(c,c,3) -> nn.Conv2d(c, n, 1)`, nn/modules/head.py:41-47):
 *           out2 = conv1x1(act(conv(in)+bias), weight2) + bias2, fp32, no activation; the intermediate never leaves the
 *           SM and `out` is not written (may be NULL).  weight2 is bf16 packed [1][ceil16(Cout2)][64], bias2 fp32
 *           [ceil16(Cout2)], out2 an fp32 NHWC [B,Ho,Wo,Cout2] slice with pixel stride out2_ld.  Requires ksize 3,
 *           stride 1, 32 < Cin <= 64, Cout == 64, Cout2 <= 64 and a map the 8x16 halo tiles cover well; anything else
 *           returns DY_ERR_UNSUPPORTED (run the two convs separately).
 * tail_decode : with weight2, decode the tail's logits in the epilogue instead of writing them (Detect._inference,
 *           head.py:100-131, for this level and branch): 1 = box branch (Cout2 == 64: DFL expectation, dist2bbox, * stride
 *           -> rows 0..3 of y), 2 = class branch (sigmoid -> rows 4..4+y_nc-1 of y).  y = fp32 [B, 4+y_nc, y_A] (the
 *           dy_detect_decode output layout), y_anchor_off = first anchor of this level; out2 is then unused.
 * pre_add : optional fp32 NHWC [B,Ho/2,Wo/2,Cout] slice (pixel stride pre_ld) added BEFORE the activation at half resolution:
 *           out = act(conv(in) + bias + pre_add[b, y/2, x/2, :]).  A 1x1 conv commutes with nearest upsampling, so the first conv
 *           behind `Concat([nn.Upsample(2x)(a), b])` (C2f.cv1 of the top-down neck, cfg/models/v8/yolov8-p2-repvgg.yaml:30-41;
 *           nn/modules/block.py:236) is W_b*b + up(W_a*a): W_a*a is computed once at LOW resolution (this tensor) and the upsampled
 *           tensor is never written or read back.  Requires ksize 1, stride 1, bf16 out, even Ho and Wo, Cout % 8 == 0.
 *           3 = the tail is a hidden Conv + SiLU (C2f.cv1 behind the stride-2 conv that feeds it, nn/modules/block.py:227-249):
 *           out2 = SiLU(conv1x1(...) + bias2) as a bf16 NHWC slice; also accepted for ksize 3, stride 2, Cin <= 32, Cout == 64.
 * ksize in {1,3}; stride in {1,2} (stride 2 needs even H and W); no dilation, no groups.
 */
typedef struct dy_conv_desc {
  const void* in;      int32_t in_ld;
  int32_t B, H, W, Cin;
  const void* weight;  const float* bias;
  int32_t Cout, ksize, stride;
  void* out;           int32_t out_ld;  int32_t out_dtype;   /* dy_dtype */
  const void* residual; int32_t res_ld;
  int32_t act;                                               /* dy_act   */
  void* up_out;        int32_t up_ld;
  const void* weight2; const float* bias2;
  int32_t Cout2;       void* out2;      int32_t out2_ld;
  int32_t tail_decode; float* y;        int32_t y_A, y_nc, y_anchor_off; float y_stride;
  const float* pre_add; int32_t pre_ld;
} dy_conv_desc;

int dy_conv2d(const dy_conv_desc* d, void* stream);

/* Stem conv: 3x3 stride-2 pad-1 Conv(3->Cout)+bias+SiLU reading the predictor's input tensor.
 * Replaces: model.0 Conv of the YAMLs (cfg/models/v8/yolov8-p2-repvgg.yaml:17) together with the
 *           dtype cast of BasePredictor.preprocess (engine/predictor.py:132-135).
 * in      : NCHW [B,3,H,W]; in_dtype DY_F32: values already in [0,1] (as LoadTensor supplies, data/loaders.py:516-584);
 *           in_dtype DY_U8: raw 0..255 pixels as the predictor uploads them for image sources — the `/ 255` of
 *           engine/predictor.py:134-135 is applied inside the kernel
 * weight  : fp32 [Cout][27] (cin-major: c*9 + ky*3 + kx), bias fp32 [Cout]; Cout % 8 == 0, <= 128
 * out     : bf16 NHWC [B,H/2,W/2,Cout] slice with pixel stride out_ld
 */
int dy_stem_conv(const void* in, int in_dtype, int B, int H, int W, const float* weight, const float* bias, int Cout,
                 void* out, int out_ld, void* stream);

/* SPPF pooling: y1=mp5(x), y2=mp5(y1), y3=mp5(y2) (5x5, stride 1, pad 2, -inf padding) in one pass.
 * Replaces: SPPF.forward's three chained MaxPool2d + cat  nn/modules/block.py:185-191.
 * buf     : bf16 NHWC [B,H,W,ld]; reads channels [0,C), writes [C,2C), [2C,3C), [3C,4C).
 */
int dy_sppf_pool(void* buf, int B, int H, int W, int C, int ld, void* stream);

/* 2x nearest upsample into a channel slice of a concat buffer.
 * Replaces: nn.Upsample(None,2,'nearest') + Concat  (yaml :30-31; conv.py:331-333).
 */
int dy_upsample2x(const void* in, int in_ld, int B, int H, int W, int C, void* out, int out_ld, void* stream);

/* Depthwise-style grouped conv of the "-sf" graph: 3x3 stride 2, groups = Cout, Cin = 2*Cout.
 * Replaces: DWConv(c1,c2,3,2) nn/modules/conv.py:102-107 (g = gcd(c1,c2)) of
 *           cfg/models/v8/yolov8-p2-repvgg-sf.yaml:32,38,44.  weight fp32 [Cout][2][9], bias fp32 [Cout].
 */
int dy_dwconv3x3s2(const void* in, int in_ld, int B, int H, int W, int Cin, const float* weight,
                   const float* bias, int Cout, void* out, int out_ld, void* stream);

/* Letterbox preprocess of ONE image: bilinear resize + constant border + BGR->RGB + HWC->CHW, uint8 in, uint8 out.
 * Replaces: LetterBox.__call__  data/augment.py:1544-1610 (cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114)) and the
 *           `[..., ::-1].transpose(0, 3, 1, 2)` of BasePredictor.preprocess  engine/predictor.py:127-131, fused.
 * src     : uint8 HWC BGR [h][w][3] on the device, row pitch src_pitch bytes (the raw frame, uploaded as is)
 * dst     : uint8 CHW RGB [3][H][W] on the device (one image of the engine's uint8 input batch)
 * new_w/new_h : size of the resized image inside the canvas, left/top : its offset (the reference's rounding of dw, dh);
 *           every other pixel is `fill` (114).  The resize follows OpenCV's 8-bit INTER_LINEAR fixed-point arithmetic
 *           (11-bit coefficients): bit-exact with cv2 when shrinking, within 1 level on < 0.1 % of the pixels when enlarging.
 */
int dy_letterbox_u8(const void* src, int h, int w, int src_pitch, void* dst, int H, int W, int new_w, int new_h,
                    int left, int top, int fill, void* stream);
/* The same for n frames of ONE geometry in one launch: frame i is read at src + i*src_image_stride bytes and written at
 * dst + i*dst_image_stride bytes (dst_image_stride a multiple of 4; 3*H*W for consecutive images of the input batch). */
int dy_letterbox_u8_batch(const void* src, int n, size_t src_image_stride, int h, int w, int src_pitch, void* dst,
                          size_t dst_image_stride, int H, int W, int new_w, int new_h, int left, int top, int fill, void* stream);

/* ----------------------------------------------------------------------------------------
 * Detect decode: DFL softmax-expectation + dist2bbox(xywh) + stride scaling + class sigmoid.
 * Replaces: Detect._inference  nn/modules/head.py:100-131, DFL.forward block.py:73-76,
 *           make_anchors utils/tal.py:333-345, dist2bbox utils/tal.py:348-357.
 * lvl[l]  : raw head map of level l, `no = 64 + nc` channels (box side*16+bin first, then classes):
 *           DY_NHWC: [B,H_l,W_l,ld[l]] (bf16 or fp32);  DY_NCHW: [B,no,H_l,W_l] (ld ignored)
 * out     : fp32 [B, 4+nc, A], A = sum H_l*W_l, rows cx,cy,w,h (input pixels), p_0..p_{nc-1};
 *           anchors level-major, row-major (y,x) inside a level — the reference's order.
 */
typedef struct dy_decode_desc {
  const void* lvl[4]; int32_t ld[4]; int32_t H[4]; int32_t W[4]; float stride[4];
  int32_t nl, B, nc, dtype /* dy_dtype */, layout /* dy_layout */;
  float* out;
  int32_t A_total;          /* anchors per image in `out`; 0 = the sum over the levels given                 */
  int32_t anchor_off[4];    /* first anchor of each level in `out` when A_total != 0 (levels decoded elsewhere) */
} dy_decode_desc;

int dy_detect_decode(const dy_decode_desc* d, void* stream);

/* ----------------------------------------------------------------------------------------
 * Batched NMS.  Replaces: ops.non_max_suppression  utils/ops.py:181-332 (confidence filter :250,
 * xywh2xyxy :259-260/:432-449, best-class / multi-label candidates :284-291, class filter :294-295,
 * max_nms truncation :301-302, class-offset boxes :305-311) and torchvision.ops.nms (:312),
 * [:max_det] (:313) and the row gather (:327).
 * pred    : fp32 [B, 4+nc, A]  rows cx,cy,w,h,p_0..p_{nc-1}
 * out     : fp32 [B, max_det, 6] rows x1,y1,x2,y2,conf,cls (only the first counts[b] rows are written)
 * counts  : int32 [B]
 * kept    : optional int64 [B, max_det]: index of each kept box in the reference's compacted candidate
 *           list `x` (ops.py:269-302) i.e. what torchvision.ops.nms returned; NULL to skip
 * classes_host : optional HOST int32 [n_classes] class filter (ops.py:294-295) or NULL
 * workspace: device scratch of dy_nms_workspace_bytes(B, nc, A, multi_label) bytes (16B aligned)
 * Ties: equal scores keep the lower candidate index first (stable order), also in the max_nms cut.
 */
typedef struct dy_nms_desc {
  const float* pred; int32_t B, nc, A;
  float conf_thres;          /* compared in fp32, as torch compares a float32 tensor with a Python scalar */
  double iou_thres;          /* torchvision's CPU kernel compares the fp32 IoU against this double        */
  int32_t max_det, max_nms; float max_wh;
  int32_t agnostic, multi_label;
  const int32_t* classes_host; int32_t n_classes;   /* HOST array (ops.py:294-295) or NULL              */
  int32_t xyxy_in_place;     /* 1: also overwrite pred rows 0..3 with x1,y1,x2,y2 (ops.py:259-260) */
  float* out; int32_t* counts; int64_t* kept;
  void* workspace; size_t workspace_bytes;
  /* Optional fused Results post-step (ops.scale_boxes + ops.clip_boxes, utils/ops.py:92-127, 335-354, as
   * DetectionPredictor.construct_result calls them, models/yolo/detect/predict.py:66-73): DEVICE fp32 [B][8] rows
   * (pad_x, pad_y, gain, w0, h0, -, -, -) read when the output rows are written: x = clamp((x - pad_x) / gain, 0, w0),
   * y = clamp((y - pad_y) / gain, 0, h0) in the reference's fp32 operation order.  NULL: rows stay in input pixels. */
  const float* rescale;
} dy_nms_desc;

size_t dy_nms_workspace_bytes(int B, int nc, int A, int multi_label);
int dy_nms(const dy_nms_desc* d, void* stream);

/* ----------------------------------------------------------------------------------------
 * Tile merge: category-aware greedy NMS over the detections of all tiles of ONE frame (frame coordinates, float64).
 * Replaces: the overlap filter of supervision.InferenceSlicer as mix6.py:84-89 configures it (Detections.with_nms ->
 *           box_non_max_suppression; third-party, not under the reference tree: restated in oracle/slicer_np.py).
 * rows    : float64 [n, 6] on the device: x1,y1,x2,y2,conf,cls
 * keep    : uint8 [n] on the device, ORIGINAL row order: 1 = survives
 * Ranking by conf descending, equal conf: higher row first; a kept row suppresses later-ranked rows of the same cls
 * (any cls when class_agnostic) with IoU > iou_thres (strict), IoU = inter / (area_a + area_b - inter) in float64.
 * workspace: device scratch of dy_box_nms_f64_workspace_bytes(n) bytes (8B aligned); n <= 16384.
 */
size_t dy_box_nms_f64_workspace_bytes(int n);
int dy_box_nms_f64(const double* rows, int n, double iou_thres, int class_agnostic, unsigned char* keep, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------
 * Program: a recorded sequence of the ops above with tensor maps encoded once, replayed per batch.
 * Replaces: the Python layer loop BaseModel._predict_once  nn/tasks.py:134-161 for a fixed
 *           (batch, H, W).  `dy_program_run` only enqueues kernels (capturable in a CUDA graph).
 */
typedef struct dy_program dy_program;
int dy_program_create(dy_program** out);
void dy_program_destroy(dy_program* p);
int dy_program_add_conv(dy_program* p, const dy_conv_desc* d);
int dy_program_add_stem(dy_program* p, const void* in, int in_dtype, int B, int H, int W, const float* weight,
                        const float* bias, int Cout, void* out, int out_ld);
int dy_program_add_sppf_pool(dy_program* p, void* buf, int B, int H, int W, int C, int ld);
int dy_program_add_upsample2x(dy_program* p, const void* in, int in_ld, int B, int H, int W, int C,
                              void* out, int out_ld);
int dy_program_add_dwconv3x3s2(dy_program* p, const void* in, int in_ld, int B, int H, int W, int Cin,
                               const float* weight, const float* bias, int Cout, void* out, int out_ld);
int dy_program_add_decode(dy_program* p, const dy_decode_desc* d);
int dy_program_add_nms(dy_program* p, const dy_nms_desc* d);
/* Lanes: ops added after dy_program_set_lane(p, k) are enqueued on the program's k-th side stream (k = 0: the stream
 * given to dy_program_run; 1 <= k <= 8: created by this call, lowest priority).  dy_program_add_sync(p, waiter, signaller)
 * makes lane `waiter` wait (event record + stream wait, capturable) for everything added so far on lane `signaller`.
 * The caller forks every side lane from lane 0 before its first op and joins it back into lane 0 after its last one.
 * Replaces: nothing in the reference (its layer loop nn/tasks.py:134-161 is strictly sequential on the default stream);
 * the branches are the per-level Detect convs of nn/modules/head.py:64-74, which only depend on their own level. */
int dy_program_set_lane(dy_program* p, int lane);
int dy_program_add_sync(dy_program* p, int waiter, int signaller);
/* in_offset_bytes / out_offset_bytes are added to the stem input pointer and to the decode output
 * pointer: the same program serves successive micro-batches of one large resident batch. */
int dy_program_run(dy_program* p, size_t in_offset_bytes, size_t out_offset_bytes, void* stream);
int dy_program_num_launches(const dy_program* p);   /* kernels enqueued by one dy_program_run */
/* Per-op device times of one eager replay: the role of the reference's per-layer profile (`BaseModel._profile_one_layer`,
 * ultralytics/nn/tasks.py:171-191, enabled by predict(profile=True), :116,151-152).  Every op is launched `reps` + 1 times in place
 * (idempotent: an op only reads its inputs and rewrites its outputs), the last `reps` bracketed by CUDA events; ms[i] = mean
 * milliseconds of op i in the order the ops were added (sync ops: 0).  Synchronises the device; not capturable. */
int dy_program_profile(dy_program* p, size_t in_offset_bytes, size_t out_offset_bytes, void* stream, int reps, float* ms, int n_ms);
int dy_program_num_ops(const dy_program* p);        /* ops (including sync ops) added so far */

/* Self-test of the tcgen05 / TMA descriptor encodings (one 128xN tile), used by tests and smoke. */
int dy_selftest_umma(int N, int K, float* max_abs_err_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* DRONEYOLO_H_ */
