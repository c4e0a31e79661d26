// conv_igemm.h — parameter block of the tcgen05 implicit-GEMM conv kernel (see conv_igemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/droneyolo.h"

namespace dy {

struct ConvParams {
  CUtensorMap tmA[4];      // activation views (1 for stride 1, 4 parity views for stride 2)
  CUtensorMap tmB;         // packed weights [tap][Cout_pad][Cin_pad]
  CUtensorMap tmO;         // output slice (TMA-store epilogue)
  CUtensorMap tmR;         // residual slice, same box as tmO (TMA-prefetched into the staging tile)
  CUtensorMap tmU[4];      // fused 2x nearest upsample: the four (dy,dx) parity views of the upsampled destination
  CUtensorMap tmW2;        // fused 1x1 tail: packed weights [1][N2][64], box = one 32-channel k-block (64B swizzle)
  CUtensorMap tmO2;        // fused 1x1 tail: fp32 output slice, box {32, TW, TH, TB}
  CUtensorMap tmP;         // half-resolution fp32 addend (dy_conv_desc.pre_add), box {32, TW/2, TH/2, TB}
  int nmaps, ntaps, kblocks;
  int BN, n_tiles, n_split;            // n_split > 1: every CTA owns ONE n tile for its whole life (weights resident per CTA)
  int TW, TH, TB, tiles_w, tiles_h, m_tiles;
  int B, Ho, Wo, Cout;
  int stages, stage_bytes, kps;        // pipeline stages, bytes per stage, 64-channel k-blocks per stage
  int nacc, b_resident, mode;
  int a_stages, a_stage_bytes;         // mode 6: the halo-tile ring in front of the weight ring (`stages` x `stage_bytes`)
  int stg_off;                         // byte offset of the epilogue staging tiles (behind resident weights and pipeline stages)
  int halo_pitch;                      // pixels per halo-tile row in shared memory (TW + 2)
  int use_tma_store;                   // chunk width CW of the TMA-store epilogue (0 = generic register->global path)
  int has_res_tma;
  int has_up;
  int fuse2, N2;                       // fused 1x1 tail (Detect output conv): N2 = Cout2 padded to 16
  const float* bias2;
  int tail_decode, y_nc, y_A; float y_stride; float* y;   // fused Detect decode of the tail's logits (tmO2 then maps the prediction tensor)
  int nbuf;                            // staging tiles per epilogue group (2, or 1 when shared memory is short)
  int bias_off;                        // byte offset of the bias table behind the staging tiles
  int eg;                              // epilogue groups: 2, or 3 for the 32-channel halo layers
  int dbg;
  unsigned long long* trace;           // debug builds only (DY_CONV_TRACE)
  void* out; int out_ld; int out_f32;
  const __nv_bfloat16* res; int res_ld;
  const float* bias; int act;
  const float* pre; int pre_ld;        // half-resolution fp32 addend in front of the activation (dy_conv_desc.pre_add), or null
  int res_off;                         // byte offset of the residual ring: [group][3] staging-sized tiles (TMA-prefetched two chunks ahead)
  int res_mode;                        // residual staging: 2 = that ring, 1 = one chunk ahead into the other staging tile, 0 = per chunk into the single staging tile
  int pre_off;                         // byte offset of its tiles in shared memory: [group][2] x (TW/2 * TH/2 * TB rows x 128 B)
};

struct ConvLaunch { int grid; int smem_bytes; };

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); strides_bytes has rank-1 entries
int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
               CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
int conv_build_params(const dy_conv_desc* d, ConvParams* p, ConvLaunch* l);
int conv_launch(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream);

}  // namespace dy
