// conv_igemm.h — parameter block of the tcgen05 implicit-GEMM conv kernel (see conv_igemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/droneyolo.h"

namespace dy {

struct ConvTap { int16_t map, dx, dy, pad; };   // which A tensor map, and the box shift in that map's pixels

struct ConvParams {
  CUtensorMap tmA[4];      // activation views (1 for stride 1, 4 parity views for stride 2)
  CUtensorMap tmB;         // packed weights [tap][Cout_pad][Cin_pad]
  CUtensorMap tmO;         // bf16 output slice (TMA-store epilogue)
  ConvTap taps[9];
  int nmaps, ntaps, kblocks;
  int BN, n_tiles;
  int TW, TH, TB, tiles_w, tiles_h, m_tiles;
  int B, Ho, Wo, Cout;
  int stages, nacc, b_resident, mode, halo_base_offset;
  int use_tma_store;
  int dbg;
  void* out; int out_ld; int out_f32;
  const __nv_bfloat16* res; int res_ld;
  const float* bias; int act;
};

struct ConvLaunch { int grid; int smem_bytes; };

int conv_pick_bn(int cout_pad);
int conv_build_params(const dy_conv_desc* d, ConvParams* p, ConvLaunch* l);
int conv_launch(const ConvParams* p, const ConvLaunch* l, cudaStream_t stream);

}  // namespace dy
