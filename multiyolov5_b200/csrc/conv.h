// Host-side description of one fused conv op (Conv / Bottleneck.cv2 / bare Conv2d+BN+SiLU / Detect.m[i] / classifier).
#pragma once
#include "common.cuh"

namespace myolo {

struct ConvTcParams {
  int B, Ho, Wo;
  int tw, th;            // output tile = tw x th pixels, tw*th == 128 (UMMA M)
  int tiles_x, tiles_y;  // per image
  int n_tiles_n, total_tiles;
  int BN;                // UMMA N (multiple of 16, <= 128)
  int Co;                // real output channels
  int kc;                // channels per K chunk: 16 / 32 / 64  (swizzle 32B / 64B / 128B)
  int cblocks;           // Ci_pad / kc
  int taps;              // k*k
  int n_chunks, chunks_per_stage, n_kstages;
  int tap_map[9], tap_dx[9], tap_dy[9];
  int act;
  int num_stages;
  int out_mode;          // 0: fp16 NHWC slice via TMA store, 1: fp32 NHWC direct stores
  int ow;                // output sub-box width in channels (16/32/64)
  int n_sub;             // sub-boxes per N tile
  int log2_tw, log2_ow;
  int ws_mode;           // weights-stationary: the whole [BN x K] weight tile stays resident in shared memory
  int b_res_bytes;       // bytes of the resident weight region (ws_mode)
  int n_stg;             // TMA-store staging buffers (1 or 2)
  const float* bias;
  const __half* residual;  // nullable; base of the residual slice (image 0, pixel 0, channel 0 of the slice)
  int res_ctot;
  float* out_f32;
  int out_f32_ctot;
  long long* dbg;        // optional timeline buffer [64 tiles][16] of clock64 stamps for CTA 0 (MYOLO_CONV_TIMELINE=1), else null
};

struct ConvOp {
  TensorView in, out, res;
  bool has_res = false;
  int k = 1, stride = 1, dil = 1, act = MYOLO_ACT_SILU;
  const __half* w = nullptr;  // packed [Co_pad][k*k*Ci_pad]
  const float* bias = nullptr;
  int Ci_pad = 0, Co_pad = 0, Co = 0;
  bool use_tc = false;
  // tcgen05 path state (built once at plan creation)
  CUtensorMap tmA[4], tmB, tmO;
  ConvTcParams p;
  int grid = 0, smem = 0;
};

// decides whether the tcgen05 path can run this op; fills tile geometry (no device pointers needed)
bool conv_tc_eligible(const ConvOp& op);
// builds tensor maps / params; requires op.in/out/res/w/bias device pointers to be final
int conv_tc_prepare(ConvOp& op, int num_sms);
int conv_tc_launch(const ConvOp& op, cudaStream_t stream);
int conv_simt_launch(const ConvOp& op, cudaStream_t stream);

// weight packing: fp32 [Co][Ci][k][k] (+BN) -> fp16 [Co_pad][k*k][Ci_pad], bias fp32 [Co_pad]
int pack_conv_weights(const float* w, int co, int ci, int k, const float* gamma, const float* beta, const float* mean,
                      const float* var, float eps, const float* bias, __half* wp, float* bp, int co_pad, int ci_pad,
                      cudaStream_t stream);

}  // namespace myolo
