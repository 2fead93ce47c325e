// Parameter blocks of the backward kernels (csrc/backward.cu); internal, not part of the C ABI.
#pragma once
#include "v2v_internal.h"

namespace v2v {

struct BwdConv {
  int N, H, W;                 // forward input extent (unpadded)
  int oh, ow;                  // forward output extent
  int Cin, Cout, kh, kw, stride, pad, transposed, pad_mode;
  ActDesc x;                   // forward input buffer (halo-padded NHWC, bf16 or split)        [weight gradient]
  const float* dy; int dy_C;   // gradient of the conv output, dense NHWC fp32, channel stride dy_C
  const float* w; const float* w2; int Cout1;   // forward weights, torch layout (second set for stacked convs)
  float* dx;                   // gradient of the conv input, dense NHWC fp32 [N][H][W][Cin], accumulated; may be null
  float* dw; float* dw2;       // weight gradients, torch layout, accumulated; may be null
  float* dbias; float* dbias2; // bias gradients; may be null
};

struct NormBwd {
  int N, H, W, C;              // value extent; C = channels of this unit (slice of the raw tensor)
  RawDesc raw; int c_off;      // forward raw conv output (full tensor) and the slice offset
  const float* scale; const float* shift; const float* mean; const float* rstd; int stat_stride;   // [N][stat_stride], slice-offset applied
  int has_norm, batch_stats;   // 0: norm-less bias unit; batch_stats: BatchNorm (one statistic over N)
  int act; float slope;
  const float* dy;             // gradient of the unit output, dense NHWC [.][C]
  float* draw; int draw_C;     // gradient of the raw tensor (full channel stride draw_C), written at c_off
  float* dadd0; float* dadd1;  // gradients of the addends (accumulated); may be null
  float* sums;                 // scratch [2][N][C]
  float* dgamma; float* dbeta; // accumulated; may be null
};

struct HeadBwd {
  int N, H, W, Cout;
  const float* out[V2V_MAX_HEAD];     // forward output plane base (caller tensor) per head channel
  const float* g_ext[V2V_MAX_HEAD];   // caller gradient tensor base per channel (may be null)
  const float* g_int[V2V_MAX_HEAD];   // plan-internal gradient (composite backward) base per channel (may be null)
  long long off[V2V_MAX_HEAD], bstride[V2V_MAX_HEAD];
  int act[V2V_MAX_HEAD]; float scale[V2V_MAX_HEAD];
  float* dz; int dz_C;                // dense NHWC fp32 [N][H][W][dz_C]
};

struct CompositeBwd {
  int N, H, W, prev_C, use_warp, align_corners;
  const float* raw; const float* flow; const float* weight; const float* prev; const float* mask;   // forward tensors (raw = head output)
  const float* g_final; const float* g_rawout;                                                      // incoming gradients (may be null)
  float* d_raw; float* d_flow; float* d_weight; float* d_fg;                                          // written
};

cudaError_t launch_conv_bwd(const BwdConv& p, cudaStream_t s);
cudaError_t launch_norm_bwd(const NormBwd& p, cudaStream_t s);
cudaError_t launch_head_bwd(const HeadBwd& p, cudaStream_t s);
cudaError_t launch_composite_bwd(const CompositeBwd& p, cudaStream_t s);
cudaError_t launch_grad_import(const float* g, float* dst, int N, int C_src, int c_off, int C, int H, int W, cudaStream_t s);
cudaError_t launch_grad_export(const float* src, float* g, int N, int C_src, int c_off, int C, int H, int W, cudaStream_t s);
cudaError_t launch_convact_bwd(const float* dy, const ActDesc& out, int act, float slope, float* dz, int C, cudaStream_t s);

}  // namespace v2v
