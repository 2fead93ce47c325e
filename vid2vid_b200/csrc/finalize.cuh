// Batch / instance-norm statistics finalisation for ONE channel by one warp (lanes stride the per-CTA partial rows, fp64
// shuffle reduction): scale / shift for the normalise pass, saved mean / rstd for training plans, running-stat side effect.
// Shared by stats_finalize_kernel (csrc/norm.cu) and the tail of conv_umma_kernel (fused finalisation).
#pragma once
#include "v2v_internal.h"

namespace v2v {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool kL2>
__device__ __forceinline__ void finalize_channel(const FinalizeParams& p, int c, int lane) {
  const int rows_img = p.num_phases * p.tiles_per_img;
  double bs = 0.0, bq = 0.0, rm_acc = 0.0, rv_acc = 0.0;
  for (int n = 0; n < p.N; ++n) {
    double s = 0.0, q = 0.0;
    for (int i = lane; i < rows_img; i += 32) {
      const int ph = i / p.tiles_per_img, t = i - ph * p.tiles_per_img;
      const size_t row = (size_t)ph * p.N * p.tiles_per_img + (size_t)n * p.tiles_per_img + t;
      const float* a = p.stats + (row * 2 + 0) * p.Cs + p.c_off + c;
      const float* b = p.stats + (row * 2 + 1) * p.Cs + p.c_off + c;
      s += (double)(kL2 ? __ldcg(a) : *a);      // kL2: partials written by other CTAs of the same launch -> read at L2
      q += (double)(kL2 ? __ldcg(b) : *b);
    }
    s = warp_sum_d(s); q = warp_sum_d(q);
    if (p.instance) {
      const double mean = s / p.count;
      double var = q / p.count - mean * mean;
      if (var < 0) var = 0;
      if (lane == 0) {
        const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
        const float sc = g * (float)(1.0 / sqrt(var + (double)p.eps));
        p.scale[(size_t)n * p.scale_stride + p.c_off + c] = sc;
        p.shift[(size_t)n * p.scale_stride + p.c_off + c] = b - (float)mean * sc;
        if (p.mean_out) {
          p.mean_out[(size_t)n * p.scale_stride + p.c_off + c] = (float)mean;
          p.rstd_out[(size_t)n * p.scale_stride + p.c_off + c] = (float)(1.0 / sqrt(var + (double)p.eps));
        }
      }
      rm_acc += mean;
      rv_acc += var * (p.count / (p.count > 1 ? p.count - 1 : 1));
    } else {
      bs += s; bq += q;
    }
  }
  if (lane != 0) return;
  double mean_run, var_run;
  if (p.instance) {
    mean_run = rm_acc / p.N; var_run = rv_acc / p.N;
  } else {
    const double cnt = p.count * p.N;
    const double mean = bs / cnt;
    double var = bq / cnt - mean * mean;
    if (var < 0) var = 0;
    const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
    const float sc = g * (float)(1.0 / sqrt(var + (double)p.eps));
    for (int n = 0; n < p.N; ++n) {
      p.scale[(size_t)n * p.scale_stride + p.c_off + c] = sc;
      p.shift[(size_t)n * p.scale_stride + p.c_off + c] = b - (float)mean * sc;
      if (p.mean_out) {
        p.mean_out[(size_t)n * p.scale_stride + p.c_off + c] = (float)mean;
        p.rstd_out[(size_t)n * p.scale_stride + p.c_off + c] = (float)(1.0 / sqrt(var + (double)p.eps));
      }
    }
    mean_run = mean; var_run = var * (cnt / (cnt > 1 ? cnt - 1 : 1));
  }
  if (p.running_mean) {   // train-mode side effect of nn.BatchNorm2d / InstanceNorm2d(track_running_stats)
    const float bias = p.conv_bias ? p.conv_bias[c] : 0.f;
    p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * ((float)mean_run + bias);
    p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)var_run;
    if (c == 0 && p.num_batches_tracked) *p.num_batches_tracked += 1;
  }
}

}  // namespace v2v
