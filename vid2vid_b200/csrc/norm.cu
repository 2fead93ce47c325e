// Batch/Instance-norm statistics finalisation and the fused
//   normalise * gamma + beta -> activation -> (+ residual / skip addends) -> write next layer's
//   halo-padded NHWC bf16 buffer (reflect or zero halo, optional stride-2 parity split)
// pass.  HBM-bound: reads 2 B/elem raw (+2 B per addend), writes 2 B/elem.
// Reference ops replaced: nn.BatchNorm2d (train-mode statistics) /
// nn.InstanceNorm2d, nn.ReLU / LeakyReLU, nn.ReflectionPad2d and the residual adds at
// models/networks.py:23-30,126,204,298-305,559-593,687-703.
#include "ptx.cuh"
#include "v2v_internal.h"

namespace v2v {

__global__ void stats_finalize_kernel(FinalizeParams p) {
  const int c = blockIdx.x;
  __shared__ double sh[2][128];
  double bs = 0.0, bq = 0.0;       // batch totals (thread 0 only)
  double rm_acc = 0.0, rv_acc = 0.0;
  for (int n = 0; n < p.N; ++n) {
    double s = 0.0, q = 0.0;
    const int rows_img = p.num_phases * p.tiles_per_img;
    for (int i = threadIdx.x; i < rows_img; i += blockDim.x) {
      const int ph = i / p.tiles_per_img, t = i - ph * p.tiles_per_img;
      const size_t row = (size_t)ph * p.N * p.tiles_per_img + (size_t)n * p.tiles_per_img + t;
      s += (double)p.stats[(row * 2 + 0) * p.Cs + c];
      q += (double)p.stats[(row * 2 + 1) * p.Cs + c];
    }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
      if (threadIdx.x < k) { sh[0][threadIdx.x] += sh[0][threadIdx.x + k]; sh[1][threadIdx.x] += sh[1][threadIdx.x + k]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      s = sh[0][0]; q = sh[1][0];
      if (p.instance) {
        const double mean = s / p.count;
        double var = q / p.count - mean * mean;
        if (var < 0) var = 0;
        const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
        const float sc = g * (float)(1.0 / sqrt(var + (double)p.eps));
        p.scale[(size_t)n * p.C + c] = sc;
        p.shift[(size_t)n * p.C + c] = b - (float)mean * sc;
        rm_acc += mean;
        rv_acc += var * (p.count / (p.count > 1 ? p.count - 1 : 1));
      } else {
        bs += s; bq += q;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
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
        p.scale[(size_t)n * p.C + c] = sc;
        p.shift[(size_t)n * p.C + c] = b - (float)mean * sc;
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
}

__device__ __forceinline__ int reflect_idx(int i, int n) {   // nn.ReflectionPad2d index map
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// one thread per (padded pixel, 8-channel vector)
__global__ void norm_apply_kernel(ApplyParams p) {
  const int vecs = p.out.C / 8;
  const int Hpad = p.out.H + p.out.pad_t + p.out.pad_b, Wpad = p.out.W + p.out.pad_l + p.out.pad_r;
  const long long total = (long long)p.out.N * Hpad * Wpad * vecs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vecs);
    long long t = idx / vecs;
    const int xp = (int)(t % Wpad); t /= Wpad;
    const int yp = (int)(t % Hpad);
    const int n = (int)(t / Hpad);
    int y = yp - p.out.pad_t, x = xp - p.out.pad_l;
    const bool halo = (y < 0 || y >= p.out.H || x < 0 || x >= p.out.W);
    uint4 o = make_uint4(0, 0, 0, 0);
    const int c0 = v * 8;
    bool zero = (c0 >= p.raw.Cvalid);
    if (halo) {
      if (p.pad_mode == PAD_REFLECT) { y = reflect_idx(y, p.out.H); x = reflect_idx(x, p.out.W); }
      else zero = true;
    }
    if (!zero) {
      const uint4 r = *reinterpret_cast<const uint4*>(p.raw.base + (((size_t)n * p.raw.H + y) * p.raw.W + x) * p.raw.C + c0);
      float f[8];
      const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 a = __bfloat1622float2(rp[j]); f[2 * j] = a.x; f[2 * j + 1] = a.y; }
      if (p.scale) {
        const float* sc = p.scale + (size_t)n * p.raw.Cvalid + c0;
        const float* sh = p.shift + (size_t)n * p.raw.Cvalid + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c0 + j < p.raw.Cvalid) ? fmaf(f[j], sc[j], sh[j]) : 0.f;
      }
      if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (p.act == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * p.slope;
      }
      for (int a = 0; a < p.n_add; ++a) {
        const ActDesc& ad = p.add[a];
        if (c0 < ad.C) {
          const uint4 q = *reinterpret_cast<const uint4*>(ad.base + ad.offset(n, y, x) + c0);
          const __nv_bfloat162* qp = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int j = 0; j < 4; ++j) { float2 b = __bfloat1622float2(qp[j]); f[2 * j] += b.x; f[2 * j + 1] += b.y; }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c0 + j >= p.raw.Cvalid) f[j] = 0.f;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    }
    *reinterpret_cast<uint4*>(p.out.base + p.out.offset(n, yp - p.out.pad_t, xp - p.out.pad_l) + c0) = o;
  }
}

static inline int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

cudaError_t launch_stats_finalize(const FinalizeParams& p, cudaStream_t stream) {
  stats_finalize_kernel<<<p.C, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_norm_apply(const ApplyParams& p, cudaStream_t stream) {
  const long long total = (long long)p.out.N * (p.out.H + p.out.pad_t + p.out.pad_b) *
                          (p.out.W + p.out.pad_l + p.out.pad_r) * (p.out.C / 8);
  norm_apply_kernel<<<grid_for(total, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace v2v
