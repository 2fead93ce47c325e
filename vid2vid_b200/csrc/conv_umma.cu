// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[128 pixels x BN channels] (fp32, TMEM) = sum over (tap, 64-channel block) A[128 x 64] * B[BN x 64]^T
//
// A tiles are boxes of the halo-padded NHWC bf16 activation buffer fetched by TMA (5-D tiled map,
// 128-byte swizzle): a box of TH x TW pixels x 64 channels lands in shared memory as 128 rows of
// 128 bytes -- exactly the K-major SWIZZLE_128B operand layout tcgen05.mma reads.  That box *is*
// the im2col slice for one filter tap; for stride-1 filters on row tiles (TH == 1) one box of
// TW + R - 1 pixels serves R horizontally adjacent taps, each tap's operand being the same smem
// patch shifted by one 128-byte row.  B tiles come from the packed weight matrix
// [Cout][tap * Cp + c] (2-D tiled map).  Reference op being replaced: nn.Conv2d / nn.ConvTranspose2d
// (+ ReflectionPad2d) at models/networks.py:132-183,247-279,571,586,685-706.
//
// Roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, then all four
// warps drain the accumulator (tcgen05.ld 32x32b) and run the fused epilogue:
//   EPI_RAW_STATS : bf16 NHWC raw output + per-tile per-channel (sum, sumsq) partials for the
//                   following Batch/InstanceNorm (deterministic: no atomics)
//   EPI_HEAD_F32  : bias + tanh/sigmoid/scale -> fp32 NCHW planes (the 7x7 image/flow/weight heads)
//   EPI_ACT_BF16  : bias + (leaky)ReLU -> interior of the next layer's padded NHWC buffer
#include "ptx.cuh"
#include "v2v_internal.h"

namespace v2v {

static constexpr int kThreads = 128;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// Sum the 32 per-lane vectors v[0..31] over the 32 lanes; lane l ends up with the total of element l.
// 31 shuffles instead of 160 (recursive halving).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      float send = upper ? v[j] : v[j + s];
      float keep = upper ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.SA * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.SB * p.b_slot_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.SA;
  uint64_t* b_full = a_empty + p.SA;
  uint64_t* b_empty = b_full + p.SB;
  uint64_t* tmem_full = b_empty + p.SB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_img = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - n_img * tiles_per_img;
  const int tile_y = t_in / p.tiles_x, tile_x = t_in - tile_y * p.tiles_x;
  const int y0 = tile_y * p.TH, x0 = tile_x * p.TW;
  const int n0 = blockIdx.y * p.BN;
  const ConvPhase ph = p.phases[blockIdx.z];
  const uint32_t tmem_cols = p.BN < 32 ? 32 : p.BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int a_tx = (p.TW + p.R - 1) * p.TH * 128;
  const int b_tx = p.BN * 128;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int g = ph.group_begin; g < ph.group_end; ++g) {
      const ConvGroup grp = p.groups[g];
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(&a_empty[sa], pa ^ 1);
        mbar_expect_tx(&a_full[sa], a_tx);
        tma_load_5d(sA + (size_t)sa * p.a_slot_bytes, &tmA, &a_full[sa], cb * 64, x0 + grp.dx, y0 + grp.dy,
                    grp.plane, n_img);
        if (++sa == p.SA) { sa = 0; pa ^= 1; }
        for (int r = 0; r < p.R; ++r) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], b_tx);
          tma_load_2d(sB + (size_t)sb * p.b_slot_bytes, &tmB, &b_full[sb], (grp.tap0 + r) * p.Cp + cb * 64, n0);
          if (++sb == p.SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    const uint32_t idesc = make_idesc_bf16(128, p.BN);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0, acc = 0;
    const int npatch = (ph.group_end - ph.group_begin) * p.cblocks;
    for (int pt = 0; pt < npatch; ++pt) {
      mbar_wait(&a_full[sa], pa);
      tcgen05_fence_after();
      const uint32_t a_base = smem_u32(sA + (size_t)sa * p.a_slot_bytes);
      for (int r = 0; r < p.R; ++r) {
        mbar_wait(&b_full[sb], pb);
        tcgen05_fence_after();
        const uint32_t b_base = smem_u32(sB + (size_t)sb * p.b_slot_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // 4 x (K = 16 bf16 = 32 bytes) per 128-byte swizzle row
          const uint64_t adesc = make_sw128_kmajor_desc(a_base + r * 128 + k * 32, p.desc_mode);
          const uint64_t bdesc = make_sw128_kmajor_desc(b_base + k * 32);
          umma_bf16(tmem_base, adesc, bdesc, idesc, acc);
          acc = 1;
        }
        umma_commit(&b_empty[sb]);         // frees the weight slot when these MMAs retire
        if (++sb == p.SB) { sb = 0; pb ^= 1; }
      }
      umma_commit(&a_empty[sa]);           // frees the activation patch
      if (++sa == p.SA) { sa = 0; pa ^= 1; }
    }
    umma_commit(tmem_full);                // accumulator complete
  }
  __syncwarp();

  // -------------------------------------------------------------- epilogue (all 4 warps)
  mbar_wait(tmem_full, 0);
  tcgen05_fence_after();

  const int row = warp * 32 + lane;
  const int ry = row / p.TW, rx = row - ry * p.TW;
  const int gy = y0 + ry, gx = x0 + rx;
  const bool valid = (gy < p.grid_h) && (gx < p.grid_w);
  const int oy = gy * p.oy_mul + ph.oy_add, ox = gx * p.ox_mul + ph.ox_add;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  if (p.epi == EPI_HEAD_F32) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(taddr, r);
    tmem_ld_wait();
    if (valid) {
      const size_t pix = (size_t)oy * p.out_W + ox;
#pragma unroll
      for (int j = 0; j < V2V_MAX_HEAD; ++j) {
        if (j < p.Cout) {
          float v = __uint_as_float(r[j]);
          if (p.bias) v += __ldg(p.bias + j);
          v = apply_act(v, p.head_act[j], p.lrelu_slope) * p.head_scale[j];
          reinterpret_cast<float*>(p.io[p.head_slot[j]])[p.head_off[j] + (size_t)n_img * p.head_bstride[j] + pix] = v;
        }
      }
    }
  } else {
    float* red = reinterpret_cast<float*>(sA);     // rings are idle once tmem_full has fired
    const bool do_stats = (p.epi == EPI_RAW_STATS) && (p.stats != nullptr);
    bf16* dst = nullptr;
    if (valid) {
      if (p.epi == EPI_RAW_STATS)
        dst = reinterpret_cast<bf16*>(p.out) + (((size_t)n_img * p.out_H + oy) * p.out_W + ox) * p.out_C;
      else
        dst = p.out_act.base + p.out_act.offset(n_img, oy, ox);
    }
    for (int c = 0; c < p.BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
      const int col0 = n0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = valid ? __uint_as_float(r[j]) : 0.f;
      if (p.epi == EPI_ACT_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b = (p.bias && col0 + j < p.Cout) ? __ldg(p.bias + col0 + j) : 0.f;
          v[j] = (col0 + j < p.Cout) ? apply_act(v[j] + b, p.act, p.lrelu_slope) : 0.f;
        }
      }
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (col0 + q * 8 < p.out_C) {
            uint4 pk;
            pk.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            pk.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            pk.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            pk.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + col0 + q * 8) = pk;
          }
        }
      }
      if (do_stats) {
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
        const float s = warp_transpose_reduce(v, lane);
        const float q = warp_transpose_reduce(sq, lane);
        red[(warp * 2 + 0) * p.BN + c * 32 + lane] = s;
        red[(warp * 2 + 1) * p.BN + c * 32 + lane] = q;
      }
    }
    if (do_stats) {
      __syncthreads();
      if (threadIdx.x < p.BN && n0 + threadIdx.x < p.stats_C) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          s += red[(w * 2 + 0) * p.BN + threadIdx.x];
          q += red[(w * 2 + 1) * p.BN + threadIdx.x];
        }
        const size_t rowi = (size_t)blockIdx.z * gridDim.x + blockIdx.x;
        p.stats[(rowi * 2 + 0) * p.stats_C + n0 + threadIdx.x] = s;
        p.stats[(rowi * 2 + 1) * p.stats_C + n0 + threadIdx.x] = q;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

cudaError_t launch_conv_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKernelParams& p,
                             cudaStream_t stream) {
  const size_t smem = (size_t)p.SA * p.a_slot_bytes + (size_t)p.SB * p.b_slot_bytes + 1024 /*align*/ +
                      (2 * (p.SA + p.SB) + 2) * sizeof(uint64_t);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  dim3 grid(p.N * p.tiles_x * p.tiles_y, (p.Cout + p.BN - 1) / p.BN, p.num_phases);
  conv_umma_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace v2v
