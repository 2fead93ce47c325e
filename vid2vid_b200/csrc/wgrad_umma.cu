// Weight gradient of a convolution on tcgen05 tensor cores (sm_100a): a GEMM whose K dimension is the PIXEL index,
//
//   G[tap][m][n] = sum over grid pixels (img, y, x) of OUT[img][y][x][m] * IN[img][(y, x) @ tap][n]
//
// where OUT is the gradient of the conv output and IN the conv input (train.py:83-90 back-propagates through every
// nn.Conv2d / nn.ConvTranspose2d of models/networks.py; for a transposed conv the roles of "input" and "output gradient"
// swap, see plan.cu).  Both operands are read where the forward pass left them: halo-padded NHWC activation buffers, a
// pixel being [hi C | lo C] bf16.  A TMA box of KP pixels x 64 channels lands in shared memory as KP rows of 128 bytes with
// the 128-byte swizzle -- the canonical *MN-major* SWIZZLE_128B operand layout (64 contiguous M / N elements per K index,
// 8-K groups 1024 bytes apart, 64-element blocks one box apart), so no transposed copy of any tensor is made: the same
// bytes serve the forward conv K-major and this kernel MN-major.
//
// One CTA per work unit (tap, M tile of 128 output-gradient channels, N tile of 64 / 128 input channels, K split); the K
// loop walks row segments of KP pixels through a ring of shared-memory stages:
//   warp 0     TMA producer (one elected lane)
//   warp 1     tcgen05.mma issuer (one elected lane) + TMEM owner; precise plans accumulate OUT_hi*IN_hi + OUT_lo*IN_hi +
//              OUT_hi*IN_lo like the forward kernel
//   warps 2-5  epilogue: tcgen05.ld the 128 x BN fp32 accumulator and add it to the staging buffer G (vector atomics: the K
//              splits of one (tap, tile) meet there); unstage_wgrad_kernel then adds G into the caller's gradient tensor in
//              the parameter's own layout [M][N][kh][kw].
#include "ptx.cuh"
#include "v2v_internal.h"
#include "backward.h"

namespace v2v {

static constexpr int kWgThreads = 192;

__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, int lbo_bytes, int sbo_bytes) {
  // MN-major SWIZZLE_128B: LBO = distance between 64-element M/N blocks, SBO = distance between groups of 8 K indices
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmIn,
                  const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int box = p.KP * 128;                                   // one TMA box: KP pixels x 64 channels
  const int nh = p.split ? 2 : 1;
  const int out_bytes = nh * 2 * box, in_bytes = nh * p.Nblocks * box;   // OUT always holds two 64-channel blocks (M = 128)
  const int stage_bytes = out_bytes + in_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.Nblocks * 64;
  const uint32_t tmem_cols = BN < 32 ? 32 : BN;

  // unit -> (K split, tap, M tile, N tile), tiles fastest: the CTAs running together walk the same pixel range, so every
  // operand row is fetched from HBM once and then served by L2 to the other (tap, tile) units
  int u = blockIdx.x;
  const int nt = u % p.n_tiles; u /= p.n_tiles;
  const int mt = u % p.m_tiles; u /= p.m_tiles;
  const int tap = u % p.ntaps;
  const int ks = u / p.ntaps;
  const int c_begin = ks * p.chunks_per_unit;
  const int c_end = min(p.chunks_total, c_begin + p.chunks_per_unit);
  const int nchunks = c_end - c_begin;                           // >= 1 by construction (launcher)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmIn);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
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

  if (warp == 0) {
    if (elect_one_sync()) {
      const WgradTap t = p.taps[tap];
      int xs = c_begin % p.xsegs;
      int row = c_begin / p.xsegs;                               // img * gh + y
      int y = row % p.gh, img = row / p.gh;
      int s = 0; uint32_t par = 0;
      for (int c = 0; c < nchunks; ++c) {
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_wait(&empty[s], par ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        for (int hf = 0; hf < nh; ++hf)
          for (int mb = 0; mb < 2; ++mb) {
            // a 64-channel OUT tensor fills both halves of the M = 128 tile with the same block (rows 64.. are not stored)
            const int blk = p.Mblocks == 2 ? mt * 2 + mb : mt;
            tma_load_5d(st + (size_t)(hf * 2 + mb) * box, &tmOut, &full[s], hf * p.out_C + blk * 64, xs * p.KP + p.out_padl,
                        y + p.out_padt, 0, img);
          }
        for (int hf = 0; hf < nh; ++hf)
          for (int nb = 0; nb < p.Nblocks; ++nb)
            tma_load_5d(st + out_bytes + (size_t)(hf * p.Nblocks + nb) * box, &tmIn, &full[s], hf * p.in_C + (nt * p.Nblocks + nb) * 64,
                        xs * p.KP + t.dx, y + t.dy, t.plane, img);
        if (++s == p.stages) { s = 0; par ^= 1; }
        if (++xs == p.xsegs) { xs = 0; if (++y == p.gh) { y = 0; ++img; } }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(128, BN) | (1u << 15) | (1u << 16);      // A and B MN-major
      const int ps_step = p.split ? 1 : 3;
      int s = 0; uint32_t par = 0;
      uint32_t first = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[s], par);
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(smem + (size_t)s * stage_bytes), b0 = a0 + out_bytes;
        for (int ps = 0; ps < 3; ps += ps_step) {
          const uint32_t a = a0 + (ps == 1 ? 2 * box : 0), b = b0 + (ps == 2 ? p.Nblocks * box : 0);
          for (int k = 0; k < p.kmma; ++k) {                      // 16 pixels = two 8-K groups = 2048 bytes per MMA
            umma_bf16(tmem_base, make_mnmajor_desc(a + k * 2048, p.lbo_bytes, p.sbo_bytes), make_mnmajor_desc(b + k * 2048, p.lbo_bytes, p.sbo_bytes), idesc, first);
            first = 1u;
          }
        }
        umma_commit(&empty[s]);
        if (++s == p.stages) { s = 0; par ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;                                       // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const int m = mt * (p.Mblocks * 64) + row;
    const bool valid = row < (p.Mblocks == 2 ? 128 : 64) && m < p.Mp;
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const int n0 = nt * BN;
    float* dst = p.stage + ((size_t)tap * p.Mp + (valid ? m : 0)) * p.Np + n0;
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + c * 32 + j * 4 < p.Np)
            atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + j * 4),
                      make_float4(__uint_as_float(r[j * 4]), __uint_as_float(r[j * 4 + 1]), __uint_as_float(r[j * 4 + 2]),
                                  __uint_as_float(r[j * 4 + 3])));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// G[tap][m][n] -> parameter gradient [M][Nv][taps] (+=); rows >= M1 go to the second weight set of a fused unit
__global__ void __launch_bounds__(256) unstage_wgrad_kernel(const float* __restrict__ stage, int Mp, int Np, int M, int M1, int Nv,
                                                            int taps, float* __restrict__ dw, float* __restrict__ dw2) {
  const long long total = (long long)M * Nv * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % taps);
    const long long r = idx / taps;
    const int n = (int)(r % Nv), m = (int)(r / Nv);
    const float v = stage[((size_t)t * Mp + m) * Np + n];
    if (m < M1) { if (dw) dw[idx] += v; }
    else if (dw2) dw2[((size_t)(m - M1) * Nv + n) * taps + t] += v;
  }
}

size_t wgrad_stage_bytes(const WgradParams& p) { return (size_t)p.ntaps * p.Mp * p.Np * sizeof(float); }

cudaError_t launch_wgrad_umma(const CUtensorMap& tmOut, const CUtensorMap& tmIn, const WgradParams& p, int M, int M1, int Nv,
                              float* dw, float* dw2, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p.stage, 0, wgrad_stage_bytes(p), s);
  if (e != cudaSuccess) return e;
  const int nh = p.split ? 2 : 1;
  const size_t stage_bytes = (size_t)nh * (2 + p.Nblocks) * p.KP * 128;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (2 * p.stages + 2) * sizeof(uint64_t);
  static size_t configured = 0;
  if (smem > configured) {
    e = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int units = p.ntaps * p.m_tiles * p.n_tiles * p.ksplit;
  wgrad_umma_kernel<<<units, kWgThreads, smem, s>>>(tmOut, tmIn, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long total = (long long)M * Nv * p.ntaps;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  unstage_wgrad_kernel<<<blocks, 256, 0, s>>>(p.stage, p.Mp, p.Np, M, M1, Nv, p.ntaps, dw, dw2);
  return cudaGetLastError();
}

// dX[n][y][x][c] += sum of the padded-extent gradient over the padded positions that mirror onto (y, x) (reflect halo) or
// the interior position alone (zero halo / no halo): gather form, deterministic.  src: dense NHWC fp32 (N, PH, PW, stride Cs)
// whose pixel (pad, pad) is input pixel (0, 0); rows / columns beyond H + 2 pad are ignored (cropped).
__global__ void __launch_bounds__(256) fold_add_kernel(const float* __restrict__ src, int Cs, int PH, int PW, float* __restrict__ dx, int N,
                                                       int H, int W, int C, int pad, int reflect) {
  const size_t total = (size_t)N * H * W * C;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    size_t t = idx / C;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + pad; xs[nx++] = x + pad;
    if (reflect) {
      if (y >= 1 && y <= pad) ys[ny++] = pad - y;
      if (y <= H - 2 && y >= H - 1 - pad) ys[ny++] = pad + 2 * (H - 1) - y;
      if (x >= 1 && x <= pad) xs[nx++] = pad - x;
      if (x <= W - 2 && x >= W - 1 - pad) xs[nx++] = pad + 2 * (W - 1) - x;
    }
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) s += src[(((size_t)n * PH + ys[a]) * PW + xs[b]) * Cs + c];
    dx[idx] += s;
  }
}

cudaError_t launch_fold_add(const float* src, int Cs, int PH, int PW, float* dx, int N, int H, int W, int C, int pad, int reflect,
                            cudaStream_t s) {
  const size_t total = (size_t)N * H * W * C;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  fold_add_kernel<<<blocks, 256, 0, s>>>(src, Cs, PH, PW, dx, N, H, W, C, pad, reflect);
  return cudaGetLastError();
}

}  // namespace v2v
