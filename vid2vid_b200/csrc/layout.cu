// Layout conversion at the module boundary and weight packing.
//   import : caller fp32 NCHW tensor  -> halo-padded NHWC bf16 activation buffer (reflect/zero halo)
//   export : activation buffer interior -> caller fp32 NCHW tensor
//   pack   : torch conv / transposed-conv weights (fp32) -> bf16 GEMM B matrix [Cout][tap*Cp + c]
// The boundary tensors are the ones Vid2VidModelG passes to netG.forward
// (models/vid2vid_model_G.py:225-226); caller pointers are read from a small device-side IO
// table so the captured CUDA graph stays valid when PyTorch hands us new tensors every frame.
#include "ptx.cuh"
#include "v2v_internal.h"

namespace v2v {

__device__ __forceinline__ int reflect_i(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// block (32, 8): tile of 32 padded-x positions x 64 channels at one (n, yp).
__global__ void import_nchw_kernel(ImportParams p) {
  __shared__ float tile[64][33];
  const float* src = reinterpret_cast<const float*>(p.io[p.slot]);
  const ActDesc& o = p.out;
  const int Wpad = o.W + o.pad_l + o.pad_r, Hpad = o.H + o.pad_t + o.pad_b;
  const int xt = blockIdx.x * 32;
  const int yp = blockIdx.y % Hpad, n = blockIdx.y / Hpad;
  const int cblk = blockIdx.z * 64;
  int y = yp - o.pad_t;
  const bool yhalo = (y < 0 || y >= o.H);
  if (p.pad_mode == PAD_REFLECT) y = reflect_i(y, o.H);
  const int xp = xt + threadIdx.x;
  int x = xp - o.pad_l;
  const bool xhalo = (x < 0 || x >= o.W);
  if (p.pad_mode == PAD_REFLECT) x = reflect_i(x, o.W);
  const bool zero_px = (xp >= Wpad) || ((yhalo || xhalo) && p.pad_mode != PAD_REFLECT);
  for (int cc = threadIdx.y; cc < 64; cc += 8) {
    const int c = cblk + cc;
    float v = 0.f;
    if (!zero_px && c < o.Cvalid)
      v = src[(((size_t)n * p.C_src + p.c_off + c) * o.H + y) * o.W + x];
    tile[cc][threadIdx.x] = v;
  }
  __syncthreads();
  // write: thread (tx, ty) -> channel pair 2*tx, pixels ty, ty+8, ...
  for (int px = threadIdx.y; px < 32; px += 8) {
    const int xo = xt + px;
    if (xo >= Wpad) continue;
    if (cblk + 2 * (int)threadIdx.x >= o.C) continue;      // C may be 16 or 32
    const uint32_t pk = pack_bf16x2(tile[2 * threadIdx.x][px], tile[2 * threadIdx.x + 1][px]);
    *reinterpret_cast<uint32_t*>(o.base + o.offset(n, yp - o.pad_t, xo - o.pad_l) + cblk + 2 * threadIdx.x) = pk;
  }
}

// block (32, 8): tile of 32 x positions x 32 channels at one (n, y).
__global__ void export_nchw_kernel(ExportParams p) {
  __shared__ float tile[32][33];
  float* dst = reinterpret_cast<float*>(p.io[p.slot]);
  const ActDesc& a = p.in;
  const int xt = blockIdx.x * 32;
  const int y = blockIdx.y % a.H, n = blockIdx.y / a.H;
  const int cblk = blockIdx.z * 32;
  for (int px = threadIdx.y; px < 32; px += 8) {
    const int x = xt + px, c = cblk + threadIdx.x;
    float v = 0.f;
    if (x < a.W && c < a.Cvalid) v = __bfloat162float(a.base[a.offset(n, y, x) + c]);
    tile[threadIdx.x][px] = v;
  }
  __syncthreads();
  for (int cc = threadIdx.y; cc < 32; cc += 8) {
    const int c = cblk + cc, x = xt + threadIdx.x;
    if (c < a.Cvalid && x < a.W) dst[(((size_t)n * a.Cvalid + c) * a.H + y) * a.W + x] = tile[cc][threadIdx.x];
  }
}

__global__ void pack_weights_kernel(PackParams p) {
  const int K = p.ntaps * p.Cp;
  const long long total = (long long)p.Cout * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K), co = (int)(idx / K);
    const int t = k / p.Cp, c = k - t * p.Cp;
    float v = 0.f;
    if (c < p.Cin) {
      const int ky = p.tap_ky[t], kx = p.tap_kx[t];
      if (p.w2 && co >= p.Cout1) {
        v = p.w2[((((size_t)(co - p.Cout1)) * p.Cin + c) * p.kh + ky) * p.kw + kx];
      } else {
        const int co1 = p.w2 ? p.Cout1 : p.Cout;
        const size_t wi = p.transposed ? ((((size_t)c * co1 + co) * p.kh + ky) * p.kw + kx)
                                       : ((((size_t)co * p.Cin + c) * p.kh + ky) * p.kw + kx);
        v = p.w[wi];
      }
    }
    p.out[idx] = __float2bfloat16_rn(v);
  }
}

cudaError_t launch_import_nchw(const ImportParams& p, cudaStream_t stream) {
  const ActDesc& o = p.out;
  const int Wpad = o.W + o.pad_l + o.pad_r, Hpad = o.H + o.pad_t + o.pad_b;
  dim3 grid((Wpad + 31) / 32, Hpad * o.N, (o.C + 63) / 64), block(32, 8);
  import_nchw_kernel<<<grid, block, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_export_nchw(const ExportParams& p, cudaStream_t stream) {
  const ActDesc& a = p.in;
  dim3 grid((a.W + 31) / 32, a.H * a.N, (a.Cvalid + 31) / 32), block(32, 8);
  export_nchw_kernel<<<grid, block, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pack_weights(const PackParams& p, cudaStream_t stream) {
  const long long total = (long long)p.Cout * p.ntaps * p.Cp;
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  pack_weights_kernel<<<(int)b, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace v2v
