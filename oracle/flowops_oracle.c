/* TEST INFRASTRUCTURE -- not product code.  Plain-C CPU restatement of the three native FlowNet2
 * operators of the reference, following the CUDA sources line by line (they have no CPU path and
 * cannot be built against the installed torch):
 *   correlation_forward  models/flownet2_pytorch/networks/correlation_package/correlation_cuda_kernel.cu:46-147
 *                        (+ output shape arithmetic correlation_cuda.cc:25-38)
 *   resample2d_forward   .../resample2d_package/resample2d_kernel.cu:15-64
 *   channelnorm_forward  .../channelnorm_package/channelnorm_kernel.cu:18-60
 * Parity pin: the reference holds no tests or vectors for these ops and the CUDA sources cannot be built or run in the build
 * container ("parity unpinned" by reference tests).  tests/test_flowops_oracle.py pins this file (i) against independent
 * PyTorch formulations (grid_sample(align_corners=True, border), shifted dot products, torch.norm) and (ii) against
 * known-answer vectors computed by hand from the CUDA sources (floor / clamp rules, displacement-to-channel order, the
 * 1 / (C k k) normalisation and the per-lane + shuffle-tree fp32 summation order).
 * Build: see oracle/Makefile -> oracle/_build/libflowops_oracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* correlation_cuda.cc:25-38 */
void corr_out_shape(int H, int W, int pad, int k, int max_disp, int s1, int s2, int* outC, int* outH, int* outW) {
  int kernel_radius = (k - 1) / 2;
  int border_radius = kernel_radius + max_disp;
  int pH = H + 2 * pad, pW = W + 2 * pad;
  int d = (max_disp / s2) * 2 + 1;
  *outC = d * d;
  *outH = (int)ceilf((float)(pH - 2 * border_radius) / (float)s1);
  *outW = (int)ceilf((float)(pW - 2 * border_radius) / (float)s1);
}

/* in1,in2: [N][C][H][W]; out: [N][outC][outH][outW].  The per-lane partial sums (ch = lane, lane+32, ..)
 * followed by the shuffle-down tree of warpReduceSum (:16-21,121-141) are reproduced so the fp32
 * summation order matches the reference kernel launched with 32 threads per block. */
int correlation_forward(const float* in1, const float* in2, float* out, int N, int C, int H, int W, int pad, int k,
                        int max_disp, int s1, int s2) {
  int outC, outH, outW;
  corr_out_shape(H, W, pad, k, max_disp, s1, s2, &outC, &outH, &outW);
  const int pH = H + 2 * pad, pW = W + 2 * pad;
  /* channels_first (:46-70): zero-padded NHWC copies */
  float* r1 = (float*)calloc((size_t)N * pH * pW * C, sizeof(float));
  float* r2 = (float*)calloc((size_t)N * pH * pW * C, sizeof(float));
  if (!r1 || !r2) return 0;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          size_t src = (((size_t)n * C + c) * H + y) * W + x;
          size_t dst = (((size_t)n * pH + (y + pad)) * pW + (x + pad)) * C + c;
          r1[dst] = in1[src];
          r2[dst] = in2[src];
        }
  const int krad = (k - 1) / 2;
  const int drad = max_disp / s2;
  const int dsize = 2 * drad + 1;
  const int nelems = k * k * C;
  for (int n = 0; n < N; ++n)
    for (int by = 0; by < outH; ++by)
      for (int bx = 0; bx < outW; ++bx) {
        const int y1 = by * s1 + max_disp, x1 = bx * s1 + max_disp;
        for (int tj = -drad; tj <= drad; ++tj)
          for (int ti = -drad; ti <= drad; ++ti) {
            const int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
            float lane_acc[32];
            for (int l = 0; l < 32; ++l) {
              float acc0 = 0.0f;
              for (int j = -krad; j <= krad; ++j)
                for (int i = -krad; i <= krad; ++i)
                  for (int ch = l; ch < C; ch += 32) {
                    size_t i1 = (((size_t)n * pH + (y1 + j)) * pW + (x1 + i)) * C + ch;
                    size_t i2 = (((size_t)n * pH + (y2 + j)) * pW + (x2 + i)) * C + ch;
                    acc0 += r1[i1] * r2[i2];
                  }
              lane_acc[l] = acc0;
            }
            for (int off = 16; off > 0; off /= 2)        /* __shfl_down_sync tree */
              for (int l = 0; l < off; ++l) lane_acc[l] += lane_acc[l + off];
            const int tc = (tj + drad) * dsize + (ti + drad);
            out[(((size_t)n * outC + tc) * outH + by) * outW + bx] = lane_acc[0] / nelems;
          }
      }
  free(r1);
  free(r2);
  return 1;
}

/* in1: [N][C][inH][inW], flow: [N][2][H][W], out: [N][C][H][W]; kernel_size as the reference passes it (1). */
int resample2d_forward(const float* in1, const float* flow, float* out, int N, int C, int H, int W, int inH, int inW,
                       int kernel_size) {
  for (int b = 0; b < N; ++b)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          float val = 0.0f;
          float dx = flow[(((size_t)b * 2 + 0) * H + y) * W + x];
          float dy = flow[(((size_t)b * 2 + 1) * H + y) * W + x];
          float xf = (float)x + dx;
          float yf = (float)y + dy;
          float alpha = xf - floorf(xf);
          float beta = yf - floorf(yf);
          int xL = (int)floorf(xf); if (xL > W - 1) xL = W - 1; if (xL < 0) xL = 0;
          int xR = (int)(floorf(xf) + 1); if (xR > W - 1) xR = W - 1; if (xR < 0) xR = 0;
          int yT = (int)floorf(yf); if (yT > H - 1) yT = H - 1; if (yT < 0) yT = 0;
          int yB = (int)(floorf(yf) + 1); if (yB > H - 1) yB = H - 1; if (yB < 0) yB = 0;
          const float* pl = in1 + ((size_t)b * C + c) * inH * inW;
          for (int fy = 0; fy < kernel_size; ++fy)
            for (int fx = 0; fx < kernel_size; ++fx) {
              val += (float)((1. - alpha) * (1. - beta) * pl[(size_t)(yT + fy) * inW + xL + fx]);
              val += (float)((alpha) * (1. - beta) * pl[(size_t)(yT + fy) * inW + xR + fx]);
              val += (float)((1. - alpha) * (beta)*pl[(size_t)(yB + fy) * inW + xL + fx]);
              val += (float)((alpha) * (beta)*pl[(size_t)(yB + fy) * inW + xR + fx]);
            }
          out[(((size_t)b * C + c) * H + y) * W + x] = val;
        }
  return 1;
}

/* in: [N][C][H][W] -> out: [N][1][H][W] */
int channelnorm_forward(const float* in, float* out, int N, int C, int H, int W, int norm_deg) {
  (void)norm_deg; /* the reference kernel ignores it too: always L2 (:53-59) */
  for (int b = 0; b < N; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float result = 0.0f;
        for (int c = 0; c < C; ++c) {
          float v = in[(((size_t)b * C + c) * H + y) * W + x];
          result += v * v;
        }
        out[((size_t)b * H + y) * W + x] = sqrtf(result);
      }
  return 1;
}
