// proj.cuh — caller-side 3x3 projections of the network: HBM-bound direct
// convolutions that also perform the NCHW fp32 <-> bf16 token layout change, so the whole forward
// stays inside this library (no cuDNN, no transposed copies).
//   input_proj : tokens[b, y*W+x, e] = LeakyReLU_0.01( b[e] + sum img[b,ci,y+ky-1,x+kx-1] w[e,ci,ky,kx] )
//                (InputProj.forward, model.py:800-805)
//   output_proj: out[b,co,y,x] = img[b,co,y,x] + b[co] + sum tok[b,(y+ky-1)*W+x+kx-1,ci] w[co,ci,ky,kx]
//                (OutputProj.forward model.py:834-842 + global residual model.py:1305)
#pragma once
#include "lewin_common.cuh"

namespace lw {

__global__ void __launch_bounds__(128) input_proj_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ tok, int B,
                                                         int Cin, int H, int W, int E) {
  __shared__ float sw[64 * 36 + 64];   // [E][Cin*9] + bias
  const int nW = E * Cin * 9;
  for (int i = threadIdx.x; i < nW; i += 128) sw[i] = w[i];
  for (int i = threadIdx.x; i < E; i += 128) sw[64 * 36 + i] = bias[i];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int b = (int)(pix / (H * W)), t = (int)(pix % (H * W)), y = t / W, x = t % W;
  float in[36];
  for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        in[ci * 9 + ky * 3 + kx] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (((size_t)b * Cin + ci) * H + yy) * W + xx) : 0.f;
      }
  const int K = Cin * 9;
  for (int e0 = 0; e0 < E; e0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = sw[64 * 36 + e0 + j];
      const float* wr = sw + (e0 + j) * K;
      for (int k = 0; k < K; ++k) s = fmaf(in[k], wr[k], s);
      acc[j] = s > 0.f ? s : 0.01f * s;
    }
    *reinterpret_cast<uint4*>(tok + (size_t)pix * E + e0) = pack8(acc);
  }
}

__global__ void __launch_bounds__(128) output_proj_kernel(const bf16* __restrict__ tok, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ img,
                                                          float* __restrict__ out, int B, int Cin, int H, int W, int Cout) {
  extern __shared__ float swo[];        // [Cout][9][Cin]
  for (int i = threadIdx.x; i < Cout * 9 * Cin; i += 128) {
    const int co = i / (9 * Cin), rem = i % (9 * Cin), tap = rem / Cin, ci = rem % Cin;
    swo[i] = w[((size_t)co * Cin + ci) * 9 + tap];
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int b = (int)(pix / (H * W)), t = (int)(pix % (H * W)), y = t / W, x = t % W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const bf16* row = tok + (((size_t)b * H + yy) * W + xx) * Cin;
      const int tap = ky * 3 + kx;
      for (int c0 = 0; c0 < Cin; c0 += 8) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(row + c0)), f);
        for (int co = 0; co < Cout; ++co) {
          const float* wr = swo + (co * 9 + tap) * Cin + c0;
          float s = acc[co];
#pragma unroll
          for (int j = 0; j < 8; ++j) s = fmaf(f[j], wr[j], s);
          acc[co] = s;
        }
      }
    }
  for (int co = 0; co < Cout; ++co) {
    const size_t o = (((size_t)b * Cout + co) * H + y) * W + x;
    out[o] = acc[co] + __ldg(bias + co) + (img ? __ldg(img + o) : 0.f);
  }
}

}  // namespace lw
