// proj.cuh — caller-side 3x3 projections of the network: HBM-bound direct
// convolutions that also perform the NCHW fp32 <-> bf16 token layout change, so the whole forward
// stays inside this library (no cuDNN, no transposed copies).
//   input_proj : tokens[b, y*W+x, e] = LeakyReLU_0.01( b[e] + sum img[b,ci,y+ky-1,x+kx-1] w[e,ci,ky,kx] )
//                (InputProj.forward, model.py:800-805)
//   output_proj: out[b,co,y,x] = img[b,co,y,x] + b[co] + sum tok[b,(y+ky-1)*W+x+kx-1,ci] w[co,ci,ky,kx]
//                (OutputProj.forward model.py:834-842 + global residual model.py:1305)
// Each thread owns TWO horizontally adjacent pixels so every broadcast weight read from shared memory
// feeds two pixels; arithmetic is packed FFMA2 where the data comes in pairs.
#pragma once
#include "lewin_common.cuh"

namespace lw {

// weights in smem as [k = ci*9+ky*3+kx][E] fp32 (k-major so one LDS.128 = 4 output channels of one tap)
__global__ void __launch_bounds__(128) input_proj_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ tok, int B,
                                                         int Cin, int H, int W, int E) {
  __shared__ __align__(16) float sw[36 * 64 + 64];
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * E; i += 128) {
    const int k = i / E, e = i % E;
    sw[i] = w[(size_t)e * K + k];
  }
  for (int i = threadIdx.x; i < E; i += 128) sw[36 * 64 + i] = bias[i];
  __syncthreads();
  const int Wp = W >> 1;                                   // pixel pairs per row (W even)
  const long long pp = (long long)blockIdx.x * 128 + threadIdx.x;
  if (pp >= (long long)B * H * Wp) return;
  const int b = (int)(pp / (H * Wp)), t = (int)(pp % (H * Wp)), y = t / Wp, x0 = (t % Wp) * 2;
  // 3 rows x 4 columns of input per channel cover both pixels
  float in[4][12];
#pragma unroll 1
  for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int yy = y + ky - 1, xx = x0 + c - 1;
        in[ci][ky * 4 + c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (((size_t)b * Cin + ci) * H + yy) * W + xx) : 0.f;
      }
  const uint32_t sw_s = smem_u32(sw);
  bf16* o0 = tok + ((size_t)(b * H + y) * W + x0) * E;
#pragma unroll 1
  for (int e0 = 0; e0 < E; e0 += 8) {
    float4 bA = lds128f(sw_s + (36 * 64 + e0) * 4), bB = lds128f(sw_s + (36 * 64 + e0 + 4) * 4);
    float a0[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
    float a1[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
    for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int k = ci * 9 + ky * 3 + kx;
          const float4 wA = lds128f(sw_s + (k * E + e0) * 4), wB = lds128f(sw_s + (k * E + e0 + 4) * 4);
          const float p0 = in[ci][ky * 4 + kx], p1 = in[ci][ky * 4 + kx + 1];
          a0[0] = fmaf(p0, wA.x, a0[0]); a0[1] = fmaf(p0, wA.y, a0[1]); a0[2] = fmaf(p0, wA.z, a0[2]); a0[3] = fmaf(p0, wA.w, a0[3]);
          a0[4] = fmaf(p0, wB.x, a0[4]); a0[5] = fmaf(p0, wB.y, a0[5]); a0[6] = fmaf(p0, wB.z, a0[6]); a0[7] = fmaf(p0, wB.w, a0[7]);
          a1[0] = fmaf(p1, wA.x, a1[0]); a1[1] = fmaf(p1, wA.y, a1[1]); a1[2] = fmaf(p1, wA.z, a1[2]); a1[3] = fmaf(p1, wA.w, a1[3]);
          a1[4] = fmaf(p1, wB.x, a1[4]); a1[5] = fmaf(p1, wB.y, a1[5]); a1[6] = fmaf(p1, wB.z, a1[6]); a1[7] = fmaf(p1, wB.w, a1[7]);
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = a0[j] > 0.f ? a0[j] : 0.01f * a0[j];
      a1[j] = a1[j] > 0.f ? a1[j] : 0.01f * a1[j];
    }
    *reinterpret_cast<uint4*>(o0 + e0) = pack8(a0);
    *reinterpret_cast<uint4*>(o0 + E + e0) = pack8(a1);
  }
}

// weights in smem as [tap][ci][4] fp32 (co padded to 4): one LDS.128 = the <=4 output weights of one (tap, ci)
__global__ void __launch_bounds__(128) output_proj_kernel(const bf16* __restrict__ tok, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ img,
                                                          float* __restrict__ out, int B, int Cin, int H, int W, int Cout) {
  extern __shared__ __align__(16) float swo[];        // [9][Cin][4]
  for (int i = threadIdx.x; i < 9 * Cin * 4; i += 128) {
    const int tap = i / (Cin * 4), ci = (i / 4) % Cin, co = i & 3;
    swo[i] = (co < Cout) ? w[((size_t)co * Cin + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int Wp = W >> 1;
  const long long pp = (long long)blockIdx.x * 128 + threadIdx.x;
  if (pp >= (long long)B * H * Wp) return;
  const int b = (int)(pp / (H * Wp)), t = (int)(pp % (H * Wp)), y = t / Wp, x0 = (t % Wp) * 2;
  const uint32_t sw_s = smem_u32(swo);
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  // For every kernel row ky the four input columns x0-1 .. x0+2 are held in registers (8 channels at a time); tap
  // (ky,kx) multiplies column kx for pixel 0 and column kx+1 for pixel 1, so each broadcast weight read feeds both.
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    const bf16* rowp = tok + (((size_t)b * H + yy) * W) * Cin;
#pragma unroll 1
    for (int c0 = 0; c0 < Cin; c0 += 8) {
      float f[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = x0 + c - 1;
        if (xx >= 0 && xx < W) unpack8(__ldg(reinterpret_cast<const uint4*>(rowp + (size_t)xx * Cin + c0)), f[c]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[c][j] = 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t wb = sw_s + (((ky * 3 + kx) * Cin + c0) * 4) * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w4 = lds128f(wb + j * 16);
          a0[0] = fmaf(f[kx][j], w4.x, a0[0]); a0[1] = fmaf(f[kx][j], w4.y, a0[1]); a0[2] = fmaf(f[kx][j], w4.z, a0[2]);
          a1[0] = fmaf(f[kx + 1][j], w4.x, a1[0]); a1[1] = fmaf(f[kx + 1][j], w4.y, a1[1]); a1[2] = fmaf(f[kx + 1][j], w4.z, a1[2]);
          a0[3] = fmaf(f[kx][j], w4.w, a0[3]); a1[3] = fmaf(f[kx + 1][j], w4.w, a1[3]);
        }
      }
    }
  }
  for (int co = 0; co < Cout; ++co) {
    const size_t o = (((size_t)b * Cout + co) * H + y) * W + x0;
    const float bb = __ldg(bias + co);
    float2 r = make_float2(a0[co] + bb, a1[co] + bb);
    if (img != nullptr) {
      const float2 im = __ldg(reinterpret_cast<const float2*>(img + o));
      r.x += im.x; r.y += im.y;
    }
    *reinterpret_cast<float2*>(out + o) = r;
  }
}

}  // namespace lw
