// train.cuh — training-step periphery, all HBM-bound streaming kernels:
//   charbonnier_kernel : loss = mean( sqrt((x-y)^2 + eps^2) ) and dloss/dx in ONE pass over x, y
//                        (CharbonnierLoss.forward, losses.py:41-52, plus its backward)
//   adamw_kernel       : one launch over the flat fp32 parameter / gradient / moment arenas
//                        (optim.AdamW(lr, betas, eps, weight_decay), train/train_denoise.py:77), fused with the
//                        1/world gradient averaging that follows the NCCL sum all-reduce and with zeroing the
//                        gradient arena for the next step.
// Algorithmic bytes: charbonnier 12 B/element (2 reads + 1 write); adamw 28 B/element (4 reads + 3 writes), +4 when
// the gradient is zeroed in the same pass.  Grids are a multiple of the SM count (grid-stride loops).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lw {

constexpr int kTrainThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[blockIdx.x] = sum over this block's elements of sqrt(d^2 + eps^2); grad = inv_n * d / sqrt(d^2 + eps^2)
__global__ void __launch_bounds__(kTrainThreads) charbonnier_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                    float* __restrict__ grad, float* __restrict__ partial,
                                                                    long long n, float eps2, float inv_n) {
  __shared__ float red[kTrainThreads / 32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * kTrainThreads;
  for (long long i = (long long)blockIdx.x * kTrainThreads + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(y) + i);
    float4 g;
    float d, r;
    d = a.x - b.x; r = sqrtf(fmaf(d, d, eps2)); acc += r; g.x = inv_n * d / r;
    d = a.y - b.y; r = sqrtf(fmaf(d, d, eps2)); acc += r; g.y = inv_n * d / r;
    d = a.z - b.z; r = sqrtf(fmaf(d, d, eps2)); acc += r; g.z = inv_n * d / r;
    d = a.w - b.w; r = sqrtf(fmaf(d, d, eps2)); acc += r; g.w = inv_n * d / r;
    if (grad) reinterpret_cast<float4*>(grad)[i] = g;
  }
  // tail (n not a multiple of 4): handled by block 0
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kTrainThreads) {
      const float d = x[i] - y[i];
      const float r = sqrtf(fmaf(d, d, eps2));
      acc += r;
      if (grad) grad[i] = inv_n * d / r;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kTrainThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// loss[0] = inv_n * sum(partial[0..m)) — one block, fixed summation order (deterministic run to run)
__global__ void __launch_bounds__(kTrainThreads) charbonnier_finish_kernel(const float* __restrict__ partial, int m, float inv_n,
                                                                           float* __restrict__ loss) {
  __shared__ float red[kTrainThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < m; i += kTrainThreads) acc += partial[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kTrainThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v * inv_n;
  }
}

struct AdamWConsts {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_corr1;        // 1 - beta1^t
  float inv_sqrt_bc2;      // 1 / sqrt(1 - beta2^t)
  float grad_scale;        // multiplies the stored gradient first (1/world after a sum all-reduce)
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWConsts& c) {
  g *= c.grad_scale;
  p *= 1.f - c.lr * c.weight_decay;                       // decoupled weight decay (torch.optim.AdamW)
  m = fmaf(c.beta1, m, (1.f - c.beta1) * g);
  v = fmaf(c.beta2, v, (1.f - c.beta2) * g * g);
  const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
  p -= (c.lr / c.bias_corr1) * (m / denom);
}

__global__ void __launch_bounds__(kTrainThreads) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                              float* __restrict__ v, long long n, AdamWConsts c, int zero_grad) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * kTrainThreads;
  for (long long i = (long long)blockIdx.x * kTrainThreads + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_one(pp.x, gg.x, mm.x, vv.x, c);
    adamw_one(pp.y, gg.y, mm.y, vv.y, c);
    adamw_one(pp.z, gg.z, mm.z, vv.z, c);
    adamw_one(pp.w, gg.w, mm.w, vv.w, c);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kTrainThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, c);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (zero_grad) g[i] = 0.f;
    }
  }
}

}  // namespace lw
