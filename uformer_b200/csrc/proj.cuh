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
#include <cuda.h>
#include "lewin_common.cuh"
#include "leff_fused.cuh"   // tma_load_4d / tma_prefetch_desc

namespace lw {

// weights in smem as [k = ci*9+ky*3+kx][E] fp32 (k-major so one LDS.128 = 4 output channels of one tap)
__global__ void __launch_bounds__(128) input_proj_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ tok, int B,
                                                         int Cin, int H, int W, int E) {
  __shared__ __align__(16) float sw[36 * 64 + 64];
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
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

// ----------------------------------------------------------------------------------------------
// output_proj on the tensor core (Cin in {32, 64}, Cout <= 3): the 3x3 convolution is re-associated as "GEMM first, taps
// after".  Z[q, tap*Cout+co] = sum_ci tok[q, ci] w[co, ci, tap] is ONE small GEMM per halo'd token tile (K = Cin, N = 64), and
//   out[p, co] = img[p, co] + b[co] + sum_tap Z[p + tap, tap*Cout + co]
// is 9*Cout shared-memory reads per pixel.  PERSISTENT CTAs (4 per SM, 53 KB each) walk 8 x 16 pixel tiles:
//   TMA    the 10 x 18 halo'd block of tokens lands with one cp.async.bulk.tensor box (4-D map (C, W, H, B); out-of-image
//          coordinates are zero-filled = the conv's zero padding) in the K-major swizzled layout of the A operand; the box of
//          tile i+1 is requested as soon as the GEMM of tile i has read the buffer and lands under the tap phase.
//   GEMM   one M=128 + one M=64 tcgen05.mma chain over the 192 (>= 180) halo rows.  The weights keep fp32 accuracy with bf16
//          operands: B rows [0, 32) carry bf16(w), rows [32, 64) carry bf16(w - bf16(w)); the epilogue adds the two halves.
//   taps   Z -> shared memory ([180][29] fp32, odd pitch: conflict-free), one thread per output pixel sums its 9 taps, adds
//          bias and the global residual and parks the result in an [Cout][8][16] fp32 tile.
//   store  the NCHW output leaves through the TMA engine: one cp.async.bulk.tensor.3d store of a 16 x 8 box per channel plane
//          (tensor map (W, H, B*Cout); boxes that overhang the image are clipped by the hardware).  Falls back to per-thread
//          stores when the row pitch is not a multiple of 16 bytes (use_tma_store = 0).
// One HBM read of the token map (neighbouring halos hit L2) instead of 9 x 4 register-blocked re-reads + 1728 FMAs per pixel.
// ----------------------------------------------------------------------------------------------
template <int CIN>
struct OutProjCfg {
  static constexpr int SW = CIN * 2;                        // row bytes = swizzle span (64 / 128)
  static constexpr int KS = CIN / 16;
  static constexpr int A_BYTES = 192 * SW;
  static constexpr int W_BYTES = 64 * SW;
  static constexpr int ZP = 29;                             // Z pitch (floats)
  static constexpr int S_A = 0;                             // ONE input buffer: four CTAs share an SM and overlap each other's phases
  static constexpr int S_W = A_BYTES;
  static constexpr int S_Z = S_W + W_BYTES;
  static constexpr int S_O = (S_Z + 180 * ZP * 4 + 127) / 128 * 128;   // [3][8][16] fp32 output planes of the tile: the source of the TMA store
  static constexpr int S_BAR = S_O + 3 * 512;
  static_assert(4 * (S_BAR + 64 + 1024 + 1024) <= 233472, "four CTAs per SM");
  static constexpr int SMEM_BYTES = S_BAR + 64 + 1024;
  static_assert(A_BYTES % 1024 == 0 && S_W % 1024 == 0, "operand alignment");
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int CIN>
__global__ void __launch_bounds__(128, 4) output_proj_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap omap,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             const float* __restrict__ img, float* __restrict__ out, int B, int H, int W, int Cout,
                                                             int tiles_x, int tiles_y, int n_tiles, int use_tma_store) {
  using Cfg = OutProjCfg<CIN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::S_BAR);       // [0] x_full, [1] mma done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::S_BAR + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t sA = smem_u32(smem + Cfg::S_A), sW = smem_u32(smem + Cfg::S_W), sZ = smem_u32(smem + Cfg::S_Z), sO = smem_u32(smem + Cfg::S_O);
  const int ntap = 9 * Cout;                                 // <= 27 columns of Z

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&xmap);
    if (use_tma_store) tma_prefetch_desc(&omap);
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  // B operand image, built once per CTA from the fp32 conv weight: row n < 32: bf16(w) of (tap, co) = (n / Cout, n % Cout);
  // row 32 + n: the bf16 remainder; unused rows zero.
  for (int i = tid; i < 64 * CIN; i += 128) {
    const int n = i / CIN, k = i % CIN, nn = n & 31;
    float v = 0.f;
    if (nn < ntap) {
      const float wf = __ldg(w + ((size_t)(nn % Cout) * CIN + k) * 9 + nn / Cout);
      const float hi = __bfloat162float(__float2bfloat16_rn(wf));
      v = (n < 32) ? hi : wf - hi;
    }
    const __nv_bfloat16 hv = __float2bfloat16_rn(v);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(sW + swz<Cfg::SW>(n, 2 * k)), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  auto load_tile = [&](int it) {                             // thread 0
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, 180 * Cfg::SW);
    tma_load_4d(sA, &xmap, 0, tx * 16 - 1, ty * 8 - 1, b, bar);
  };
  if (tid == 0 && n_my > 0) load_tile(0);
  constexpr uint32_t idesc_a = make_idesc_bf16(128, 64), idesc_b = make_idesc_bf16(64, 64);
  const int py = tid >> 4, px = tid & 15;
  float bco[3];
#pragma unroll
  for (int co = 0; co < 3; ++co) bco[co] = (co < Cout) ? __ldg(bias + co) : 0.f;

  for (int it = 0; it < n_my; ++it) {
    const uint32_t ph = it & 1;
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int Y = ty * 8 + py, X = tx * 16 + px;
    const bool inside = Y < H && X < W;
    // this pixel's residual + bias: requested now, consumed after the taps (its latency hides under the GEMM and the Z pass)
    float acc[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      acc[co] = bco[co];
      if (co < Cout && inside && img != nullptr) acc[co] += __ldg(img + (((size_t)b * Cout + co) * H + Y) * W + X);
    }
    if (warp == 0) {
      mbar_wait(smem_u32(&bars[0]), ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < Cfg::KS; ++ks) {
          const uint64_t bd = kmajor_desc<Cfg::SW>(sW + ks * 32);
          umma_ss(tb, kmajor_desc<Cfg::SW>(sA + ks * 32), bd, idesc_a, ks != 0);
          umma_ss(tb + 64, kmajor_desc<Cfg::SW>(sA + 128 * Cfg::SW + ks * 32), bd, idesc_b, ks != 0);
        }
        umma_commit(smem_u32(&bars[1]));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bars[1]), ph);
    tc_fence_after();
    if (tid == 0 && it + 1 < n_my) load_tile(it + 1);       // the GEMM has read the buffer: refill it under the rest of this tile
    // ---- Z = hi + lo halves of the accumulators -> shared memory (thread = TMEM lane = halo row) ----
    {
      uint32_t hi[32], lo[32];
      const uint32_t tl = (uint32_t)(32 * warp) << 16;
      tmem_ld32(tb + tl, hi);
      tmem_ld32(tb + tl + 32, lo);
      tmem_wait_ld();
      const uint32_t zr = sZ + (32 * warp + lane) * (Cfg::ZP * 4);
#pragma unroll
      for (int n = 0; n < 27; ++n) sts32f(zr + n * 4, __uint_as_float(hi[n]) + __uint_as_float(lo[n]));
      // M = 64 part: its rows 16q .. 16q+15 sit in lanes 32q .. 32q+15 (halo rows 128 + 16q + lane)
      tmem_ld32(tb + tl + 64, hi);
      tmem_ld32(tb + tl + 96, lo);
      tmem_wait_ld();
      if (lane < 16 && 128 + 16 * warp + lane < 180) {       // (Z holds the 180 halo rows only)
        const uint32_t zr2 = sZ + (128 + 16 * warp + lane) * (Cfg::ZP * 4);
#pragma unroll
        for (int n = 0; n < 27; ++n) sts32f(zr2 + n * 4, __uint_as_float(hi[n]) + __uint_as_float(lo[n]));
      }
    }
    if (tid == 0 && use_tma_store) bulk_wait_read0();       // the previous tile's TMA stores (issued a GEMM ago) have read the output planes
    tc_fence_before();
    __syncthreads();
    // ---- 9 taps per pixel (+ bias + global residual, already in acc), NCHW fp32 ----
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t zr = sZ + ((py + ky) * 18 + px + kx) * (Cfg::ZP * 4) + (ky * 3 + kx) * Cout * 4;
#pragma unroll
        for (int co = 0; co < 3; ++co)
          if (co < Cout) acc[co] += lds32f(zr + co * 4);
      }
    if (use_tma_store) {
#pragma unroll
      for (int co = 0; co < 3; ++co)
        if (co < Cout) sts32f(sO + co * 512 + tid * 4, acc[co]);
      fence_async_smem();
    } else if (inside) {
#pragma unroll
      for (int co = 0; co < 3; ++co)
        if (co < Cout) out[(((size_t)b * Cout + co) * H + Y) * W + X] = acc[co];
    }
    __syncthreads();                                        // Z is free again; the output planes are complete
    if (tid == 0 && use_tma_store) {
      for (int co = 0; co < Cout; ++co) tma_store_3d(&omap, sO + co * 512, tx * 16, ty * 8, b * Cout + co);
      bulk_commit();
    }
  }
  if (tid == 0 && use_tma_store) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

// ----------------------------------------------------------------------------------------------
// input_proj on the tensor core (Cin = 3, E in {16, 32}): the 3 x 3 x 3 im2col row of a pixel (K = 27) is built in shared memory
// and multiplied with the weight by tcgen05.  Both operands keep fp32 accuracy with bf16 inputs through a three-term split
//   x w ~ xh wh + xh wl + xl wh      (xh = bf16(x), xl = bf16(x - xh); the dropped xl wl term is 2^-18 relative)
// laid out as ONE K = 96 chain: A k-block 0 = [xh | xh], k-block 1 = [xl | -];  B k-block 0 = [wh | wl], k-block 1 = [wh | -].
// PERSISTENT CTAs (4 per SM) walk 8 x 16 pixel tiles: the fp32 image halo (3 x 10 x 18, zero outside the image = the conv's
// padding) is staged in shared memory — the loads of tile i+1 are issued before the GEMM of tile i is waited for — every
// thread builds the im2col row of its pixel (12 16-byte stores), one elected thread issues six M = 128 UMMAs, and the epilogue
// (thread = TMEM lane = pixel) adds the bias, applies LeakyReLU(0.01) and writes the pixel's E bf16 channels.
// Replaces 864 fp32 FMAs per pixel of the SIMT kernel (which stays for other widths).
// ----------------------------------------------------------------------------------------------
template <int E>
struct InProjCfg {
  static constexpr int S_A = 0;                             // 2 k-blocks x 128 rows x 128 B
  static constexpr int S_W = 2 * 16384;                     // 2 k-blocks x E rows x 128 B
  static constexpr int S_HALO = S_W + 2 * E * 128;          // 3 x 10 x 18 fp32
  static constexpr int S_BAR = S_HALO + 2176;
  static constexpr int SMEM_BYTES = S_BAR + 64 + 1024;
  static_assert(S_W % 1024 == 0 && (E * 128) % 1024 == 0, "operand alignment");
};

template <int E>
__global__ void __launch_bounds__(128, 4) input_proj_tc_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                                            bf16* __restrict__ tok, int B, int H, int W, int tiles_x, int tiles_y, int n_tiles) {
  using Cfg = InProjCfg<E>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::S_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::S_BAR + 32);
  float* halo = reinterpret_cast<float*>(smem + Cfg::S_HALO);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t sA = smem_u32(smem + Cfg::S_A), sW = smem_u32(smem + Cfg::S_W);

  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 32);
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  // B operand images, built once per CTA: k-block 0 row e = [wh(27) 0.. | wl(27) 0..], k-block 1 row e = [wh(27) 0.. | 0]
  for (int i = tid; i < 2 * E * 64; i += 128) {
    const int kb = i / (E * 64), e = (i / 64) % E, c = i % 64, k = c & 31;
    float v = 0.f;
    if (k < 27 && !(kb == 1 && c >= 32)) {
      const float wf = __ldg(w + e * 27 + k);
      const float hi = __bfloat162float(__float2bfloat16_rn(wf));
      v = (kb == 0 && c >= 32) ? wf - hi : hi;
    }
    const __nv_bfloat16 hv = __float2bfloat16_rn(v);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(sW + kb * (E * 128) + swz<128>(e, 2 * c)), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, E);
  const int py = tid >> 4, px = tid & 15;
  float bv[E];
#pragma unroll
  for (int e = 0; e < E; ++e) bv[e] = __ldg(bias + e);

  // halo elements i = tid + 128 j (j < 5) of a tile: channel i / 180, row (i % 180) / 18, column i % 18
  float pre[5];
  auto prefetch = [&](int it) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = tid + 128 * j;
      const int c = i / 180, r = (i % 180) / 18, col = i % 18;
      const int Y = ty * 8 - 1 + r, X = tx * 16 - 1 + col;
      pre[j] = (i < 540 && Y >= 0 && Y < H && X >= 0 && X < W) ? __ldg(img + (((size_t)b * 3 + c) * H + Y) * W + X) : 0.f;
    }
  };
  if (n_my > 0) prefetch(0);

  for (int it = 0; it < n_my; ++it) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (tid + 128 * j < 540) halo[tid + 128 * j] = pre[j];
    tc_fence_before();                                       // (the previous tile's TMEM reads precede the barrier, the next GEMM follows it)
    __syncthreads();
    // ---- im2col row of this thread's pixel: k = ci*9 + ky*3 + kx, hi / lo bf16 split ----
    {
      uint32_t hi[16], lo[16];                               // 32 bf16 each (k = 27..31 zero)
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        float v0 = 0.f, v1 = 0.f;
        const int k0 = 2 * k2, k1 = 2 * k2 + 1;
        if (k0 < 27) v0 = halo[(k0 / 9) * 180 + (py + (k0 % 9) / 3) * 18 + px + (k0 % 3)];
        if (k1 < 27) v1 = halo[(k1 / 9) * 180 + (py + (k1 % 9) / 3) * 18 + px + (k1 % 3)];
        const uint32_t h = pack_bf16(v0, v1);
        hi[k2] = h;
        lo[k2] = pack_bf16(v0 - __uint_as_float(h << 16), v1 - __uint_as_float(h & 0xffff0000u));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 hv = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        sts128(sA + swz<128>(tid, 16 * j), hv);
        sts128(sA + swz<128>(tid, 64 + 16 * j), hv);
        sts128(sA + 16384 + swz<128>(tid, 16 * j), make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]));
      }
    }
    fence_async_smem();
    if (it + 1 < n_my) prefetch(it + 1);                     // next tile's halo: in flight under the GEMM and the epilogue
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss(tb, kmajor_desc<128>(sA + ks * 32), kmajor_desc<128>(sW + ks * 32), idesc, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) umma_ss(tb, kmajor_desc<128>(sA + 16384 + ks * 32), kmajor_desc<128>(sW + E * 128 + ks * 32), idesc, 1u);
        umma_commit(smem_u32(&bars[0]));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bars[0]), it & 1);
    tc_fence_after();
    // ---- epilogue: thread = TMEM lane = pixel: + bias, LeakyReLU(0.01), E bf16 channels ----
    {
      uint32_t v[E];
      if (E == 32) tmem_ld32(tb + ((uint32_t)(32 * warp) << 16), v);
      else tmem_ld16(tb + ((uint32_t)(32 * warp) << 16), v);
      tmem_wait_ld();
      const int Y = ty * 8 + py, X = tx * 16 + px;
      if (Y < H && X < W) {
        bf16* o = tok + ((size_t)(b * H + Y) * W + X) * E;
#pragma unroll
        for (int e0 = 0; e0 < E; e0 += 8) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = __uint_as_float(v[e0 + j]) + bv[e0 + j];
            f[j] = a > 0.f ? a : 0.01f * a;
          }
          *reinterpret_cast<uint4*>(o + e0) = pack8(f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 32);
}

}  // namespace lw
