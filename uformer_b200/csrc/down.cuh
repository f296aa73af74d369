// down.cuh — Downsample: Conv2d(k4, s2, p1) on the token map as an implicit GEMM (model.py:739-746).
//   out[b, (y,x), co] = bias[co] + sum_{ky,kx,ci} in[b, 2y-1+ky, 2x-1+kx, ci] * W[co, ci, ky, kx]
// One CTA = 128 consecutive output tokens, K = 16*Cin walked in 64-wide k-blocks (k = tap*Cin + ci).
// The im2col A k-blocks are gathered straight into the swizzled UMMA tile with cp.async (zero-fill = the
// conv's zero padding), four k-blocks in flight; no registers, no separate staging pass.  The accumulator
// D[128 x Cout] stays in TMEM for the whole tile; the epilogue uses the fast 16x256b TMEM shape and the
// staged, coalesced copy-out.
#pragma once
#include "lewin_common.cuh"
#include "leff.cuh"
#include "leff2.cuh"

namespace lw {

struct DownCfg {
  static constexpr int ABUF = 4;        // A k-block buffers
  static constexpr int S_A = 0;
  static constexpr int S_MISC = ABUF * 16384;
  static constexpr int S_RING = S_MISC + 1024;
  // weight ring depth (kernel argument): 2 stages where the accumulator is <= 256 TMEM columns — two CTAs then share an SM and
  // one runs its epilogue under the other's main loop; 4 stages for wider outputs (Cout = 512 streams four 16 KB chunks per
  // k-block), whose 133 KB keep one CTA per SM (a second one could only block in tcgen05.alloc).
  static constexpr int smem_bytes(int stages) { return S_RING + stages * kStageBytes + 1024; }
};
static_assert(DownCfg::S_RING % 1024 == 0, "ring alignment");

struct DownMisc {
  int row_tok[128];
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_a_full[4], bar_a_empty[4];
  uint64_t bar_d_full;
  uint32_t tmem_base;
};
static_assert(sizeof(DownMisc) <= 1024, "misc too large");

__device__ __forceinline__ void cp_async_wait_group2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int OCC>      // CTAs per SM the register allocation is bounded for (2: <= 96 registers; 1: the deep-K Cout = 512 case keeps its 142)
__global__ void __launch_bounds__(kThreads8, OCC) down_kernel(const AStreamArgs a, const int t_alloc, const int stages) {
  using Cfg = DownCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  DownMisc& ms = *reinterpret_cast<DownMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int KB = a.K / 64;
  const int NC = a.N / a.nch;
  const int Ho = a.H / 2, Wo = a.W / 2;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1);
      mbar_init(smem_u32(&ms.bar_a_full[s]), kWorkers8); mbar_init(smem_u32(&ms.bar_a_empty[s]), 1);
    }
    mbar_init(smem_u32(&ms.bar_d_full), 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&ms.tmem_base), t_alloc);
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const uint32_t chunk_bytes = a.nch * 128;
  const uint32_t sA0 = smem_u32(smem + Cfg::S_A);

  if (warp == 8) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), stages, 0};
      for (int kb = 0; kb < KB; ++kb)
        for (int nc = 0; nc < NC; ++nc)
          ring.load(a.w_img + (size_t)(kb * NC + nc) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 9) {
    Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), stages, 0};
    const uint32_t idesc = make_idesc_bf16(128, a.nch);
    const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
    const uint64_t b_desc0 = kmajor_desc<128>(ring_base);
    for (int kb = 0; kb < KB; ++kb) {
      const int ab = kb & 3;
      mbar_wait(smem_u32(&ms.bar_a_full[ab]), (kb >> 2) & 1);
      tc_fence_after();
      const uint64_t ad = kmajor_desc<128>(sA0 + ab * 16384);
      for (int nc = 0; nc < NC; ++nc) {
        const uint32_t wst = ring.acquire();
        const uint64_t bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tb + nc * a.nch, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) != 0);
        }
        __syncwarp();
        ring.release();
      }
      if (elect_one()) umma_commit(smem_u32(&ms.bar_a_empty[ab]));
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full));
    __syncwarp();
  } else {
    // ---------------- workers: im2col gather with cp.async ----------------
    // thread -> 16-byte chunk v = tid & 7 of tile rows (tid >> 3) + 32 * it, it = 0..3
    const int v = tid & 7;
    int base_off[4], y2[4], x2[4];     // image base (elements), 2*oy-1, 2*ox-1 ; base_off < 0: padding row
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = (tid >> 3) + 32 * it;
      const int row = tile * 128 + rr;
      if (row < a.B * Ho * Wo) {
        const int b = row / (Ho * Wo), t = row % (Ho * Wo);
        base_off[it] = b * a.H * a.W;
        y2[it] = 2 * (t / Wo) - 1;
        x2[it] = 2 * (t % Wo) - 1;
      } else {
        base_off[it] = -1; y2[it] = 0; x2[it] = 0;
      }
    }
    if (tid < 128) {
      const int row = tile * 128 + tid;
      ms.row_tok[tid] = (row < a.B * Ho * Wo) ? row : -1;
    }
    auto gather = [&](int kb) {
      const int ab = kb & 3;
      mbar_wait(smem_u32(&ms.bar_a_empty[ab]), ((kb >> 2) & 1) ^ 1);
      const int k0 = kb * 64 + v * 8;
      const int tap = k0 / a.Cin, ci = k0 - tap * a.Cin;
      const int ky = tap >> 2, kx = tap & 3;
      const uint32_t dst0 = sA0 + ab * 16384;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = (tid >> 3) + 32 * it;
        const int iy = y2[it] + ky, ix = x2[it] + kx;
        const bool ok = base_off[it] >= 0 && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        const bf16* g = ok ? a.src + ((size_t)(base_off[it] + iy * a.W + ix)) * a.src_stride + ci : a.src;
        cp_async16(dst0 + swz<128>(rr, v * 16), g, ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    auto publish = [&](int kb) {       // this thread's part of k-block kb has landed
      fence_async_smem();
      mbar_arrive(smem_u32(&ms.bar_a_full[kb & 3]));
    };
    for (int kb = 0; kb < KB; ++kb) {
      gather(kb);
      if (kb >= 2) { cp_async_wait_group2(); publish(kb - 2); }
    }
    if (KB >= 2) { cp_async_wait_group1(); publish(KB - 2); }
    cp_async_wait_all();
    publish(KB - 1);

    // ---------------- epilogue: + bias -> bf16 -> staging (A buffers are free) -> coalesced copy-out ----------------
    const uint32_t stage_s = sA0;
    const int sub_cols = a.N < 128 ? a.N : 128;
    const int pitch = sub_cols * 2 + 16;
    int sub_log2 = 4;
    while ((1 << sub_log2) < sub_cols) ++sub_log2;
    const int row16 = (warp & 3) * 32 + (warp >> 2) * 16;
    mbar_wait(smem_u32(&ms.bar_d_full), 0);
    tc_fence_after();
    for (int sc = 0; sc < a.N; sc += 128) {
      const uint32_t tcol = tb + ((uint32_t)row16 << 16) + sc;
      if (sub_cols == 128) { epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0); epi_cols<8>(tcol + 64, a.bias + sc + 64, stage_s, pitch, row16, 64); }
      else if (sub_cols == 64) epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      else if (sub_cols == 32) epi_cols<4>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      else epi_cols<2>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      worker_bar8();
      store_staged_rows(stage_s, pitch, sub_log2, ms.row_tok, a.out, nullptr, (size_t)a.N, sc, tid, kWorkers8);
      worker_bar8();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, t_alloc);
}

}  // namespace lw
