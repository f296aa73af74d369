// leff2.cuh — LeFF part 2 (C = 512; narrower blocks use leff_fused.cuh): out = resid + GELU(dwconv3x3(h1) + bd) W2^T + b2
// (model.py:674-682).  h1, the conv arithmetic (HFMA2), the GELU and the GEMM operands are HALF precision (fp16).
//
// PERSISTENT kernel: one CTA per SM walks 8 x 16 spatial tiles (128 tokens) round-robin.  512 threads in four
// warpgroups; setmaxnreg moves registers from the light warpgroups (88) to the conv warpgroups (168):
//   warps 0-7   conv     : depthwise conv + GELU producing the GEMM A operand, back to back across tiles
//   warp  8     producer : W2 chunk images -> smem ring (cp.async.bulk); owns the TMEM allocation
//   warp  9     issuer   : tcgen05.mma; the D[128 x C] accumulator of a tile lives in TMEM, double-buffered
//                          when 2C <= 512 columns so tile i+1 accumulates while tile i is drained
//   warps 12-15 epilogue : TMEM -> +bias -> bf16 -> staging -> coalesced store (+ residual), in the background
// The hidden dimension is walked in 64-channel slices.  For each slice the 10 x 18 halo'd h1 tile is fetched
// with cp.async (zero-filled outside the image = the conv's zero padding of h1, model.py:659) one slice ahead
// — across tile boundaries too.  One conv warp per channel octet: its 3x3 taps are warp-uniform broadcast
// reads; lanes are 16 columns x 2 row-halves and slide down their column, so every h1 vector is read from
// shared memory 3x (not 9x); arithmetic is packed FFMA2 on fp32 pairs; the halo tile is XOR-swizzled by
// token so quarter-warps never conflict.
#pragma once
#include "lewin_common.cuh"
#include "leff.cuh"

namespace lw {

constexpr int kL2Conv = 256;          // conv threads (warps 0-7)
constexpr int kL2Threads = 512;       // + warpgroup 2 (producer, issuer, 2 idle warps) + warpgroup 3 (4 epilogue warps)

struct Leff2Cfg {
  static constexpr int STAGES = 4;
  static constexpr int HALO_TOK = 180;                          // 10 x 18
  static constexpr int S_A = 0;                                 // 2 x 16 KB A k-block buffers
  static constexpr int S_HALO = 2 * 16384;                      // 2 x 23552 (180 x 128 B, padded)
  static constexpr int S_WD = S_HALO + 2 * 23552;               // 2 x (10 x 64 fp16) = 2 x 1280 (region sized 2 x 2560 + 1024 pad)
  static constexpr int S_RING = S_WD + 2 * 2560 + 1024;         // = 86016, 1024-aligned
  static constexpr int S_STAGE = S_RING + STAGES * kStageBytes; // epilogue staging tile 128 x 272 B
  static constexpr int S_MISC = S_STAGE + 35840;
  static constexpr int SMEM_BYTES = S_MISC + 1024 + 1024;
};
static_assert(Leff2Cfg::S_RING % 1024 == 0, "ring alignment");
static_assert(Leff2Cfg::SMEM_BYTES <= 232448, "smem budget");

struct Leff2Misc {
  int row_tok[128];
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_a_full[2], bar_a_empty[2];
  uint64_t bar_d_full[2], bar_d_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Leff2Misc) <= 1024, "misc too large");

__device__ __forceinline__ void l2_conv_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void l2_epi_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// accumulator columns [tcol, +8*NB) of 16 TMEM lanes -> + bias -> bf16 -> staging tile rows row16..+16, columns col0..
template <int NB>
__device__ __forceinline__ void epi_cols(uint32_t tcol, const float* __restrict__ bias, uint32_t stage_s, int pitch, int row16, int col0) {
  const int tq = threadIdx.x & 3;
  uint32_t v[4 * NB];
  if (NB == 8) tmem_ld_16x256b_x8(tcol, v); else if (NB == 4) tmem_ld_16x256b_x4(tcol, v); else tmem_ld_16x256b_x2(tcol, v);
  f2 bb[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + 8 * i + 2 * tq));
    bb[i] = f2_pack(b2.x, b2.y);
  }
  tmem_wait_ld();
  uint32_t pk[2 * NB];
  frag_bias_act_pack<NB, false>(v, bb, pk);
  stage_frag<NB>(stage_s, pitch, row16, col0, pk);
}

__global__ void __launch_bounds__(kL2Threads, 1) leff2_kernel(const AStreamArgs a, const int t_alloc, const int n_tiles, const int nbuf_d) {
  using Cfg = Leff2Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Leff2Misc& ms = *reinterpret_cast<Leff2Misc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = a.K / 64;
  const int NC = a.N / a.nch;
  const int tiles_y = a.H / 8;                 // tiles never straddle images (H % 8 == 0)

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ms.bar_a_full[i]), kL2Conv); mbar_init(smem_u32(&ms.bar_a_empty[i]), 1);
      mbar_init(smem_u32(&ms.bar_d_full[i]), 1); mbar_init(smem_u32(&ms.bar_d_empty[i]), 128);
    }
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
  // Register re-balancing per warpgroup (the branches below never re-merge before the final barrier, so ptxas
  // allocates each role under its own budget): every scheduler partition hosts 2 conv warps + 2 light warps,
  // 2*168 + 2*88 = 512 registers per lane.
  const int wg = warp >> 2;
  if (wg == 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == 8) {
    // ============================== producer ==============================
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int kb = 0; kb < KB; ++kb)
          for (int nc = 0; nc < NC; ++nc)
            ring.load(a.w_img + (size_t)(kb * NC + nc) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 9) {
    // ============================== issuer (warp-uniform; one elected lane issues) ==============================
    Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
    const uint32_t idesc = make_idesc_f16(128, a.nch);          // fp16 A (conv output) x fp16 W2 image
    const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
    const uint64_t b_desc0 = kmajor_desc<128>(ring_base);
    int g = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int buf = (nbuf_d == 2) ? (j & 1) : 0;
      const int use = (nbuf_d == 2) ? (j >> 1) : j;            // how many times this buffer was used before
      mbar_wait(smem_u32(&ms.bar_d_empty[buf]), (use & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int ab = g & 1;
        mbar_wait(smem_u32(&ms.bar_a_full[ab]), (g >> 1) & 1);
        tc_fence_after();
        const uint64_t ad = kmajor_desc<128>(smem_u32(smem + Cfg::S_A + ab * 16384));
        for (int nc = 0; nc < NC; ++nc) {
          const uint32_t wst = ring.acquire();
          const uint64_t bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss(tb + buf * a.N + nc * a.nch, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_a_empty[ab]));
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full[buf]));
      __syncwarp();
    }
  }
  } else if (wg < 2) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    // ============================== conv warps ==============================
    const int v = warp;                  // channel octet inside the 64-channel slice
    const int cx = lane & 15;            // tile column
    const int hf = lane >> 4;            // row half: output rows hf*4 .. hf*4+3
    const uint32_t halo0 = smem_u32(smem + Cfg::S_HALO), wd0 = smem_u32(smem + Cfg::S_WD);
    // slice-independent parts of the prefetch: smem offsets of this thread's <= 6 halo vectors, taps vector
    uint32_t pf_soff[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int idx = tid + i * kL2Conv;
      const int t = idx >> 3, vv = idx & 7;
      pf_soff[i] = (idx < Cfg::HALO_TOK * 8) ? (uint32_t)(t * 128 + ((vv ^ (t & 7)) * 16)) : 0xffffffffu;   // chunk XOR-swizzled by token
    }
    const int tw_row = tid >> 3, tw_vec = tid & 7;                   // taps: 10 rows x 8 vectors of 8 halves per 64-channel slice
    const uint16_t* tw_src = a.taps + (size_t)(tw_row < 10 ? tw_row : 0) * a.K + tw_vec * 8;
    GeluH2 gelu;
    gelu.init();

    // global element offsets of the halo vectors of a tile (0xffffffff: outside the image -> zero fill)
    auto tile_desc = [&](int tile, uint32_t* goff) {
      const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % tiles_y, b = tile / (a.tiles_x * tiles_y);
      const int y0 = ty * 8, x0 = tx * 16;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int idx = tid + i * kL2Conv;
        const int t = idx >> 3, vv = idx & 7;
        const int y = y0 - 1 + t / 18, x = x0 - 1 + t % 18;
        const bool ok = (idx < Cfg::HALO_TOK * 8) && (y >= 0) && (y < a.H) && (x >= 0) && (x < a.W);
        goff[i] = ok ? (uint32_t)(((b * a.H + y) * a.W + x) * a.K + vv * 8) : 0xffffffffu;
      }
    };
    auto prefetch = [&](const uint32_t* goff, int kb, int bufi) {
      const uint32_t hb = halo0 + bufi * 23552;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (pf_soff[i] != 0xffffffffu) {
          const bool ok = goff[i] != 0xffffffffu;
          cp_async16(hb + pf_soff[i], ok ? a.src + goff[i] + kb * 64 : a.src, ok ? 16u : 0u);
        }
      }
      if (tid < 80) cp_async16(wd0 + bufi * 2560 + tid * 16, tw_src + kb * 64, 16u);
      cp_async_commit();
    };

    uint32_t gcur[6], gnext[6];
    int g = 0;
    if ((int)blockIdx.x < n_tiles) { tile_desc(blockIdx.x, gcur); prefetch(gcur, 0, 0); }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int next_tile = tile + gridDim.x;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int ab = g & 1;
        cp_async_wait_all();
        l2_conv_bar();                         // slice g halo + taps visible; buffers of slice g-1 free
        if (kb + 1 < KB) prefetch(gcur, kb + 1, (g + 1) & 1);
        else if (next_tile < n_tiles) { tile_desc(next_tile, gnext); prefetch(gnext, 0, (g + 1) & 1); }
        const uint32_t sH = halo0 + ab * 23552;
        const uint32_t sW = wd0 + ab * 2560;
        // accumulators start at the conv bias of this octet; packed-half arithmetic (h1 is fp16: no unpacking, 2 channels per op)
        h2 acc[4][4];
        {
          const uint4 b0 = lds128(sW + (9 * 64 + v * 8) * 2);
#pragma unroll
          for (int o = 0; o < 4; ++o) { acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w; }
        }
        // column by column (dx): pull the 6 halo rows of that column into registers once, then apply the three
        // taps (ky) that touch them; halo row r feeds output row o = r - ky.
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          uint4 h[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            const int t = (hf * 4 + r) * 18 + cx + dx;
            h[r] = lds128(sH + t * 128 + ((v ^ (t & 7)) * 16));
          }
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint4 w = lds128(sW + ((ky * 3 + dx) * 64 + v * 8) * 2);            // warp-uniform address: broadcast
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              acc[o][0] = h2_fma(h[o + ky].x, w.x, acc[o][0]);
              acc[o][1] = h2_fma(h[o + ky].y, w.y, acc[o][1]);
              acc[o][2] = h2_fma(h[o + ky].z, w.z, acc[o][2]);
              acc[o][3] = h2_fma(h[o + ky].w, w.w, acc[o][3]);
            }
          }
        }
        // A buffer ab must have been consumed by the MMAs of slice g-2
        mbar_wait(smem_u32(&ms.bar_a_empty[ab]), ((g >> 1) & 1) ^ 1);
        const uint32_t sA = smem_u32(smem + Cfg::S_A + ab * 16384);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          uint4 pk;
          pk.x = gelu(acc[o][0]);
          pk.y = gelu(acc[o][1]);
          pk.z = gelu(acc[o][2]);
          pk.w = gelu(acc[o][3]);
          const int rr = (hf * 4 + o) * 16 + cx;
          sts128(sA + swz<128>(rr, v * 16), pk);
        }
        fence_async_smem();
        mbar_arrive(smem_u32(&ms.bar_a_full[ab]));
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) gcur[i] = gnext[i];
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // ============================== epilogue warps 12-15 (TMEM lane quadrant = warp & 3) ==============================
    const int et = tid - 12 * 32;                         // 0..127
    const int q = warp & 3;
    const uint32_t stage_s = smem_u32(smem + Cfg::S_STAGE);
    const int sub_cols = a.N < 128 ? a.N : 128;
    const int pitch = sub_cols * 2 + 16;
    int sub_log2 = 4;
    while ((1 << sub_log2) < sub_cols) ++sub_log2;
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int buf = (nbuf_d == 2) ? (j & 1) : 0;
      const int use = (nbuf_d == 2) ? (j >> 1) : j;
      {   // destination token of every tile row
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % tiles_y, b = tile / (a.tiles_x * tiles_y);
        const int y = ty * 8 + (et >> 4), x = tx * 16 + (et & 15);
        ms.row_tok[et] = (x < a.W) ? ((b * a.H + y) * a.W + x) : -1;
      }
      mbar_wait(smem_u32(&ms.bar_d_full[buf]), use & 1);
      tc_fence_after();
      for (int sc = 0; sc < a.N; sc += 128) {
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          const int row16 = q * 32 + hl * 16;
          const uint32_t tcol = tb + ((uint32_t)row16 << 16) + buf * a.N + sc;
          if (sub_cols == 128) { epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0); epi_cols<8>(tcol + 64, a.bias + sc + 64, stage_s, pitch, row16, 64); }
          else if (sub_cols == 64) epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
          else if (sub_cols == 32) epi_cols<4>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
          else epi_cols<2>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
        }
        if (sc + 128 >= a.N) {                            // accumulator fully read: hand the TMEM buffer back
          tc_fence_before();
          mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
        }
        l2_epi_bar();
        if (a.resid_fp32 | a.out_fp32) store_staged_rows_mixed<128>(stage_s, pitch, sub_log2, smem_u32(ms.row_tok), a.out, a.resid, nullptr, a.resid_fp32 != 0,
                                                                    a.out_fp32 != 0, (size_t)a.N, sc, et);
        else store_staged_rows128(stage_s, pitch, sub_log2, ms.row_tok, a.out, a.resid, (size_t)a.N, sc, et);
        l2_epi_bar();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, t_alloc);
}

}  // namespace lw
