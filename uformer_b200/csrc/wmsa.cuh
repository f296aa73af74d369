// wmsa.cuh — fused shifted-window multi-head self-attention for one LeWin block.
//
// One CTA = 128 tokens = two 8x8 windows.  Everything between the block input and the first
// residual sum of LeWinTransformerBlock.forward (model.py:951-986) happens in this kernel:
//   gather (roll by -shift + window_partition folded into the load addresses, model.py:956-963)
//   -> LayerNorm (fp32) -> + modulator -> bf16 A operand in shared memory (SWIZZLE_128B)
//   per head: QKV projection on tcgen05 (D in TMEM) -> Q,K,V tiles in shared memory
//             S = Q K^T (tcgen05) -> + relative-position bias (+ shift / explicit mask) -> softmax
//             (fp32, registers) -> P written to TMEM as bf16 A operand -> O = P V (tcgen05, V
//             consumed MN-major straight from its row-major tile) -> O/rowsum to TMEM (bf16)
//   output projection with A = O resident in TMEM -> + bias + shortcut -> scatter
//   (window_reverse + roll by +shift folded into the store addresses, model.py:975-983).
// The (nW,64,64) shift mask the reference rebuilds every forward (model.py:924-942) is replaced by
// its closed form on per-token region ids.
#pragma once
#include "lewin_common.cuh"
#include "../../include/lewin_b200.h"

namespace lw {

template <int C, int HD>
struct WmsaCfg {
  static constexpr int NH = C / HD;
  static constexpr int KB = (C + 63) / 64;                 // 64-channel k-blocks of the A operand
  static constexpr int KSTEPS = C / 16;                    // UMMA K steps over the channels
  static constexpr int SWH = 2 * HD;                       // swizzle bytes of the per-head tiles
  static constexpr int QKV_N = 3 * HD;
  static constexpr int QKV_CHUNK_BYTES = QKV_N * 128;      // per (head, k-block) weight image
  static constexpr int NCH = C < 128 ? C : 128;            // proj N chunk
  static constexpr int NC = C / NCH;
  static constexpr int PROJ_CHUNK_BYTES = NCH * 128;
  static constexpr int STAGES = (C >= 512) ? 3 : (C == 128 ? 2 : 4);   // C=128: 2 stages so two CTAs fit per SM
  // TMEM columns
  static constexpr int T_OALL = 0;                         // O for all heads, bf16 packed: C/2 cols
  static constexpr int T_WORK = (C / 2 < 32) ? 32 : C / 2; // S / D_qkv (aliased), 128 cols
  static constexpr int T_DO = T_WORK + 128;                // D_o, HD cols (<= 32)
  static constexpr int T_NEED = T_WORK + ((NC > 1) ? 256 : 160);
  static constexpr int T_ALLOC = T_NEED <= 256 ? 256 : 512;
  // shared memory map (bytes)
  static constexpr int S_X = 0;
  static constexpr int S_Q = KB * 16384;
  static constexpr int TILE_B = 128 * SWH;
  static constexpr int S_K = S_Q + TILE_B;
  static constexpr int S_V = S_K + TILE_B;
  static constexpr int S_RING = S_V + TILE_B;
  static constexpr int S_MISC = S_RING + STAGES * kStageBytes;
  static constexpr int SMEM_BYTES = S_MISC + 4096 + 1024;  // + slack for 1024 B alignment
};

struct WmsaMisc {
  float relpos[232];
  int row_tok[128];
  uint8_t region[128];
  int win_mixed[2];
  float xmax[2][128];      // softmax row max / sum halves exchanged between the two column-half threads
  float xsum[2][128];
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_xn, bar_qkv_full, bar_qkv_staged, bar_s_full, bar_p_ready, bar_o_full, bar_oall;
  uint64_t bar_d_full[2], bar_d_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(WmsaMisc) <= 4096, "misc too large");

template <int C, int HD>
__global__ void __launch_bounds__(kThreads8, (C <= 128) ? 2 : 1) wmsa_kernel(const lw_wmsa_args a) {
  using Cfg = WmsaCfg<C, HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WmsaMisc& ms = *reinterpret_cast<WmsaMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const bf16* __restrict__ xin = reinterpret_cast<const bf16*>(a.x);

  // ---------------- setup ----------------
  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    mbar_init(smem_u32(&ms.bar_xn), kWorkers8);
    mbar_init(smem_u32(&ms.bar_qkv_full), 1);
    mbar_init(smem_u32(&ms.bar_qkv_staged), kWorkers8);
    mbar_init(smem_u32(&ms.bar_s_full), 1);
    mbar_init(smem_u32(&ms.bar_p_ready), kWorkers8);
    mbar_init(smem_u32(&ms.bar_o_full), 1);
    mbar_init(smem_u32(&ms.bar_oall), kWorkers8);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_d_full[i]), 1); mbar_init(smem_u32(&ms.bar_d_empty[i]), kWorkers8); }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&ms.tmem_base), Cfg::T_ALLOC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;

  const uint32_t sX = smem_u32(smem + Cfg::S_X), sQ = smem_u32(smem + Cfg::S_Q), sK = smem_u32(smem + Cfg::S_K),
                 sV = smem_u32(smem + Cfg::S_V);

  if (warp == 8) {
    // ======================= producer: weight chunk images =======================
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      const uint8_t* wq = reinterpret_cast<const uint8_t*>(a.wqkv_img);
      for (int h = 0; h < Cfg::NH; ++h)
        for (int kb = 0; kb < Cfg::KB; ++kb)
          ring.load(wq + (size_t)(h * Cfg::KB + kb) * Cfg::QKV_CHUNK_BYTES, Cfg::QKV_CHUNK_BYTES);
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(a.wproj_img);
      for (int nc = 0; nc < Cfg::NC; ++nc)
        for (int kb = 0; kb < Cfg::KB; ++kb)
          ring.load(wp + (size_t)(nc * Cfg::KB + kb) * Cfg::PROJ_CHUNK_BYTES, Cfg::PROJ_CHUNK_BYTES);
    }
  } else if (warp == 9) {
    // ======================= issuer: all tcgen05.mma =======================
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      constexpr uint32_t idesc_qkv = make_idesc_bf16(128, Cfg::QKV_N);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
      constexpr uint32_t idesc_proj = make_idesc_bf16(128, Cfg::NCH);
      mbar_wait(smem_u32(&ms.bar_xn), 0);
      tc_fence_after();
      for (int h = 0; h < Cfg::NH; ++h) {
        const uint32_t ph = h & 1;
        if (h > 0) { mbar_wait(smem_u32(&ms.bar_o_full), ph ^ 1); tc_fence_after(); }  // PV(h-1) done: S/P columns reusable
        // --- D_qkv[128 x 3HD] = Xn * Wqkv_h^T ---
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          constexpr int KS = (C >= 64) ? 4 : C / 16;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t ad = kmajor_desc<128>(sX + kb * 16384 + ks * 32);
            const uint64_t bd = kmajor_desc<128>(wst + ks * 32);
            umma_ss(tb + Cfg::T_WORK, ad, bd, idesc_qkv, (kb | ks) != 0);
          }
          ring.release();
        }
        umma_commit(smem_u32(&ms.bar_qkv_full));
        // --- S[128 x 128] = Q_h K_h^T (block-diagonal 64x64 halves are the two windows) ---
        mbar_wait(smem_u32(&ms.bar_qkv_staged), ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t ad = kmajor_desc<Cfg::SWH>(sQ + ks * 32);
          const uint64_t bd = kmajor_desc<Cfg::SWH>(sK + ks * 32);
          umma_ss(tb + Cfg::T_WORK, ad, bd, idesc_s, ks != 0);
        }
        umma_commit(smem_u32(&ms.bar_s_full));
        // --- D_o[128 x HD] = P[128 x 128 keys] (TMEM, bf16) * V_h (MN-major tile) ---
        mbar_wait(smem_u32(&ms.bar_p_ready), ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bd = mnmajor_desc<Cfg::SWH>(sV + ks * 16 * Cfg::SWH, 8 * Cfg::SWH);
          umma_ts(tb + Cfg::T_DO, tb + Cfg::T_WORK + ks * 8, bd, idesc_pv, ks != 0);
        }
        umma_commit(smem_u32(&ms.bar_o_full));
      }
      // --- output projection: D_out[128 x C] = O_all (TMEM) * Wp^T, N chunks of NCH ---
      mbar_wait(smem_u32(&ms.bar_oall), 0);
      tc_fence_after();
      for (int nc = 0; nc < Cfg::NC; ++nc) {
        const int buf = nc & 1;
        mbar_wait(smem_u32(&ms.bar_d_empty[buf]), ((nc >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          constexpr int KS = (C >= 64) ? 4 : C / 16;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t bd = kmajor_desc<128>(wst + ks * 32);
            umma_ts(tb + Cfg::T_WORK + buf * 128, tb + Cfg::T_OALL + kb * 32 + ks * 8, bd, idesc_proj, (kb | ks) != 0);
          }
          ring.release();
        }
        umma_commit(smem_u32(&ms.bar_d_full[buf]));
      }
    }
  } else {
    // ======================= workers (8 warps) =======================
    // thread -> tile row r = (warp&3)*32 + lane (== TMEM lane), column half hf = warp>>2
    const int r = (warp & 3) * 32 + lane;
    const int hf = warp >> 2;
    const int wl = r >> 6, i = r & 63;   // window within the tile, token within the window
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // ---- source token of every row + region id for the shift mask (one thread per row) ----
    if (hf == 0) {
      const int w = tile * 2 + wl;
      int tok = -1;
      uint8_t reg = 0;
      if (w < a.n_windows) {
        if (a.windowed) {
          tok = w * 64 + i;
        } else {
          const int nwx = a.W >> 3, nwy = a.H >> 3;
          const int b = w / (nwx * nwy), wi = w % (nwx * nwy);
          const int ry = (wi / nwx) * 8 + (i >> 3), rx = (wi % nwx) * 8 + (i & 7);   // rolled coordinates
          int y = ry + a.shift, x = rx + a.shift;
          if (y >= a.H) y -= a.H;
          if (x >= a.W) x -= a.W;
          tok = (b * a.H + y) * a.W + x;
          if (a.shift > 0) {
            const int gy = (ry >= a.H - 8) + (ry >= a.H - a.shift);
            const int gx = (rx >= a.W - 8) + (rx >= a.W - a.shift);
            reg = (uint8_t)(3 * gy + gx);
          }
        }
      }
      ms.row_tok[r] = tok;
      ms.region[r] = reg;
      if (i == 0) ms.win_mixed[wl] = 0;
    }
    worker_bar8();
    if (hf == 0 && a.shift > 0 && !a.windowed && ms.region[r] != ms.region[wl * 64]) ms.win_mixed[wl] = 1;
    // ---- A operand: LN(x) + modulator ----
    stage_rows_ln<C, 8>(smem + Cfg::S_X, xin, ms.row_tok, a.ln_w, a.ln_b, a.ln_eps, a.modulator);
    fence_async_smem();
    mbar_arrive(smem_u32(&ms.bar_xn));

    const int tok = ms.row_tok[r];
    const int yi = i >> 3, xi = i & 7;
    const int rp_base = (yi + 7) * 15 + xi + 7;
    const uint32_t relpos_s = smem_u32(&ms.relpos[0]);
    const uint32_t region_s = smem_u32(&ms.region[0]);
    (void)region_s;

    for (int h = 0; h < Cfg::NH; ++h) {
      const uint32_t ph = h & 1;
      // relative-position bias row of this head -> smem (readers of the previous head are past their softmax barrier)
      if (tid < 225) ms.relpos[tid] = __ldg(a.relpos + h * 225 + tid);
      // ---- QKV epilogue: + bias -> bf16 tiles Q,K (K-major) and V (row-major = MN-major B); 16-col groups alternate halves ----
      mbar_wait(smem_u32(&ms.bar_qkv_full), ph);
      tc_fence_after();
      {
        const float* bq = a.bqkv + h * Cfg::QKV_N;
        constexpr int NG = Cfg::QKV_N / 16;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if ((g & 1) != hf) continue;
          const int col = g * 16;                      // column inside D_qkv
          const int part = col / HD, c0 = col % HD;    // 0:q 1:k 2:v, channel inside the head
          const uint32_t tile_s = (part == 0 ? sQ : part == 1 ? sK : sV);
          uint32_t v[16];
          tmem_ld16(tb + lane_base + Cfg::T_WORK + col, v);
          tmem_wait_ld();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bq + col + j));
            f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          sts128(tile_s + swz<Cfg::SWH>(r, c0 * 2), pack8(f));
          sts128(tile_s + swz<Cfg::SWH>(r, c0 * 2 + 16), pack8(f + 8));
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_qkv_staged));

      // ---- softmax: this thread owns keys [hf*32, hf*32+32) of its row ----
      mbar_wait(smem_u32(&ms.bar_s_full), ph);
      tc_fence_after();
      float s[32];
      {
        uint32_t v[32];
        tmem_ld32(tb + lane_base + Cfg::T_WORK + wl * 64 + hf * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = hf * 32 + j;                   // hf is warp-uniform
        s[j] += lds32f(relpos_s + (rp_base - (key >> 3) * 15 - (key & 7)) * 4);
      }
      if (ms.win_mixed[wl]) {
        const uint8_t myreg = ms.region[r];
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] += (ms.region[wl * 64 + hf * 32 + j] != myreg) ? -100.0f : 0.0f;
      }
      if (a.mask != nullptr && tok >= 0) {
        const int w = tile * 2 + wl;
        const float* mrow = a.mask + ((size_t)(w % a.n_mask_windows) * 64 + i) * 64 + hf * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + j));
          s[j] += m4.x; s[j + 1] += m4.y; s[j + 2] += m4.z; s[j + 3] += m4.w;
        }
      }
      float mx = s[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) mx = fmaxf(mx, s[j]);
      ms.xmax[hf][r] = mx;
      worker_bar8();
      mx = fmaxf(mx, ms.xmax[hf ^ 1][r]);
      float sum = 0.f;
      const float mxs = mx * kLog2e;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = exp2f(fmaf(s[j], kLog2e, -mxs));
        sum += s[j];
      }
      ms.xsum[hf][r] = sum;
      // P (unnormalised, bf16) over the S columns: this row's keys at packed cols [wl*32 + hf*16, +16), zeros in the other window's cols
      {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(s[2 * j], s[2 * j + 1]);
        tmem_st16(tb + lane_base + Cfg::T_WORK + wl * 32 + hf * 16, pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
        tmem_st16(tb + lane_base + Cfg::T_WORK + (1 - wl) * 32 + hf * 16, pk);
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_p_ready));

      // ---- O epilogue: normalise, park as bf16 A operand of the projection (16-col groups alternate halves) ----
      mbar_wait(smem_u32(&ms.bar_o_full), ph);
      tc_fence_after();
      if (hf * 16 < HD) {
        const float inv = 1.0f / (ms.xsum[0][r] + ms.xsum[1][r]);
        const int c0 = hf * 16;
        uint32_t v[16];
        tmem_ld16(tb + lane_base + Cfg::T_DO + c0, v);
        tmem_wait_ld();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
        tmem_st8(tb + lane_base + Cfg::T_OALL + (h * HD + c0) / 2, pk);
        tmem_wait_st();
      }
      tc_fence_before();
    }
    mbar_arrive(smem_u32(&ms.bar_oall));

    // ---- projection epilogue: + bias + shortcut, scatter to the (un-rolled) token positions ----
    bf16* __restrict__ outp = reinterpret_cast<bf16*>(a.out);
    const bf16* __restrict__ resid = reinterpret_cast<const bf16*>(a.resid);
    for (int nc = 0; nc < Cfg::NC; ++nc) {
      const int buf = nc & 1;
      mbar_wait(smem_u32(&ms.bar_d_full[buf]), (nc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = hf * 16; c0 < Cfg::NCH; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(tb + lane_base + Cfg::T_WORK + buf * 128 + c0, v);
        tmem_wait_ld();
        if (tok >= 0) {
          const int col = nc * Cfg::NCH + c0;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bproj + col + j));
            f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          if (resid != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(resid + (size_t)tok * C + col);
            float g[8];
            unpack8(__ldg(rp), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += g[j];
            unpack8(__ldg(rp + 1), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[8 + j] += g[j];
          }
          uint4* op = reinterpret_cast<uint4*>(outp + (size_t)tok * C + col);
          op[0] = pack8(f);
          op[1] = pack8(f + 8);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
    }
  }
  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

}  // namespace lw
