// wmsa16.cuh — fused shifted-window attention for 16 x 16 windows (BASELINE configs[4]; WindowAttention / LeWinTransformerBlock
// are generic in win_size, model.py:453-492, :863-865).
//
// One CTA = ONE window = 256 tokens = two 128-row M tiles; the arithmetic chain is wmsa.cuh's (model.py:951-986), re-tiled:
//   gather (roll + window_partition folded into addresses) -> LayerNorm -> + modulator (256 x C) -> bf16 A operand, both M tiles
//   per head:  for each M tile: QKV GEMM (tcgen05, D in TMEM) -> + bias -> Q, K (K-major) and V (row-major) tiles of 256 rows
//              for each M tile: S[128 x 256] = Q_mt K^T (one N = 256 UMMA chain) -> + relative-position bias (961-entry table)
//                 (+ shift-region / explicit mask) -> softmax over 256 keys in two passes over 64-key TMEM chunks -> P (bf16) back
//                 to TMEM as the A operand -> O = P V with K = 256 (V consumed MN-major) -> O / rowsum parked in TMEM (bf16)
//   output projection per M tile with A = O resident in TMEM -> + bias + shortcut -> coalesced scatter (window_reverse + roll).
// TMEM (512 columns): [0, C) O of all heads for both M tiles | [C, C+256) S / P / D_qkv / D_out | D_o sits in the upper half of
// the S region (free once the softmax has read it).  The steps of a head are serial (every accumulator shares the S region);
// one CTA per SM.  Supported: C <= 256, head_dim 16 / 32 / 64, except (C = 256, head_dim 64) whose tiles exceed shared memory.
#pragma once
#include "lewin_common.cuh"
#include "../../include/lewin_b200.h"

namespace lw {

template <int C, int HD>
struct Wmsa16Cfg {
  static constexpr int NH = C / HD;
  static constexpr int KB = (C + 63) / 64;
  static constexpr int KS = (C >= 64) ? 4 : C / 16;         // UMMA K steps per k-block
  static constexpr int SWH = 2 * HD;
  static constexpr int QKV_N = 3 * HD;
  static constexpr int QKV_CHUNK_BYTES = QKV_N * 128;
  static constexpr int NCH = C < 128 ? C : 128;
  static constexpr int NC = C / NCH;
  static constexpr int PROJ_CHUNK_BYTES = NCH * 128;
  static constexpr int STAGES = 2;
  static constexpr int STAGE_BYTES = (HD == 64) ? 24576 : kStageBytes;
  static constexpr int T_OALL = 0;                          // [mt][C/2] packed bf16 columns
  static constexpr int T_S = (C < 32) ? 32 : C;             // 256 columns
  static constexpr int T_DO = T_S + 128;
  static constexpr int T_ALLOC = 512;
  static_assert(T_S + 256 <= 512, "TMEM budget");
  static constexpr int XT_BYTES = KB * 16384;               // one M tile of the A operand
  static constexpr int S_X = 0;
  static constexpr int TILE_B = 256 * SWH;
  static constexpr int S_Q = 2 * XT_BYTES;
  static constexpr int S_K = S_Q + TILE_B;
  static constexpr int S_V = S_K + TILE_B;
  static constexpr int S_RING = S_V + TILE_B;
  static constexpr int S_MISC = S_RING + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = S_MISC + 12288 + 1024;
  static constexpr int PITCH = NCH * 2 + 16;
  static_assert(128 * PITCH <= 2 * XT_BYTES, "staging tile must fit in the dead A operand");
  static_assert(S_Q % 1024 == 0 && S_K % 1024 == 0 && S_V % 1024 == 0 && S_RING % 1024 == 0, "operand alignment");
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

struct Wmsa16Misc {
  float relpos[2][968];    // 31 x 31 bias table of the current / next head
  int row_tok[256];
  uint8_t region[256];
  float bqkv[2][192];
  int win_mixed;
  uint64_t bar_full[2], bar_empty[2];
  uint64_t bar_xn, bar_qkv_full, bar_qkv_staged, bar_s_full, bar_p_ready, bar_o_full, bar_o_done;
  uint64_t bar_d_full[2], bar_d_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Wmsa16Misc) <= 12288, "misc too large");

template <int C, int HD>
__global__ void __launch_bounds__(kThreads8, 1) wmsa16_kernel(const lw_wmsa_args a) {
  using Cfg = Wmsa16Cfg<C, HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Wmsa16Misc& ms = *reinterpret_cast<Wmsa16Misc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int win = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    mbar_init(smem_u32(&ms.bar_xn), kWorkers8);
    mbar_init(smem_u32(&ms.bar_qkv_full), 1);
    mbar_init(smem_u32(&ms.bar_qkv_staged), kWorkers8);
    mbar_init(smem_u32(&ms.bar_s_full), 1);
    mbar_init(smem_u32(&ms.bar_p_ready), kWorkers8);
    mbar_init(smem_u32(&ms.bar_o_full), 1);
    mbar_init(smem_u32(&ms.bar_o_done), kWorkers8);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_d_full[i]), 1); mbar_init(smem_u32(&ms.bar_d_empty[i]), kWorkers8); }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&ms.tmem_base), Cfg::T_ALLOC);
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const uint32_t sX = smem_u32(smem + Cfg::S_X), sQ = smem_u32(smem + Cfg::S_Q), sK = smem_u32(smem + Cfg::S_K), sV = smem_u32(smem + Cfg::S_V);

  if (warp == 8) {
    // ======================= producer: weight chunk images, in the issuer's consumption order =======================
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
      const uint8_t* wq = reinterpret_cast<const uint8_t*>(a.wqkv_img);
      for (int h = 0; h < Cfg::NH; ++h)
        for (int mt = 0; mt < 2; ++mt)
          for (int kb = 0; kb < Cfg::KB; ++kb)
            ring.load(wq + (size_t)(h * Cfg::KB + kb) * Cfg::QKV_CHUNK_BYTES, Cfg::QKV_CHUNK_BYTES);
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(a.wproj_img);
      for (int mt = 0; mt < 2; ++mt)
        for (int nc = 0; nc < Cfg::NC; ++nc)
          for (int kb = 0; kb < Cfg::KB; ++kb)
            ring.load(wp + (size_t)(nc * Cfg::KB + kb) * Cfg::PROJ_CHUNK_BYTES, Cfg::PROJ_CHUNK_BYTES);
    }
  } else if (warp == 9) {
    // ======================= issuer =======================
    Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
    constexpr uint32_t idesc_qkv = make_idesc_bf16(128, Cfg::QKV_N);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 256);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
    constexpr uint32_t idesc_proj = make_idesc_bf16(128, Cfg::NCH);
    const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
    const uint64_t b_desc0 = kmajor_desc<128>(ring_base);
    int n_staged = 0, n_pready = 0, n_odone = 0;              // completed phases of the worker -> issuer barriers
    mbar_wait(smem_u32(&ms.bar_xn), 0);
    tc_fence_after();
    for (int h = 0; h < Cfg::NH; ++h) {
      for (int mt = 0; mt < 2; ++mt) {
        // the S region is free: the previous QKV accumulator was drained (qkv_staged) / the previous head's last O was read (o_done)
        if (mt == 1) { mbar_wait(smem_u32(&ms.bar_qkv_staged), n_staged & 1); ++n_staged; tc_fence_after(); }
        else if (h > 0) { mbar_wait(smem_u32(&ms.bar_o_done), n_odone & 1); ++n_odone; tc_fence_after(); }
        const uint64_t a_desc0 = kmajor_desc<128>(sX + mt * Cfg::XT_BYTES);
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          const uint64_t ad = a_desc0 + (uint64_t)(kb * 1024), bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < Cfg::KS; ++ks) umma_ss(tb + Cfg::T_S, ad + 2 * ks, bd + 2 * ks, idesc_qkv, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_qkv_full));
        __syncwarp();
      }
      for (int mt = 0; mt < 2; ++mt) {
        if (mt == 0) { mbar_wait(smem_u32(&ms.bar_qkv_staged), n_staged & 1); ++n_staged; }      // Q, K, V of the whole window staged
        else { mbar_wait(smem_u32(&ms.bar_o_done), n_odone & 1); ++n_odone; }                    // D_o / P of M tile 0 consumed
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_ss(tb + Cfg::T_S, kmajor_desc<Cfg::SWH>(sQ + mt * 128 * Cfg::SWH + ks * 32), kmajor_desc<Cfg::SWH>(sK + ks * 32), idesc_s, ks != 0);
          umma_commit(smem_u32(&ms.bar_s_full));
        }
        __syncwarp();
        mbar_wait(smem_u32(&ms.bar_p_ready), n_pready & 1); ++n_pready;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_ts(tb + Cfg::T_DO, tb + Cfg::T_S + ks * 8, mnmajor_desc<Cfg::SWH>(sV + ks * 16 * Cfg::SWH, 8 * Cfg::SWH), idesc_pv, ks != 0);
          umma_commit(smem_u32(&ms.bar_o_full));
        }
        __syncwarp();
      }
    }
    // output projection: D_out[128 x NCH] = O_all[mt] (TMEM) * Wp^T per N chunk, two accumulator buffers in the S region
    mbar_wait(smem_u32(&ms.bar_o_done), n_odone & 1); ++n_odone;
    tc_fence_after();
    for (int mt = 0; mt < 2; ++mt)
      for (int nc = 0; nc < Cfg::NC; ++nc) {
        const int idx = mt * Cfg::NC + nc, buf = idx & 1;
        mbar_wait(smem_u32(&ms.bar_d_empty[buf]), ((idx >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          const uint64_t bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < Cfg::KS; ++ks)
              umma_ts(tb + Cfg::T_S + buf * 128, tb + Cfg::T_OALL + mt * (C / 2) + kb * 32 + ks * 8, bd + 2 * ks, idesc_proj, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full[buf]));
        __syncwarp();
      }
  } else {
    // ======================= workers (8 warps): warp w owns rows row16 .. +16 of BOTH M tiles =======================
    const int row16 = (warp & 3) * 32 + (warp >> 2) * 16;
    const int t4 = lane >> 2, tq = lane & 3;
    const uint32_t tl = (uint32_t)row16 << 16;
    // ---- source token of every window row (natural order i = y*16 + x) + region id of the shift mask ----
    {
      const int i = tid;
      int tok = -1;
      uint8_t reg = 0;
      if (a.windowed) {
        tok = win * 256 + i;
      } else {
        const int nwx = a.W >> 4, nwy = a.H >> 4;
        const int b = win / (nwx * nwy), wi = win % (nwx * nwy);
        const int ry = (wi / nwx) * 16 + (i >> 4), rx = (wi % nwx) * 16 + (i & 15);     // rolled coordinates
        int y = ry + a.shift, x = rx + a.shift;
        if (y >= a.H) y -= a.H;
        if (x >= a.W) x -= a.W;
        tok = (b * a.H + y) * a.W + x;
        if (a.shift > 0) {
          const int gy = (ry >= a.H - 16) + (ry >= a.H - a.shift);
          const int gx = (rx >= a.W - 16) + (rx >= a.W - a.shift);
          reg = (uint8_t)(3 * gy + gx);
        }
      }
      ms.row_tok[i] = tok;
      ms.region[i] = reg;
      if (i == 0) ms.win_mixed = 0;
    }
    worker_bar8();
    if (a.shift > 0 && !a.windowed && ms.region[tid] != ms.region[0]) ms.win_mixed = 1;
    // ---- A operand: LN(x) + modulator, both M tiles ----
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const float* mod = a.modulator ? a.modulator + (size_t)mt * 128 * C : nullptr;
      if (a.x_fp32) stage_rows_ln<C, 8, true, 127>(smem + Cfg::S_X + mt * Cfg::XT_BYTES, a.x, ms.row_tok + mt * 128, a.ln_w, a.ln_b, a.ln_eps, mod);
      else stage_rows_ln<C, 8, false, 127>(smem + Cfg::S_X + mt * Cfg::XT_BYTES, a.x, ms.row_tok + mt * 128, a.ln_w, a.ln_b, a.ln_eps, mod);
    }
    fence_async_smem();
    mbar_arrive(smem_u32(&ms.bar_xn));
    if (tid < Cfg::QKV_N) ms.bqkv[0][tid] = __ldg(a.bqkv + tid);
    for (int i = tid; i < 961; i += kWorkers8) ms.relpos[0][i] = __ldg(a.relpos + i);
    worker_bar8();
    constexpr int NBH = HD / 8;
    int n_qkv = 0, n_s = 0, n_o = 0;                           // completed phases of the issuer -> worker barriers
    const bool mixed_win = ms.win_mixed != 0;

    for (int h = 0; h < Cfg::NH; ++h) {
      const bool nxt = h + 1 < Cfg::NH;
      const uint32_t bqkv_s = smem_u32(&ms.bqkv[h & 1][0]), relpos_s = smem_u32(&ms.relpos[h & 1][0]);
      // ---- QKV epilogue per M tile: + bias -> bf16 -> stmatrix into the Q, K (K-major) and V (row-major) tiles ----
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        mbar_wait(smem_u32(&ms.bar_qkv_full), n_qkv & 1); ++n_qkv;
        tc_fence_after();
        const int m = lane >> 3, rr = lane & 7;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          uint32_t v[4 * NBH];
          if (NBH == 8) tmem_ld_16x256b_x8(tb + tl + Cfg::T_S + part * HD, v);
          else if (NBH == 4) tmem_ld_16x256b_x4(tb + tl + Cfg::T_S + part * HD, v);
          else tmem_ld_16x256b_x2(tb + tl + Cfg::T_S + part * HD, v);
          f2 bb[NBH];
#pragma unroll
          for (int i = 0; i < NBH; ++i) {
            const float2 b2 = lds64f(bqkv_s + (part * HD + 8 * i + 2 * tq) * 4);
            bb[i] = f2_pack(b2.x, b2.y);
          }
          tmem_wait_ld();
          uint32_t pk[2 * NBH];
          frag_bias_act_pack<NBH, false>(v, bb, pk);
          const uint32_t tile_s = (part == 0 ? sQ : part == 1 ? sK : sV) + mt * 128 * Cfg::SWH;
          const int row = row16 + (m & 1) * 8 + rr;
#pragma unroll
          for (int i2 = 0; i2 < NBH / 2; ++i2)
            stsm_x4(tile_s + swz<Cfg::SWH>(row, (2 * i2 + (m >> 1)) * 16), pk[4 * i2], pk[4 * i2 + 1], pk[4 * i2 + 2], pk[4 * i2 + 3]);
        }
        if (mt == 1 && nxt) {                                  // tables of the next head (their readers, head h-1, are long done)
          if (tid < Cfg::QKV_N) ms.bqkv[(h + 1) & 1][tid] = __ldg(a.bqkv + (h + 1) * Cfg::QKV_N + tid);
          for (int i = tid; i < 961; i += kWorkers8) ms.relpos[(h + 1) & 1][i] = __ldg(a.relpos + (h + 1) * 961 + i);
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(smem_u32(&ms.bar_qkv_staged));
      }
      // ---- attention per M tile ----
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = mt * 128 + row16 + t4, r1 = r0 + 8;    // this thread's two window rows
        const int rp0 = ((r0 >> 4) + 15) * 31 + (r0 & 15) + 15, rp1 = ((r1 >> 4) + 15) * 31 + (r1 & 15) + 15;
        const uint8_t g0 = ms.region[r0], g1 = ms.region[r1];
        const float* m0 = nullptr;
        const float* m1 = nullptr;
        if (a.mask != nullptr) {
          m0 = a.mask + ((size_t)(win % a.n_mask_windows) * 256 + r0) * 256;
          m1 = a.mask + ((size_t)(win % a.n_mask_windows) * 256 + r1) * 256;
        }
        mbar_wait(smem_u32(&ms.bar_s_full), n_s & 1); ++n_s;
        tc_fence_after();
        // scores of key chunk c (64 keys) for rows r0 / r1: s0/s1[2b+e] = key 64c + 8b + 2tq + e
        auto load_scores = [&](int c, float* s0, float* s1) {
          uint32_t v[32];
          tmem_ld_16x256b_x8(tb + tl + Cfg::T_S + c * 64, v);
          tmem_wait_ld();
#pragma unroll
          for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = c * 64 + 8 * b + 2 * tq + e;
              const int koff = (key >> 4) * 31 + (key & 15);
              s0[2 * b + e] = __uint_as_float(v[4 * b + e]) + lds32f(relpos_s + (rp0 - koff) * 4);
              s1[2 * b + e] = __uint_as_float(v[4 * b + 2 + e]) + lds32f(relpos_s + (rp1 - koff) * 4);
            }
          if (mixed_win) {
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const uint8_t gk = ms.region[c * 64 + 8 * b + 2 * tq + e];
                s0[2 * b + e] += (gk != g0) ? -100.0f : 0.0f;
                s1[2 * b + e] += (gk != g1) ? -100.0f : 0.0f;
              }
          }
          if (m0 != nullptr) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              const float2 a0 = __ldg(reinterpret_cast<const float2*>(m0 + c * 64 + 8 * b + 2 * tq));
              const float2 a1 = __ldg(reinterpret_cast<const float2*>(m1 + c * 64 + 8 * b + 2 * tq));
              s0[2 * b] += a0.x; s0[2 * b + 1] += a0.y;
              s1[2 * b] += a1.x; s1[2 * b + 1] += a1.y;
            }
          }
        };
        // pass 1: row maxima
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float s0[16], s1[16];
          load_scores(c, s0, s1);
#pragma unroll
          for (int j = 0; j < 16; ++j) { mx0 = fmaxf(mx0, s0[j]); mx1 = fmaxf(mx1, s1[j]); }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float ms0 = mx0 * kLog2e, ms1 = mx1 * kLog2e;
        // pass 2: exp, row sums, P (unnormalised bf16) over the S columns already consumed: chunk c -> packed columns [32c, 32c+32)
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float s0[16], s1[16];
          load_scores(c, s0, s1);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s0[j] = exp2_approx(fmaf(s0[j], kLog2e, -ms0)); sum0 += s0[j];
            s1[j] = exp2_approx(fmaf(s1[j], kLog2e, -ms1)); sum1 += s1[j];
          }
          uint32_t pk[16];
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            pk[2 * b] = pack_bf16(s0[2 * b], s0[2 * b + 1]);
            pk[2 * b + 1] = pack_bf16(s1[2 * b], s1[2 * b + 1]);
          }
          tmem_st_16x128b_x8(tb + tl + Cfg::T_S + c * 32, pk);
        }
        tmem_wait_st();
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        tc_fence_before();
        mbar_arrive(smem_u32(&ms.bar_p_ready));

        // ---- O epilogue: normalise, park as the bf16 A operand of the projection ----
        mbar_wait(smem_u32(&ms.bar_o_full), n_o & 1); ++n_o;
        tc_fence_after();
        {
          uint32_t v[4 * NBH];
          if (NBH == 8) tmem_ld_16x256b_x8(tb + tl + Cfg::T_DO, v);
          else if (NBH == 4) tmem_ld_16x256b_x4(tb + tl + Cfg::T_DO, v);
          else tmem_ld_16x256b_x2(tb + tl + Cfg::T_DO, v);
          tmem_wait_ld();
          const float i0 = 1.0f / sum0, i1 = 1.0f / sum1;
          uint32_t pk[2 * NBH];
#pragma unroll
          for (int i = 0; i < NBH; ++i) {
            pk[2 * i] = pack_bf16(__uint_as_float(v[4 * i]) * i0, __uint_as_float(v[4 * i + 1]) * i0);
            pk[2 * i + 1] = pack_bf16(__uint_as_float(v[4 * i + 2]) * i1, __uint_as_float(v[4 * i + 3]) * i1);
          }
          const uint32_t to = tb + tl + Cfg::T_OALL + mt * (C / 2) + (h * HD) / 2;
          if (NBH == 8) tmem_st_16x128b_x8(to, pk);
          else if (NBH == 4) tmem_st_16x128b_x4(to, pk);
          else tmem_st_16x128b_x2(to, pk);
          tmem_wait_st();
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&ms.bar_o_done));
      }
    }

    // ---- projection epilogue per (M tile, N chunk): + bias -> bf16 -> staging tile (the dead A operand) -> coalesced scatter
    // to the (un-rolled) token positions with the shortcut added on the way ----
    bf16* __restrict__ outp = reinterpret_cast<bf16*>(a.out);
    const bf16* __restrict__ resid = reinterpret_cast<const bf16*>(a.resid);
    const bool mixed = (a.x_fp32 | a.out_fp32) != 0 || a.out_b != nullptr;
    constexpr int NCH_LOG2 = Cfg::NCH == 128 ? 7 : Cfg::NCH == 64 ? 6 : Cfg::NCH == 32 ? 5 : 4;
    constexpr int NBP = (Cfg::NCH >= 64) ? 8 : Cfg::NCH / 8;
    const uint32_t stage_s = sX;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll 1
      for (int nc = 0; nc < Cfg::NC; ++nc) {
        const int idx = mt * Cfg::NC + nc, buf = idx & 1;
        mbar_wait(smem_u32(&ms.bar_d_full[buf]), (idx >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < Cfg::NCH; c0 += 8 * NBP) {
          uint32_t v[4 * NBP];
          const uint32_t ta = tb + tl + Cfg::T_S + buf * 128 + c0;
          if (NBP == 8) tmem_ld_16x256b_x8(ta, v); else if (NBP == 4) tmem_ld_16x256b_x4(ta, v); else tmem_ld_16x256b_x2(ta, v);
          f2 bb[NBP];
#pragma unroll
          for (int i = 0; i < NBP; ++i) {
            const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.bproj + nc * Cfg::NCH + c0 + 8 * i + 2 * tq));
            bb[i] = f2_pack(b2.x, b2.y);
          }
          tmem_wait_ld();
          uint32_t pk[2 * NBP];
          frag_bias_act_pack<NBP, false>(v, bb, pk);
          stage_frag<NBP>(stage_s, Cfg::PITCH, row16, c0, pk);
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
        worker_bar8();
        if (mixed) store_staged_rows_mixed<kWorkers8>(stage_s, Cfg::PITCH, NCH_LOG2, smem_u32(ms.row_tok + mt * 128), a.out, a.resid,
                                                       reinterpret_cast<bf16*>(a.out_b), a.x_fp32 != 0, a.out_fp32 != 0, (size_t)C, nc * Cfg::NCH, tid);
        else store_staged_rows(stage_s, Cfg::PITCH, NCH_LOG2, ms.row_tok + mt * 128, outp, resid, (size_t)C, nc * Cfg::NCH, tid, kWorkers8);
        worker_bar8();
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

}  // namespace lw
