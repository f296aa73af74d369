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
  static constexpr int STAGES = (C == 128) ? 2 : 4;   // C=128: 2 stages so two CTAs fit per SM; C=512: four 16 KB stages just fit (227 KB)
  static constexpr int STAGE_BYTES = (HD == 64) ? 24576 : kStageBytes; // head_dim 64: a (head, k-block) QKV chunk is 192 rows
  // TMEM columns
  static constexpr int T_OALL = 0;                         // O for all heads, bf16 packed: C/2 cols
  static constexpr int T_WORK = (C / 2 < 32) ? 32 : C / 2; // S / P (128 cols); also D_out buffer 0
  // head_dim 64: D_qkv is 192 columns wide; aliased onto S (128 columns) it reaches 64 columns further, so D_o sits behind it
  static constexpr int T_DO = T_WORK + (HD == 64 ? 192 : 128);   // D_o, HD cols
  // C >= 256 (1 CTA/SM anyway): D_qkv gets its own 96 columns so the issuer can run the QKV GEMM of head h+1
  // underneath the softmax of head h.  C <= 128 (and head_dim 64, whose 192 columns do not fit beside the rest): D_qkv
  // aliases S so that two CTAs fit in TMEM.
  static constexpr bool PIPE = (C >= 256) && (HD <= 32);
  static constexpr int T_QKV = PIPE ? T_DO + 32 : T_WORK;
  static constexpr int T_NEED = (HD == 64) ? T_WORK + 256 : T_WORK + ((NC > 1 || PIPE) ? 256 : 160);
  static_assert(T_NEED <= 512, "TMEM budget");
  static constexpr int T_ALLOC = T_NEED <= 256 ? 256 : 512;
  // shared memory map (bytes)
  static constexpr int S_X = 0;
  static constexpr int S_Q = KB * 16384;
  static constexpr int TILE_B = 128 * SWH;
  static constexpr int S_K = S_Q + TILE_B;
  static constexpr int S_V = S_K + TILE_B;
  static constexpr int S_RING = S_V + TILE_B;
  static constexpr int S_MISC = S_RING + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = S_MISC + 5120 + 1024;  // + slack for 1024 B alignment
};

struct WmsaMisc {
  float relpos[2][232];    // bias table of the current / next head
  int row_tok[128];
  uint8_t region[128];
  int win_mixed[2];
  float bqkv[2][192];      // q|k|v bias of the current / next head (written one head ahead)
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_xn, bar_qkv_full, bar_qkv_staged, bar_s_full, bar_p_ready, bar_o_full, bar_oall;
  uint64_t bar_d_full[2], bar_d_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(WmsaMisc) <= 5120, "misc too large");

template <int C, int HD>
__global__ void __launch_bounds__(kThreads8, (C <= 128 && HD <= 32) ? 2 : 1) wmsa_kernel(const lw_wmsa_args a) {
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
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;

  const uint32_t sX = smem_u32(smem + Cfg::S_X), sQ = smem_u32(smem + Cfg::S_Q), sK = smem_u32(smem + Cfg::S_K),
                 sV = smem_u32(smem + Cfg::S_V);

  if (warp == 8) {
    // ======================= producer: weight chunk images =======================
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
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
    // ======================= issuer: all tcgen05.mma (warp-uniform; one elected lane issues) =======================
    {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
      constexpr uint32_t idesc_qkv = make_idesc_bf16(128, Cfg::QKV_N);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
      constexpr uint32_t idesc_proj = make_idesc_bf16(128, Cfg::NCH);
      const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
      const uint64_t a_desc0 = kmajor_desc<128>(sX), b_desc0 = kmajor_desc<128>(ring_base);   // address field += bytes >> 4
      auto issue_qkv = [&]() {      // D_qkv[128 x 3HD] = Xn * Wqkv_h^T for the next head in the weight stream
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          constexpr int KS = (C >= 64) ? 4 : C / 16;
          const uint64_t ad = a_desc0 + (uint64_t)(kb * 1024), bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) umma_ss(tb + Cfg::T_QKV, ad + 2 * ks, bd + 2 * ks, idesc_qkv, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_qkv_full));
        __syncwarp();
      };
      mbar_wait(smem_u32(&ms.bar_xn), 0);
      tc_fence_after();
      if (Cfg::PIPE) issue_qkv();
      for (int h = 0; h < Cfg::NH; ++h) {
        const uint32_t ph = h & 1;
        if (!Cfg::PIPE) {
          if (h > 0) { mbar_wait(smem_u32(&ms.bar_o_full), ph ^ 1); tc_fence_after(); }  // PV(h-1) done: S/P columns reusable
          issue_qkv();
        }
        // --- S[128 x 128] = Q_h K_h^T (block-diagonal 64x64 halves are the two windows) ---
        mbar_wait(smem_u32(&ms.bar_qkv_staged), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t ad = kmajor_desc<Cfg::SWH>(sQ + ks * 32);
            const uint64_t bd = kmajor_desc<Cfg::SWH>(sK + ks * 32);
            umma_ss(tb + Cfg::T_WORK, ad, bd, idesc_s, ks != 0);
          }
          umma_commit(smem_u32(&ms.bar_s_full));
        }
        __syncwarp();
        // QKV GEMM of the next head runs while the workers do this head's softmax (D_qkv(h) was drained before qkv_staged)
        if (Cfg::PIPE && h + 1 < Cfg::NH) issue_qkv();
        // --- D_o[128 x HD] = P[128 x 128 keys] (TMEM, bf16) * V_h (MN-major tile) ---
        mbar_wait(smem_u32(&ms.bar_p_ready), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bd = mnmajor_desc<Cfg::SWH>(sV + ks * 16 * Cfg::SWH, 8 * Cfg::SWH);
            umma_ts(tb + Cfg::T_DO, tb + Cfg::T_WORK + ks * 8, bd, idesc_pv, ks != 0);
          }
          umma_commit(smem_u32(&ms.bar_o_full));
        }
        __syncwarp();
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
          const uint64_t bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              umma_ts(tb + Cfg::T_WORK + buf * 128, tb + Cfg::T_OALL + kb * 32 + ks * 8, bd + 2 * ks, idesc_proj, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full[buf]));
        __syncwarp();
      }
    }
  } else {
    // ======================= workers (8 warps) =======================
    // warp w owns the 16 tile rows row16 = (w&3)*32 + (w>>2)*16 .. +16 (TMEM lanes) for every per-head step.
    // TMEM is read with the 16x256b shape: thread t holds rows row16 + t/4 (+8) and column pairs 8i + 2(t%4).
    const int row16 = (warp & 3) * 32 + (warp >> 2) * 16;
    const int wl = row16 >> 6;                       // window of this warp's rows (warp-uniform)
    const int t4 = lane >> 2, tq = lane & 3;
    const int r0 = row16 + t4, r1 = r0 + 8;          // this thread's two rows
    const uint32_t tl = (uint32_t)row16 << 16;       // TMEM lane field
    // ---- source token of every row + region id for the shift mask (one thread per row) ----
    if (tid < 128) {
      const int r = tid, wlr = r >> 6, i = r & 63;
      const int w = tile * 2 + wlr;
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
      if (i == 0) ms.win_mixed[wlr] = 0;
    }
    worker_bar8();
    if (tid < 128 && a.shift > 0 && !a.windowed && ms.region[tid] != ms.region[(tid >> 6) * 64]) ms.win_mixed[tid >> 6] = 1;
    // ---- A operand: LN(x) + modulator ----
    if (a.x_fp32) stage_rows_ln<C, 8, true>(smem + Cfg::S_X, a.x, ms.row_tok, a.ln_w, a.ln_b, a.ln_eps, a.modulator);
    else stage_rows_ln<C, 8, false>(smem + Cfg::S_X, xin, ms.row_tok, a.ln_w, a.ln_b, a.ln_eps, a.modulator);
    fence_async_smem();
    mbar_arrive(smem_u32(&ms.bar_xn));

    // tables of head 0 (later heads are fetched one head ahead, see the head loop)
    if (tid < Cfg::QKV_N) ms.bqkv[0][tid] = __ldg(a.bqkv + tid);
    if (tid < 225) ms.relpos[0][tid] = __ldg(a.relpos + tid);
    worker_bar8();
    // relative-position index base of this thread's two rows (token i = row & 63: yi = i>>3, xi = i&7)
    const int rp0 = (((r0 & 63) >> 3) + 7) * 15 + (r0 & 7) + 7;
    const int rp1 = (((r1 & 63) >> 3) + 7) * 15 + (r1 & 7) + 7;
    constexpr int NBH = HD / 8;                      // 8-column blocks per head slice (8, 4 or 2)

    LW_TRACE_STMT(const bool trw = (a.dbg & 16) && blockIdx.x == 0 && tid == 0 && a.trace != nullptr; int tw = 0;)
    LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
    for (int h = 0; h < Cfg::NH; ++h) {
      const uint32_t ph = h & 1;
      // tables of the next head: the loads are issued here and parked in shared memory just before this head's qkv_staged
      // arrival (their latency hides under the QKV epilogue; every thread past s_full of this head sees them; the slot's
      // previous readers, head h-1, all arrived on p_ready(h-1) before anyone got here)
      const bool nxt = h + 1 < Cfg::NH;
      float rp_n = 0.f, bq_n = 0.f;
      if (nxt && tid < 225) rp_n = __ldg(a.relpos + (h + 1) * 225 + tid);
      if (nxt && tid < Cfg::QKV_N) bq_n = __ldg(a.bqkv + (h + 1) * Cfg::QKV_N + tid);
      const uint32_t bqkv_s = smem_u32(&ms.bqkv[h & 1][0]), relpos_s = smem_u32(&ms.relpos[h & 1][0]);
      // ---- QKV epilogue: + bias -> bf16 -> stmatrix into the Q,K (K-major) and V (row-major = MN-major B) tiles ----
      mbar_wait(smem_u32(&ms.bar_qkv_full), ph);
      tc_fence_after();
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)
      {
        const int m = lane >> 3, rr = lane & 7;
        auto qkv_part = [&](const uint32_t* vp, int part) {
          f2 bb[NBH];
#pragma unroll
          for (int i = 0; i < NBH; ++i) {
            const float2 b2 = lds64f(bqkv_s + (part * HD + 8 * i + 2 * tq) * 4);
            bb[i] = f2_pack(b2.x, b2.y);
          }
          uint32_t pk[2 * NBH];
          frag_bias_act_pack<NBH, false>(vp, bb, pk);
          const uint32_t tile_s = (part == 0 ? sQ : part == 1 ? sK : sV);
          const int row = row16 + (m & 1) * 8 + rr;
#pragma unroll
          for (int i2 = 0; i2 < NBH / 2; ++i2)
            stsm_x4(tile_s + swz<Cfg::SWH>(row, (2 * i2 + (m >> 1)) * 16), pk[4 * i2], pk[4 * i2 + 1], pk[4 * i2 + 2], pk[4 * i2 + 3]);
        };
        if (NBH == 8) {                    // head_dim 64: 32 registers per part, one part in flight
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            uint32_t v[4 * NBH];
            tmem_ld_16x256b_x8(tb + tl + Cfg::T_QKV + part * HD, v);
            tmem_wait_ld();
            qkv_part(v, part);
          }
        } else {
          uint32_t v[3][4 * NBH];
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            if (NBH == 4) tmem_ld_16x256b_x4(tb + tl + Cfg::T_QKV + part * HD, v[part]);
            else tmem_ld_16x256b_x2(tb + tl + Cfg::T_QKV + part * HD, v[part]);
          }
          tmem_wait_ld();
#pragma unroll
          for (int part = 0; part < 3; ++part) qkv_part(v[part], part);
        }
      }
      if (nxt && tid < 225) ms.relpos[(h + 1) & 1][tid] = rp_n;
      if (nxt && tid < Cfg::QKV_N) ms.bqkv[(h + 1) & 1][tid] = bq_n;
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_qkv_staged));
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)

      // ---- softmax over the 64 keys of this warp's 16 rows; a row lives in the 4 threads of a quad ----
      mbar_wait(smem_u32(&ms.bar_s_full), ph);
      tc_fence_after();
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)
      float sum0, sum1;
      {
        uint32_t v[32];
        tmem_ld_16x256b_x8(tb + tl + Cfg::T_WORK + wl * 64, v);
        tmem_wait_ld();
        float s0[16], s1[16];                        // row r0 / r1, keys 8b + 2tq + e  (index 2b + e)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int koff = b * 15 + 2 * tq + e;    // (key>>3)*15 + (key&7)
            s0[2 * b + e] = __uint_as_float(v[4 * b + e]) + lds32f(relpos_s + (rp0 - koff) * 4);
            s1[2 * b + e] = __uint_as_float(v[4 * b + 2 + e]) + lds32f(relpos_s + (rp1 - koff) * 4);
          }
        }
        if (ms.win_mixed[wl]) {
          const uint8_t g0 = ms.region[r0], g1 = ms.region[r1];
#pragma unroll
          for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const uint8_t gk = ms.region[wl * 64 + 8 * b + 2 * tq + e];
              s0[2 * b + e] += (gk != g0) ? -100.0f : 0.0f;
              s1[2 * b + e] += (gk != g1) ? -100.0f : 0.0f;
            }
        }
        if (a.mask != nullptr) {
          const int w = tile * 2 + wl;
          if (w < a.n_windows) {
            const float* m0 = a.mask + ((size_t)(w % a.n_mask_windows) * 64 + (r0 & 63)) * 64;
            const float* m1 = a.mask + ((size_t)(w % a.n_mask_windows) * 64 + (r1 & 63)) * 64;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              const float2 a0 = __ldg(reinterpret_cast<const float2*>(m0 + 8 * b + 2 * tq));
              const float2 a1 = __ldg(reinterpret_cast<const float2*>(m1 + 8 * b + 2 * tq));
              s0[2 * b] += a0.x; s0[2 * b + 1] += a0.y;
              s1[2 * b] += a1.x; s1[2 * b + 1] += a1.y;
            }
          }
        }
        float mx0 = s0[0], mx1 = s1[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) { mx0 = fmaxf(mx0, s0[j]); mx1 = fmaxf(mx1, s1[j]); }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float ms0 = mx0 * kLog2e, ms1 = mx1 * kLog2e;
        sum0 = 0.f; sum1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          s0[j] = exp2_approx(fmaf(s0[j], kLog2e, -ms0)); sum0 += s0[j];
          s1[j] = exp2_approx(fmaf(s1[j], kLog2e, -ms1)); sum1 += s1[j];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        // P (unnormalised bf16) as the A operand of P·V: packed key pairs -> 16x128b fragments.  This window's
        // keys go to packed columns [wl*32, +32); the other window's columns are zero (block-diagonal P).
        uint32_t pk[16];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          pk[2 * b] = pack_bf16(s0[2 * b], s0[2 * b + 1]);
          pk[2 * b + 1] = pack_bf16(s1[2 * b], s1[2 * b + 1]);
        }
        tmem_st_16x128b_x8(tb + tl + Cfg::T_WORK + wl * 32, pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
        tmem_st_16x128b_x8(tb + tl + Cfg::T_WORK + (1 - wl) * 32, pk);
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_p_ready));
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)

      // ---- O epilogue: normalise, park as bf16 A operand of the projection ----
      mbar_wait(smem_u32(&ms.bar_o_full), ph);
      tc_fence_after();
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)
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
        if (NBH == 8) tmem_st_16x128b_x8(tb + tl + Cfg::T_OALL + (h * HD) / 2, pk);
        else if (NBH == 4) tmem_st_16x128b_x4(tb + tl + Cfg::T_OALL + (h * HD) / 2, pk);
        else tmem_st_16x128b_x2(tb + tl + Cfg::T_OALL + (h * HD) / 2, pk);
        tmem_wait_st();
      }
      tc_fence_before();
      LW_TRACE_STMT(if (trw && h < 4) a.trace[tw++] = clock64();)
    }
    mbar_arrive(smem_u32(&ms.bar_oall));
    LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)

    // ---- projection epilogue: + bias -> bf16 -> staging tile (the A operand / Q,K,V tiles are dead now) ->
    // coalesced scatter to the (un-rolled) token positions with the shortcut added on the way ----
    bf16* __restrict__ outp = reinterpret_cast<bf16*>(a.out);
    const bf16* __restrict__ resid = reinterpret_cast<const bf16*>(a.resid);
    const bool mixed = (a.x_fp32 | a.out_fp32) != 0 || a.out_b != nullptr;        // fp32 residual-stream mode
    constexpr int PITCH = Cfg::NCH * 2 + 16;
    static_assert(128 * PITCH <= Cfg::S_RING, "staging tile must fit in the dead A/QKV region");
    constexpr int NCH_LOG2 = Cfg::NCH == 128 ? 7 : Cfg::NCH == 64 ? 6 : Cfg::NCH == 32 ? 5 : 4;
    constexpr int NBP = (Cfg::NCH >= 64) ? 8 : Cfg::NCH / 8;          // column blocks per TMEM load (64 / 32 / 16 columns)
    const uint32_t stage_s = sX;
    for (int nc = 0; nc < Cfg::NC; ++nc) {
      const int buf = nc & 1;
      mbar_wait(smem_u32(&ms.bar_d_full[buf]), (nc >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < Cfg::NCH; c0 += 8 * NBP) {
        uint32_t v[4 * NBP];
        const uint32_t ta = tb + tl + Cfg::T_WORK + buf * 128 + c0;
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
        stage_frag<NBP>(stage_s, PITCH, row16, c0, pk);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
      worker_bar8();
      if (mixed) store_staged_rows_mixed<kWorkers8>(stage_s, PITCH, NCH_LOG2, smem_u32(ms.row_tok), a.out, a.resid, reinterpret_cast<bf16*>(a.out_b), a.x_fp32 != 0,
                                                     a.out_fp32 != 0, (size_t)C, nc * Cfg::NCH, tid);
      else store_staged_rows(stage_s, PITCH, NCH_LOG2, ms.row_tok, outp, resid, (size_t)C, nc * Cfg::NCH, tid, kWorkers8);
      worker_bar8();
      LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
    }
    LW_TRACE_STMT(if (trw) a.trace[tw++] = -1;)
  }
  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

}  // namespace lw
