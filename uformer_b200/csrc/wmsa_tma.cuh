// wmsa_tma.cuh — persistent W-MSA kernel with a TMA window gather (token-map inputs, C <= 256, head_dim 16 / 32).
//
// Same arithmetic chain as wmsa.cuh (model.py:951-986), different plumbing:
//   * window_partition + roll (model.py:956-963) is a TMA box gather: the (C, W, H, B) tensor map of the token map is read in
//     4x4-pixel boxes of <= 64 channels (cp.async.bulk.tensor.4d), eight boxes per 128-token tile and k-block.  With shift = 4
//     every 4x4 quarter of a rolled 8x8 window lies on one side of the wrap, so the roll is a coordinate offset per box.
//     The boxes land in SWIZZLE_{128,64,32}B rows, i.e. directly in the K-major A-operand layout of the QKV GEMM: no register
//     pass between HBM and the tensor core.  Rows of a window are therefore in QUARTER-MAJOR order
//        row = 16*(2*(y>>2) + (x>>2)) + 4*(y&3) + (x&3)
//     (attention is permutation-equivariant inside a window; the bias-table index, shift regions, explicit mask and the
//     scatter use the same order).
//   * LayerNorm (norm1, model.py:953) is folded into the projection, as in the fused LeFF:
//        LN(x) Wqkv^T + b = rstd * (x Wg^T) - rstd*mean*cs + bf,   Wg = Wqkv diag(gamma) (bf16), cs = row sums of Wg,
//        bf = b + Wqkv beta.  The row statistics are computed from the landed tile while the first QKV GEMM runs.
//   * the CTA is persistent (tile = blockIdx.x + k*gridDim.x): barriers, TMEM and the weight ring are set up once, the
//     gather of tile i+1 (two buffers; C = 128 keeps one so that two CTAs fit per SM) and — for C = 256 — the QKV GEMM of
//     its first head run underneath tile i.
// The scatter (window_reverse + roll back, model.py:975-983) is the staged copy-out of wmsa.cuh with the quarter-major row
// table.
//   * the window modulator (model.py:966-969: LN(x) + m[pos] before the projection) rides on the tensor core as well: its
//     contribution m[pos] W^T is a per-position constant, so the A operand gets one more 64-wide k-block that is a one-hot of
//     the row's position scaled by sigma = 1/rstd (128 two-byte stores per tile), multiplied with the per-head image of
//     (m W^T)^T; the epilogue's rstd * acc then yields rstd*(x Wg^T) + m W^T.  No SIMT work per head.
// Window-major inputs, C = 512 and head_dim 64 stay on wmsa_kernel.
#pragma once
#include <cuda.h>
#include "lewin_common.cuh"
#include "leff_fused.cuh"   // tma_load_4d / tma_prefetch_desc

namespace lw {

struct WmsaTArgs {
  void* out;
  const void* resid;       // bf16 or fp32 (resid_fp32), or null
  bf16* out_b;             // optional bf16 copy of out
  const uint8_t* wqkv_img; // LN-folded, packed as for wmsa_kernel
  const float* bqkv;       // folded bias (heads*3*hd)
  const float* cs;         // row sums of the bf16 folded weight (heads*3*hd)
  const uint8_t* wmod_img; // per head [3hd rows x 64 positions (quarter-major)] bf16 image of (modulator W^T)^T, or null
  const uint8_t* wproj_img;
  const float* bproj;
  const float* relpos;
  const float* mask;
  int n_mask_windows, n_windows, H, W, shift;
  float ln_eps;
  int resid_fp32, out_fp32;
  int n_tiles;
  int dbg;
  long long* trace;
};

// Waits of the worker / issuer roles.  -DLW_WMSA_SPIN: non-suspending polls (mbarrier.test_wait) instead of try_wait with a
// suspend hint — an experiment knob for the per-phase hand-off latency (tools/wmsa_tma_trace.py).
__device__ __forceinline__ void wt_wait(uint32_t bar, uint32_t parity) {
#ifdef LW_WMSA_SPIN
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++spins > (1u << 26)) __trap();
  } while (!ok);
#else
  mbar_wait(bar, parity);
#endif
}

template <int C, int HD>
struct WmsaTCfg {
  static constexpr int NH = C / HD;
  static constexpr int KB = (C + 63) / 64;
  static constexpr int KS = (C >= 64) ? 4 : C / 16;          // UMMA K steps per k-block
  static constexpr int SWH = 2 * HD;
  static constexpr int QKV_N = 3 * HD;
  static constexpr int QKV_CHUNK_BYTES = QKV_N * 128;
  static constexpr int NCH = C < 128 ? C : 128;              // projection N chunk (weight image granularity)
  static constexpr int NC = C / NCH;
  static constexpr int PROJ_CHUNK_BYTES = NCH * 128;
  static constexpr int SWX = (C >= 64) ? 128 : 2 * C;        // row bytes / swizzle span of the gathered tile
  static constexpr int XKB_BYTES = 128 * SWX;
  static constexpr int X_BYTES = KB * XKB_BYTES;
  static constexpr int NXB = (C == 128) ? 1 : 2;             // gather buffers
  static constexpr int STAGES = (C == 128) ? 2 : 3;
  static constexpr int STAGE_BYTES = kStageBytes;
  static constexpr bool PIPE = (C >= 256);                   // D_qkv has its own columns: the QKV GEMM stream runs one head ahead
  static constexpr int T_OALL = 0;
  static constexpr int T_WORK = (C / 2 < 32) ? 32 : C / 2;   // S / P, and D_out of the projection
  static constexpr int T_DO = T_WORK + 128;
  static constexpr int T_QKV = PIPE ? T_DO + 32 : T_WORK;
  static constexpr int T_NEED = PIPE ? T_QKV + 96 : T_DO + 32;
  static_assert(T_NEED <= 512, "TMEM budget");
  static constexpr int T_ALLOC = T_NEED <= 256 ? 256 : 512;
  static constexpr int NCHS = (NCH < (HD == 32 ? 64 : 32)) ? NCH : (HD == 32 ? 64 : 32);   // columns per copy-out round
  static constexpr int PITCH = NCHS * 2 + 16;
  static constexpr int S_X = 0;
  static constexpr int S_Q = NXB * X_BYTES;
  static constexpr int TILE_B = 128 * SWH;
  static constexpr int S_K = S_Q + TILE_B;
  static constexpr int S_V = S_K + TILE_B;
  static constexpr int S_RING = S_V + TILE_B;
  // modulator blocks: a 128 x 64 one-hot A-operand tile (16 KB).  C = 64 has no room beside three ring stages and two CTAs per
  // SM: there the tile takes the place of the third stage (the ring then runs two deep).
  static constexpr int STAGES_MOD = (C == 64) ? 2 : STAGES;
  static constexpr int S_OH = S_RING + STAGES_MOD * STAGE_BYTES;
  static constexpr int S_MISC = (S_RING + STAGES * STAGE_BYTES > S_OH + 16384) ? S_RING + STAGES * STAGE_BYTES : S_OH + 16384;
  static constexpr int SMEM_BYTES = S_MISC + 6144 + 1024;
  static_assert(S_Q % 1024 == 0 && S_RING % 1024 == 0, "operand alignment");
  static_assert(128 * PITCH <= 3 * TILE_B, "staging tile must fit in the Q/K/V tiles");
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

struct WmsaTMisc {
  float relpos[2][232];    // bias table of the current / next head in the stream
  int row_tok[2][128];     // destination token of every tile row, this / next tile (written one tile ahead)
  uint8_t region[2][128];
  int win_mixed[2][2];
  float bqkv[2][96];       // folded q|k|v bias of the current / next head in the stream
  float csq[2][96];        // row sums of the folded weight
  float2 stats[128];       // (rstd, -mean*rstd) per tile row
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_x_full[2], bar_x_empty[2];
  uint64_t bar_qkv_full, bar_qkv_staged, bar_s_full, bar_p_ready, bar_o_full, bar_oall;
  uint64_t bar_d_full, bar_d_empty;
  uint64_t bar_oh_ready;
  uint32_t tmem_base;
};
static_assert(sizeof(WmsaTMisc) <= 6144, "misc too large");

template <int C, int HD>
__global__ void __launch_bounds__(kThreads8, C <= 128 ? 2 : 1) wmsa_tma_kernel(const __grid_constant__ CUtensorMap xmap, const WmsaTArgs a) {
  using Cfg = WmsaTCfg<C, HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WmsaTMisc& ms = *reinterpret_cast<WmsaTMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // grid <= n_tiles

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_x_full[i]), 1); mbar_init(smem_u32(&ms.bar_x_empty[i]), kWorkers8 + 1); }
    mbar_init(smem_u32(&ms.bar_qkv_full), 1);
    mbar_init(smem_u32(&ms.bar_qkv_staged), kWorkers8);
    mbar_init(smem_u32(&ms.bar_s_full), 1);
    mbar_init(smem_u32(&ms.bar_p_ready), kWorkers8);
    mbar_init(smem_u32(&ms.bar_o_full), 1);
    mbar_init(smem_u32(&ms.bar_oall), kWorkers8);
    mbar_init(smem_u32(&ms.bar_d_full), 1);
    mbar_init(smem_u32(&ms.bar_d_empty), kWorkers8);
    mbar_init(smem_u32(&ms.bar_oh_ready), kWorkers8);
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
  const int nwx = a.W >> 3, nwin_img = nwx * (a.H >> 3);
  const bool has_mod = a.wmod_img != nullptr;
  const int ring_stages = has_mod ? Cfg::STAGES_MOD : Cfg::STAGES;
  const uint32_t sOH = smem_u32(smem + Cfg::S_OH);

  if (warp == 8) {
    // ======================= producer: window gather (TMA boxes) + weight chunk images =======================
    if (lane == 0) {
      tma_prefetch_desc(&xmap);
      auto load_x = [&](int it) {
        const int buf = it % Cfg::NXB;
        const uint32_t full = smem_u32(&ms.bar_x_full[buf]);
        mbar_wait(smem_u32(&ms.bar_x_empty[buf]), ((it / Cfg::NXB) & 1) ^ 1);
        mbar_expect_tx(full, 128 * C * 2);
        const int tile = blockIdx.x + it * gridDim.x;
        const uint32_t dst0 = sX + buf * Cfg::X_BYTES;
#pragma unroll 1
        for (int wl = 0; wl < 2; ++wl) {
          int w = tile * 2 + wl;
          if (w >= a.n_windows) w = a.n_windows - 1;           // odd window count: the spare half tile re-reads a real window (finite values)
          const int b = w / nwin_img, wi = w - b * nwin_img;
          const int wy = wi / nwx, wx = wi - wy * nwx;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int y = wy * 8 + (q >> 1) * 4 + a.shift, x = wx * 8 + (q & 1) * 4 + a.shift;
            if (y >= a.H) y -= a.H;
            if (x >= a.W) x -= a.W;
#pragma unroll
            for (int kb = 0; kb < Cfg::KB; ++kb)
              tma_load_4d(dst0 + kb * Cfg::XKB_BYTES + (wl * 64 + q * 16) * Cfg::SWX, &xmap, kb * 64, x, y, b, full);
          }
        }
      };
      for (int it = 0; it < Cfg::NXB && it < n_my; ++it) load_x(it);
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), ring_stages, 0, Cfg::STAGE_BYTES};
      // chunk order = the issuer's consumption order.  PIPE: the QKV GEMM stream runs one head ahead, across tiles: head 0 of
      // tile it+1 is consumed before the projection of tile it.
      auto load_qkv = [&](int h0, int h1) {
        for (int h = h0; h < h1; ++h) {
          for (int kb = 0; kb < Cfg::KB; ++kb)
            ring.load(a.wqkv_img + (size_t)(h * Cfg::KB + kb) * Cfg::QKV_CHUNK_BYTES, Cfg::QKV_CHUNK_BYTES);
          if (has_mod) ring.load(a.wmod_img + (size_t)h * Cfg::QKV_CHUNK_BYTES, Cfg::QKV_CHUNK_BYTES);
        }
      };
      const bool ahead = Cfg::PIPE && !has_mod;          // (a modulated head 0 needs its tile's statistics: no cross-tile issue)
      if (ahead) load_qkv(0, Cfg::NH);
      for (int it = 0; it < n_my; ++it) {
        if (!ahead) load_qkv(0, Cfg::NH);
        else if (it + 1 < n_my) load_qkv(0, 1);
        if (it + Cfg::NXB < n_my) load_x(it + Cfg::NXB);      // its buffer is released by the last QKV GEMM of tile `it`
        for (int nc = 0; nc < Cfg::NC; ++nc)
          for (int kb = 0; kb < Cfg::KB; ++kb)
            ring.load(a.wproj_img + (size_t)(nc * Cfg::KB + kb) * Cfg::PROJ_CHUNK_BYTES, Cfg::PROJ_CHUNK_BYTES);
        if (ahead && it + 1 < n_my) load_qkv(1, Cfg::NH);
      }
    }
  } else if (warp == 9) {
    // ======================= issuer =======================
    Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), ring_stages, 0, Cfg::STAGE_BYTES};
    constexpr uint32_t idesc_qkv = make_idesc_bf16(128, Cfg::QKV_N);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
    constexpr uint32_t idesc_proj = make_idesc_bf16(128, Cfg::NCH);
    const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
    const uint64_t b_desc0 = kmajor_desc<128>(ring_base);
    // QKV GEMM of head hq of this CTA's tile number itq: D_qkv[128 x 3HD] = X * Wg_h^T (raw x: LayerNorm is applied in the epilogue)
    auto issue_qkv = [&](int itq, int hq) {
      const int buf = itq % Cfg::NXB;
      if (hq == 0) { wt_wait(smem_u32(&ms.bar_x_full[buf]), (itq / Cfg::NXB) & 1); tc_fence_after(); }
      const uint64_t a_desc0 = kmajor_desc<Cfg::SWX>(sX + buf * Cfg::X_BYTES);
      for (int kb = 0; kb < Cfg::KB; ++kb) {
        const uint32_t wst = ring.acquire();
        const uint64_t ad = a_desc0 + (uint64_t)(kb * (Cfg::XKB_BYTES >> 4)), bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < Cfg::KS; ++ks) umma_ss(tb + Cfg::T_QKV, ad + 2 * ks, bd + 2 * ks, idesc_qkv, (kb | ks) != 0);
        }
        __syncwarp();
        ring.release();
      }
      if (has_mod) {                                  // + sigma * onehot(pos) x (m W_h^T)^T: the modulator term, un-scaled by rstd
        if (hq == 0) { wt_wait(smem_u32(&ms.bar_oh_ready), itq & 1); tc_fence_after(); }
        const uint32_t wst = ring.acquire();
        const uint64_t ad = kmajor_desc<128>(sOH), bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss(tb + Cfg::T_QKV, ad + 2 * ks, bd + 2 * ks, idesc_qkv, 1u);
        }
        __syncwarp();
        ring.release();
      }
      if (elect_one()) {
        umma_commit(smem_u32(&ms.bar_qkv_full));
        if (hq == Cfg::NH - 1) umma_commit(smem_u32(&ms.bar_x_empty[buf]));     // the gathered tile is dead: refill
      }
      __syncwarp();
    };
    int g = 0;                                        // heads processed so far (barrier phases)
    const bool ahead = Cfg::PIPE && !has_mod;
    if (ahead) issue_qkv(0, 0);
    for (int it = 0; it < n_my; ++it) {
      if (Cfg::PIPE && !ahead) issue_qkv(it, 0);
      for (int h = 0; h < Cfg::NH; ++h, ++g) {
        const uint32_t ph = g & 1;
        if (!Cfg::PIPE) {
          // D_qkv aliases S / P / D_out: the previous user of those columns must be done
          if (h > 0) { wt_wait(smem_u32(&ms.bar_o_full), ph ^ 1); tc_fence_after(); }
          else if (it > 0) { wt_wait(smem_u32(&ms.bar_d_empty), (it * Cfg::NC - 1) & 1); tc_fence_after(); }
          issue_qkv(it, h);
        }
        wt_wait(smem_u32(&ms.bar_qkv_staged), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_ss(tb + Cfg::T_WORK, kmajor_desc<Cfg::SWH>(sQ + ks * 32), kmajor_desc<Cfg::SWH>(sK + ks * 32), idesc_s, ks != 0);
          umma_commit(smem_u32(&ms.bar_s_full));
        }
        __syncwarp();
        if (Cfg::PIPE) {                              // next QKV GEMM of the stream (next head, or head 0 of the next tile)
          if (h + 1 < Cfg::NH) issue_qkv(it, h + 1);
          else if (ahead && it + 1 < n_my) issue_qkv(it + 1, 0);
        }
        wt_wait(smem_u32(&ms.bar_p_ready), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ts(tb + Cfg::T_DO, tb + Cfg::T_WORK + ks * 8, mnmajor_desc<Cfg::SWH>(sV + ks * 16 * Cfg::SWH, 8 * Cfg::SWH), idesc_pv, ks != 0);
          umma_commit(smem_u32(&ms.bar_o_full));
        }
        __syncwarp();
      }
      // output projection: D_out[128 x NCH] = O_all (TMEM) * Wp^T per N chunk, single accumulator buffer (the S / P columns)
      wt_wait(smem_u32(&ms.bar_oall), it & 1);
      tc_fence_after();
      for (int nc = 0; nc < Cfg::NC; ++nc) {
        const int pc = it * Cfg::NC + nc;
        wt_wait(smem_u32(&ms.bar_d_empty), (pc & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          const uint64_t bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < Cfg::KS; ++ks)
              umma_ts(tb + Cfg::T_WORK, tb + Cfg::T_OALL + kb * 32 + ks * 8, bd + 2 * ks, idesc_proj, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full));
        __syncwarp();
      }
    }
  } else {
    // ======================= workers (8 warps) =======================
    const int row16 = (warp & 3) * 32 + (warp >> 2) * 16;
    const int wl = row16 >> 6;
    const int t4 = lane >> 2, tq = lane & 3;
    const int r0 = row16 + t4, r1 = r0 + 8;
    const uint32_t tl = (uint32_t)row16 << 16;
    // quarter-major row -> window coordinates
    auto row_y = [](int r) { return ((r >> 5) & 1) * 4 + ((r >> 2) & 3); };
    auto row_x = [](int r) { return ((r >> 4) & 1) * 4 + (r & 3); };
    constexpr int NBH = HD / 8;
    constexpr int NBP = Cfg::NCHS / 8;
    const uint32_t stage_s = sQ;

    if (tid < Cfg::QKV_N) { ms.bqkv[0][tid] = __ldg(a.bqkv + tid); ms.csq[0][tid] = __ldg(a.cs + tid); }
    if (tid < 225) ms.relpos[0][tid] = __ldg(a.relpos + tid);

    // destination token of every row + region id of the shift mask for this CTA's tile number itn -> table buffer itn & 1.
    // Tables are built one tile ahead, underneath the projection GEMM of the previous tile (threads 0..127, one row each).
    auto build_rows = [&](int itn) {
      const int tile_n = blockIdx.x + itn * gridDim.x, nb = itn & 1;
      const int r = tid, wlr = r >> 6;
      const int w = tile_n * 2 + wlr;
      int tok = -1;
      uint8_t reg = 0;
      if (w < a.n_windows) {
        const int b = w / nwin_img, wi = w - b * nwin_img;
        const int ry = (wi / nwx) * 8 + row_y(r), rx = (wi % nwx) * 8 + row_x(r);       // rolled coordinates
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
      ms.row_tok[nb][r] = tok;
      ms.region[nb][r] = reg;
      if ((r & 63) == 0) ms.win_mixed[nb][wlr] = 0;
    };
    auto mark_mixed = [&](int itn) {                     // after a barrier behind build_rows: does a window straddle shift regions?
      const int nb = itn & 1;
      if (a.shift > 0 && ms.region[nb][tid] != ms.region[nb][(tid >> 6) * 64]) ms.win_mixed[nb][tid >> 6] = 1;
    };
    if (has_mod) {                                       // one-hot tile: zero once, the diagonal is rewritten per tile
#pragma unroll
      for (int k = 0; k < 4; ++k) sts128(sOH + (tid * 4 + k) * 16, make_uint4(0, 0, 0, 0));
    }
    if (tid < 128) build_rows(0);
    worker_bar8();
    if (tid < 128) mark_mixed(0);
    worker_bar8();

    LW_TRACE_STMT(const bool trw = (a.dbg & 16) && blockIdx.x == 0 && tid == 0 && a.trace != nullptr; int tw = 0;)
    int g = 0;
    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int buf = it % Cfg::NXB;
      LW_TRACE_STMT(if (trw && it < 3) a.trace[tw++] = clock64();)
      const int rb = it & 1;                             // row-table buffer of this tile
      // ---- LayerNorm statistics of this warp's 16 rows from the landed tile (one shifted pass, as in leff_fused.cuh) ----
      wt_wait(smem_u32(&ms.bar_x_full[buf]), (it / Cfg::NXB) & 1);
      {
        constexpr int LPR = Cfg::SWX / 16;               // lanes per row: 8 / 4 / 2
        constexpr int RPP = 32 / LPR;                    // rows per pass: 4 / 8 / 16
        constexpr int NPASS = 16 / RPP;
        const int sub = lane % LPR, rin = lane / LPR;
        const uint32_t xs = sX + buf * Cfg::X_BYTES;
        float s1[NPASS], s2[NPASS], x0[NPASS];
#pragma unroll
        for (int u = 0; u < NPASS; ++u) {
          const int r = warp * 16 + u * RPP + rin;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int kb = 0; kb < Cfg::KB; ++kb) {
            float v[8];
            unpack8(lds128(xs + kb * Cfg::XKB_BYTES + swz<Cfg::SWX>(r, sub * 16)), v);
            if (kb == 0) x0[u] = __shfl_sync(0xffffffffu, v[0], lane - sub);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float d0 = v[i] - x0[u], d1 = v[i + 1] - x0[u];
              a0 += d0; a1 += d1;
              b0 = fmaf(d0, d0, b0); b1 = fmaf(d1, d1, b1);
            }
          }
          s1[u] = a0 + a1; s2[u] = b0 + b1;
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1)
#pragma unroll
          for (int u = 0; u < NPASS; ++u) {
            s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
            s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
          }
        if (sub == 0) {
#pragma unroll
          for (int u = 0; u < NPASS; ++u) {
            const float md = s1[u] * (1.0f / C);                             // mean - x0
            const float var = fmaxf(s2[u] * (1.0f / C) - md * md, 0.f);
            const float rstd = rsqrtf(var + a.ln_eps);
            const int row = warp * 16 + u * RPP + rin;
            ms.stats[row] = make_float2(rstd, -(x0[u] + md) * rstd);
            if (has_mod) {                               // sigma on the diagonal of the one-hot k-block (column = position in the window)
              const __nv_bfloat16 sg = __float2bfloat16_rn((var + a.ln_eps) * rstd);
              asm volatile("st.shared.b16 [%0], %1;" ::"r"(sOH + swz<128>(row, (row & 63) * 2)), "h"(*reinterpret_cast<const unsigned short*>(&sg)) : "memory");
            }
          }
        }
      }
      mbar_arrive(smem_u32(&ms.bar_x_empty[buf]));       // this thread is done reading the raw tile
      if (has_mod) { fence_async_smem(); mbar_arrive(smem_u32(&ms.bar_oh_ready)); }
      worker_bar8();
      LW_TRACE_STMT(if (trw && it < 3) a.trace[tw++] = clock64();)

      for (int h = 0; h < Cfg::NH; ++h, ++g) {
        const uint32_t ph = g & 1;
        // tables of the next head of the stream (next head, or head 0 of the next tile): the loads are issued here and parked
        // in shared memory just before this head's qkv_staged arrival, so that their latency hides under the QKV epilogue and
        // every thread that has passed s_full of this head sees them
        const bool nxt = Cfg::NH > 1 && (h + 1 < Cfg::NH || it + 1 < n_my);
        const int hn = (h + 1 < Cfg::NH) ? h + 1 : 0;
        float rp_n = 0.f, bq_n = 0.f, cs_n = 0.f;
        if (nxt && tid < 225) rp_n = __ldg(a.relpos + hn * 225 + tid);
        if (nxt && tid < Cfg::QKV_N) { bq_n = __ldg(a.bqkv + hn * Cfg::QKV_N + tid); cs_n = __ldg(a.cs + hn * Cfg::QKV_N + tid); }
        const int slot = (Cfg::NH == 1) ? 0 : (g & 1);            // a single head never changes its tables
        const uint32_t bqkv_s = smem_u32(&ms.bqkv[slot][0]), cs_s = smem_u32(&ms.csq[slot][0]), relpos_s = smem_u32(&ms.relpos[slot][0]);
        // ---- QKV epilogue: LayerNorm fold + bias -> bf16 -> Q, K (K-major) and V (row-major) tiles ----
        wt_wait(smem_u32(&ms.bar_qkv_full), ph);
        tc_fence_after();
        LW_TRACE_STMT(if (trw && it < 3 && h < 2) a.trace[tw++] = clock64();)
        {
          const int m = lane >> 3, rr = lane & 7;
          const float2 st0 = ms.stats[r0], st1 = ms.stats[r1];
          // parts q, k, v: the TMEM load of part p+1 is in flight while part p is folded and stored
          uint32_t v[2][4 * NBH];
          auto ld_part = [&](int part, uint32_t* dst) {
            if (NBH == 4) tmem_ld_16x256b_x4(tb + tl + Cfg::T_QKV + part * HD, dst);
            else tmem_ld_16x256b_x2(tb + tl + Cfg::T_QKV + part * HD, dst);
          };
          ld_part(0, v[0]);
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            tmem_wait_ld();
            if (part + 1 < 3) ld_part(part + 1, v[(part + 1) & 1]);
            const uint32_t* vp = v[part & 1];
            uint32_t pk[2 * NBH];
#pragma unroll
            for (int i = 0; i < NBH; ++i) {
              const int n = part * HD + 8 * i + 2 * tq;
              const float2 b2 = lds64f(bqkv_s + n * 4), c2 = lds64f(cs_s + n * 4);
              const f2 bbv = f2_pack(b2.x, b2.y), ccv = f2_pack(c2.x, c2.y);
              const f2 d0 = f2_pack(__uint_as_float(vp[4 * i]), __uint_as_float(vp[4 * i + 1]));
              const f2 d1 = f2_pack(__uint_as_float(vp[4 * i + 2]), __uint_as_float(vp[4 * i + 3]));
              pk[2 * i] = f2_to_bf2(f2_fma(d0, f2_pack(st0.x, st0.x), f2_fma(f2_pack(st0.y, st0.y), ccv, bbv)));
              pk[2 * i + 1] = f2_to_bf2(f2_fma(d1, f2_pack(st1.x, st1.x), f2_fma(f2_pack(st1.y, st1.y), ccv, bbv)));
            }
            const uint32_t tile_s = (part == 0 ? sQ : part == 1 ? sK : sV);
            const int row = row16 + (m & 1) * 8 + rr;
#pragma unroll
            for (int i2 = 0; i2 < NBH / 2; ++i2)
              stsm_x4(tile_s + swz<Cfg::SWH>(row, (2 * i2 + (m >> 1)) * 16), pk[4 * i2], pk[4 * i2 + 1], pk[4 * i2 + 2], pk[4 * i2 + 3]);
          }
        }
        if (nxt && tid < 225) ms.relpos[slot ^ 1][tid] = rp_n;
        if (nxt && tid < Cfg::QKV_N) { ms.bqkv[slot ^ 1][tid] = bq_n; ms.csq[slot ^ 1][tid] = cs_n; }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(smem_u32(&ms.bar_qkv_staged));
        LW_TRACE_STMT(if (trw && it < 3 && h < 2) a.trace[tw++] = clock64();)

        // ---- softmax over the 64 keys of this warp's 16 rows ----
        wt_wait(smem_u32(&ms.bar_s_full), ph);
        tc_fence_after();
        LW_TRACE_STMT(if (trw && it < 3 && h < 2) a.trace[tw++] = clock64();)
        float sum0, sum1;
        {
          uint32_t v[32];
          tmem_ld_16x256b_x8(tb + tl + Cfg::T_WORK + wl * 64, v);
          tmem_wait_ld();
          float s0[16], s1[16];
          const int kq = (tq >> 1) * 15 + 2 * (tq & 1);        // per-thread part of the key offset (key = 8b + 2tq + e)
          const int rp0 = (row_y(r0) + 7) * 15 + row_x(r0) + 7 - kq, rp1 = rp0 + 2 * 15;     // row r0 + 8: y + 2
#pragma unroll
          for (int b = 0; b < 8; ++b) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              // key 8b + 2tq + e (quarter-major): y = 4(b>>2) + 2(b&1) + (tq>>1), x = 4((b>>1)&1) + 2(tq&1) + e
              const int koff = ((b >> 2) * 4 + (b & 1) * 2) * 15 + ((b >> 1) & 1) * 4 + e;
              s0[2 * b + e] = __uint_as_float(v[4 * b + e]) + lds32f(relpos_s + (rp0 - koff) * 4);
              s1[2 * b + e] = __uint_as_float(v[4 * b + 2 + e]) + lds32f(relpos_s + (rp1 - koff) * 4);
            }
          }
          if (ms.win_mixed[rb][wl]) {
            const uint8_t g0 = ms.region[rb][r0], g1 = ms.region[rb][r1];
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const uint8_t gk = ms.region[rb][wl * 64 + 8 * b + 2 * tq + e];
                s0[2 * b + e] += (gk != g0) ? -100.0f : 0.0f;
                s1[2 * b + e] += (gk != g1) ? -100.0f : 0.0f;
              }
          }
          if (a.mask != nullptr) {
            const int w = tile * 2 + wl;
            if (w < a.n_windows) {
              const int kn = (tq >> 1) * 8 + 2 * (tq & 1);     // per-thread key offset in natural (row-major 8x8) order
              const float* m0 = a.mask + ((size_t)(w % a.n_mask_windows) * 64 + row_y(r0) * 8 + row_x(r0)) * 64 + kn;
              const float* m1 = a.mask + ((size_t)(w % a.n_mask_windows) * 64 + row_y(r1) * 8 + row_x(r1)) * 64 + kn;
#pragma unroll
              for (int b = 0; b < 8; ++b) {
                const int kb8 = ((b >> 2) * 4 + (b & 1) * 2) * 8 + ((b >> 1) & 1) * 4;       // natural index of key 8b (+ kn per thread)
                const float2 a0 = __ldg(reinterpret_cast<const float2*>(m0 + kb8));
                const float2 a1 = __ldg(reinterpret_cast<const float2*>(m1 + kb8));
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
        LW_TRACE_STMT(if (trw && it < 3 && h < 2) a.trace[tw++] = clock64();)

        // ---- O epilogue: normalise, park as bf16 A operand of the projection ----
        wt_wait(smem_u32(&ms.bar_o_full), ph);
        tc_fence_after();
        LW_TRACE_STMT(if (trw && it < 3 && h < 2) a.trace[tw++] = clock64();)
        {
          uint32_t v[4 * NBH];
          if (NBH == 4) tmem_ld_16x256b_x4(tb + tl + Cfg::T_DO, v);
          else tmem_ld_16x256b_x2(tb + tl + Cfg::T_DO, v);
          tmem_wait_ld();
          const float i0 = 1.0f / sum0, i1 = 1.0f / sum1;
          uint32_t pk[2 * NBH];
#pragma unroll
          for (int i = 0; i < NBH; ++i) {
            pk[2 * i] = pack_bf16(__uint_as_float(v[4 * i]) * i0, __uint_as_float(v[4 * i + 1]) * i0);
            pk[2 * i + 1] = pack_bf16(__uint_as_float(v[4 * i + 2]) * i1, __uint_as_float(v[4 * i + 3]) * i1);
          }
          if (NBH == 4) tmem_st_16x128b_x4(tb + tl + Cfg::T_OALL + (h * HD) / 2, pk);
          else tmem_st_16x128b_x2(tb + tl + Cfg::T_OALL + (h * HD) / 2, pk);
          tmem_wait_st();
        }
        tc_fence_before();
      }
      mbar_arrive(smem_u32(&ms.bar_oall));
      LW_TRACE_STMT(if (trw && it < 3) a.trace[tw++] = clock64();)

      // ---- projection epilogue: + bias -> bf16 -> staging tile (the dead Q/K/V tiles) -> coalesced scatter with the shortcut.
      // The shortcut rows of a round are fetched into registers before the round's accumulator is waited for. ----
      if (tid < 128 && it + 1 < n_my) build_rows(it + 1);          // next tile's row tables, under the projection GEMM
      constexpr int VPR = Cfg::NCHS / 8;                             // 16-byte vectors per row and round
      constexpr int VPT = 128 * VPR / kWorkers8;                     // vectors per thread and round: 4 / 2 / 1
      int tok[VPT];
      uint4 rv[VPT][2];
      auto fetch_resid = [&](int col0) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
          const int i = tid + k * kWorkers8, row = i / VPR, vec = i % VPR;
          tok[k] = ms.row_tok[rb][row];
          if (tok[k] >= 0 && a.resid != nullptr) {
            const size_t off = (size_t)tok[k] * C + col0 + vec * 8;
            if (a.resid_fp32) {
              rv[k][0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.resid) + off));
              rv[k][1] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.resid) + off + 4));
            } else {
              rv[k][0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.resid) + off));
            }
          }
        }
      };
      auto store_round = [&](int col0) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
          if (tok[k] < 0) continue;
          const int i = tid + k * kWorkers8, row = i / VPR, vec = i % VPR;
          float f[8];
          unpack8(lds128(stage_s + row * Cfg::PITCH + vec * 16), f);
          if (a.resid != nullptr) {
            if (a.resid_fp32) {
              const uint4 p = rv[k][0], q = rv[k][1];
              f[0] += __uint_as_float(p.x); f[1] += __uint_as_float(p.y); f[2] += __uint_as_float(p.z); f[3] += __uint_as_float(p.w);
              f[4] += __uint_as_float(q.x); f[5] += __uint_as_float(q.y); f[6] += __uint_as_float(q.z); f[7] += __uint_as_float(q.w);
            } else {
              float r[8];
              unpack8(rv[k][0], r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += r[e];
            }
          }
          const size_t off = (size_t)tok[k] * C + col0 + vec * 8;
          if (a.out_fp32) {
            float* op = reinterpret_cast<float*>(a.out) + off;
            *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.out) + off) = pack8(f);
          }
          if (a.out_b != nullptr) *reinterpret_cast<uint4*>(a.out_b + off) = pack8(f);
        }
      };
      fetch_resid(0);
      for (int nc = 0; nc < Cfg::NC; ++nc) {
        const int pc = it * Cfg::NC + nc;
        wt_wait(smem_u32(&ms.bar_d_full), pc & 1);
        tc_fence_after();
        LW_TRACE_STMT(if (trw && it < 3) a.trace[tw++] = clock64();)
#pragma unroll 1
        for (int c0 = 0; c0 < Cfg::NCH; c0 += Cfg::NCHS) {
          uint32_t v[4 * NBP];
          const uint32_t ta = tb + tl + Cfg::T_WORK + c0;
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
          stage_frag<NBP>(stage_s, Cfg::PITCH, row16, 0, pk);
          const bool last_round = c0 + Cfg::NCHS >= Cfg::NCH;
          if (last_round) { tc_fence_before(); mbar_arrive(smem_u32(&ms.bar_d_empty)); }
          worker_bar8();                                   // (also: build_rows of the next tile is complete)
          if (nc == 0 && c0 == 0 && tid < 128 && it + 1 < n_my) mark_mixed(it + 1);
          const int col0 = nc * Cfg::NCH + c0;
          store_round(col0);
          if (!(last_round && nc + 1 == Cfg::NC)) fetch_resid(col0 + Cfg::NCHS);      // next round's shortcut rows
          worker_bar8();
        }
      }
      LW_TRACE_STMT(if (trw && it < 3) a.trace[tw++] = clock64();)
    }
    LW_TRACE_STMT(if (trw) a.trace[tw++] = -1;)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

}  // namespace lw
