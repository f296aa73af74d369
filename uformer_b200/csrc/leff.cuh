// leff.cuh — the GEMM-shaped kernels around the attention: LeFF (two kernels), Downsample, Upsample.
//
// Two CTA skeletons, both 128 output rows per CTA with the worker/producer/issuer roles of
// lewin_common.cuh:
//  * "A resident": the whole 128 x K A operand is staged once in shared memory; the N dimension is
//    streamed in chunks with a double-buffered TMEM accumulator so the epilogue of chunk n overlaps
//    the MMAs of chunk n+1.            -> LeFF linear1 (+LayerNorm, +GELU), Upsample (+scatter)
//  * "A streamed": the A operand is produced k-block by k-block (64 channels) by the workers into a
//    double-buffered shared tile while the full 128 x N accumulator stays resident in TMEM.
//                                      -> LeFF dwconv3x3+GELU+linear2 (+residual), Downsample (im2col)
#pragma once
#include "lewin_common.cuh"
#include "../../include/lewin_b200.h"

namespace lw {

struct GemmMisc {
  int row_tok[128];
  uint64_t bar_full[8], bar_empty[8];
  uint64_t bar_a_ready;
  uint64_t bar_a_full[2], bar_a_empty[2];
  uint64_t bar_d_full[2], bar_d_empty[2];
  uint32_t tmem_base;
};

// =================================================================================================
// A-resident skeleton.  EPI = 0: LeFF linear1 (bias + exact GELU -> bf16 h1 rows)
//                       EPI = 1: Upsample (bias + 2x2 pixel-shuffle scatter)
// =================================================================================================
struct AResArgs {
  const bf16* x; int n_rows; int K;        // A rows (tokens) and channels
  const float* ln_w; const float* ln_b; float ln_eps;
  const uint8_t* w_img; int n_total; int nch;   // N, chunk rows; image [N/nch][KB][nch*128B]
  const float* bias;
  bf16* out;
  // EPI 0: out row stride = n_total.  EPI 1: upsample geometry
  int H, W, Cout, out_stride;
  long long* trace;   // dbg&16: CTA 0 writes clock64 timestamps here (profiling aid)
  int dbg;   // profiling knobs (env LW_DEBUG): 1 skip epilogue stores, 2 skip GELU, 4 skip weight loads, 8 skip A staging loads
  int nsplit;         // CTAs per 128-row tile (grid = tiles * nsplit): each takes NC / nsplit consecutive N chunks.  Small token
                      // maps (the 16 x 16 and 32 x 32 stages) otherwise launch fewer CTAs than there are SMs; staging the A tile
                      // twice is cheap next to streaming the whole weight through one CTA.
};

template <int K>
struct AResCfg {
  static constexpr int KB = (K + 63) / 64;
  // N chunk per accumulator buffer: 256-wide UMMAs halve the A re-reads from shared memory (an SS UMMA
  // re-reads its whole 128 x 16 A slice every instruction); only K == 256 has the capacity for it.
  static constexpr int NCH_MAX = (K == 256) ? 256 : 128;
  static constexpr int STAGE_BYTES = NCH_MAX * 128;
  // Ring depth.  The ring is latency-bound rather than bandwidth-bound: one producer thread keeps STAGES bulk copies of 16 KB in
  // flight against ~1.3 us of loaded L2 latency (~12 GB/s per stage and SM).  K = 512 streams 2 MB of weights per 128-row tile
  // (28 GB/s at the measured 75 us per tile): a fourth stage is paid for by epilogue passes of 64 instead of 128 accumulator
  // columns (SUB_COLS), whose staging tiles are 20 KB instead of 36 KB.
  static constexpr int STAGES = (K <= 128) ? 2 : (K == 512 ? 4 : 3);
  static constexpr int SUB_COLS = (K == 512) ? 64 : 128;             // accumulator columns per epilogue pass (two column halves)
  static constexpr int STAGE_HALF = 128 * (SUB_COLS + 16);           // one half's staging tile: 128 rows x (SUB_COLS/2 bf16 + 16 B)
  static constexpr int S_X = 0;
  static constexpr int S_RING = KB * 16384;
  static constexpr int S_STAGE = S_RING + STAGES * STAGE_BYTES;      // epilogue staging tiles of the two column halves
  static constexpr int S_BIAS = S_STAGE + 2 * STAGE_HALF;            // bias of the whole N range (<= 2048 fp32)
  static constexpr int S_MISC = S_BIAS + 8192;
  static constexpr int SMEM_BYTES = S_MISC + 1024 + 1024;
  static constexpr int T_ALLOC = 2 * NCH_MAX;                        // two accumulator buffers
};
static_assert(sizeof(GemmMisc) <= 1024, "misc too large");

// Phase A of the A-resident epilogue for one sub-chunk of 16*NB columns: warp (q, half) owns TMEM lanes q*32..+32
// (two 16-lane fragments) and columns half*8NB .. +8NB.  tcol = TMEM address of the sub-chunk's first column,
// ncol = its global N index.
template <int NB, int EPI>
__device__ __forceinline__ void ares_phase_a(const AResArgs& a, uint32_t tcol, int ncol, uint32_t bias_s, uint32_t stage_s, int pitch, const GeluH2& gelu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2, tq = lane & 3;
  constexpr int CPH = 8 * NB;
  uint32_t v[2][4 * NB];
#pragma unroll
  for (int hl = 0; hl < 2; ++hl) {
    const uint32_t ta = tcol + half * CPH + ((uint32_t)(q * 32 + hl * 16) << 16);
    if (NB == 8) tmem_ld_16x256b_x8(ta, v[hl]); else tmem_ld_16x256b_x4(ta, v[hl]);
  }
  f2 bb[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int n0 = ncol + half * CPH + 8 * i + 2 * tq;
    const float2 b2 = lds64f(bias_s + ((EPI == 0) ? n0 : (n0 % a.Cout)) * 4);
    bb[i] = f2_pack(b2.x, b2.y);
  }
  tmem_wait_ld();
#pragma unroll
  for (int hl = 0; hl < 2; ++hl) {
    uint32_t pk[2 * NB];
    if (EPI == 0) frag_bias_gelu_h2<NB>(v[hl], bb, pk, gelu);          // LeFF hidden map: fp16 in HBM (internal buffer)
    else frag_bias_act_pack<NB, false>(v[hl], bb, pk);
    stage_frag<NB>(stage_s, pitch, q * 32 + hl * 16, 0, pk);       // the column half owns a private staging tile
  }
}

template <int K, int EPI>
__global__ void __launch_bounds__(kThreads8, (K <= 128) ? 2 : 1) ares_kernel(const AResArgs a) {
  using Cfg = AResCfg<K>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmMisc& ms = *reinterpret_cast<GemmMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.nsplit;
  const int NCT = a.n_total / a.nch / a.nsplit;                 // N chunks of this CTA: [nc0, nc0 + NCT)
  const int nc0 = (blockIdx.x % a.nsplit) * NCT;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    mbar_init(smem_u32(&ms.bar_a_ready), kWorkers8);
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
  const uint32_t sX = smem_u32(smem + Cfg::S_X);
  const uint32_t chunk_bytes = a.nch * 128;

  if (warp == 8) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
      for (int nc = nc0; nc < nc0 + NCT; ++nc)
        for (int kb = 0; kb < Cfg::KB; ++kb)
          if (LW_DBG(a, 4)) { mbar_wait(ring.empty(), ring.phase() ^ 1); mbar_arrive(ring.full()); ++ring.idx; }
          else ring.load(a.w_img + (size_t)(nc * Cfg::KB + kb) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 9) {
    {   // issuer warp: warp-uniform control flow, one elected lane issues (operands stay in uniform registers)
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0, Cfg::STAGE_BYTES};
      const uint32_t idesc = make_idesc_bf16(128, a.nch);
      const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
      const uint64_t a_desc0 = kmajor_desc<128>(sX), b_desc0 = kmajor_desc<128>(ring_base);   // address field += bytes >> 4
      LW_TRACE_STMT(const bool tr = (a.dbg & 16) && blockIdx.x == 0 && a.trace != nullptr; int ti = 0;)
      LW_TRACE_STMT(if (tr) a.trace[ti++] = clock64();)
      mbar_wait(smem_u32(&ms.bar_a_ready), 0);
      tc_fence_after();
      LW_TRACE_STMT(if (tr) a.trace[ti++] = clock64();)
      for (int nc = 0; nc < NCT; ++nc) {                      // (local chunk counter: buffers and phases)
        const int buf = nc & 1;
        mbar_wait(smem_u32(&ms.bar_d_empty[buf]), ((nc >> 1) & 1) ^ 1);
        tc_fence_after();
        LW_TRACE_STMT(if (tr) a.trace[ti++] = clock64();)
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          LW_TRACE_STMT(if (tr) a.trace[ti++] = clock64();)
          constexpr int KS = (K >= 64) ? 4 : K / 16;
          const uint64_t ad = a_desc0 + (uint64_t)(kb * 1024), bd = b_desc0 + (uint64_t)((wst - ring_base) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              umma_ss(tb + buf * Cfg::NCH_MAX, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) != 0);
          }
          __syncwarp();
          ring.release();
          LW_TRACE_STMT(if (tr) a.trace[ti++] = clock64();)
        }
        if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full[buf]));
        __syncwarp();
      }
      LW_TRACE_STMT(if (tr) a.trace[ti++] = -1;)
    }
  } else {
    // 8 worker warps: stage 16 A rows each; epilogue: lane quadrant warp&3, column half warp>>2
    if (tid < 128) {
      const int row0 = tile * 128 + tid;
      ms.row_tok[tid] = (row0 < a.n_rows) ? row0 : -1;
    }
    const uint32_t bias_s = smem_u32(smem + Cfg::S_BIAS);
    {
      const int nb = (EPI == 0) ? a.n_total : a.Cout;      // bias entries
      for (int i4 = tid * 4; i4 < nb; i4 += kWorkers8 * 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + i4));
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(bias_s + i4 * 4), "f"(b4.x), "f"(b4.y), "f"(b4.z), "f"(b4.w) : "memory");
      }
    }
    worker_bar8();
    if (!LW_DBG(a, 8)) stage_rows_ln<K, 8>(smem + Cfg::S_X, a.x, ms.row_tok, a.ln_w, a.ln_b, a.ln_eps, nullptr);
    fence_async_smem();
    mbar_arrive(smem_u32(&ms.bar_a_ready));
    const uint32_t stage_s = smem_u32(smem + Cfg::S_STAGE);
    if (EPI == 1) {
      // destination base row of every input token: (b, 2y, 2x) in the (B, 2H, 2W) output map
      worker_bar8();
      if (tid < 128) {
        const int row0 = ms.row_tok[tid];
        if (row0 >= 0) {
          const int ub = row0 / (a.H * a.W), t = row0 % (a.H * a.W), uy = t / a.W, ux = t % a.W;
          ms.row_tok[tid] = (ub * 2 * a.H + 2 * uy) * (2 * a.W) + 2 * ux;
        }
      }
      worker_bar8();
    }
    GeluH2 gelu;
    gelu.init();
    const int sub_cols = a.nch < Cfg::SUB_COLS ? a.nch : Cfg::SUB_COLS;   // accumulator columns per pass: half of them per column half
    const int grp = warp >> 2;                               // column half == synchronisation group (128 threads)
    const int cph = sub_cols >> 1;
    int cph_log2 = 4;
    while ((1 << cph_log2) < cph) ++cph_log2;
    const int pitch_g = cph * 2 + 16;
    const uint32_t stage_g = stage_s + grp * Cfg::STAGE_HALF;
    auto group_bar = [](int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 2) : "memory"); };
    LW_TRACE_STMT(const bool trw = (a.dbg & 16) && blockIdx.x == 0 && tid == 0 && a.trace != nullptr; int tw = 512, tw2 = 1024;)
    LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
    for (int ncl = 0; ncl < NCT; ++ncl) {
      const int buf = ncl & 1, nc = nc0 + ncl;
      mbar_wait(smem_u32(&ms.bar_d_full[buf]), (ncl >> 1) & 1);
      tc_fence_after();
      LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
      for (int sc = 0; sc < a.nch; sc += sub_cols) {
        // ---- phase A: TMEM (16x256b fragments) -> (+bias, GELU) -> bf16 -> stmatrix into this column half's private
        // staging tile.  The two halves (warps 0-3 / 4-7) synchronise only among themselves, so one half's
        // copy-out overlaps the other half's GELU math. ----
        if (sub_cols == 128) ares_phase_a<8, EPI>(a, tb + buf * Cfg::NCH_MAX + sc, nc * a.nch + sc, bias_s, stage_g, pitch_g, gelu);
        else ares_phase_a<4, EPI>(a, tb + buf * Cfg::NCH_MAX + sc, nc * a.nch + sc, bias_s, stage_g, pitch_g, gelu);
        if (sc + sub_cols >= a.nch) {       // accumulator fully read: hand the buffer back to the issuer
          tc_fence_before();
          mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
        }
        group_bar(grp);
        // ---- phase B: coalesced copy-out of this half's columns by its 128 threads ----
        if (!LW_DBG(a, 1)) {
          const int colg = nc * a.nch + sc + grp * cph;
          if (EPI == 0) {
            store_staged_rows128(stage_g, pitch_g, cph_log2, ms.row_tok, a.out, nullptr, (size_t)a.n_total, colg, tid & 127);
          } else {
            const int vshift = cph_log2 - 3;
            const int total = 128 << vshift;
            for (int i2 = (tid & 127); i2 < total; i2 += 128) {
              const int row = i2 >> vshift, vec = i2 & ((1 << vshift) - 1);
              const int obase = ms.row_tok[row];
              if (obase < 0) continue;
              const int n0 = colg + vec * 8;
              const int q = n0 / a.Cout, co = n0 % a.Cout;
              const size_t orow = (size_t)obase + (q >> 1) * (2 * a.W) + (q & 1);
              *reinterpret_cast<uint4*>(a.out + orow * a.out_stride + co) = lds128(stage_g + row * pitch_g + vec * 16);
            }
          }
        }
        group_bar(grp);                     // this half's staging tile is free for its next phase A
        LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
      }
    }
    LW_TRACE_STMT(if (trw) { a.trace[tw++] = -1; a.trace[tw2++] = -1; })
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

// =================================================================================================
// Arguments of the A-streamed kernels (leff2.cuh: depthwise conv producer; down.cuh: im2col producer)
// =================================================================================================
struct AStreamArgs {
  const bf16* src;        // PROD 0: h1 (B,H,W,K) — FP16 bits (the LeFF hidden map is half precision)   PROD 1: x (B,H,W,Cin) bf16
  int B, H, W;            // geometry of src
  int K;                  // PROD 0: hidden; PROD 1: 16*Cin
  int Cin;                // PROD 1 only
  int src_stride;         // PROD 1: row stride of src in elements (>= Cin: src may be a column slice of a wider buffer)
  const uint16_t* taps;   // PROD 0: (10, K) fp16: 9 depthwise taps (tap = ky*3+kx), then the conv bias
  const uint8_t* w_img;   // [K/64][N/nch][nch*128B]
  int N, nch;
  const float* bias;      // (N)
  const bf16* resid;      // (rows, N) or null (fp32 if resid_fp32)
  bf16* out;              // (rows, N) (fp32 if out_fp32)
  int resid_fp32, out_fp32;
  int TW, TH;             // PROD 0 tile shape (TW*TH == 128)
  int tiles_x;
  long long* trace;       // profiling aid (LW_DEBUG & 16)
  int dbg;
};

}  // namespace lw
