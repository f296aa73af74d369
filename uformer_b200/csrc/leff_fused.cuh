// leff_fused.cuh — LeFF in ONE kernel (model.py:666-685 with norm2 of :987 folded in):
//   out = resid + GELU( dwconv3x3( GELU( LN(x) W1^T + b1 ) ) + bd ) W2^T + b2
// The 4C-wide hidden map never exists in HBM: it lives, 64 channels at a time, in shared memory.
//
// PERSISTENT kernel, one CTA per SM, 8 x 16 spatial output tiles (128 tokens) walked round-robin.  Per tile:
//   TMA      the 10 x 18 halo'd block of RAW input tokens lands with one cp.async.bulk.tensor box per 64-channel
//            k-block (4-D tensor map (C, W, H, B), out-of-image coordinates are zero-filled by the hardware) directly
//            in the K-major swizzled layout tcgen05 consumes: 180 rows, row = halo token.
//   stats    two warps compute the LayerNorm statistics of the 180 rows from shared memory.  LayerNorm itself is
//            FOLDED into GEMM-1: W1' = W1 diag(gamma) (host packing), so
//              LN(x) W1^T + b1 = rstd * (x W1'^T) - rstd*mu * colsum(W1') + (b1 + W1 beta)
//            and the raw tile is the A operand as it landed (no normalise-and-rewrite pass; also one bf16 rounding less).
//   GEMM-1   per 64-channel hidden slice j: D1_j[192 x 64] = X[192 x C] W1'_j^T as one M=128 + one M=64 tcgen05.mma chain
//            (192 = 180 halo rows rounded up), accumulators double-buffered in TMEM.
//   E1       4 epilogue warps: D1 (16x256b TMEM fragments) -> LN fold + bias -> GELU -> bf16 -> stmatrix into the halo
//            tile [192 tokens][64 ch] (zero for out-of-image tokens = the conv's zero padding of h1, model.py:659).
//   conv     8 warps: depthwise 3x3 + bias + GELU on the halo tile (one warp per channel octet, lanes = 16 columns x
//            2 row halves sliding down their column, packed FFMA2) -> bf16 A operand of GEMM-2.
//   GEMM-2   D2[128 x C] += A2_j[128 x 64] W2_j^T, accumulator double-buffered in TMEM across tiles.
//   E2       (the E1 warps, one tile behind) D2 + b2 -> staging -> coalesced store with the residual added.
// Every stage is double-buffered and mbarrier-linked, so GEMM-1(j+1), E1(j), conv(j-1), GEMM-2(j-2) run concurrently and
// the pipeline does not drain between tiles.
#pragma once
#include <cuda.h>
#include "lewin_common.cuh"
#include "leff2.cuh"

namespace lw {

// per-role timeline slots of CTA 0 (only compiled with -DLW_TRACE): role r owns trace[r*512 ..]
#define LF_TRACE(role, idx) LW_TRACE_STMT(if (a.trace != nullptr && blockIdx.x == 0 && (idx) < 512) a.trace[(role) * 512 + (idx)] = clock64();)

constexpr int kLFThreads = 640;       // 8 conv warps | 4 E1/E2 warps | producer, issuer, 2 stats warps | 4 E1/E2 warps
constexpr int kLFConv = 256;

struct LeffFArgs {
  const bf16* x;           // the tensor-map source again: the LayerNorm statistics are read with plain loads
  int x_stride;
  int B, H, W, hidden;
  const uint8_t* w1_img;   // [hidden/SL][KB1][SL rows x SW bytes] bf16, gamma folded, swizzle SW = 2*min(C,64)
  const float* b1f;        // (hidden)  b1 + W1 beta
  const float* cs;         // (hidden)  row sums of the bf16-rounded W1' (mean correction of the folded LayerNorm)
  const uint8_t* taps;     // [hidden/SL][10][SL] f16: 9 depthwise taps (tap = ky*3+kx) + conv bias per slice
  const uint8_t* w2_img;   // [hidden/SL][C rows x 2*SL bytes] f16, swizzle 2*SL
  const float* b2;         // (C)
  const void* resid;       // (B*H*W rows, stride resid_stride) bf16 or fp32, or null
  void* out;               // (B*H*W rows, stride out_stride) bf16 or fp32
  int resid_stride, out_stride, resid_fp32, out_fp32;
  int has_ln;
  float ln_eps;
  int tiles_x, tiles_y, n_tiles;
  bf16* out_b;             // optional bf16 copy of out (row stride out_b_stride), or null
  int out_b_stride;
  long long* trace;        // -DLW_TRACE builds: CTA 0 writes per-role clock64 timelines here (tools/leff_fused_trace.py)
};

template <int C>
struct LeffFCfg {
  static constexpr int SL = (C <= 128) ? 64 : 32;           // hidden channels per slice (smem / TMEM budget at C = 256)
  static constexpr int SWH = SL * 2;                        // row bytes / swizzle span of the halo tile, A2 and W2 chunks
  static constexpr int CB = C < 64 ? C : 64;               // channels per k-block
  static constexpr int SW = CB * 2;                         // bytes per A / W1 row inside a k-block = swizzle span
  static constexpr int KB1 = (C + 63) / 64;
  static constexpr int KS1 = CB / 16;
  static constexpr int A1_ROWS = 192;
  static constexpr int A1_KB_BYTES = A1_ROWS * SW;
  static constexpr int A1_BYTES = KB1 * A1_KB_BYTES;
  static constexpr int NA1 = (C <= 64) ? 2 : 1;             // input tiles in flight
  static constexpr int X_BOX_BYTES = 180 * SW;              // one TMA box (one k-block)
  // C = 256 (SL = 32): GEMM-1 is issued for PAIRS of slices (N = 64: half as many tcgen05.mma as two N = 32 chains; the
  // 32-wide MMAs and the weight ring behind them were the bound of that configuration, tools/leff_fused_trace.py), its W1
  // operand arrives as two ring chunks of two k-blocks each.  Everything downstream stays 32 channels wide.
  static constexpr bool PAIR = (SL == 32);
  static constexpr int G1N = PAIR ? 64 : SL;                // N of the GEMM-1 instructions
  static constexpr int G1_CHUNKS = PAIR ? 2 : 1;            // ring chunks per GEMM-1
  static constexpr int W1_CHUNK = PAIR ? (KB1 / 2) * 64 * SW : KB1 * SL * SW;
  static constexpr int W2_CHUNK = C * SWH;
  static constexpr int STAGES = 3;
  static constexpr int HALO_BYTES = 192 * SWH;
  static constexpr int A2_BYTES = 128 * SWH;
  static constexpr int NTAP = 4;
  static constexpr int TAP_BYTES = 10 * SL * 2;             // [10][SL] f16
  static constexpr int STAGE_PITCH = 80;                    // E2 staging: 32 bf16 columns + 16 B pad
  static constexpr int S_A1 = 0;
  static constexpr int S_HALO = (NA1 * A1_BYTES + 1023) / 1024 * 1024;
  static constexpr int S_A2 = S_HALO + 2 * HALO_BYTES;
  static constexpr int S_RING = S_A2 + 2 * A2_BYTES;
  static constexpr int S_STAGE = S_RING + STAGES * kStageBytes;
  static constexpr int S_TAPS = S_STAGE + 2 * 128 * STAGE_PITCH;      // (two staging tiles: one per epilogue group)          // NTAP x [10][SL] fp32
  static constexpr int S_B1 = S_TAPS + NTAP * TAP_BYTES;              // b1f[hidden], cs[hidden], hidden <= 1024
  static constexpr int S_B2 = S_B1 + 2 * 1024 * 4;                    // b2[C]
  static constexpr int S_STATS = S_B2 + 1024;                         // [2][192] float2
  static constexpr int S_MISC = S_STATS + 2 * 192 * 8;
  static constexpr int SMEM_BYTES = S_MISC + 2048 + 1024;
  static constexpr int T_D1 = 0;                            // 2 x (G1N cols M=128 part | G1N cols M=64 part)
  static constexpr int D1_COLS = 2 * G1N;
  static constexpr int T_D2 = 2 * D1_COLS;                  // ND2 x C columns
  static constexpr int ND2 = (T_D2 + 2 * C <= 512) ? 2 : 1;
  static constexpr int T_ALLOC = 512;
  static_assert(C <= 256 && C % 16 == 0, "fused LeFF: C in {16,32,64,128,256}");
  static_assert(W1_CHUNK <= kStageBytes && W2_CHUNK <= kStageBytes, "ring stage");
  static_assert(S_RING % 1024 == 0 && S_A2 % 1024 == 0 && S_HALO % 1024 == 0, "operand alignment");
  static_assert(T_D2 + ND2 * C <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

struct LeffFMisc {
  uint64_t bar_full[4], bar_empty[4];
  uint64_t bar_x_full[2], bar_x_empty[2];
  uint64_t bar_st_full[2], bar_st_empty[2];
  uint64_t bar_d1_full[2], bar_d1_empty[2];
  uint64_t bar_h_full[2], bar_h_empty[2];
  uint64_t bar_a2_full[2], bar_a2_empty[2];
  uint64_t bar_d2_full[2], bar_d2_empty[2];
  uint64_t bar_tap_full[4], bar_tap_empty[4];
  uint32_t tmem_base;
  int row_out[256];
};
static_assert(sizeof(LeffFMisc) <= 2048, "misc too large");

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int C>
__global__ void __launch_bounds__(kLFThreads, 1) leff_fused_kernel(const __grid_constant__ CUtensorMap xmap, const LeffFArgs a) {
  using Cfg = LeffFCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  LeffFMisc& ms = *reinterpret_cast<LeffFMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int SL = Cfg::SL;
  const int NS = a.hidden / SL;
  const int my_tiles = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_tiles * NS;                    // hidden slices this CTA walks

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ms.bar_x_full[i]), 1);  mbar_init(smem_u32(&ms.bar_x_empty[i]), 257);    // GEMM-1 commit + 256 epilogue threads (stats)
      mbar_init(smem_u32(&ms.bar_st_full[i]), 256); mbar_init(smem_u32(&ms.bar_st_empty[i]), 8);   // every epilogue thread (each
                                                                                                   // releases its own writes)
      mbar_init(smem_u32(&ms.bar_d1_full[i]), 1); mbar_init(smem_u32(&ms.bar_d1_empty[i]), Cfg::PAIR ? 512 : 256);
      mbar_init(smem_u32(&ms.bar_h_full[i]), 256); mbar_init(smem_u32(&ms.bar_h_empty[i]), kLFConv);
      mbar_init(smem_u32(&ms.bar_a2_full[i]), kLFConv); mbar_init(smem_u32(&ms.bar_a2_empty[i]), 1);
      mbar_init(smem_u32(&ms.bar_d2_full[i]), 1); mbar_init(smem_u32(&ms.bar_d2_empty[i]), (C >= 64) ? 256 : 128);
    }
    for (int i = 0; i < Cfg::NTAP; ++i) { mbar_init(smem_u32(&ms.bar_tap_full[i]), 1); mbar_init(smem_u32(&ms.bar_tap_empty[i]), kLFConv); }
    fence_mbar_init();
    tma_prefetch_desc(&xmap);
  }
  if (warp == 12) tmem_alloc(smem_u32(&ms.tmem_base), Cfg::T_ALLOC);
  pdl_launch_dependents();
  pdl_wait();                      // nothing above touches global memory
  // resident per-channel tables: b1f, cs, b2; the never-loaded A rows 180..191 are zeroed once
  {
    for (int i = tid; i < a.hidden; i += kLFThreads) {
      reinterpret_cast<float*>(smem + Cfg::S_B1)[i] = __ldg(a.b1f + i);
      reinterpret_cast<float*>(smem + Cfg::S_B1)[a.hidden + i] = __ldg(a.cs + i);
    }
    for (int i = tid; i < C; i += kLFThreads) reinterpret_cast<float*>(smem + Cfg::S_B2)[i] = __ldg(a.b2 + i);
    constexpr int PAD_WORDS = 12 * Cfg::SW / 4;
    for (int i = tid; i < Cfg::NA1 * Cfg::KB1 * PAD_WORDS; i += kLFThreads) {
      const int buf = i / PAD_WORDS, w = i % PAD_WORDS;
      reinterpret_cast<uint32_t*>(smem + Cfg::S_A1 + buf * Cfg::A1_KB_BYTES + 180 * Cfg::SW)[w] = 0u;
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const int wg = warp >> 2;

  auto tile_of = [&](int it) { return (int)blockIdx.x + it * (int)gridDim.x; };

  // Register re-balancing (setmaxnreg moves registers inside the CTA's launch allocation of 640 x 96): control warpgroup
  // 96 -> 64, the two epilogue warpgroups 96 -> 88, the two conv warpgroups 96 -> 120:  128*(32 + 8 + 8) = 2*128*24.
  if (wg == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 12) {
      // ============================== producer: input tiles (TMA boxes) + weight ring ==============================
      if (lane == 0) {
        Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
        auto load_x = [&](int it) {
          const int ab = it % Cfg::NA1, use = it / Cfg::NA1;
          mbar_wait(smem_u32(&ms.bar_x_empty[ab]), (use & 1) ^ 1);
          const int tile = tile_of(it);
          const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / (a.tiles_x * a.tiles_y);
          const uint32_t bar = smem_u32(&ms.bar_x_full[ab]);
          mbar_expect_tx(bar, Cfg::KB1 * Cfg::X_BOX_BYTES);
          LF_TRACE(5, it)
#pragma unroll
          for (int kb = 0; kb < Cfg::KB1; ++kb)
            tma_load_4d(smem_u32(smem + Cfg::S_A1 + ab * Cfg::A1_BYTES + kb * Cfg::A1_KB_BYTES), &xmap, kb * 64, tx * 16 - 1, ty * 8 - 1, b, bar);
        };
        if (my_tiles > 0) load_x(0);
        // (slice, tile) counters are advanced incrementally: NS is a runtime value and an integer division per iteration costs
        // ~25 instructions on warps that issue one instruction every ~8 cycles
        int j = 0, it = 0, j2 = 0;
        for (int k = 0; k < total + 2; ++k) {
          if (k < total) {
            if (j == 0) {
              if (Cfg::NA1 == 2) { if (it + 1 < my_tiles) load_x(it + 1); }
              else if (it > 0) load_x(it);
            }
            {   // depthwise taps + conv bias of this slice (consumed by the conv warps two pipeline slots later)
              const int tbuf = k % Cfg::NTAP;
              mbar_wait(smem_u32(&ms.bar_tap_empty[tbuf]), ((k / Cfg::NTAP) & 1) ^ 1);
              mbar_expect_tx(smem_u32(&ms.bar_tap_full[tbuf]), Cfg::TAP_BYTES);
              bulk_g2s(smem_u32(smem + Cfg::S_TAPS + tbuf * Cfg::TAP_BYTES), a.taps + (size_t)j * Cfg::TAP_BYTES, Cfg::TAP_BYTES, smem_u32(&ms.bar_tap_full[tbuf]));
            }
            if (!Cfg::PAIR) ring.load(a.w1_img + (size_t)j * Cfg::W1_CHUNK, Cfg::W1_CHUNK);
            else if ((k & 1) == 0) {
              ring.load(a.w1_img + (size_t)j * Cfg::W1_CHUNK, Cfg::W1_CHUNK);              // pair j/2: k-blocks 0-1 ...
              ring.load(a.w1_img + (size_t)(j + 1) * Cfg::W1_CHUNK, Cfg::W1_CHUNK);        // ... and 2-3
            }
          }
          if (k >= 2) { ring.load(a.w2_img + (size_t)j2 * Cfg::W2_CHUNK, Cfg::W2_CHUNK); if (++j2 == NS) j2 = 0; }
          if (k < total && ++j == NS) { j = 0; ++it; }
        }
      }
    } else if (warp == 13) {
      // ============================== issuer (warp-uniform; one elected lane issues) ==============================
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      constexpr uint32_t idesc_g1a = make_idesc_bf16(128, Cfg::G1N), idesc_g1b = make_idesc_bf16(64, Cfg::G1N), idesc_g2 = make_idesc_f16(128, C);
      int j1 = 0, it1 = 0, j2 = 0, it2 = 0;                 // (slice, tile) of the GEMM-1 / GEMM-2 being issued
      for (int k = 0; k < total + 2; ++k) {
        if (k < total && (!Cfg::PAIR || (k & 1) == 0)) {
          // ---- GEMM-1 of slice k (PAIR: of slices k, k+1): D1[buffer] = X W1'^T ----
          const int j = j1, it = it1, ab = it % Cfg::NA1;
          const int db = Cfg::PAIR ? ((k >> 1) & 1) : (k & 1), use = Cfg::PAIR ? (k >> 2) : (k >> 1);
          if (j == 0) { mbar_wait(smem_u32(&ms.bar_x_full[ab]), (it / Cfg::NA1) & 1); }
          mbar_wait(smem_u32(&ms.bar_d1_empty[db]), (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t xs = smem_u32(smem + Cfg::S_A1 + ab * Cfg::A1_BYTES);
          constexpr int KBC = Cfg::KB1 / Cfg::G1_CHUNKS;       // k-blocks per ring chunk
#pragma unroll
          for (int ch = 0; ch < Cfg::G1_CHUNKS; ++ch) {
            const uint32_t wst = ring.acquire();
            if (elect_one()) {
#pragma unroll
              for (int kb2 = 0; kb2 < KBC; ++kb2)
#pragma unroll
                for (int ks = 0; ks < Cfg::KS1; ++ks) {
                  const int kb = ch * KBC + kb2;
                  const uint64_t bd = kmajor_desc<Cfg::SW>(wst + kb2 * Cfg::G1N * Cfg::SW + ks * 32);
                  umma_ss(tb + Cfg::T_D1 + db * Cfg::D1_COLS, kmajor_desc<Cfg::SW>(xs + kb * Cfg::A1_KB_BYTES + ks * 32), bd, idesc_g1a, (kb | ks) != 0);
                  umma_ss(tb + Cfg::T_D1 + db * Cfg::D1_COLS + Cfg::G1N, kmajor_desc<Cfg::SW>(xs + kb * Cfg::A1_KB_BYTES + 128 * Cfg::SW + ks * 32), bd,
                          idesc_g1b, (kb | ks) != 0);
                }
            }
            __syncwarp();
            ring.release();
          }
          if (lane == 0) { LF_TRACE(3, 2 * k) }
          if (elect_one()) {
            umma_commit(smem_u32(&ms.bar_d1_full[db]));
            if (j == NS - (Cfg::PAIR ? 2 : 1)) umma_commit(smem_u32(&ms.bar_x_empty[ab]));
          }
          __syncwarp();
        }
        if (k >= 2) {
          // ---- GEMM-2 of slice g = k-2: D2[tile & 1] += A2[g&1] W2_j^T ----
          const int g = k - 2, j = j2, it = it2, ob = it % Cfg::ND2, ab2 = g & 1;
          if (j == 0) { mbar_wait(smem_u32(&ms.bar_d2_empty[ob]), ((it / Cfg::ND2) & 1) ^ 1); }
          mbar_wait(smem_u32(&ms.bar_a2_full[ab2]), (g >> 1) & 1);
          tc_fence_after();
          const uint32_t wst = ring.acquire();
          if (elect_one()) {
            const uint64_t ad = kmajor_desc<Cfg::SWH>(smem_u32(smem + Cfg::S_A2 + ab2 * Cfg::A2_BYTES)), bd = kmajor_desc<Cfg::SWH>(wst);
#pragma unroll
            for (int ks = 0; ks < SL / 16; ++ks) umma_ss(tb + Cfg::T_D2 + ob * C, ad + 2 * ks, bd + 2 * ks, idesc_g2, (j | ks) != 0);
          }
          __syncwarp();
          ring.release();
          if (lane == 0) { LF_TRACE(3, 2 * (k - 2) + 1) }
          if (elect_one()) {
            umma_commit(smem_u32(&ms.bar_a2_empty[ab2]));
            if (j == NS - 1) umma_commit(smem_u32(&ms.bar_d2_full[ob]));
          }
          __syncwarp();
          if (++j2 == NS) { j2 = 0; ++it2; }
        }
        if (k < total && ++j1 == NS) { j1 = 0; ++it1; }
      }
    }
    // (warps 14, 15 are spares: the LayerNorm statistics are computed by the epilogue warps, see below)
  } else if (wg == 2 || wg == 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // ============================== E1: two epilogue groups (warps 8-11 and 16-19), TMEM lane quadrant q ==============================
    // Each group takes half of the 32-column pieces of a hidden slice (and, at a tile's first slice, a share of the
    // LayerNorm statistics).  The output epilogue E2 runs on the conv warps, which have the slack.
    const int grp = (wg == 4) ? 1 : 0;
    const int q = warp & 3;
    const int t4 = lane >> 2, tq = lane & 3;
    const int m = lane >> 3, rr = lane & 7;
    const uint32_t halo0 = smem_u32(smem + Cfg::S_HALO);
    const uint32_t b1_s = smem_u32(smem + Cfg::S_B1), stats_s = smem_u32(smem + Cfg::S_STATS);
    GeluH2 gelu;
    gelu.init();

    // E1 pieces of 32 hidden columns (16 TMEM lanes each).  Fragment f < 2: rows 32q + 16f of the M=128 part; f == 2: rows
    // 128 + 16q of the M=64 part (its 64 rows sit 16 per lane quadrant).  SL = 64: 6 pieces, 3 per group.  SL = 32: 3 pieces:
    // group g takes fragment g whole and one 16-column half of fragment 2.
    constexpr int HP = SL / 32;
    int j = 0, it = 0;
    for (int k = 0; k < total; ++k, j = (j + 1 == NS ? 0 : j + 1), it += (j == 0)) {
      const int db = k & 1, sb = it & 1;                                              // db: halo buffer
      const int d1b = Cfg::PAIR ? ((k >> 1) & 1) : (k & 1), d1use = Cfg::PAIR ? (k >> 2) : (k >> 1);   // D1 accumulator buffer
      if (j == 0) {
        // ---- LayerNorm statistics of the new tile's 192 rows, 24 rows per epilogue warp, straight from the landed A tile.
        // One pass: sums of d = x - x0 and d^2 with x0 = the row's first element (shifting by a value of the row keeps
        // E[d^2] - E[d]^2 free of cancellation).  Row validity (inside the image) is folded in: rstd = -1 marks a row whose
        // hidden activations must be zero (the conv's zero padding of h1, model.py:659). ----
        const int ab = it % Cfg::NA1;
        mbar_wait(smem_u32(&ms.bar_x_full[ab]), (it / Cfg::NA1) & 1);
        if (q == 0 && lane == 0) { LF_TRACE(4, 2 * it) }
        constexpr int LPR = Cfg::SW / 16;                    // lanes per row inside a k-block (16-byte vectors): 8 / 4 / 2
        constexpr int RPP = 32 / LPR;                        // rows per pass: 4 / 8 / 16
        constexpr int NPASS = (24 + RPP - 1) / RPP;
        const int sub = lane % LPR, rin = lane / LPR;
        const int ew = grp * 4 + q;
        const int tile = tile_of(it);
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y;
        const uint32_t xs = smem_u32(smem + Cfg::S_A1 + ab * Cfg::A1_BYTES);
        float2* st = reinterpret_cast<float2*>(smem + Cfg::S_STATS) + sb * 192;
        float s1[NPASS], s2[NPASS], x0[NPASS];
#pragma unroll
        for (int u = 0; u < NPASS; ++u) {
          const int rl = u * RPP + rin;                      // row inside this warp's 24
          const int r = ew * 24 + (rl < 24 ? rl : 23);
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int kb = 0; kb < Cfg::KB1; ++kb) {
            float v[8];
            unpack8(lds128(xs + kb * Cfg::A1_KB_BYTES + swz<Cfg::SW>(r, sub * 16)), v);
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
            const int rl = u * RPP + rin;
            if (rl < 24) {
              const int r = ew * 24 + rl;
              const int hy = r / 18, hx = r - hy * 18;
              const int y = ty * 8 - 1 + hy, x = tx * 16 - 1 + hx;
              const bool valid = (r < 180) && (y >= 0) && (y < a.H) && (x >= 0) && (x < a.W);
              float rstd = 1.0f, nm = 0.0f;
              if (a.has_ln) {
                const float md = s1[u] * (1.0f / C);                        // mean - x0
                const float var = fmaxf(s2[u] * (1.0f / C) - md * md, 0.f);
                rstd = rsqrtf(var + a.ln_eps);
                nm = -(x0[u] + md) * rstd;
              }
              st[r] = valid ? make_float2(rstd, nm) : make_float2(-1.0f, 0.0f);
            }
          }
        }
        mbar_arrive(smem_u32(&ms.bar_x_empty[ab]));          // this thread is done reading the raw tile
        mbar_arrive(smem_u32(&ms.bar_st_full[sb]));          // ... and has published its statistics (release)
        mbar_wait(smem_u32(&ms.bar_st_full[sb]), (it >> 1) & 1);             // all eight warps' rows are in
        if (q == 0 && lane == 0) { LF_TRACE(4, 2 * it + 1) }
      }
      mbar_wait(smem_u32(&ms.bar_h_empty[db]), ((k >> 1) & 1) ^ 1);          // conv finished reading this halo buffer (slice k-2)
      mbar_wait(smem_u32(&ms.bar_d1_full[d1b]), d1use & 1);
      tc_fence_after();
      if (q == 0 && lane == 0) { LF_TRACE(grp, 2 * k) }
      const uint32_t hb = halo0 + db * Cfg::HALO_BYTES;
      const uint32_t bsl = b1_s + j * SL * 4, csl = b1_s + (a.hidden + j * SL) * 4;
      auto frag_lanes = [&](int f) { return (uint32_t)(q * 32 + (f == 1 ? 16 : 0)) << 16; };
      auto frag_row0 = [&](int f) { return (f < 2) ? q * 32 + f * 16 : 128 + q * 16; };
      // one piece: NBP column blocks of 8 starting at column c0 of fragment f (values already in v)
      auto do_piece = [&](const uint32_t* v, int f, int c0, auto nbp_tag) {
        constexpr int NBP = decltype(nbp_tag)::value;
        const int row0 = frag_row0(f);
        const float2 sa = lds64f(stats_s + (sb * 192 + row0 + t4) * 8), sbb = lds64f(stats_s + (sb * 192 + row0 + t4 + 8) * 8);
        // out-of-image / padding rows are zeroed with a bit mask: a select would be compiled into a branch around each GELU
        // chain and serialise the independent chains of a piece
        const uint32_t ma = sa.x > 0.f ? 0xffffffffu : 0u, mb = sbb.x > 0.f ? 0xffffffffu : 0u;
        const f2 ra = f2_pack(sa.x, sa.x), na = f2_pack(sa.y, sa.y), rb = f2_pack(sbb.x, sbb.x), nb = f2_pack(sbb.y, sbb.y);
        uint32_t pk[2 * NBP];                      // f16x2: the hidden map lives in half precision on chip
#pragma unroll
        for (int i = 0; i < NBP; ++i) {
          const int n = c0 + 8 * i + 2 * tq;
          const float2 b2 = lds64f(bsl + n * 4), c2 = lds64f(csl + n * 4);
          const f2 bbv = f2_pack(b2.x, b2.y), ccv = f2_pack(c2.x, c2.y);
          const f2 d0 = f2_pack(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]));
          const f2 d1 = f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          const f2 x0 = f2_fma(d0, ra, f2_fma(na, ccv, bbv));
          const f2 x1 = f2_fma(d1, rb, f2_fma(nb, ccv, bbv));
          pk[2 * i] = gelu(h2_from_f2(x0)) & ma;
          pk[2 * i + 1] = gelu(h2_from_f2(x1)) & mb;
        }
        const int row = row0 + (m & 1) * 8 + rr;
        if (NBP == 4) {
#pragma unroll
          for (int i2 = 0; i2 < 2; ++i2)
            stsm_x4(hb + swz<Cfg::SWH>(row, (c0 / 8 + 2 * i2 + (m >> 1)) * 16), pk[4 * i2], pk[4 * i2 + 1], pk[4 * i2 + 2], pk[4 * i2 + 3]);
        } else {
          stsm_x4(hb + swz<Cfg::SWH>(row, (c0 / 8 + (m >> 1)) * 16), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      // columns of this slice inside the accumulator buffer: M=128 part at +0 (PAIR: + 32*(k&1)), M=64 part G1N further
      const uint32_t tcol = tb + Cfg::T_D1 + d1b * Cfg::D1_COLS + (Cfg::PAIR ? (k & 1) * SL : 0);
      if (SL == 64) {
        // pieces p = 3*grp .. 3*grp+2 of (f = p / 2, half = p % 2), software-pipelined TMEM loads
        uint32_t v[2][16];
        const int p0 = 3 * grp;
        auto paddr = [&](int p) { const int f = p / HP; return tcol + frag_lanes(f) + (f == 2 ? Cfg::G1N : 0) + (p % HP) * 32; };
        tmem_ld_16x256b_x4(paddr(p0), v[0]);
#pragma unroll
        for (int pi = 0; pi < 3; ++pi) {
          tmem_wait_ld();
          if (pi + 1 < 3) tmem_ld_16x256b_x4(paddr(p0 + pi + 1), v[(pi + 1) & 1]);
          const int p = p0 + pi;
          do_piece(v[pi & 1], p / HP, (p % HP) * 32, std::integral_constant<int, 4>());
        }
      } else {
        uint32_t v0[16], v1[8];
        tmem_ld_16x256b_x4(tcol + frag_lanes(grp), v0);                                   // fragment grp, 32 columns
        tmem_ld_16x256b_x2(tcol + frag_lanes(2) + Cfg::G1N + grp * 16, v1);               // fragment 2, columns 16*grp .. +16
        tmem_wait_ld();
        do_piece(v0, grp, 0, std::integral_constant<int, 4>());
        do_piece(v1, 2, grp * 16, std::integral_constant<int, 2>());
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_d1_empty[d1b]));
      mbar_arrive(smem_u32(&ms.bar_h_full[db]));
      if (q == 0 && lane == 0) { LF_TRACE(grp, 2 * k + 1) }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    // ============================== conv warps 0-7 ==============================
    // SL = 64: warp = channel octet, lanes = 16 columns x 2 row halves, 4 output rows per lane (6 halo rows).
    // SL = 32: warp = (octet, row half), lanes = 16 columns x 2 row pairs, 2 output rows per lane (4 halo rows).
    constexpr int NV = SL / 8;                     // octets per slice
    constexpr int RPL = (SL == 64) ? 4 : 2;        // output rows per lane
    const int v = warp % NV;
    const int cx = lane & 15;
    const int rbase = (SL == 64) ? (lane >> 4) * 4 : (warp / NV) * 4 + (lane >> 4) * 2;     // first output row of this lane
    const uint32_t halo0 = smem_u32(smem + Cfg::S_HALO), taps0 = smem_u32(smem + Cfg::S_TAPS);
    GeluH2 gelu;
    gelu.init();
    // ---- E2, the output epilogue (D2 + b2 -> staging -> coalesced store with the residual added), also lives on these warps:
    // two groups of four warps (one per TMEM lane quadrant), each drains half of the output columns through its own staging
    // tile (C >= 64; narrower outputs are drained by group 0 alone). ----
    const int grp = warp >> 2, q = warp & 3, et = tid & 127, tq = lane & 3;
    const uint32_t stage_s = smem_u32(smem + Cfg::S_STAGE) + grp * (128 * Cfg::STAGE_PITCH), b2_s = smem_u32(smem + Cfg::S_B2);
    auto grp_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory"); };
    constexpr bool SPLIT_E2 = (C >= 64);
    constexpr int E2_COLS = SPLIT_E2 ? C / 2 : C;             // output columns this group drains
    constexpr int E2_PASS = (E2_COLS >= 32) ? 32 : 16;        // columns per staging pass
    constexpr int VPR = E2_PASS / 8;                          // 16-byte vectors per staged row
    constexpr int VPT = VPR;                                  // vectors per thread per pass (128 rows x VPR vectors / 128 threads)
    int* row_out = ms.row_out + grp * 128;
    auto epilogue2 = [&](int it) {
      if (!SPLIT_E2 && grp == 1) return;
      const int ob = it % Cfg::ND2;
      const int tile = tile_of(it);
      if (tid == 0) { LF_TRACE(6, 3 * it) }
      {
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / (a.tiles_x * a.tiles_y);
        const int y = ty * 8 + (et >> 4), x = tx * 16 + (et & 15);
        row_out[et] = (x < a.W) ? ((b * a.H + y) * a.W + x) : -1;
      }
      grp_bar();
      // this thread copies out vectors i = et + p*128 of every pass: row i / VPR, vector i % VPR (fixed over the passes)
      int tok[VPT];
#pragma unroll
      for (int p2 = 0; p2 < VPT; ++p2) tok[p2] = row_out[(et + p2 * 128) / VPR];
      mbar_wait(smem_u32(&ms.bar_d2_full[ob]), (it / Cfg::ND2) & 1);
      tc_fence_after();
      if (tid == 0) { LF_TRACE(6, 3 * it + 1) }
      const int cbase = grp * (SPLIT_E2 ? C / 2 : 0);
#pragma unroll 1
      for (int sc = 0; sc < E2_COLS; sc += E2_PASS) {
        // residual vectors of this pass: issued first so that their (L2) latency hides under the TMEM read + staging below
        uint4 rv[VPT][2];
        if (a.resid != nullptr) {
#pragma unroll
          for (int p2 = 0; p2 < VPT; ++p2) {
            const int c = cbase + sc + ((et + p2 * 128) % VPR) * 8;
            const size_t off = (size_t)(tok[p2] < 0 ? 0 : tok[p2]) * a.resid_stride + c;
            if (a.resid_fp32) {
              rv[p2][0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.resid) + off));
              rv[p2][1] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.resid) + off + 4));
            } else {
              rv[p2][0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.resid) + off));
            }
          }
        }
        constexpr int NB = E2_PASS / 8;
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          const int row16 = q * 32 + hl * 16;
          uint32_t v[4 * NB];
          const uint32_t ta = tb + ((uint32_t)row16 << 16) + Cfg::T_D2 + ob * C + cbase + sc;
          if (NB == 4) tmem_ld_16x256b_x4(ta, v); else tmem_ld_16x256b_x2(ta, v);
          f2 bb[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) { const float2 b2 = lds64f(b2_s + (cbase + sc + 8 * i + 2 * tq) * 4); bb[i] = f2_pack(b2.x, b2.y); }
          tmem_wait_ld();
          uint32_t pk[2 * NB];
          frag_bias_act_pack<NB, false>(v, bb, pk);
          stage_frag<NB>(stage_s, Cfg::STAGE_PITCH, row16, 0, pk);
        }
        if (sc + E2_PASS >= E2_COLS) { tc_fence_before(); mbar_arrive(smem_u32(&ms.bar_d2_empty[ob])); }
        grp_bar();
#pragma unroll
        for (int p2 = 0; p2 < VPT; ++p2) {
          if (tok[p2] < 0) continue;
          const int i = et + p2 * 128;
          const int row = i / VPR, vec = i % VPR;
          float f[8];
          unpack8(lds128(stage_s + row * Cfg::STAGE_PITCH + vec * 16), f);
          const int c = cbase + sc + vec * 8;
          if (a.resid != nullptr) {
            if (a.resid_fp32) {
              const float4 r0 = *reinterpret_cast<const float4*>(&rv[p2][0]), r1 = *reinterpret_cast<const float4*>(&rv[p2][1]);
              f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w; f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
            } else {
              float r[8];
              unpack8(rv[p2][0], r);
#pragma unroll
              for (int k2 = 0; k2 < 8; ++k2) f[k2] += r[k2];
            }
          }
          if (a.out_fp32) {
            float* op = reinterpret_cast<float*>(a.out) + (size_t)tok[p2] * a.out_stride + c;
            *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.out) + (size_t)tok[p2] * a.out_stride + c) = pack8(f);
          }
          if (a.out_b != nullptr) *reinterpret_cast<uint4*>(a.out_b + (size_t)tok[p2] * a.out_b_stride + c) = pack8(f);
        }
        grp_bar();
      }
      if (tid == 0) { LF_TRACE(6, 3 * it + 2) }
    };

    int cj = 0, cit = 0;                                    // (slice, tile) of the conv iteration
    for (int k = 0; k < total; ++k) {
      const int hbi = k & 1, tbuf = k % Cfg::NTAP;
      const uint32_t sH = halo0 + hbi * Cfg::HALO_BYTES;
      const uint32_t sW = taps0 + tbuf * Cfg::TAP_BYTES + v * 16;       // tap t at + t * SL * 2 (f16x2 per channel pair)
      constexpr uint32_t tstride = SL * 2;
      mbar_wait(smem_u32(&ms.bar_tap_full[tbuf]), (k / Cfg::NTAP) & 1);
      h2 acc[RPL][4];                                // half2 accumulators: 2 channels per register
      {
        const uint4 b0 = lds128(sW + 9 * tstride);
#pragma unroll
        for (int o = 0; o < RPL; ++o) { acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w; }
      }
      mbar_wait(smem_u32(&ms.bar_h_full[hbi]), (k >> 1) & 1);
      if (tid == 0) { LF_TRACE(2, 4 * k) }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        uint4 h[RPL + 2];
#pragma unroll
        for (int r = 0; r < RPL + 2; ++r) {
          const int t = (rbase + r) * 18 + cx + dx;
          h[r] = lds128(sH + swz<Cfg::SWH>(t, v * 16));
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint4 w = lds128(sW + (ky * 3 + dx) * tstride);            // warp-uniform address: broadcast
#pragma unroll
          for (int o = 0; o < RPL; ++o) {
            acc[o][0] = h2_fma(h[o + ky].x, w.x, acc[o][0]);
            acc[o][1] = h2_fma(h[o + ky].y, w.y, acc[o][1]);
            acc[o][2] = h2_fma(h[o + ky].z, w.z, acc[o][2]);
            acc[o][3] = h2_fma(h[o + ky].w, w.w, acc[o][3]);
          }
        }
      }
      mbar_arrive(smem_u32(&ms.bar_h_empty[hbi]));                 // halo buffer free for E1 of slice k+2
      mbar_arrive(smem_u32(&ms.bar_tap_empty[tbuf]));
      if (tid == 0) { LF_TRACE(2, 4 * k + 1) }
      mbar_wait(smem_u32(&ms.bar_a2_empty[hbi]), ((k >> 1) & 1) ^ 1);   // GEMM-2 of slice k-2 has consumed this A buffer
      if (tid == 0) { LF_TRACE(2, 4 * k + 2) }
      const uint32_t sA = smem_u32(smem + Cfg::S_A2 + hbi * Cfg::A2_BYTES);
#pragma unroll
      for (int o = 0; o < RPL; ++o) {
        uint4 pk;
        pk.x = gelu(acc[o][0]);
        pk.y = gelu(acc[o][1]);
        pk.z = gelu(acc[o][2]);
        pk.w = gelu(acc[o][3]);
        const int rw = (rbase + o) * 16 + cx;
        sts128(sA + swz<Cfg::SWH>(rw, v * 16), pk);
      }
      fence_async_smem();
      mbar_arrive(smem_u32(&ms.bar_a2_full[hbi]));
      if (tid == 0) { LF_TRACE(2, 4 * k + 3) }
      // the previous tile's output epilogue runs one or two slices into this tile (its GEMM-2 chain has drained by then;
      // with a single D2 buffer it must run before this tile's second GEMM-2 can be issued)
      if (cit > 0 && cj == ((NS > 1 && Cfg::ND2 == 2) ? 1 : 0)) epilogue2(cit - 1);
      if (++cj == NS) { cj = 0; ++cit; }
    }
    if (my_tiles > 0) epilogue2(my_tiles - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tb, Cfg::T_ALLOC);
}

}  // namespace lw
