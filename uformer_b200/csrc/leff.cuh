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
  uint64_t bar_full[4], bar_empty[4];
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
};

template <int K>
struct AResCfg {
  static constexpr int KB = (K + 63) / 64;
  static constexpr int STAGES = 4;
  static constexpr int S_X = 0;
  static constexpr int S_RING = KB * 16384;
  static constexpr int S_MISC = S_RING + STAGES * kStageBytes;
  static constexpr int SMEM_BYTES = S_MISC + 1024 + 1024;
  static constexpr int T_ALLOC = 256;      // two 128-column accumulators
};
static_assert(sizeof(GemmMisc) <= 1024, "misc too large");

template <int K, int EPI>
__global__ void __launch_bounds__(kThreads8, (K <= 128) ? 2 : 1) ares_kernel(const AResArgs a) {
  using Cfg = AResCfg<K>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmMisc& ms = *reinterpret_cast<GemmMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int NC = a.n_total / a.nch;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    mbar_init(smem_u32(&ms.bar_a_ready), kWorkers8);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_d_full[i]), 1); mbar_init(smem_u32(&ms.bar_d_empty[i]), kWorkers8); }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&ms.tmem_base), Cfg::T_ALLOC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const uint32_t sX = smem_u32(smem + Cfg::S_X);
  const uint32_t chunk_bytes = a.nch * 128;

  if (warp == 8) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      for (int nc = 0; nc < NC; ++nc)
        for (int kb = 0; kb < Cfg::KB; ++kb)
          ring.load(a.w_img + (size_t)(nc * Cfg::KB + kb) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 9) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      const uint32_t idesc = make_idesc_bf16(128, a.nch);
      mbar_wait(smem_u32(&ms.bar_a_ready), 0);
      tc_fence_after();
      for (int nc = 0; nc < NC; ++nc) {
        const int buf = nc & 1;
        mbar_wait(smem_u32(&ms.bar_d_empty[buf]), ((nc >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          const uint32_t wst = ring.acquire();
          constexpr int KS = (K >= 64) ? 4 : K / 16;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
            umma_ss(tb + buf * 128, kmajor_desc<128>(sX + kb * 16384 + ks * 32), kmajor_desc<128>(wst + ks * 32), idesc,
                    (kb | ks) != 0);
          ring.release();
        }
        umma_commit(smem_u32(&ms.bar_d_full[buf]));
      }
    }
  } else {
    // 8 worker warps: stage 16 rows each; epilogue: lane quadrant warp&3, column half warp>>2
    if (tid < 128) {
      const int row0 = tile * 128 + tid;
      ms.row_tok[tid] = (row0 < a.n_rows) ? row0 : -1;
    }
    worker_bar8();
    stage_rows_ln<K, 8>(smem + Cfg::S_X, a.x, ms.row_tok, a.ln_w, a.ln_b, a.ln_eps, nullptr);
    fence_async_smem();
    mbar_arrive(smem_u32(&ms.bar_a_ready));
    const int r = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int row = tile * 128 + r;
    const bool valid = row < a.n_rows;
    int ub = 0, uy = 0, ux = 0;      // upsample geometry of this input token
    if (EPI == 1 && valid) { ub = row / (a.H * a.W); int t = row % (a.H * a.W); uy = t / a.W; ux = t % a.W; }
    for (int nc = 0; nc < NC; ++nc) {
      const int buf = nc & 1;
      mbar_wait(smem_u32(&ms.bar_d_full[buf]), (nc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 16; c0 < a.nch; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(tb + lane_base + buf * 128 + c0, v);
        tmem_wait_ld();
        if (valid) {
          const int n0 = nc * a.nch + c0;
          uint4 o0, o1;
          if (EPI == 0) {
            const float4* bp = reinterpret_cast<const float4*>(a.bias + n0);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = __ldg(bp + j);
              const f2 x0 = f2_add(f2_pack(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), f2_pack(b4.x, b4.y));
              const f2 x1 = f2_add(f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), f2_pack(b4.z, b4.w));
              pk[2 * j] = f2_to_bf2(gelu2(x0));
              pk[2 * j + 1] = f2_to_bf2(gelu2(x1));
            }
            o0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            uint4* op = reinterpret_cast<uint4*>(a.out + (size_t)row * a.n_total + n0);
            op[0] = o0;
            op[1] = o1;
          } else {
            const int q = n0 / a.Cout, co = n0 % a.Cout;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + co + j));
              f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
            const size_t orow = ((size_t)ub * (2 * a.H) + (2 * uy + (q >> 1))) * (2 * a.W) + (2 * ux + (q & 1));
            uint4* op = reinterpret_cast<uint4*>(a.out + orow * a.out_stride + co);
            op[0] = pack8(f);
            op[1] = pack8(f + 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ms.bar_d_empty[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, Cfg::T_ALLOC);
}

// =================================================================================================
// A-streamed skeleton.  PROD = 0: depthwise 3x3 conv + GELU over a 2D token tile (LeFF part 2)
//                       PROD = 1: im2col gather of a 4x4/stride-2 convolution (Downsample)
// =================================================================================================
struct AStreamArgs {
  const bf16* src;        // PROD 0: h1 (B,H,W,K)   PROD 1: x (B,H,W,Cin)
  int B, H, W;            // geometry of src
  int K;                  // PROD 0: hidden; PROD 1: 16*Cin
  int Cin;                // PROD 1 only
  const float* wd;        // PROD 0: (9, K) taps
  const float* bd;        // PROD 0: (K)
  const uint8_t* w_img;   // [K/64][N/nch][nch*128B]
  int N, nch;
  const float* bias;      // (N)
  const bf16* resid;      // (rows, N) or null
  bf16* out;              // (rows, N)
  int TW, TH;             // PROD 0 tile shape (TW*TH == 128)
  int tiles_x;
};

constexpr int kHaloMaxTok = 18 * 10;   // (TW+2)*(TH+2) upper bound for TW in {8,16}: 10*18 = 180

struct AStreamCfg {
  static constexpr int STAGES = 4;
  static constexpr int S_A = 0;                               // 2 x 16 KB A k-block buffers
  static constexpr int S_HALO = 2 * 16384;                    // halo'd h1 slice: 180 tokens x 128 B
  static constexpr int S_WD = S_HALO + 23552;                 // 9 x 64 fp32 taps + 64 bias = 2560 B
  static constexpr int S_RING = S_WD + 3072;
  static constexpr int S_MISC = S_RING + STAGES * kStageBytes;
  static constexpr int SMEM_BYTES = S_MISC + 1024 + 1024;
};

template <int PROD>
__global__ void __launch_bounds__(kThreads, 1) astream_kernel(const AStreamArgs a, const int t_alloc) {
  using Cfg = AStreamCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmMisc& ms = *reinterpret_cast<GemmMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int KB = a.K / 64;
  const int NC = a.N / a.nch;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_a_full[i]), kWorkers); mbar_init(smem_u32(&ms.bar_a_empty[i]), 1); }
    mbar_init(smem_u32(&ms.bar_d_full[0]), 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&ms.tmem_base), t_alloc);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const uint32_t chunk_bytes = a.nch * 128;

  if (warp == 4) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      for (int kb = 0; kb < KB; ++kb)
        for (int nc = 0; nc < NC; ++nc)
          ring.load(a.w_img + (size_t)(kb * NC + nc) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      const uint32_t idesc = make_idesc_bf16(128, a.nch);
      for (int kb = 0; kb < KB; ++kb) {
        const int ab = kb & 1;
        mbar_wait(smem_u32(&ms.bar_a_full[ab]), (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + Cfg::S_A + ab * 16384);
        for (int nc = 0; nc < NC; ++nc) {
          const uint32_t wst = ring.acquire();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tb + nc * a.nch, kmajor_desc<128>(sA + ks * 32), kmajor_desc<128>(wst + ks * 32), idesc, (kb | ks) != 0);
          ring.release();
        }
        umma_commit(smem_u32(&ms.bar_a_empty[ab]));
      }
      umma_commit(smem_u32(&ms.bar_d_full[0]));
    }
  } else {
    const int r = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // ---------------- geometry of this tile ----------------
    int out_row = -1;          // flat output token of tile row r (or -1)
    int gy0 = 0, x0 = 0;       // PROD 0: tile origin in (stacked row, column)
    int db = 0, dy = 0, dx = 0; // PROD 1: output pixel of row r
    if (PROD == 0) {
      const int ty = tile / a.tiles_x, tx = tile % a.tiles_x;
      gy0 = ty * a.TH; x0 = tx * a.TW;
      const int gy = gy0 + r / a.TW, x = x0 + r % a.TW;
      if (gy < a.B * a.H && x < a.W) out_row = gy * a.W + x;
    } else {
      const int Ho = a.H / 2, Wo = a.W / 2;
      const int row = tile * 128 + r;
      if (row < a.B * Ho * Wo) { out_row = row; db = row / (Ho * Wo); int t = row % (Ho * Wo); dy = t / Wo; dx = t % Wo; }
    }
    ms.row_tok[r] = out_row;
    worker_bar();

    for (int kb = 0; kb < KB; ++kb) {
      const int ab = kb & 1;
      uint8_t* sA = smem + Cfg::S_A + ab * 16384;
      if (PROD == 0) {
        // ---- stage the halo'd h1 slice [ (TH+2) x (TW+2) tokens ][64 ch] and this slice's taps ----
        const int HW2 = a.TW + 2, HT = (a.TH + 2) * HW2;
        uint8_t* sH = smem + Cfg::S_HALO;
        float* sW = reinterpret_cast<float*>(smem + Cfg::S_WD);
        for (int idx = tid; idx < HT * 8; idx += kWorkers) {
          const int t = idx >> 3, v = idx & 7;
          const int gy = gy0 - 1 + t / HW2, x = x0 - 1 + t % HW2;
          uint4 val = make_uint4(0, 0, 0, 0);
          // zero padding applies per image: the row above image b's first row belongs to image b-1
          if (gy >= 0 && gy < a.B * a.H && x >= 0 && x < a.W)
            val = __ldg(reinterpret_cast<const uint4*>(a.src + ((size_t)gy * a.W + x) * a.K + kb * 64 + v * 8));
          *reinterpret_cast<uint4*>(sH + t * 128 + v * 16) = val;
        }
        for (int idx = tid; idx < 10 * 64; idx += kWorkers) {
          const int tap = idx >> 6, c = idx & 63;
          sW[idx] = (tap < 9) ? __ldg(a.wd + (size_t)tap * a.K + kb * 64 + c) : __ldg(a.bd + kb * 64 + c);
        }
        worker_bar();
        mbar_wait(smem_u32(&ms.bar_a_empty[ab]), ((kb >> 1) & 1) ^ 1);
        // ---- depthwise conv + GELU: thread handles channel octet v of 8 tile rows ----
        const int v = tid & 7;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
          const int rr = (tid >> 3) + it * 16;
          const int ly = rr / a.TW, lx = rr % a.TW;              // position inside the tile
          const int gy = gy0 + ly;
          const int yimg = gy % a.H;                             // row inside its image
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = sW[9 * 64 + v * 8 + j];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = yimg + ky - 1;
            if (yy < 0 || yy >= a.H) continue;                   // zero padding at the image border
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int t = (ly + ky) * HW2 + (lx + kx);
              float hv[8];
              unpack8(*reinterpret_cast<const uint4*>(sH + t * 128 + v * 16), hv);
              const float* wt = sW + (ky * 3 + kx) * 64 + v * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = fmaf(hv[j], wt[j], acc[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = gelu_erf(acc[j]);
          *reinterpret_cast<uint4*>(sA + swz<128>(rr, v * 16)) = pack8(acc);
        }
      } else {
        // ---- im2col: k index = tap*Cin + ci, 64 consecutive k per k-block ----
        mbar_wait(smem_u32(&ms.bar_a_empty[ab]), ((kb >> 1) & 1) ^ 1);
        const int v = tid & 7;
        const int k0 = kb * 64 + v * 8;
        const int tap = k0 / a.Cin, ci = k0 % a.Cin;
        const int ky = tap >> 2, kx = tap & 3;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
          const int rr = (tid >> 3) + it * 16;
          const int orow = ms.row_tok[rr];
          uint4 val = make_uint4(0, 0, 0, 0);
          if (orow >= 0) {
            const int Ho = a.H / 2, Wo = a.W / 2;
            const int b = orow / (Ho * Wo), t = orow % (Ho * Wo);
            const int iy = 2 * (t / Wo) - 1 + ky, ix = 2 * (t % Wo) - 1 + kx;
            if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
              val = __ldg(reinterpret_cast<const uint4*>(a.src + (((size_t)b * a.H + iy) * a.W + ix) * a.Cin + ci));
          }
          *reinterpret_cast<uint4*>(sA + swz<128>(rr, v * 16)) = val;
        }
      }
      fence_async_smem();
      mbar_arrive(smem_u32(&ms.bar_a_full[ab]));
      if (PROD == 0) worker_bar();   // halo / tap buffers are rewritten by the next slice
    }
    (void)db; (void)dy; (void)dx;

    // ---------------- epilogue: + bias (+ residual) ----------------
    mbar_wait(smem_u32(&ms.bar_d_full[0]), 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < a.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tb + lane_base + c0, v);
      tmem_wait_ld();
      if (out_row >= 0) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __ldg(a.bias + c0 + j);
        if (a.resid != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.resid + (size_t)out_row * a.N + c0);
          float g[8];
          unpack8(__ldg(rp), g);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += g[j];
          unpack8(__ldg(rp + 1), g);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[8 + j] += g[j];
        }
        uint4* op = reinterpret_cast<uint4*>(a.out + (size_t)out_row * a.N + c0);
        op[0] = pack8(f);
        op[1] = pack8(f + 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, t_alloc);
}

}  // namespace lw
