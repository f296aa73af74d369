// leff2.cuh — LeFF part 2: out = resid + GELU(dwconv3x3(h1) + bd) W2^T + b2   (model.py:674-682)
//
// One CTA = an 8 x 16 spatial tile (128 tokens) of one image; 320 threads:
//   warps 0-7  workers : depthwise conv + GELU producing the GEMM A operand, then the epilogue
//   warp  8    producer: W2 chunk images -> smem ring (cp.async.bulk)
//   warp  9    issuer  : tcgen05.mma, accumulator D[128 x C] resident in TMEM for the whole tile
// The hidden dimension is walked in 64-channel slices.  For each slice the 10 x 18 halo'd h1 tile
// is fetched with cp.async (zero-filled outside the image = the conv's zero padding of h1,
// model.py:659) one slice ahead; each worker owns (row-half, column x, channel octet) and slides
// down its column keeping the 3x3 taps of its 8 channels in registers, so every h1 vector is read
// from shared memory 3x (not 9x) and all arithmetic is packed FFMA2 on fp32 pairs.
#pragma once
#include "lewin_common.cuh"
#include "leff.cuh"

namespace lw {

constexpr int kL2Workers = 256;
constexpr int kL2Threads = 320;

struct Leff2Cfg {
  static constexpr int STAGES = 4;
  static constexpr int HALO_TOK = 180;                          // 10 x 18
  static constexpr int HALO_BYTES = 23040;                      // 180 x 128 B (64 bf16 channels)
  static constexpr int S_A = 0;                                 // 2 x 16 KB A k-block buffers
  static constexpr int S_HALO = 2 * 16384;                      // 2 x 23040 (padded to 23552)
  static constexpr int S_WD = S_HALO + 2 * 23552;               // 2 x (10 x 64 fp32) = 2 x 2560
  static constexpr int S_RING = S_WD + 2 * 2560 + 1024;         // keep 1024 B alignment: 32768+47104+6144 = 86016
  static constexpr int S_MISC = S_RING + STAGES * kStageBytes;
  static constexpr int SMEM_BYTES = S_MISC + 1024 + 1024;
};
static_assert(Leff2Cfg::S_RING % 1024 == 0, "ring alignment");

__device__ __forceinline__ void l2_worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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

__global__ void __launch_bounds__(kL2Threads, 1) leff2_kernel(const AStreamArgs a, const int t_alloc) {
  using Cfg = Leff2Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmMisc& ms = *reinterpret_cast<GemmMisc*>(smem + Cfg::S_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = a.K / 64;
  const int NC = a.N / a.nch;
  // tile origin: tiles never straddle images (H % 8 == 0, TH == 8)
  const int tiles_y = a.H / 8;
  const int tx = blockIdx.x % a.tiles_x;
  const int ty = (blockIdx.x / a.tiles_x) % tiles_y;
  const int b = blockIdx.x / (a.tiles_x * tiles_y);
  const int y0 = ty * 8, x0 = tx * 16;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(smem_u32(&ms.bar_full[s]), 1); mbar_init(smem_u32(&ms.bar_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&ms.bar_a_full[i]), kL2Workers); mbar_init(smem_u32(&ms.bar_a_empty[i]), 1); }
    mbar_init(smem_u32(&ms.bar_d_full[0]), 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&ms.tmem_base), t_alloc);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ms.tmem_base;
  const uint32_t chunk_bytes = a.nch * 128;

  if (warp == 8) {
    if (lane == 0) {
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      for (int kb = 0; kb < KB; ++kb)
        for (int nc = 0; nc < NC; ++nc)
          ring.load(a.w_img + (size_t)(kb * NC + nc) * chunk_bytes, chunk_bytes);
    }
  } else if (warp == 9) {
    {   // issuer warp (warp-uniform; one elected lane issues so descriptors stay in uniform registers)
      Ring ring{smem_u32(smem + Cfg::S_RING), smem_u32(&ms.bar_full[0]), smem_u32(&ms.bar_empty[0]), Cfg::STAGES, 0};
      const uint32_t idesc = make_idesc_bf16(128, a.nch);
      const uint32_t ring_base = smem_u32(smem + Cfg::S_RING);
      const uint64_t b_desc0 = kmajor_desc<128>(ring_base);
      for (int kb = 0; kb < KB; ++kb) {
        const int ab = kb & 1;
        mbar_wait(smem_u32(&ms.bar_a_full[ab]), (kb >> 1) & 1);
        tc_fence_after();
        const uint64_t ad = kmajor_desc<128>(smem_u32(smem + Cfg::S_A + ab * 16384));
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
      if (elect_one()) umma_commit(smem_u32(&ms.bar_d_full[0]));
      __syncwarp();
    }
  } else {
    // ============================== workers ==============================
    // one warp per channel octet: the 3x3 taps of the octet are warp-uniform (broadcast LDS at the point of
    // use, no tap registers); the 32 lanes are 16 tile columns x 2 row halves.
    const int v = warp;                  // channel octet inside the 64-channel slice
    const int cx = lane & 15;            // tile column
    const int hf = lane >> 4;            // row half: output rows hf*4 .. hf*4+3
    const bf16* __restrict__ src = a.src + (size_t)b * a.H * a.W * a.K;

    // Per-thread prefetch descriptors (independent of the slice): up to 6 halo vectors + 1 tap vector.
    // A halo vector = 16 B (8 channels) of one halo token; zero-filled (src-size 0) outside the image,
    // which is exactly the conv's zero padding of h1.
    uint32_t pf_soff[6];      // smem byte offset inside the halo buffer
    uint32_t pf_goff[6];      // global element offset of the token (without the slice term), or 0xffffffff
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int idx = tid + i * kL2Workers;
      const int t = idx >> 3, vv = idx & 7;
      const int y = y0 - 1 + t / 18, x = x0 - 1 + t % 18;
      const bool in = (idx < Cfg::HALO_TOK * 8);
      const bool ok = in && (y >= 0) && (y < a.H) && (x >= 0) && (x < a.W);
      pf_soff[i] = in ? (uint32_t)(t * 128 + ((vv ^ (t & 7)) * 16)) : 0xffffffffu;   // 16 B chunk XOR-swizzled by token
      pf_goff[i] = ok ? (uint32_t)((y * a.W + x) * a.K + vv * 8) : 0xffffffffu;
    }
    // taps: 10 rows (9 taps + bias) x 64 fp32 = 160 16-byte vectors, one per thread for tid < 160
    const int tw_row = tid >> 4, tw_vec = tid & 15;
    const float* tw_src = (tw_row < 9) ? a.wd + (size_t)tw_row * a.K + tw_vec * 4 : a.bd + tw_vec * 4;
    const uint32_t halo0 = smem_u32(smem + Cfg::S_HALO), wd0 = smem_u32(smem + Cfg::S_WD);

    auto prefetch = [&](int kb) {
      const uint32_t hb = halo0 + (kb & 1) * 23552;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (pf_soff[i] != 0xffffffffu) {
          const bool ok = pf_goff[i] != 0xffffffffu;
          cp_async16(hb + pf_soff[i], ok ? src + pf_goff[i] + kb * 64 : src, ok ? 16u : 0u);
        }
      }
      if (tid < 160) cp_async16(wd0 + (kb & 1) * 2560 + tid * 16, tw_src + kb * 64, 16u);
      cp_async_commit();
    };

    LW_TRACE_STMT(const bool trw = (a.dbg & 16) && blockIdx.x == 0 && tid == 0 && a.trace != nullptr; int tw = 0;)
    LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
    prefetch(0);
    for (int kb = 0; kb < KB; ++kb) {
      const int ab = kb & 1;
      cp_async_wait_all();
      l2_worker_bar();                         // slice kb halo + taps visible; buffers of slice kb-1 free
      LW_TRACE_STMT(if (trw && kb < 6) a.trace[tw++] = clock64();)
      if (kb + 1 < KB) prefetch(kb + 1);
      LW_TRACE_STMT(if (trw && kb < 6) a.trace[tw++] = clock64();)
      const uint32_t sH = halo0 + ab * 23552;
      const uint32_t sW = wd0 + ab * 2560;
      // accumulators start at the conv bias of this octet
      f2 acc[4][4];
      {
        const float4 b0 = lds128f(sW + (9 * 64 + v * 8) * 4);
        const float4 b1 = lds128f(sW + (9 * 64 + v * 8 + 4) * 4);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          acc[o][0] = f2_pack(b0.x, b0.y); acc[o][1] = f2_pack(b0.z, b0.w);
          acc[o][2] = f2_pack(b1.x, b1.y); acc[o][3] = f2_pack(b1.z, b1.w);
        }
      }
      // column by column (dx): pull the 6 halo rows of that column into registers once, then apply the three
      // taps (ky) that touch them; halo row r feeds output row o = r - ky.
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        f2 h[6][4];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const int t = (hf * 4 + r) * 18 + cx + dx;
          const uint4 raw = lds128(sH + t * 128 + ((v ^ (t & 7)) * 16));
          h[r][0] = bf2_to_f2(raw.x); h[r][1] = bf2_to_f2(raw.y); h[r][2] = bf2_to_f2(raw.z); h[r][3] = bf2_to_f2(raw.w);
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4 w0 = lds128f(sW + ((ky * 3 + dx) * 64 + v * 8) * 4);        // warp-uniform address: broadcast
          const float4 w1 = lds128f(sW + ((ky * 3 + dx) * 64 + v * 8 + 4) * 4);
          const f2 wa = f2_pack(w0.x, w0.y), wb = f2_pack(w0.z, w0.w), wc = f2_pack(w1.x, w1.y), wd = f2_pack(w1.z, w1.w);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            acc[o][0] = f2_fma(h[o + ky][0], wa, acc[o][0]);
            acc[o][1] = f2_fma(h[o + ky][1], wb, acc[o][1]);
            acc[o][2] = f2_fma(h[o + ky][2], wc, acc[o][2]);
            acc[o][3] = f2_fma(h[o + ky][3], wd, acc[o][3]);
          }
        }
      }
      LW_TRACE_STMT(if (trw && kb < 6) a.trace[tw++] = clock64();)
      // A buffer ab must have been consumed by the MMAs of slice kb-2
      mbar_wait(smem_u32(&ms.bar_a_empty[ab]), ((kb >> 1) & 1) ^ 1);
      LW_TRACE_STMT(if (trw && kb < 6) a.trace[tw++] = clock64();)
      const uint32_t sA = smem_u32(smem + Cfg::S_A + ab * 16384);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        uint4 pk;
        pk.x = f2_to_bf2(gelu2(acc[o][0]));
        pk.y = f2_to_bf2(gelu2(acc[o][1]));
        pk.z = f2_to_bf2(gelu2(acc[o][2]));
        pk.w = f2_to_bf2(gelu2(acc[o][3]));
        const int rr = (hf * 4 + o) * 16 + cx;
        sts128(sA + swz<128>(rr, v * 16), pk);
      }
      fence_async_smem();
      mbar_arrive(smem_u32(&ms.bar_a_full[ab]));
      LW_TRACE_STMT(if (trw && kb < 6) a.trace[tw++] = clock64();)
    }
    LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)

    // ---------------- epilogue: + bias -> bf16 -> staging tile (the halo buffers are free now) ->
    // coalesced copy-out with the residual added on the way (warp w: lane quadrant w&3, column half w>>2)
    const int r = (warp & 3) * 32 + lane;
    if (warp < 4) {
      const int y = y0 + (r >> 4), x = x0 + (r & 15);
      ms.row_tok[r] = (x < a.W) ? (int)(((size_t)b * a.H + y) * a.W + x) : -1;
    }
    const uint32_t stage_s = halo0;
    const int sub_cols = a.N < 128 ? a.N : 128;
    const int pitch = sub_cols * 2 + 16;
    int sub_log2 = 4;
    while ((1 << sub_log2) < sub_cols) ++sub_log2;
    mbar_wait(smem_u32(&ms.bar_d_full[0]), 0);
    tc_fence_after();
    const int row16 = (warp & 3) * 32 + (warp >> 2) * 16;     // this warp's 16 TMEM lanes / tile rows
    for (int sc = 0; sc < a.N; sc += 128) {
      const uint32_t tcol = tb + ((uint32_t)row16 << 16) + sc;
      if (sub_cols == 128) { epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0); epi_cols<8>(tcol + 64, a.bias + sc + 64, stage_s, pitch, row16, 64); }
      else if (sub_cols == 64) epi_cols<8>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      else if (sub_cols == 32) epi_cols<4>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      else epi_cols<2>(tcol, a.bias + sc, stage_s, pitch, row16, 0);
      l2_worker_bar();
      store_staged_rows(stage_s, pitch, sub_log2, ms.row_tok, a.out, a.resid, (size_t)a.N, sc, tid, kL2Workers);
      l2_worker_bar();
      LW_TRACE_STMT(if (trw) a.trace[tw++] = clock64();)
    }
    LW_TRACE_STMT(if (trw) a.trace[tw++] = -1;)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, t_alloc);
}

}  // namespace lw
