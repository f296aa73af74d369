// lewin_common.cuh — CTA organisation and device helpers shared by the LeWin kernels.
//
// Every GEMM-type kernel processes 128-row tiles (128 tokens) with specialised warps:
//   worker warps (8)  : thread <-> (TMEM lane quadrant = warp & 3, 16-row group / column half = warp >> 2).
//                       They gather / normalise / convolve the A operand into shared memory and run the epilogues
//                       on 16x256b TMEM fragments (stmatrix-staged, coalesced copy-out).
//   producer warp     : lane 0 streams pre-swizzled weight chunk images global -> shared with cp.async.bulk
//                       (UBLKCP) into a ring of 16/32 KB stages guarded by full/empty mbarriers; owns the TMEM allocation.
//   issuer warp       : warp-uniform loop; one elected lane issues tcgen05.mma / tcgen05.commit.
// The persistent LeFF-2 kernel adds a dedicated epilogue warpgroup (leff2.cuh).
#pragma once
#include "umma.cuh"

// Profiling aids (per-CTA clock64 timelines, LW_DEBUG knobs) are compiled in only with -DLW_TRACE: their
// branches split the hot loops into small basic blocks and cost real time in production builds.
#ifdef LW_TRACE
#define LW_TRACE_STMT(...) __VA_ARGS__
#define LW_DBG(a, bit) ((a).dbg & (bit))
#else
#define LW_TRACE_STMT(...)
#define LW_DBG(a, bit) 0
#endif

namespace lw {

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr int kWorkers = 128;     // legacy 4-worker-warp skeleton (downsample)
constexpr int kThreads = 192;
constexpr int kWorkers8 = 256;    // 8 worker warps: warp w -> TMEM lane quadrant w & 3, column half w >> 2
constexpr int kThreads8 = 320;    // + producer warp 8 + issuer warp 9
constexpr int kStageBytes = 16384;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void worker_bar8() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Weight-chunk ring shared by producer and issuer.  Both walk the same static chunk schedule.
struct Ring {
  uint32_t base;       // smem address of stage 0
  uint32_t full0;      // smem address of full[0]; full[s] = full0 + 8*s
  uint32_t empty0;
  int stages;
  int idx;             // running chunk counter
  int stage_bytes = kStageBytes;
  __device__ __forceinline__ uint32_t stage_addr() const { return base + (idx % stages) * stage_bytes; }
  __device__ __forceinline__ uint32_t full() const { return full0 + 8 * (idx % stages); }
  __device__ __forceinline__ uint32_t empty() const { return empty0 + 8 * (idx % stages); }
  __device__ __forceinline__ uint32_t phase() const { return (idx / stages) & 1; }
  // producer side
  __device__ __forceinline__ void load(const void* src, uint32_t bytes) {
    mbar_wait(empty(), phase() ^ 1);
    mbar_expect_tx(full(), bytes);
    bulk_g2s(stage_addr(), src, bytes, full());
    ++idx;
  }
  // issuer side: wait for the chunk, returns its smem address; call release() after the MMAs
  __device__ __forceinline__ uint32_t acquire() {
    mbar_wait(full(), phase());
    tc_fence_after();
    return stage_addr();
  }
  // issuer side is warp-uniform: every lane tracks the ring, one elected lane commits
  __device__ __forceinline__ void release() {
    if (elect_one()) umma_commit(empty());
    __syncwarp();
    ++idx;
  }
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// ---- packed fp32 pairs (FFMA2 / FMUL2 on sm_100: two fp32 lanes per instruction) ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2 r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two bf16 packed in a 32-bit word -> two fp32 (exact): lo = u << 16, hi = u & 0xffff0000
__device__ __forceinline__ f2 bf2_to_f2(uint32_t u) {
  return f2_pack(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ float exp2_approx(float x) {       // x <= 0 in the softmax: single MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU(x) = x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))): a least-squares fit of
// the EXACT erf form (nn.GELU default, model.py:658), |error| <= 4.2e-5 for exact tanh (the common
// "tanh approximation" with 0.044715 is 10x worse); MUFU.TANH adds <= 2^-11 relative on tanh.
// The polynomial argument is clamped to |x| <= 8 where tanh has saturated (keeps it monotone).
__device__ __forceinline__ f2 gelu2(f2 x) {
  float a, b;
  f2_unpack(x, a, b);
  const f2 xc = f2_pack(fminf(fmaxf(a, -8.f), 8.f), fminf(fmaxf(b, -8.f), 8.f));
  const f2 x2 = f2_mul(xc, xc);
  f2 p = f2_fma(x2, f2_pack(-3.72804244e-4f, -3.72804244e-4f), f2_pack(3.71494616e-2f, 3.71494616e-2f));
  p = f2_fma(x2, p, f2_pack(0.797344279f, 0.797344279f));
  float qa, qb;
  f2_unpack(f2_mul(xc, p), qa, qb);
  const f2 t = f2_pack(tanh_approx(qa), tanh_approx(qb));
  const f2 hx = f2_mul(x, f2_pack(0.5f, 0.5f));
  return f2_fma(hx, t, hx);
}
// (A tanh.approx.f16x2 variant with one MUFU per pair was measured 2 % slower end to end: conversion overhead.)
__device__ __forceinline__ float gelu1(float x) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float x2 = xc * xc;
  const float q = xc * fmaf(x2, fmaf(x2, -3.72804244e-4f, 3.71494616e-2f), 0.797344279f);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(q), hx);
}
// ---- packed half precision (f16x2): one instruction = two elements at the full FMA-pipe rate.  (The fp32 pair forms above
// save instruction slots but occupy the FMA pipe for two cycles each; measured on the fused LeFF kernel, whose SM
// sub-partitions are issue / pipe bound, only the genuine half2 forms halve the cost.) ----
typedef uint32_t h2;
__device__ __forceinline__ h2 h2_fma(h2 a, h2 b, h2 c) {
  h2 d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ h2 h2_mul(h2 a, h2 b) {
  h2 d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ h2 h2_min(h2 a, h2 b) {
  h2 d;
  asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ h2 h2_tanh(h2 a) {
  h2 d;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ h2 h2_from_f32(float lo, float hi) {       // round-to-nearest pack, saturating at +-65504 (one F2FP)
  h2 d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ h2 h2_from_f2(f2 v) {
  float a, b;
  f2_unpack(v, a, b);
  return h2_from_f32(a, b);
}
// GELU on a half2 pair, same fitted tanh form as gelu2 (x * 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4)))).  x^2 is clamped at 100:
// beyond |x| = 10 tanh has saturated in half precision and the polynomial stays positive (monotone argument); it also absorbs
// the overflow of x^2 to +inf for |x| > 255.  8 instructions per pair.
struct GeluH2 {
  h2 c0, c1, c2, half, cap;
  __device__ __forceinline__ void init() {
    c0 = h2_from_f32(0.797344279f, 0.797344279f);
    c1 = h2_from_f32(3.71494616e-2f, 3.71494616e-2f);
    c2 = h2_from_f32(-3.72804244e-4f, -3.72804244e-4f);
    half = h2_from_f32(0.5f, 0.5f);
    cap = h2_from_f32(100.f, 100.f);
  }
  __device__ __forceinline__ h2 operator()(h2 x) const {
    const h2 x2 = h2_min(h2_mul(x, x), cap);
    const h2 p = h2_fma(x2, h2_fma(x2, c2, c1), c0);
    const h2 t = h2_tanh(h2_mul(x, p));
    const h2 hx = h2_mul(x, half);
    return h2_fma(hx, t, hx);
  }
};

__device__ __forceinline__ uint32_t f2_to_bf2(f2 v) {
  float a, b;
  f2_unpack(v, a, b);
  return pack_bf16(a, b);
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// asynchronous bulk store shared -> global (TMA engine, no registers, no LSU store traffic)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]);
  v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]);
  v.w = pack_bf16(f[6], f[7]);
  return v;
}

// ----------------------------------------------------------------------------------------------
// Epilogue staging.  Thread-per-row TMEM reads produce 32 B per thread per 16 columns; writing those
// straight to global memory touches 32 different 128 B lines per warp instruction.  Instead every
// epilogue first parks its bf16 results in a shared [128 rows][pitch] tile (pitch = 2*ncols + 16 B, so
// the per-row 16 B writes of a warp spread over all banks), then all worker threads copy the tile out
// with lanes running ALONG rows: fully coalesced 16 B stores, and fully coalesced residual loads.
// ----------------------------------------------------------------------------------------------
// Park a bf16-packed 16-row x (8*NB)-column accumulator fragment (pk[2i] = row t/4 of column block i,
// pk[2i+1] = row t/4+8) in a row-major staging tile with stmatrix: NB/2 instructions, conflict-free.
template <int NB>
__device__ __forceinline__ void stage_frag(uint32_t stage_s, int pitch, int row16, int col0, const uint32_t* pk) {
  const int lane = threadIdx.x & 31;
  const int m = lane >> 3, rr = lane & 7;
  const uint32_t base = stage_s + (row16 + (m & 1) * 8 + rr) * pitch + (col0 + (m >> 1) * 8) * 2;
  if (NB == 1) {
    stsm_x2(stage_s + (row16 + (m & 1) * 8 + rr) * pitch + col0 * 2, pk[0], pk[1]);
  } else {
#pragma unroll
    for (int i2 = 0; i2 < NB / 2; ++i2) stsm_x4(base + i2 * 32, pk[4 * i2], pk[4 * i2 + 1], pk[4 * i2 + 2], pk[4 * i2 + 3]);
  }
}

// Branch-free epilogue body for one 16-lane x (8*NB)-column fp32 fragment: + bias (per column pair, fp32 pairs
// preloaded in bb[NB]) -> optional GELU -> bf16 pack.  Fully unrolled straight-line code so the scheduler can
// interleave the NB*2 independent dependency chains.
template <int NB, bool GELU>
__device__ __forceinline__ void frag_bias_act_pack(const uint32_t* v, const f2* bb, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const f2 x0 = f2_add(f2_pack(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), bb[i]);
    const f2 x1 = f2_add(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), bb[i]);
    pk[2 * i] = f2_to_bf2(GELU ? gelu2(x0) : x0);
    pk[2 * i + 1] = f2_to_bf2(GELU ? gelu2(x1) : x1);
  }
}

// The same for the LeFF hidden map, which is kept in HALF precision (fp16, 11-bit significand) between linear1 and the
// depthwise conv: + bias (fp32) -> f16x2 -> packed-half GELU.  Output words are f16x2.
template <int NB>
__device__ __forceinline__ void frag_bias_gelu_h2(const uint32_t* v, const f2* bb, uint32_t* pk, const GeluH2& gelu) {
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const f2 x0 = f2_add(f2_pack(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), bb[i]);
    const f2 x1 = f2_add(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), bb[i]);
    pk[2 * i] = gelu(h2_from_f2(x0));
    pk[2 * i + 1] = gelu(h2_from_f2(x1));
  }
}

__device__ __forceinline__ uint4 add_bf16x8(const uint4& a, const uint4& b) {
  float fa[8], fb[8];
  unpack8(a, fa);
  unpack8(b, fb);
#pragma unroll
  for (int i = 0; i < 8; ++i) fa[i] += fb[i];
  return pack8(fa);
}
// copy a staged tile to global rows with 256 threads.  row_tok (smem) gives the destination row of each tile
// row (-1: skip).  All shared/global loads of a thread are issued before its first store (latency overlap).
template <int VSHIFT, int NT = 256>   // log2(16-byte vectors per row) = log2(ncols) - 3 ; NT cooperating threads
__device__ __forceinline__ void store_staged_rows_t(uint32_t stage_s, int pitch, uint32_t row_tok_s,
                                                    bf16* __restrict__ out, const bf16* __restrict__ resid,
                                                    size_t row_stride, int col0, int tid) {
  constexpr int TOTAL = 128 << VSHIFT;
  constexpr int ITER_ALL = (TOTAL + NT - 1) / NT;
  constexpr int CAP = (NT == 128) ? 4 : 8;                    // vectors in flight per pass (register budget)
  constexpr int ITER = ITER_ALL > CAP ? CAP : ITER_ALL;
#pragma unroll 1
  for (int p0 = 0; p0 < ITER_ALL; p0 += ITER) {
    uint4 v[ITER], rv[ITER];
    size_t g[ITER];
    int tok[ITER];
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      const int i = tid + (p0 + k) * NT;
      const int row = i >> VSHIFT, vec = i & ((1 << VSHIFT) - 1);
      const bool in = i < TOTAL;
      int t;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(row_tok_s + (in ? row : 0) * 4));
      tok[k] = in ? t : -1;
      v[k] = lds128(stage_s + (in ? row : 0) * pitch + vec * 16);
      g[k] = (size_t)(tok[k] < 0 ? 0 : tok[k]) * row_stride + col0 + vec * 8;
    }
    if (resid != nullptr) {
#pragma unroll
      for (int k = 0; k < ITER; ++k) rv[k] = (tok[k] >= 0) ? __ldg(reinterpret_cast<const uint4*>(resid + g[k])) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < ITER; ++k) v[k] = add_bf16x8(v[k], rv[k]);
    }
#pragma unroll
    for (int k = 0; k < ITER; ++k)
      if (tok[k] >= 0) *reinterpret_cast<uint4*>(out + g[k]) = v[k];
  }
}
// Mixed-precision copy-out (fp32 residual-stream mode): residual bf16 or fp32, output bf16 or fp32, optional bf16 copy of
// the result (the GEMM operand of the kernel that follows).  The branch is rounded to bf16 in the staging tile, the
// residual sum is formed in fp32.  NT cooperating threads; same row_tok convention as store_staged_rows_t.
template <int NT, bool RF32>
__device__ __forceinline__ void store_staged_rows_mixed_t(uint32_t stage_s, int pitch, int ncols_log2, uint32_t row_tok_s, void* __restrict__ out,
                                                          const void* __restrict__ resid, bf16* __restrict__ out_b, bool out_fp32,
                                                          size_t row_stride, int col0, int tid) {
  const int vshift = ncols_log2 - 3;
  const int total = 128 << vshift;
  constexpr int ITER = (NT == 128) ? 2 : 4;     // vectors in flight per thread (all loads of a pass are issued before any use); the
                                                // 128-thread callers are the 88-register epilogue warps of the persistent kernels
#pragma unroll 1
  for (int i0 = tid; i0 < total; i0 += NT * ITER) {
    int tok[ITER];
    size_t g[ITER];
    uint4 sv[ITER], r0[ITER], r1[RF32 ? ITER : 1];
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      const int i = i0 + k * NT;
      const bool in = i < total;
      const int row = (in ? i : 0) >> vshift, vec = (in ? i : 0) & ((1 << vshift) - 1);
      int t;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(row_tok_s + row * 4));
      tok[k] = in ? t : -1;
      sv[k] = lds128(stage_s + row * pitch + vec * 16);
      g[k] = (size_t)(tok[k] < 0 ? 0 : tok[k]) * row_stride + col0 + vec * 8;
    }
    if (resid != nullptr) {
#pragma unroll
      for (int k = 0; k < ITER; ++k) {
        if (RF32) {
          r0[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(resid) + g[k]));
          r1[RF32 ? k : 0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(resid) + g[k] + 4));
        } else {
          r0[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(resid) + g[k]));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      if (tok[k] < 0) continue;
      float f[8];
      unpack8(sv[k], f);
      if (resid != nullptr) {
        if (RF32) {
          const uint4 a = r0[k], b = r1[RF32 ? k : 0];
          f[0] += __uint_as_float(a.x); f[1] += __uint_as_float(a.y); f[2] += __uint_as_float(a.z); f[3] += __uint_as_float(a.w);
          f[4] += __uint_as_float(b.x); f[5] += __uint_as_float(b.y); f[6] += __uint_as_float(b.z); f[7] += __uint_as_float(b.w);
        } else {
          float r[8];
          unpack8(r0[k], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += r[e];
        }
      }
      if (out_fp32) {
        float* op = reinterpret_cast<float*>(out) + g[k];
        *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + g[k]) = pack8(f);
      }
      if (out_b != nullptr) *reinterpret_cast<uint4*>(out_b + g[k]) = pack8(f);
    }
  }
}
template <int NT>
__device__ __forceinline__ void store_staged_rows_mixed(uint32_t stage_s, int pitch, int ncols_log2, uint32_t row_tok_s, void* __restrict__ out,
                                                        const void* __restrict__ resid, bf16* __restrict__ out_b, bool resid_fp32,
                                                        bool out_fp32, size_t row_stride, int col0, int tid) {
  if (resid_fp32) store_staged_rows_mixed_t<NT, true>(stage_s, pitch, ncols_log2, row_tok_s, out, resid, out_b, out_fp32, row_stride, col0, tid);
  else store_staged_rows_mixed_t<NT, false>(stage_s, pitch, ncols_log2, row_tok_s, out, resid, out_b, out_fp32, row_stride, col0, tid);
}
// 128-thread variant (dedicated epilogue warps of the persistent kernels)
__device__ __forceinline__ void store_staged_rows128(uint32_t stage_s, int pitch, int ncols_log2, const int* row_tok,
                                                     bf16* __restrict__ out, const bf16* __restrict__ resid,
                                                     size_t row_stride, int col0, int tid) {
  const uint32_t rts = smem_u32(row_tok);
  switch (ncols_log2) {
    case 7: store_staged_rows_t<4, 128>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    case 6: store_staged_rows_t<3, 128>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    case 5: store_staged_rows_t<2, 128>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    default: store_staged_rows_t<1, 128>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
  }
}
__device__ __forceinline__ void store_staged_rows(uint32_t stage_s, int pitch, int ncols_log2, const int* row_tok,
                                                  bf16* __restrict__ out, const bf16* __restrict__ resid,
                                                  size_t row_stride, int col0, int tid, int nthreads) {
  const uint32_t rts = smem_u32(row_tok);
  (void)nthreads;   // 256 worker threads
  switch (ncols_log2) {
    case 7: store_staged_rows_t<4>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    case 6: store_staged_rows_t<3>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    case 5: store_staged_rows_t<2>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
    default: store_staged_rows_t<1>(stage_s, pitch, rts, out, resid, row_stride, col0, tid); break;
  }
}

// ----------------------------------------------------------------------------------------------
// Stage a 128 x C bf16 A operand into shared memory in the K-major SWIZZLE_128B layout
// ([C/64 k-blocks][128 rows][128 B]), optionally applying LayerNorm (fp32 statistics, biased
// variance, eps inside the sqrt — nn.LayerNorm as at model.py:881,888) and adding a per-row
// fp32 table (the window modulator, model.py:966-969).  row_tok[r] is the source token index of
// tile row r (-1: row is padding -> zeros).  Executed by the 4 worker warps; warp w stages rows
// [32w, 32w+32).  Lanes run along channels, so global reads are coalesced 16 B vectors.
// ----------------------------------------------------------------------------------------------
template <int C, int NW = 4, bool X32 = false, int TABMASK = 63>   // TABMASK: rows of the per-row table minus one (window positions)
__device__ __forceinline__ void stage_rows_ln(uint8_t* sX, const void* __restrict__ xv,
                                              const int* row_tok, const float* __restrict__ ln_w,
                                              const float* __restrict__ ln_b, float eps,
                                              const float* __restrict__ addtab /* [64][C] or null */) {
  constexpr int VPL = (C >= 256) ? C / 256 : 1;   // 16-byte vectors per lane
  constexpr int LPT = (C >= 256) ? 32 : C / 8;    // lanes per token
  constexpr int TPP = 32 / LPT;                   // tokens per pass
  constexpr int RPW = 128 / NW;                   // rows per worker warp
  constexpr int PASSES = RPW / TPP;
  constexpr int BATCH = (PASSES * VPL > 8) ? (8 / VPL) : PASSES;   // passes whose loads are in flight together
  static_assert(PASSES % BATCH == 0, "batching");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const uint32_t sX_s = smem_u32(sX);
  const bf16* __restrict__ x = reinterpret_cast<const bf16*>(xv);
  const float* __restrict__ x32 = reinterpret_cast<const float*>(xv);     // X32: the fp32 residual stream is the input
  for (int p0 = 0; p0 < PASSES; p0 += BATCH) {
    // ---- issue every global load of the batch before touching any (memory-level parallelism) ----
    uint4 raw[BATCH][VPL];
    uint4 raw2[X32 ? BATCH : 1][X32 ? VPL : 1];
    int toks[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int r = warp * RPW + (p0 + b) * TPP + lane / LPT;
      toks[b] = row_tok[r];
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int c0 = j * 256 + sub * 8;
        if (X32) {
          raw[b][j] = (toks[b] >= 0) ? __ldg(reinterpret_cast<const uint4*>(x32 + (size_t)toks[b] * C + c0)) : make_uint4(0, 0, 0, 0);
          raw2[X32 ? b : 0][X32 ? j : 0] = (toks[b] >= 0) ? __ldg(reinterpret_cast<const uint4*>(x32 + (size_t)toks[b] * C + c0 + 4)) : make_uint4(0, 0, 0, 0);
        } else {
          raw[b][j] = (toks[b] >= 0) ? __ldg(reinterpret_cast<const uint4*>(x + (size_t)toks[b] * C + c0)) : make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int r = warp * RPW + (p0 + b) * TPP + lane / LPT;
      float v[VPL][8];
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        if (X32) {
          const uint4 lo = raw[b][j], hi = raw2[X32 ? b : 0][X32 ? j : 0];
          v[j][0] = __uint_as_float(lo.x); v[j][1] = __uint_as_float(lo.y); v[j][2] = __uint_as_float(lo.z); v[j][3] = __uint_as_float(lo.w);
          v[j][4] = __uint_as_float(hi.x); v[j][5] = __uint_as_float(hi.y); v[j][6] = __uint_as_float(hi.z); v[j][7] = __uint_as_float(hi.w);
        } else {
          unpack8(raw[b][j], v[j]);
        }
      }
      if (ln_w != nullptr) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) s += v[j][i];
#pragma unroll
        for (int o = 1; o < LPT; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float d = v[j][i] - mean;
            q += d * d;
          }
#pragma unroll
        for (int o = 1; o < LPT; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / C) + eps);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int c0 = j * 256 + sub * 8;
          float4 g0 = __ldg(reinterpret_cast<const float4*>(ln_w + c0));
          float4 g1 = __ldg(reinterpret_cast<const float4*>(ln_w + c0 + 4));
          float4 b0 = __ldg(reinterpret_cast<const float4*>(ln_b + c0));
          float4 b1 = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + 4));
          const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] = (v[j][i] - mean) * rstd * g[i] + bb[i];
        }
      }
      if (addtab != nullptr && toks[b] >= 0) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int c0 = j * 256 + sub * 8;
          const float* t = addtab + (size_t)(r & TABMASK) * C + c0;
          float4 a0 = __ldg(reinterpret_cast<const float4*>(t));
          float4 a1 = __ldg(reinterpret_cast<const float4*>(t + 4));
          v[j][0] += a0.x; v[j][1] += a0.y; v[j][2] += a0.z; v[j][3] += a0.w;
          v[j][4] += a1.x; v[j][5] += a1.y; v[j][6] += a1.z; v[j][7] += a1.w;
        }
      }
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int c0 = j * 256 + sub * 8;
        sts128(sX_s + (c0 >> 6) * (128 * 128) + swz<128>(r, (c0 & 63) * 2), pack8(v[j]));
      }
    }
  }
}

}  // namespace lw
