// umma.cuh — sm_100a device primitives shared by every LeWin kernel.
//
// Hand-written wrappers around the Blackwell tensor-core path: tcgen05.mma (UMMA) with
// shared-memory matrix descriptors, TMEM allocation / load / store, mbarrier completion,
// and the 128B/64B/32B shared-memory swizzle used by the canonical UMMA operand layouts.
// Nothing here exists in the reference (it is pure PyTorch); this is the B200-native substrate.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lw {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------
// Swizzle<log2(SW/16), 4, 3> on byte offsets of a tile whose base is aligned to 8*SW bytes.
// A K-major operand tile is stored as rows of SW bytes (SW/2 bf16 along K); a row group of 8
// rows is 8*SW bytes.  The 16-byte chunk index inside a row is XOR-ed with bits [7..] of the
// linear offset, which is what the UMMA/TMA hardware applies for SWIZZLE_{32,64,128}B.
// ----------------------------------------------------------------------------------------
template <int SW>
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t byte_in_row) {
  uint32_t lin = row * SW + byte_in_row;
  constexpr uint32_t mask = SW / 16 - 1;
  return lin ^ (((lin >> 7) & mask) << 4);
}
__device__ __forceinline__ uint32_t swz_rt(uint32_t sw, uint32_t row, uint32_t byte_in_row) {
  uint32_t lin = row * sw + byte_in_row;
  uint32_t mask = sw / 16 - 1;
  return lin ^ (((lin >> 7) & mask) << 4);
}

// layout_type field of the sm_100 shared-memory matrix descriptor
__host__ __device__ constexpr uint32_t layout_type_of(int sw_bytes) {
  return sw_bytes == 128 ? 2u : sw_bytes == 64 ? 4u : sw_bytes == 32 ? 6u : 0u;
}

// 64-bit shared-memory matrix descriptor (sm_100 format, version field = 1).
//   [0,14)  start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1      [61,64) swizzle mode
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// K-major operand tile with rows of SW bytes: 8-row groups are 8*SW bytes apart.
template <int SW>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
  return make_smem_desc(saddr, 16, 8 * SW, layout_type_of(SW));
}
// MN-major operand tile (rows are K indices, SW bytes of contiguous MN per row).
// 8-K-row groups are 8*SW bytes apart; `atom_stride` separates SW/2-element atoms along MN.
template <int SW>
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr, uint32_t atom_stride_bytes) {
  return make_smem_desc(saddr, atom_stride_bytes, 8 * SW, layout_type_of(SW));
}

// 32-bit instruction descriptor for tcgen05.mma.kind::f16 with BF16 A/B and FP32 D.
//   [4,6) D fmt (1=f32)  [7,10) A fmt (1=bf16)  [10,13) B fmt  bit15 A MN-major  bit16 B MN-major
//   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major = false,
                                                       bool b_mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// the same with FP16 A/B operands (format code 0) and FP32 D
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool a_mn_major = false, bool b_mn_major = false) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// The same with a suspend-time hint (ns): a waiting thread may sleep in hardware up to that long per call (it is woken when
// the phase completes) instead of burning issue slots of the SM sub-partition it shares with the compute warps.  Measured
// (tools/leff_fused_trace.py): the hinted form costs several hundred cycles even when the phase has already completed, so it
// is only used after a plain probe has failed.
constexpr uint32_t kMbarSuspendHintNs = 20000;
__device__ __forceinline__ uint32_t mbar_try_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(kMbarSuspendHintNs)
      : "memory");
  return ok;
}
// Wait until the phase with the given parity completes.  Bounded so that a protocol bug traps the kernel (sticky error on
// the host) instead of wedging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_sleep(bar, parity)) {
    if (++spins > (1u << 18)) { __trap(); }
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA read smem through it)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// TMEM
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// one full warp; writes the TMEM base address into *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 bit, 8/16/32 consecutive columns: thread t of the warp reads TMEM lane
// (lane field of taddr)+t.  A warp may only touch the lane quadrant 32*(warp_id%4).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------------------
// Fast TMEM access shapes.  Measured on B200 (tests/cuda/tmem_bw_probe.cu): the thread-per-lane 32x32b
// shape reads TMEM at ~30 B/clk/SM, the 16x256b / 16x128b shapes at 220-400 B/clk.  Their register
// layout is the mma m16n8 accumulator fragment (tests/cuda/tmem_shape_probe.cu):
//   16x256b.xN load : regs [4i,4i+1] = (lane L+t/4,   cols 8i+2(t%4), +1), regs [4i+2,4i+3] = (lane L+t/4+8, same cols)
//   16x128b.xN ld/st: reg  [2i]      = (lane L+t/4,   col 4i+t%4),         reg  [2i+1]      = (lane L+t/4+8, same col)
// so a bf16x2-packed 16x256b fragment is exactly a 16x128b fragment.  L is the lane field of taddr
// (multiple of 16 inside the warp's 32-lane quadrant).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {   // 16 lanes x 64 cols
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {   // 16 lanes x 32 cols
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* r) {   // 16 lanes x 16 cols
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t* r) {   // 16 lanes x 32 cols
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x4(uint32_t taddr, const uint32_t* r) {   // 16 lanes x 16 cols
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, const uint32_t* r) {   // 16 lanes x 8 cols
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
// four 8x8 b16 matrices: register j carries the fragment of matrix j (row t/4, elements 2(t%4), +1);
// thread t supplies the shared address of row t%8 of matrix t/8
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void stsm_x2(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1,%2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// ----------------------------------------------------------------------------------------
// UMMA issue (one thread).  D[tmem] (+)= A * B, FP32 accumulate.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand resident in TMEM (lane = row, 2 bf16 of consecutive K per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued UMMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// ----------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library calls pdl_launch_dependents() at its top, so the next
// kernel in the stream / graph may be scheduled onto SMs as they free up, and pdl_wait() after its own on-chip set-up
// (mbarrier init, TMEM allocation, tensor-map prefetch) and BEFORE its first global-memory access: the wait returns when every
// prerequisite grid has completed and its writes are visible.  Launched without the programmatic attribute both are no-ops.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// explicit shared-space accesses on 32-bit shared addresses (keeps LDS/STS instead of generic LD/ST)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace lw
