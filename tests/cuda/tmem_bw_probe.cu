// tmem_bw_probe.cu — measures TMEM -> register read bandwidth per SM for several tcgen05.ld widths and warp counts.
#include <cstdio>
#include "../../uformer_b200/csrc/umma.cuh"
using namespace lw;

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
        "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
        "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
        "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}

// 16 lanes x 256 bit per step, x8 steps: 16 lanes x 64 columns = 32 registers per thread
__device__ __forceinline__ void ld16x256_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// 16 lanes x 128 bit per step, x16: 16 lanes x 64 columns = 32 registers
__device__ __forceinline__ void ld16x128_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x128b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// shape test: MODE 0: 16x256b.x8, MODE 1: 16x128b.x16; each instruction covers 16 lanes x 64 columns
template <int MODE>
__global__ void __launch_bounds__(512, 1) bw_kernel16(long long* out, int nwarps, int iters) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase_s;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      const int share = 512 / ((nwarps + 3) / 4);
      const int cbeg = (warp >> 2) * share;
      for (int c = cbeg; c < cbeg + share; c += 64) {
        uint32_t r[2][32];
        for (int hl = 0; hl < 2; ++hl) {          // the two 16-lane halves of this warp's quadrant
          const uint32_t la = (uint32_t)((warp & 3) * 32 + hl * 16) << 16;
          if (MODE == 0) ld16x256_x8(tb + la + c, r[hl]); else ld16x128_x16(tb + la + c, r[hl]);
        }
        tmem_wait_ld();
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[hl][j];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int W>   // columns per load
__global__ void __launch_bounds__(512, 1) bw_kernel(long long* out, int nwarps, int iters, int inflight) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase_s;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int it = 0; it < iters; ++it) {
      // each pass reads this warp's share of the 512 columns
      const int share = 512 / ((nwarps + 3) / 4);
      const int cbeg = (warp >> 2) * share;
      for (int c = cbeg; c < cbeg + share; c += W * inflight) {
        uint32_t r[2][W];
        for (int u = 0; u < inflight; ++u) {
          if (W == 16) tmem_ld16(tb + lane_base + c + u * W, r[u]);
          else if (W == 32) tmem_ld32(tb + lane_base + c + u * W, r[u]);
          else ld64(tb + lane_base + c + u * W, r[u]);
        }
        tmem_wait_ld();
        for (int u = 0; u < inflight; ++u)
#pragma unroll
          for (int j = 0; j < W; ++j) acc ^= r[u][j];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int iters = 50;
  for (int W : {16, 32, 64})
    for (int nw : {4, 8, 16})
      for (int inf : {1, 2}) {
        if (W == 64 && inf == 2) continue;
        if (W == 16) bw_kernel<16><<<1, 512>>>(d, nw, iters, inf);
        if (W == 32) bw_kernel<32><<<1, 512>>>(d, nw, iters, inf);
        if (W == 64) bw_kernel<64><<<1, 512>>>(d, nw, iters, inf);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        double bytes = 128.0 * 512 * 4 * iters;     // the whole TMEM read once per pass
        printf("ld x%-2d warps %2d inflight %d : %8lld cycles  %.1f B/clk\n", W, nw, inf, cyc, bytes / cyc);
      }
  for (int mode : {0, 1})
    for (int nw : {4, 8, 16}) {
      if (mode == 0) bw_kernel16<0><<<1, 512>>>(d, nw, iters); else bw_kernel16<1><<<1, 512>>>(d, nw, iters);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      double bytes = 128.0 * 512 * 4 * iters;
      printf("ld %s warps %2d : %8lld cycles  %.1f B/clk\n", mode == 0 ? "16x256b.x8 " : "16x128b.x16", nw, cyc, bytes / cyc);
    }
  return 0;
}
