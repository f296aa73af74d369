// tmem_shape_probe.cu — prints the thread/register -> (TMEM lane, column) mapping of tcgen05.ld/st shapes.
#include <cstdio>
#include "../../uformer_b200/csrc/umma.cuh"
using namespace lw;

__global__ void __launch_bounds__(128, 1) shape_kernel(uint32_t* out) {
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase_s;
  // fill: cell(lane l, col c) = l*1024 + c, via the known-good 32x32b store
  {
    uint32_t v[16];
    for (int c0 = 0; c0 < 64; c0 += 16) {
      for (int j = 0; j < 16; ++j) v[j] = (uint32_t)tid * 1024u + c0 + j;
      tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) {
    uint32_t r[8];
    // ---- 16x256b.x2: 16 lanes x 16 columns, 8 regs ----
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tb) : "memory");
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) out[0 * 256 + lane * 8 + j] = r[j];
    // ---- 16x128b.x2: 16 lanes x 8 columns, 4 regs ----
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tb) : "memory");
    tmem_wait_ld();
    for (int j = 0; j < 4; ++j) out[1 * 256 + lane * 8 + j] = r[j];
    // ---- 16x64b.x2: 16 lanes x 4 columns, 2 regs ----
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(tb) : "memory");
    tmem_wait_ld();
    for (int j = 0; j < 2; ++j) out[2 * 256 + lane * 8 + j] = r[j];
    // ---- 16x256b.x2 at lane offset 16 (second half of the quadrant) ----
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tb + (16u << 16) + 8) : "memory");
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) out[3 * 256 + lane * 8 + j] = r[j];
    // ---- store test: 16x128b.x2 store of tag values into columns 32.., read back with 32x32b ----
    uint32_t w[4];
    for (int j = 0; j < 4; ++j) w[j] = 0x80000000u + lane * 16 + j;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(tb + 32), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    tmem_wait_st();
    uint32_t q[8];
    tmem_ld8(tb + 32, q);
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) out[4 * 256 + lane * 8 + j] = q[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 5 * 256 * 4); cudaMemset(d, 0, 5 * 256 * 4);
  shape_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[5 * 256];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[4] = {"16x256b.x2", "16x128b.x2", "16x64b.x2", "16x256b.x2 @lane16,col8"};
  const int nreg[4] = {8, 4, 2, 8};
  for (int s = 0; s < 4; ++s) {
    printf("== %s: thread: reg=(lane,col)\n", names[s]);
    for (int t = 0; t < 32; ++t) {
      printf("t%02d:", t);
      for (int j = 0; j < nreg[s]; ++j) printf(" (%u,%u)", h[s * 256 + t * 8 + j] >> 10, h[s * 256 + t * 8 + j] & 1023);
      printf("\n");
    }
  }
  printf("== after 16x128b.x2 STORE at col 32: 32x32b read-back, thread t = lane t: cols 32..39 hold (src thread, src reg) or raw\n");
  for (int t = 0; t < 32; ++t) {
    printf("lane%02d:", t);
    for (int j = 0; j < 8; ++j) {
      uint32_t v = h[4 * 256 + t * 8 + j];
      if (v & 0x80000000u) printf(" (t%u,r%u)", (v & 0xffff) / 16, v & 15); else printf(" raw");
    }
    printf("\n");
  }
  return 0;
}
