// umma_probe.cu — hardware probe for the tcgen05 primitives in csrc/umma.cuh.
//
// Runs one CTA per case: stages A/B (bf16) into shared memory (or A into TMEM) in the canonical
// UMMA layouts produced by lw::swz*, issues tcgen05.mma over K, reads the FP32 accumulator back
// with tcgen05.ld and compares with a CPU reference.  Each operand-layout variant the LeWin
// kernels rely on has a case here, so a descriptor mistake shows up as one failing line, not as
// a wrong image.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include "../../uformer_b200/csrc/umma.cuh"

using namespace lw;

struct Case {
  const char* name;
  int M, N, K;        // K = full K extent staged in the tiles
  int k_begin, k_len; // sub-range of K actually multiplied (multiple of 16)
  int a_tmem;         // 1: A lives in TMEM
  int a_sw;           // swizzle bytes of A tile rows (K-major)
  int b_mn;           // 1: B is MN-major ([K][N] rows), 0: K-major ([N][K] rows)
  int b_sw;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ D,
             int M, int N, int K, int k_begin, int k_len, int a_tmem, int a_sw, int b_mn, int b_sw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;                 // up to 64 KB
  uint8_t* sB = smem + 64 * 1024;     // up to 128 KB

  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t tD = tbase, tA = tbase + 256;

  // ---- stage A ----
  if (!a_tmem) {
    const int KB = a_sw / 2;
    for (int i = tid; i < M * K; i += 128) {
      int r = i / K, k = i % K;
      uint32_t off = (k / KB) * (M * a_sw) + swz_rt(a_sw, r, (k % KB) * 2);
      *reinterpret_cast<bf16*>(sA + off) = A[i];
    }
  } else {
    // lane = row; 2 bf16 per column
    const int r = tid;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const bf16* p = A + (size_t)r * K + 2 * (c0 + j);
        uint32_t lo = *reinterpret_cast<const uint16_t*>(p);
        uint32_t hi = *reinterpret_cast<const uint16_t*>(p + 1);
        v[j] = lo | (hi << 16);
      }
      tmem_st8(tA + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_wait_st();
  }
  // ---- stage B ----
  if (!b_mn) {
    const int KB = b_sw / 2;
    for (int i = tid; i < N * K; i += 128) {
      int n = i / K, k = i % K;
      uint32_t off = (k / KB) * (N * b_sw) + swz_rt(b_sw, n, (k % KB) * 2);
      *reinterpret_cast<bf16*>(sB + off) = B[i];   // B given as [N][K]
    }
  } else {
    const int NB = b_sw / 2;
    for (int i = tid; i < K * N; i += 128) {
      int k = i / N, n = i % N;
      uint32_t off = (n / NB) * (K * b_sw) + swz_rt(b_sw, k, (n % NB) * 2);
      *reinterpret_cast<bf16*>(sB + off) = B[i];   // B given as [K][N]
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(M, N, false, b_mn != 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int k = k_begin; k < k_begin + k_len; k += 16) {
      uint64_t bdesc;
      if (!b_mn) {
        const int KB = b_sw / 2;
        uint32_t s = b0 + (k / KB) * (N * b_sw) + (k % KB) * 2;
        bdesc = make_smem_desc(s, 16, 8 * b_sw, layout_type_of(b_sw));
      } else {
        uint32_t s = b0 + k * b_sw;
        bdesc = make_smem_desc(s, K * b_sw, 8 * b_sw, layout_type_of(b_sw));
      }
      if (!a_tmem) {
        const int KB = a_sw / 2;
        uint32_t s = a0 + (k / KB) * (M * a_sw) + (k % KB) * 2;
        uint64_t adesc = make_smem_desc(s, 16, 8 * a_sw, layout_type_of(a_sw));
        umma_ss(tD, adesc, bdesc, idesc, k > k_begin);
      } else {
        umma_ts(tD, tA + k / 2, bdesc, idesc, k > k_begin);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();

  // read back all 128 lanes x N columns
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    tmem_ld8(tD + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<Case> cases = {
      {"ss_k128_sw128_N128_K128", 128, 128, 128, 0, 128, 0, 128, 0, 128},
      {"ss_k128_sw128_N256_K64", 128, 256, 64, 0, 64, 0, 128, 0, 128},
      {"ss_sw128_N96_K256", 128, 96, 256, 0, 256, 0, 128, 0, 128},
      {"ss_sw128_N16_K64", 128, 16, 64, 0, 64, 0, 128, 0, 128},
      {"ss_sw128_N32_K32of64", 128, 32, 64, 0, 32, 0, 128, 0, 128},
      {"ss_sw128_subatom_k32..64", 128, 128, 64, 32, 32, 0, 128, 0, 128},
      {"ss_sw128_subatom_k16..32", 128, 128, 64, 16, 16, 0, 128, 0, 128},
      {"ss_sw64_N128_K32", 128, 128, 32, 0, 32, 0, 64, 0, 64},
      {"ss_sw64_N128_K64(2kb)", 128, 128, 64, 0, 64, 0, 64, 0, 64},
      {"ss_sw32_N128_K16", 128, 128, 16, 0, 16, 0, 32, 0, 32},
      {"ss_Bmn_sw64_N32_K128", 128, 32, 128, 0, 128, 0, 128, 1, 64},
      {"ss_Bmn_sw128_N64_K128", 128, 64, 128, 0, 128, 0, 128, 1, 128},
      {"ss_Bmn_sw128_N128_K64(2atoms)", 128, 128, 64, 0, 64, 0, 128, 1, 128},
      {"ss_Bmn_sw32_N16_K128", 128, 16, 128, 0, 128, 0, 128, 1, 32},
      {"ts_Atmem_N256_K128", 128, 256, 128, 0, 128, 1, 128, 0, 128},
      {"ts_Atmem_Bmn_sw64_N32_K128", 128, 32, 128, 0, 128, 1, 128, 1, 64},
      {"ts_Atmem_subK_k64..128", 128, 64, 128, 64, 64, 1, 128, 0, 128},
      {"m64_ss_sw128_N64_K64", 64, 64, 64, 0, 64, 0, 128, 0, 128},
  };
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
  int fails = 0;
  for (auto& c : cases) {
    std::vector<float> hA((size_t)c.M * c.K), hB((size_t)c.N * c.K);
    std::vector<bf16> bA(hA.size()), bB(hB.size());
    srand(1234);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); bA[i] = __float2bfloat16(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); bB[i] = __float2bfloat16(hB[i]); }
    // reference: D[m][n] = sum_k A[m][k] * Bmat(n,k);  B storage [N][K] (K-major) or [K][N] (MN-major)
    std::vector<float> ref((size_t)c.M * c.N, 0.f);
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double s = 0;
        for (int k = c.k_begin; k < c.k_begin + c.k_len; ++k) {
          float b = c.b_mn ? hB[(size_t)k * c.N + n] : hB[(size_t)n * c.K + k];
          s += (double)hA[(size_t)m * c.K + k] * b;
        }
        ref[(size_t)m * c.N + n] = (float)s;
      }
    bf16 *dA, *dB; float* dD;
    cudaMalloc(&dA, bA.size() * 2); cudaMalloc(&dB, bB.size() * 2); cudaMalloc(&dD, 128 * c.N * 4);
    cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * c.N * 4);
    probe_kernel<<<1, 128, 192 * 1024>>>(dA, dB, dD, c.M, c.N, c.K, c.k_begin, c.k_len, c.a_tmem, c.a_sw, c.b_mn, c.b_sw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CASE %-34s CUDA ERROR %s\n", c.name, cudaGetErrorString(e)); return 2; }
    std::vector<float> hD((size_t)128 * c.N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    if (c.M == 128) {
      double maxerr = 0, maxref = 0;
      for (size_t i = 0; i < ref.size(); ++i) { maxerr = fmax(maxerr, fabs(hD[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
      bool ok = maxerr <= 1e-3 * fmax(1.0, maxref);
      printf("CASE %-34s %s maxerr=%.3e maxref=%.3e\n", c.name, ok ? "PASS" : "FAIL", maxerr, maxref);
      if (!ok) ++fails;
    } else {
      // M=64: report which TMEM lane holds which row
      printf("CASE %-34s lane->row map:", c.name);
      int found = 0;
      for (int lane = 0; lane < 128; ++lane) {
        int best = -1;
        for (int m = 0; m < c.M; ++m) {
          double err = 0;
          for (int n = 0; n < c.N; ++n) err = fmax(err, fabs(hD[(size_t)lane * c.N + n] - ref[(size_t)m * c.N + n]));
          if (err < 1e-3) { best = m; break; }
        }
        if (best >= 0) { ++found; if (lane % 16 == 0) printf(" L%d=r%d", lane, best); }
      }
      printf("  (matched %d lanes)\n", found);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf("PROBE %s (%d failing)\n", fails ? "FAILED" : "OK", fails);
  return 0;
}
