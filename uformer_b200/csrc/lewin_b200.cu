// lewin_b200.cu — C ABI (include/lewin_b200.h) over the sm_100a kernels.
// Validation first, then a template dispatch on the channel count; no allocation, no sync.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "wmsa.cuh"
#include "wmsa_tma.cuh"
#include "wmsa16.cuh"
#include "leff.cuh"
#include "leff2.cuh"
#include "leff_fused.cuh"
#include "down.cuh"
#include "proj.cuh"
#include "train.cuh"
#include <cmath>
#include <atomic>

using namespace lw;

static thread_local char g_err[256] = "";

static int cuda_fail(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString(e));
  return LW_ERR_CUDA;
}
#define LW_TRY(expr)                               \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return cuda_fail(_e);   \
  } while (0)

// Every pointer the kernels touch with 16-byte vector accesses or cp.async.bulk must be 16-byte aligned (NULL passes).
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
template <typename... P>
static bool all_aligned16(P... ps) {
  bool ok = true;
  const void* v[] = {ps...};
  for (const void* p : v) ok = ok && aligned16(p);
  return ok;
}

extern "C" int lw_abi_version(void) { return 6; }
extern "C" const char* lw_last_cuda_error(void) { return g_err; }
extern "C" int lw_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return LW_ERR_ARCH;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return LW_ERR_ARCH;
  return major == 10 ? LW_OK : LW_ERR_ARCH;
}

extern "C" int lw_struct_size(int id) {
  switch (id) {
    case 0: return (int)sizeof(lw_wmsa_args);
    case 1: return (int)sizeof(lw_leff1_args);
    case 2: return (int)sizeof(lw_leff2_args);
    case 3: return (int)sizeof(lw_leff_args);
    case 4: return (int)sizeof(lw_down_args);
    case 5: return (int)sizeof(lw_up_args);
    case 6: return (int)sizeof(lw_adamw_args);
    default: return -1;
  }
}

// Profiling knobs (env LW_DEBUG, see leff.cuh) exist only in -DLW_TRACE builds (liblewin_b200_trace.so, tools/*_trace.py); the
// production library never touches the environment on the launch path.
static int debug_flags() {
#ifdef LW_TRACE
  const char* e = getenv("LW_DEBUG");
  return e ? atoi(e) : 0;
#else
  return 0;
#endif
}

// Opt a kernel into its dynamic shared-memory size once per device (the attribute is per device and sticky), not per launch.
#define LW_ENSURE_SMEM(kernel, bytes)                                                                         \
  do {                                                                                                        \
    static std::atomic<unsigned long long> _done{0};                                                          \
    int _dev = 0;                                                                                             \
    cudaGetDevice(&_dev);                                                                                     \
    const unsigned long long _bit = 1ull << (_dev & 63);                                                      \
    if (!(_done.load(std::memory_order_relaxed) & _bit)) {                                                    \
      LW_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));               \
      _done.fetch_or(_bit, std::memory_order_relaxed);                                                        \
    }                                                                                                         \
  } while (0)

// N-chunk (rows per weight image chunk) of the A-resident kernels; packing.py mirrors this rule.
extern "C" int lw_nch_ares(int K, int n_total) {
  const int cap = (K == 256) ? 256 : 128;
  return n_total < cap ? n_total : cap;
}

static int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

// Test hook (lw_set_max_ctas): caps the grid of the persistent kernels so that small test inputs still walk several tiles per
// CTA (the multi-tile paths: barrier phases, buffer rotation, cross-tile prefetch).  0 = no cap.
static std::atomic<int> g_max_ctas{0};
extern "C" void lw_set_max_ctas(int n) { g_max_ctas.store(n < 0 ? 0 : n); }
static int cap_grid(int grid) {
  const int c = g_max_ctas.load(std::memory_order_relaxed);
  return (c > 0 && c < grid) ? c : grid;
}

// Programmatic dependent launch (PDL, umma.cuh pdl_*), OFF by default: with lw_set_pdl(1) every launch carries the
// programmatic-stream-serialization attribute, so a kernel may be scheduled while its predecessor drains and runs its on-chip
// set-up there; griddepcontrol.wait orders its first global access behind the predecessor's completion (captured into a CUDA
// graph the launches become programmatic dependency edges).  Measured on the Uformer-B forward graph (100 launches, B200):
// 17.74 ms with PDL against 17.39 ms without — early CTAs of the next persistent grid land unevenly on the SMs that free up
// first — so the default stays the plain serialised launch; the switch is kept for A/B measurements.
static std::atomic<int> g_pdl{0};
extern "C" void lw_set_pdl(int on) { g_pdl.store(on != 0); }
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------
template <int C, int HD>
static int launch_wmsa(const lw_wmsa_args* a, cudaStream_t st) {
  using Cfg = WmsaCfg<C, HD>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "smem budget");
  LW_ENSURE_SMEM((wmsa_kernel<C, HD>), Cfg::SMEM_BYTES);
  const int tiles = (a->n_windows + 1) / 2;
  LW_TRY(launch_k(wmsa_kernel<C, HD>, tiles, kThreads8, Cfg::SMEM_BYTES, st, *a));
  return LW_OK;
}

static bool wmsa_tma_eligible(const lw_wmsa_args* a);
static int launch_wmsa_tma_any(const lw_wmsa_args* a, cudaStream_t st);

// ---- 16 x 16 windows (wmsa16.cuh): one CTA per window ----
extern "C" int lw_wmsa16_supported(int C, int head_dim) {
  if (!(head_dim == 16 || head_dim == 32 || head_dim == 64) || C < head_dim || C > 256 || C % head_dim) return 0;
  if (!(C == 16 || C == 32 || C == 64 || C == 128 || C == 256)) return 0;
  return !(C == 256 && head_dim == 64);
}
template <int C, int HD>
static int launch_wmsa16(const lw_wmsa_args* a, cudaStream_t st) {
  using Cfg = Wmsa16Cfg<C, HD>;
  LW_ENSURE_SMEM((wmsa16_kernel<C, HD>), Cfg::SMEM_BYTES);
  LW_TRY(launch_k(wmsa16_kernel<C, HD>, a->n_windows, kThreads8, Cfg::SMEM_BYTES, st, *a));
  return LW_OK;
}
static int launch_wmsa16_any(const lw_wmsa_args* a, cudaStream_t st) {
  if (!lw_wmsa16_supported(a->C, a->head_dim)) return LW_ERR_BAD_SHAPE;
#define WMSA16_CASE(c, hd) \
  if (a->C == c && a->head_dim == hd) return launch_wmsa16<c, hd>(a, st);
  WMSA16_CASE(16, 16) WMSA16_CASE(32, 16) WMSA16_CASE(64, 16) WMSA16_CASE(128, 16) WMSA16_CASE(256, 16)
  WMSA16_CASE(32, 32) WMSA16_CASE(64, 32) WMSA16_CASE(128, 32) WMSA16_CASE(256, 32)
  WMSA16_CASE(64, 64) WMSA16_CASE(128, 64)
#undef WMSA16_CASE
  return LW_ERR_BAD_SHAPE;
}

extern "C" int lw_wmsa_fwd(const lw_wmsa_args* a, lw_stream_t stream) {
  if (!a || !a->x || !a->out || !a->wqkv_img || !a->bqkv || !a->wproj_img || !a->bproj || !a->relpos) return LW_ERR_NULL;
  if ((a->ln_w == nullptr) != (a->ln_b == nullptr)) return LW_ERR_NULL;
  if (a->n_windows <= 0) return LW_ERR_BAD_SHAPE;
  if (a->win_size != 0 && a->win_size != 8 && a->win_size != 16) return LW_ERR_BAD_SHAPE;
  const int ws = a->win_size == 16 ? 16 : 8;
  if (!a->windowed) {
    if (a->H <= 0 || a->W <= 0 || (a->H % ws) || (a->W % ws)) return LW_ERR_BAD_SHAPE;
    if (a->n_windows % ((a->H / ws) * (a->W / ws))) return LW_ERR_BAD_SHAPE;
    if (a->shift < 0 || a->shift >= ws) return LW_ERR_BAD_SHAPE;
  } else if (a->shift != 0) {
    return LW_ERR_BAD_SHAPE;
  }
  if (a->mask && a->n_mask_windows <= 0) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(a->x, a->out, a->out_b, a->resid, a->ln_w, a->ln_b, a->modulator, a->wqkv_img, a->wproj_img, a->bproj)) return LW_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  lw_wmsa_args aa = *a;
  aa.dbg = debug_flags();
  aa.trace = nullptr;
  if (aa.dbg & 16) {
    const char* e = getenv("LW_TRACE_PTR");
    aa.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  a = &aa;
  if (ws == 16) return launch_wmsa16_any(a, st);
  if (wmsa_tma_eligible(a)) return launch_wmsa_tma_any(a, st);
#define WMSA_CASE(c, hd) \
  if (a->C == c && a->head_dim == hd) return launch_wmsa<c, hd>(a, st);
  WMSA_CASE(32, 32) WMSA_CASE(64, 32) WMSA_CASE(128, 32) WMSA_CASE(256, 32) WMSA_CASE(512, 32)
  WMSA_CASE(16, 16) WMSA_CASE(32, 16) WMSA_CASE(64, 16) WMSA_CASE(128, 16) WMSA_CASE(256, 16)
  WMSA_CASE(64, 64) WMSA_CASE(128, 64) WMSA_CASE(256, 64)
#undef WMSA_CASE
  return LW_ERR_BAD_SHAPE;
}

// ------------------------------------------------------------------------------------------------
template <int K>
static int cap_ares_ctas() { return sm_count() * (K <= 128 ? 2 : 1); }
template <int K, int EPI>
static int launch_ares(const AResArgs& a, cudaStream_t st) {
  using Cfg = AResCfg<K>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "smem budget");
  LW_ENSURE_SMEM((ares_kernel<K, EPI>), Cfg::SMEM_BYTES);
  const int tiles = (a.n_rows + 127) / 128;
  // few tiles (the 16 x 16 / 32 x 32 token maps): split the N range of a tile over 2 or 4 CTAs while the grid still fits the machine
  AResArgs aa = a;
  aa.nsplit = 1;
  const int cap = cap_ares_ctas<K>();
  const int NC = a.n_total / a.nch;
  while (aa.nsplit < 4 && tiles * aa.nsplit * 2 <= cap && NC % (aa.nsplit * 2) == 0) aa.nsplit *= 2;
  LW_TRY(launch_k(ares_kernel<K, EPI>, tiles * aa.nsplit, kThreads8, Cfg::SMEM_BYTES, st, aa));
  return LW_OK;
}
template <int EPI>
static int dispatch_ares(const AResArgs& a, cudaStream_t st) {
  switch (a.K) {
    case 16: return launch_ares<16, EPI>(a, st);
    case 32: return launch_ares<32, EPI>(a, st);
    case 64: return launch_ares<64, EPI>(a, st);
    case 128: return launch_ares<128, EPI>(a, st);
    case 256: return launch_ares<256, EPI>(a, st);
    case 512: return launch_ares<512, EPI>(a, st);
    default: return LW_ERR_BAD_SHAPE;
  }
}

extern "C" int lw_leff1_fwd(const lw_leff1_args* p, lw_stream_t stream) {
  if (!p || !p->x || !p->h1 || !p->w1_img || !p->b1) return LW_ERR_NULL;
  if ((p->ln_w == nullptr) != (p->ln_b == nullptr)) return LW_ERR_NULL;
  if (p->n_tokens <= 0 || p->hidden % 64) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(p->x, p->h1, p->ln_w, p->ln_b, p->w1_img, p->b1)) return LW_ERR_ALIGN;
  AResArgs a{};
  a.x = reinterpret_cast<const bf16*>(p->x); a.n_rows = p->n_tokens; a.K = p->C;
  a.ln_w = p->ln_w; a.ln_b = p->ln_b; a.ln_eps = p->ln_eps;
  a.w_img = reinterpret_cast<const uint8_t*>(p->w1_img); a.n_total = p->hidden; a.nch = lw_nch_ares(p->C, p->hidden);
  a.bias = p->b1; a.out = reinterpret_cast<bf16*>(p->h1);
  a.dbg = debug_flags();
  if (a.dbg & 16) {   // profiling aid: env LW_TRACE_PTR carries a device buffer address (>= 8 KB) for CTA-0 timestamps
    const char* e = getenv("LW_TRACE_PTR");
    a.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  return dispatch_ares<0>(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int lw_upsample_fwd(const lw_up_args* p, lw_stream_t stream) {
  if (!p || !p->x || !p->out || !p->w_img || !p->bias) return LW_ERR_NULL;
  if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->Cout % 16 || p->out_stride < p->Cout || p->out_stride % 8) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(p->x, p->out, p->w_img, p->bias)) return LW_ERR_ALIGN;
  AResArgs a{};
  a.x = reinterpret_cast<const bf16*>(p->x); a.n_rows = p->B * p->H * p->W; a.K = p->Cin;
  a.w_img = reinterpret_cast<const uint8_t*>(p->w_img); a.n_total = 4 * p->Cout; a.nch = lw_nch_ares(p->Cin, 4 * p->Cout);
  a.bias = p->bias; a.out = reinterpret_cast<bf16*>(p->out);
  a.H = p->H; a.W = p->W; a.Cout = p->Cout; a.out_stride = p->out_stride;
  return dispatch_ares<1>(a, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------
extern "C" int lw_leff2_fwd(const lw_leff2_args* p, lw_stream_t stream) {
  if (!p || !p->h1 || !p->out || !p->taps || !p->w2_img || !p->b2) return LW_ERR_NULL;
  if (p->B <= 0 || p->H <= 0 || p->W < 8 || p->hidden % 64 || p->C % 16 || p->C > 512) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(p->h1, p->out, p->resid, p->taps, p->w2_img, p->b2)) return LW_ERR_ALIGN;
  AStreamArgs a{};
  a.src = reinterpret_cast<const bf16*>(p->h1); a.B = p->B; a.H = p->H; a.W = p->W; a.K = p->hidden;
  a.taps = reinterpret_cast<const uint16_t*>(p->taps); a.w_img = reinterpret_cast<const uint8_t*>(p->w2_img);
  a.N = p->C; a.nch = p->C < 128 ? p->C : 128; a.bias = p->b2;
  a.resid = reinterpret_cast<const bf16*>(p->resid); a.out = reinterpret_cast<bf16*>(p->out);
  a.resid_fp32 = p->resid_fp32; a.out_fp32 = p->out_fp32;
  if (p->H % 8) return LW_ERR_BAD_SHAPE;
  a.dbg = debug_flags();
  if (a.dbg & 16) {
    const char* e = getenv("LW_TRACE_PTR");
    a.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  a.TW = 16; a.TH = 8;
  a.tiles_x = (p->W + 15) / 16;
  const int tiles = a.tiles_x * (p->H / 8) * p->B;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((long long)p->B * p->H * p->W * p->hidden >= (1ll << 32)) return LW_ERR_BAD_SHAPE;   // 32-bit element offsets in the halo prefetch
  const int nbuf_d = (2 * a.N <= 512) ? 2 : 1;                 // double-buffer the TMEM accumulator when it fits
  LW_ENSURE_SMEM(leff2_kernel, Leff2Cfg::SMEM_BYTES);
  const int grid = cap_grid(tiles < sm_count() ? tiles : sm_count());    // persistent: one CTA per SM
  LW_TRY(launch_k(leff2_kernel, grid, kL2Threads, Leff2Cfg::SMEM_BYTES, st, a, pow2_cols(nbuf_d * a.N), tiles, nbuf_d));
  return LW_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA tensor maps are encoded by the driver (cuTensorMapEncodeTiled); the entry point is fetched through the runtime so the
// library links against libcudart only.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_tiled_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}
// bf16 token map (B, H, W, C) with row stride `stride` elements -> 4-D map (C, W, H, B), box (cb, bw, bh, 1), swizzle = 2*cb bytes
static int make_token_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int stride, int cb, int bw, int bh) {
  PFN_encodeTiled enc = encode_tiled_fn();
  if (!enc) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled unavailable"); return LW_ERR_CUDA; }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)stride * 2, (cuuint64_t)stride * 2 * W, (cuuint64_t)stride * 2 * W * H};
  const cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cb * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : cb * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return LW_ERR_CUDA; }
  return LW_OK;
}

// fp32 planes (P, H, W) contiguous -> 3-D map (W, H, P), box (bw, bh, 1), no swizzle (the NCHW output of OutputProj)
static int make_plane_map(CUtensorMap* m, const void* base, int P, int H, int W, int bw, int bh) {
  PFN_encodeTiled enc = encode_tiled_fn();
  if (!enc) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled unavailable"); return LW_ERR_CUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * 4 * H};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return LW_ERR_CUDA; }
  return LW_OK;
}

// ---- W-MSA with the TMA window gather (wmsa_tma.cuh) ----
extern "C" int lw_wmsa_tma_supported(int C, int head_dim) {
  return (head_dim == 16 || head_dim == 32) && (C == 16 || C == 32 || C == 64 || C == 128 || C == 256) && C % head_dim == 0 && C >= head_dim;
}
static bool wmsa_tma_eligible(const lw_wmsa_args* a) {
  if (!a->wqkv_fold_img || !a->bqkv_fold || !a->cs_qkv) return false;
  if (a->windowed || (a->modulator && !a->wmod_fold_img) || !a->ln_w || (a->shift % 4) != 0) return false;
  if (a->x_fp32 && !a->x_b) return false;
  if (!all_aligned16(a->wqkv_fold_img, a->x_b, a->wmod_fold_img)) return false;
  return lw_wmsa_tma_supported(a->C, a->head_dim) != 0;
}
template <int C, int HD>
static int launch_wmsa_tma(const CUtensorMap& map, const WmsaTArgs& a, cudaStream_t st) {
  using Cfg = WmsaTCfg<C, HD>;
  LW_ENSURE_SMEM((wmsa_tma_kernel<C, HD>), Cfg::SMEM_BYTES);
  const int cap = sm_count() * (C <= 128 ? 2 : 1);
  const int grid = cap_grid(a.n_tiles < cap ? a.n_tiles : cap);
  LW_TRY(launch_k(wmsa_tma_kernel<C, HD>, grid, kThreads8, Cfg::SMEM_BYTES, st, map, a));
  return LW_OK;
}
static int launch_wmsa_tma_any(const lw_wmsa_args* p, cudaStream_t st) {
  const int nwin_img = (p->H / 8) * (p->W / 8);
  const int B = p->n_windows / nwin_img;
  CUtensorMap map;
  const int cb = p->C < 64 ? p->C : 64;
  const int rc = make_token_map(&map, p->x_fp32 ? p->x_b : p->x, B, p->H, p->W, p->C, p->C, cb, 4, 4);
  if (rc != LW_OK) return rc;
  WmsaTArgs a{};
  a.out = p->out; a.resid = p->resid; a.out_b = reinterpret_cast<bf16*>(p->out_b);
  a.wqkv_img = reinterpret_cast<const uint8_t*>(p->wqkv_fold_img); a.bqkv = p->bqkv_fold; a.cs = p->cs_qkv;
  a.wmod_img = p->modulator ? reinterpret_cast<const uint8_t*>(p->wmod_fold_img) : nullptr;
  a.wproj_img = reinterpret_cast<const uint8_t*>(p->wproj_img); a.bproj = p->bproj; a.relpos = p->relpos;
  a.mask = p->mask; a.n_mask_windows = p->n_mask_windows; a.n_windows = p->n_windows; a.H = p->H; a.W = p->W; a.shift = p->shift;
  a.ln_eps = p->ln_eps; a.resid_fp32 = p->x_fp32; a.out_fp32 = p->out_fp32;
  a.n_tiles = (p->n_windows + 1) / 2;
  a.dbg = p->dbg; a.trace = p->trace;
#define WMSA_TCASE(c, hd) \
  if (p->C == c && p->head_dim == hd) return launch_wmsa_tma<c, hd>(map, a, st);
  WMSA_TCASE(32, 32) WMSA_TCASE(64, 32) WMSA_TCASE(128, 32) WMSA_TCASE(256, 32)
  WMSA_TCASE(16, 16) WMSA_TCASE(32, 16) WMSA_TCASE(64, 16) WMSA_TCASE(128, 16) WMSA_TCASE(256, 16)
#undef WMSA_TCASE
  return LW_ERR_BAD_SHAPE;
}

extern "C" int lw_leff_fused_supported(int C, int hidden) {
  return (C == 16 || C == 32 || C == 64 || C == 128 || C == 256) && hidden % 64 == 0 && hidden >= 64 && hidden <= 1024;
}

extern "C" int lw_leff_slice(int C) { return C <= 128 ? 64 : 32; }

template <int C>
static int launch_leff_fused(const CUtensorMap& map, const LeffFArgs& a, cudaStream_t st) {
  using Cfg = LeffFCfg<C>;
  LW_ENSURE_SMEM(leff_fused_kernel<C>, Cfg::SMEM_BYTES);
  const int grid = cap_grid(a.n_tiles < sm_count() ? a.n_tiles : sm_count());
  LW_TRY(launch_k(leff_fused_kernel<C>, grid, kLFThreads, Cfg::SMEM_BYTES, st, map, a));
  return LW_OK;
}

extern "C" int lw_leff_fwd(const lw_leff_args* p, lw_stream_t stream) {
  if (!p || !p->x || !p->out || !p->w1_img || !p->b1f || !p->cs || !p->taps || !p->w2_img || !p->b2) return LW_ERR_NULL;
  if (!lw_leff_fused_supported(p->C, p->hidden)) return LW_ERR_BAD_SHAPE;
  if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->H % 8 || p->W % 8) return LW_ERR_BAD_SHAPE;
  if (p->x_stride < p->C || p->x_stride % 8 || p->out_stride < p->C || p->out_stride % 8) return LW_ERR_BAD_SHAPE;
  if (p->resid && (p->resid_stride < p->C || p->resid_stride % 8)) return LW_ERR_BAD_SHAPE;
  if (p->out == p->x) return LW_ERR_BAD_SHAPE;                       // halo rows of x are read after neighbouring tiles were written
  if ((long long)p->B * p->H * p->W >= (1ll << 31)) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(p->x, p->out, p->out_b, p->resid, p->w1_img, p->b1f, p->cs, p->taps, p->w2_img, p->b2)) return LW_ERR_ALIGN;
  CUtensorMap map;
  const int cb = p->C < 64 ? p->C : 64;
  const int rc = make_token_map(&map, p->x, p->B, p->H, p->W, p->C, p->x_stride, cb, 18, 10);
  if (rc != LW_OK) return rc;
  LeffFArgs a{};
  a.x = reinterpret_cast<const bf16*>(p->x); a.x_stride = p->x_stride;
  a.B = p->B; a.H = p->H; a.W = p->W; a.hidden = p->hidden;
  a.w1_img = reinterpret_cast<const uint8_t*>(p->w1_img); a.b1f = p->b1f; a.cs = p->cs; a.taps = reinterpret_cast<const uint8_t*>(p->taps);
  a.w2_img = reinterpret_cast<const uint8_t*>(p->w2_img); a.b2 = p->b2;
  a.resid = p->resid; a.out = p->out; a.resid_stride = p->resid_stride; a.out_stride = p->out_stride;
  a.resid_fp32 = p->resid_fp32; a.out_fp32 = p->out_fp32; a.has_ln = p->has_ln; a.ln_eps = p->ln_eps;
  a.out_b = reinterpret_cast<bf16*>(p->out_b); a.out_b_stride = p->C;
  a.tiles_x = (p->W + 15) / 16; a.tiles_y = p->H / 8; a.n_tiles = a.tiles_x * a.tiles_y * p->B;
  a.trace = nullptr;
  if (debug_flags() & 16) {   // profiling aid (-DLW_TRACE builds): env LW_TRACE_PTR = device buffer of >= 7*512 int64
    const char* e = getenv("LW_TRACE_PTR");
    a.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (p->C) {
    case 16: return launch_leff_fused<16>(map, a, st);
    case 32: return launch_leff_fused<32>(map, a, st);
    case 64: return launch_leff_fused<64>(map, a, st);
    case 128: return launch_leff_fused<128>(map, a, st);
    case 256: return launch_leff_fused<256>(map, a, st);
    default: return LW_ERR_BAD_SHAPE;
  }
}

extern "C" int lw_downsample_fwd(const lw_down_args* p, lw_stream_t stream) {
  if (!p || !p->x || !p->out || !p->w_img || !p->bias) return LW_ERR_NULL;
  if (p->B <= 0 || p->H % 2 || p->W % 2 || p->Cin % 8 || (16 * p->Cin) % 64 || p->Cout % 16 || p->Cout > 512) return LW_ERR_BAD_SHAPE;
  if (!((64 % p->Cin == 0) || (p->Cin % 64 == 0))) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(p->x, p->out, p->w_img, p->bias)) return LW_ERR_ALIGN;
  AStreamArgs a{};
  a.src = reinterpret_cast<const bf16*>(p->x); a.B = p->B; a.H = p->H; a.W = p->W; a.K = 16 * p->Cin; a.Cin = p->Cin;
  a.src_stride = p->x_stride > 0 ? p->x_stride : p->Cin;
  if (a.src_stride < p->Cin || a.src_stride % 8) return LW_ERR_BAD_SHAPE;
  a.w_img = reinterpret_cast<const uint8_t*>(p->w_img); a.N = p->Cout; a.nch = p->Cout < 128 ? p->Cout : 128;
  a.bias = p->bias; a.out = reinterpret_cast<bf16*>(p->out);
  const int rows = p->B * (p->H / 2) * (p->W / 2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int t_alloc = pow2_cols(a.N);
  const int stages = t_alloc > 256 ? 4 : 2;            // see DownCfg
  if (stages == 4) {
    LW_ENSURE_SMEM(down_kernel<1>, DownCfg::smem_bytes(4));
    LW_TRY(launch_k(down_kernel<1>, (rows + 127) / 128, kThreads8, DownCfg::smem_bytes(4), st, a, t_alloc, stages));
  } else {
    LW_ENSURE_SMEM(down_kernel<2>, DownCfg::smem_bytes(2));
    LW_TRY(launch_k(down_kernel<2>, (rows + 127) / 128, kThreads8, DownCfg::smem_bytes(2), st, a, t_alloc, stages));
  }
  return LW_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int lw_input_proj_fwd(const float* img, const float* w, const float* b, void* tokens, int32_t B, int32_t Cin,
                                 int32_t H, int32_t W, int32_t E, lw_stream_t stream) {
  if (!img || !w || !b || !tokens) return LW_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (W & 1) || E % 8 || E > 64 || Cin < 1 || Cin > 4) return LW_ERR_BAD_SHAPE;
  if (Cin == 3 && (E == 16 || E == 32) && aligned16(tokens)) {           // tensor-core path (im2col rows in shared memory): proj.cuh
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8;
    const long long n_tiles = (long long)tiles_x * tiles_y * B;
    if (n_tiles >= (1ll << 31)) return LW_ERR_BAD_SHAPE;
    const int cap = 4 * sm_count();
    const int grid = cap_grid(n_tiles < cap ? (int)n_tiles : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (E == 32) {
      LW_ENSURE_SMEM(input_proj_tc_kernel<32>, InProjCfg<32>::SMEM_BYTES);
      LW_TRY(launch_k(input_proj_tc_kernel<32>, grid, 128, InProjCfg<32>::SMEM_BYTES, st, img, w, b, reinterpret_cast<bf16*>(tokens), B, H, W, tiles_x, tiles_y,
                      (int)n_tiles));
    } else {
      LW_ENSURE_SMEM(input_proj_tc_kernel<16>, InProjCfg<16>::SMEM_BYTES);
      LW_TRY(launch_k(input_proj_tc_kernel<16>, grid, 128, InProjCfg<16>::SMEM_BYTES, st, img, w, b, reinterpret_cast<bf16*>(tokens), B, H, W, tiles_x, tiles_y,
                      (int)n_tiles));
    }
    return LW_OK;
  }
  const long long npix = (long long)B * H * (W / 2);     // one thread per horizontal pixel pair
  const int blocks = (int)((npix + 127) / 128);
  LW_TRY(launch_k(input_proj_kernel, blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream), img, w, b, reinterpret_cast<bf16*>(tokens), B, Cin, H, W, E));
  return LW_OK;
}

extern "C" int lw_output_proj_fwd(const void* tokens, const float* w, const float* b, const float* img, float* out, int32_t B,
                                  int32_t Cin, int32_t H, int32_t W, int32_t Cout, lw_stream_t stream) {
  if (!tokens || !w || !b || !out) return LW_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (W & 1) || Cin % 8 || Cin > 128 || Cout < 1 || Cout > 4) return LW_ERR_BAD_SHAPE;
  if ((Cin == 32 || Cin == 64) && Cout <= 3 && aligned16(tokens)) {      // tensor-core path (GEMM first, taps after): proj.cuh
    CUtensorMap map;
    int rc = make_token_map(&map, tokens, B, H, W, Cin, Cin, Cin, 18, 10);
    if (rc != LW_OK) return rc;
    // the output leaves through TMA stores when its rows are 16-byte pitched and aligned (else per-thread stores)
    const int use_tma_store = (W % 4 == 0) && aligned16(out);
    CUtensorMap omap;
    memset(&omap, 0, sizeof(omap));
    if (use_tma_store) {
      rc = make_plane_map(&omap, out, B * Cout, H, W, 16, 8);
      if (rc != LW_OK) return rc;
    }
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8;
    const long long n_tiles = (long long)tiles_x * tiles_y * B;
    if (n_tiles >= (1ll << 31)) return LW_ERR_BAD_SHAPE;
    const int cap = 4 * sm_count();
    const int grid = cap_grid(n_tiles < cap ? (int)n_tiles : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (Cin == 64) {
      LW_ENSURE_SMEM(output_proj_tc_kernel<64>, OutProjCfg<64>::SMEM_BYTES);
      LW_TRY(launch_k(output_proj_tc_kernel<64>, grid, 128, OutProjCfg<64>::SMEM_BYTES, st, map, omap, w, b, img, out, B, H, W, Cout, tiles_x, tiles_y, (int)n_tiles, use_tma_store));
    } else {
      LW_ENSURE_SMEM(output_proj_tc_kernel<32>, OutProjCfg<32>::SMEM_BYTES);
      LW_TRY(launch_k(output_proj_tc_kernel<32>, grid, 128, OutProjCfg<32>::SMEM_BYTES, st, map, omap, w, b, img, out, B, H, W, Cout, tiles_x, tiles_y, (int)n_tiles, use_tma_store));
    }
    LW_TRY(cudaGetLastError());
    return LW_OK;
  }
  const long long npix = (long long)B * H * (W / 2);     // one thread per horizontal pixel pair
  const int blocks = (int)((npix + 127) / 128);
  const size_t smem = (size_t)9 * Cin * 4 * sizeof(float);
  LW_TRY(launch_k(output_proj_kernel, blocks, 128, smem, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const bf16*>(tokens), w, b, img, out, B,
                  Cin, H, W, Cout));
  return LW_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int lw_charbonnier_fwd_bwd(const float* x, const float* y, float* grad, float* loss, float* partial, int64_t n, float eps,
                                      lw_stream_t stream) {
  if (!x || !y || !loss || !partial) return LW_ERR_NULL;
  if (n <= 0) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(x, y, grad)) return LW_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int grid = 4 * sm_count();
  if (grid > LW_CHARBONNIER_PARTIALS) grid = LW_CHARBONNIER_PARTIALS;
  const long long work = (n / 4 + kTrainThreads - 1) / kTrainThreads;
  if (work < grid) grid = work > 0 ? (int)work : 1;
  const float inv_n = (float)(1.0 / (double)n);
  charbonnier_kernel<<<grid, kTrainThreads, 0, st>>>(x, y, grad, partial, n, eps * eps, inv_n);
  LW_TRY(cudaGetLastError());
  charbonnier_finish_kernel<<<1, kTrainThreads, 0, st>>>(partial, grid, inv_n, loss);
  LW_TRY(cudaGetLastError());
  return LW_OK;
}

extern "C" int lw_adamw_step(const lw_adamw_args* a, lw_stream_t stream) {
  if (!a || !a->p || !a->g || !a->m || !a->v) return LW_ERR_NULL;
  if (a->n <= 0 || a->step < 1) return LW_ERR_BAD_SHAPE;
  if (!all_aligned16(a->p, a->g, a->m, a->v)) return LW_ERR_ALIGN;
  if (!(a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f)) return LW_ERR_BAD_SHAPE;
  AdamWConsts c;
  c.lr = a->lr; c.beta1 = a->beta1; c.beta2 = a->beta2; c.eps = a->eps; c.weight_decay = a->weight_decay;
  c.bias_corr1 = (float)(1.0 - std::pow((double)a->beta1, (double)a->step));
  c.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(1.0 - std::pow((double)a->beta2, (double)a->step)));
  c.grad_scale = a->grad_scale;
  int grid = 8 * sm_count();
  const long long work = (a->n / 4 + kTrainThreads - 1) / kTrainThreads;
  if (work < grid) grid = work > 0 ? (int)work : 1;
  adamw_kernel<<<grid, kTrainThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a->p, a->g, a->m, a->v, a->n, c, a->zero_grad);
  LW_TRY(cudaGetLastError());
  return LW_OK;
}
