/* lewin_b200.h — C ABI of the B200-native LeWin-block engine (liblewin_b200.so).
 *
 * The reference (ZhendongWang6/Uformer) has no FFI: its hot path is the nn.Module surface of
 * model.py.  Each entry point below replaces the arithmetic of one reference forward; the Python
 * modules in uformer_b200/modules.py keep the reference's constructor signatures and state-dict
 * keys and call these functions with raw device pointers.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer
 * unless stated; all activations are bf16, row-major "token" layout (B, H*W, C) exactly as the
 * reference passes them between modules; parameters that enter GEMMs are pre-packed bf16 "operand
 * images" (see lw_pack_* in uformer_b200/packing.py), small per-channel parameters are fp32.
 * Calls are asynchronous on `stream`, never allocate, never synchronise, and return 0 or a negative
 * LW_ERR_* code (argument validation happens before any launch).  Re-entrant: no global mutable
 * state, so it is safe under the reference's DataParallel threading (train/train_denoise.py:83).
 */
#ifndef LEWIN_B200_H
#define LEWIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LW_OK 0
#define LW_ERR_BAD_SHAPE (-1)   /* unsupported / inconsistent dimensions */
#define LW_ERR_NULL (-2)        /* required pointer is NULL */
#define LW_ERR_CUDA (-3)        /* launch failed; see lw_last_cuda_error() */
#define LW_ERR_ARCH (-4)        /* device is not sm_100 */
#define LW_ERR_ALIGN (-5)       /* a pointer read/written with 16-byte vectors or bulk copies is not 16-byte aligned
                                   (activations, packed images, fp32 per-channel vectors; torch allocations always are) */

typedef void* lw_stream_t;      /* cudaStream_t */

/* Library/ABI version (bumped on any signature change). */
int lw_abi_version(void);
/* cudaGetLastError() text of the calling thread's most recent failure ("" if none). */
const char* lw_last_cuda_error(void);
/* 0 if the current device is a B200-class (sm_100) part, LW_ERR_ARCH otherwise. */
int lw_check_device(void);

/* sizeof() of the argument struct `id` as this library was compiled (0 wmsa, 1 leff1, 2 leff2, 3 leff, 4 down, 5 up, 6 adamw):
 * lets a binding (uformer_b200/_lib.py mirrors the structs with ctypes) verify its layout; -1 for an unknown id. */
int lw_struct_size(int id);

/* Test hook: cap the grid of the persistent kernels (fused LeFF, LeFF part 2, TMA-gather W-MSA) at n CTAs so that small test
 * inputs walk several tiles per CTA; 0 restores the default (one or two CTAs per SM). */
void lw_set_max_ctas(int n);
/* Programmatic dependent launch between the library's kernels (default OFF: measured 2 % slower on the Uformer-B forward graph,
 * DESIGN §8): with 1 every launch carries the programmatic stream serialization attribute, each kernel runs its on-chip set-up
 * under the tail of its predecessor and executes griddepcontrol.wait before its first global access. */
void lw_set_pdl(int on);

/* Rows per weight-image chunk the A-resident GEMM kernels (lw_leff1_fwd, lw_upsample_fwd) expect for
 * reduction depth K and output width n_total: the host packer must cut w1_img / w_img with this value. */
int lw_nch_ares(int K, int n_total);

/* ---- fused W-MSA: replaces LeWinTransformerBlock.forward's attention half (model.py:951-986)
 * and, with ln_w=NULL / resid=NULL / windowed=1, WindowAttention.forward (model.py:494-522).
 *   out = resid + reverse( proj( softmax( q k^T * hd^-1/2 + relpos_bias + mask ) v ) )
 *   q,k,v = Linear( partition( roll( LN(x), -shift ) ) + modulator )
 * Window size is 8x8 (64 tokens); head_dim in {16, 32, 64}; C in {16,...,512}, C % head_dim == 0 (head_dim 16: C <= 256;
 * head_dim 64: C <= 256). */
typedef struct lw_wmsa_args {
  const void* x;           /* bf16 (B, H*W, C) token map, or (n_windows, 64, C) if windowed */
  void* out;               /* bf16, same layout as x */
  const void* resid;       /* bf16 residual (same layout) added to the output, or NULL */
  const float* ln_w;       /* LayerNorm weight/bias (C), or NULL for no normalisation */
  const float* ln_b;
  const float* modulator;  /* (64, C) fp32 added after LN inside each window, or NULL */
  const void* wqkv_img;    /* packed bf16 [heads][KB][3*hd rows x 128B]; q rows pre-scaled */
  const float* bqkv;       /* (heads, 3*hd) fp32: q (pre-scaled) | k | v bias per head */
  const void* wproj_img;   /* packed bf16 [C/nch][KB][nch rows x 128B] */
  const float* bproj;      /* (C) */
  const float* relpos;     /* (heads, 225) fp32: relative_position_bias_table transposed */
  const float* mask;       /* optional explicit additive mask (n_mask_windows, 64, 64) fp32 */
  int32_t n_mask_windows;
  int32_t n_windows;       /* total windows = B * (H/8) * (W/8) */
  int32_t H, W;            /* token-map size (ignored if windowed) */
  int32_t C, head_dim;
  int32_t shift;           /* cyclic shift (0 or 4); the {0,-100} region mask is computed in-kernel */
  int32_t windowed;        /* 1: x/out are already window-major (WindowAttention standalone) */
  float ln_eps;
  int32_t dbg;             /* profiling aid (env LW_DEBUG & 16): CTA 0 writes clock64 timestamps to `trace` */
  long long* trace;
  /* fp32 residual-stream mode (0 / NULL = everything bf16): */
  int32_t x_fp32;          /* x and resid are fp32 (the residual stream); the GEMM operand is still rounded to bf16 after LayerNorm */
  int32_t out_fp32;        /* out is fp32 */
  void* out_b;             /* optional bf16 copy of out (same layout): the GEMM operand of the LeFF kernel that follows */
  /* TMA-gather path (persistent kernel, csrc/wmsa_tma.cuh).  Taken when wqkv_fold_img != NULL and the call is eligible:
   * token-map input (windowed = 0), LayerNorm present, lw_wmsa_tma_supported(C, head_dim), shift % 4 == 0, a bf16 source for
   * the gather (x itself, or x_b when x_fp32) and, if there is a modulator, its folded image.  Otherwise the fields are ignored
   * and wmsa_kernel runs.
   * LayerNorm is folded into the projection: LN(x) Wqkv^T + b = rstd*(x Wg^T) - rstd*mean*cs + bf. */
  const void* wqkv_fold_img; /* Wg = Wqkv diag(ln_w) rounded to bf16, packed like wqkv_img */
  const float* bqkv_fold;    /* bf = bqkv + Wqkv ln_b, (heads, 3*hd) */
  const float* cs_qkv;       /* row sums of the bf16 Wg, (heads, 3*hd) */
  const void* x_b;           /* bf16 copy of an fp32 x (what the previous kernel wrote as its out_b), or NULL */
  const void* wmod_fold_img; /* with a modulator: packed bf16 [heads][3*hd rows x 64 positions]: (modulator Wqkv^T)^T per head,
                                positions in the kernel's quarter-major window order (packing.pack_qkv_fold); else NULL */
  /* Window size: 0 or 8 = 8x8 windows (everything above); 16 = 16x16 windows (csrc/wmsa16.cuh: one CTA per 256-token window).
   * With 16: x is (n_windows, 256, C) if windowed; n_windows = B*(H/16)*(W/16); relpos is (heads, 961); modulator (256, C);
   * mask (n_mask_windows, 256, 256); shift in [0, 16); H, W multiples of 16; C <= 256 and not (C = 256, head_dim = 64)
   * (lw_wmsa16_supported).  The TMA-gather fields are ignored. */
  int32_t win_size;
} lw_wmsa_args;
int lw_wmsa_fwd(const lw_wmsa_args* a, lw_stream_t stream);
/* 1 if the TMA-gather W-MSA kernel is built for (C, head_dim): C in {16,32,64,128,256}, head_dim in {16,32}. */
int lw_wmsa_tma_supported(int C, int head_dim);
/* 1 if the 16x16-window kernel is built for (C, head_dim): head_dim in {16,32,64}, C = head_dim * 2^k <= 256, not (256, 64). */
int lw_wmsa16_supported(int C, int head_dim);

/* ---- LeFF part 1 (two-kernel path, C = 512): h1 = GELU( LN(x) W1^T + b1 )  (model.py:671 with norm2 of :987 folded in).
 * h1 is written in HALF precision (fp16): it is an internal buffer between the two kernels. */
typedef struct lw_leff1_args {
  const void* x;           /* bf16 (n_tokens, C) */
  void* h1;                /* fp16 (n_tokens, 4C) */
  const float* ln_w;       /* NULL: no LayerNorm (LeFF standalone) */
  const float* ln_b;
  const void* w1_img;      /* packed bf16 [hidden/nch][KB][nch x 128B] */
  const float* b1;         /* (hidden) */
  int32_t n_tokens, C, hidden;
  float ln_eps;
} lw_leff1_args;
int lw_leff1_fwd(const lw_leff1_args* a, lw_stream_t stream);

/* ---- LeFF part 2: out = resid + GELU( dwconv3x3(h1) + bd ) W2^T + b2  (model.py:674-682);
 * the depthwise conv is staged in shared memory (zero padding on h1) and feeds the GEMM directly. */
typedef struct lw_leff2_args {
  const void* h1;          /* fp16 (B, H, W, hidden), as written by lw_leff1_fwd */
  void* out;               /* bf16 (B, H*W, C) (fp32 if out_fp32) */
  const void* resid;       /* bf16 (B, H*W, C) (fp32 if resid_fp32) or NULL */
  const void* taps;        /* fp16 (10, hidden): 9 depthwise taps (tap = ky*3+kx), then the conv bias */
  const void* w2_img;      /* packed FP16 [hidden/64][C/nch][nch x 128B] */
  const float* b2;         /* (C) */
  int32_t B, H, W, C, hidden;
  int32_t resid_fp32, out_fp32;   /* fp32 residual-stream mode (0 = bf16) */
} lw_leff2_args;
int lw_leff2_fwd(const lw_leff2_args* a, lw_stream_t stream);

/* ---- fused LeFF (one launch; the 4C-wide hidden map stays in shared memory): replaces LeFF.forward (model.py:666-685) with
 * norm2 of LeWinTransformerBlock.forward (model.py:987) folded in:
 *   out = resid + GELU( dwconv3x3( GELU( LN(x) W1^T + b1 ) ) + bd ) W2^T + b2
 * LayerNorm is folded into the first GEMM by the host packer (uformer_b200/packing.py pack_leff_fused):
 *   w1_img = bf16( W1 diag(gamma) ),  b1f = b1 + W1 beta,  cs[n] = sum_k float(w1_img[n,k])
 * and the kernel applies  rstd*(x w1_img^T) - rstd*mean*cs + b1f  per token (has_ln = 0: x is used as is; cs is ignored).
 * x is read through a 4-D TMA tensor map built per call ((C, W, H, B), row stride x_stride), 10x18 halo'd boxes per 8x16 tile.
 * The hidden dimension is walked in slices of SL = lw_leff_slice(C) channels (64 for C <= 128, 32 for C = 256).
 * Supported: C in {16, 32, 64, 128, 256}, hidden % 64 == 0, hidden <= 1024 (lw_leff_fused_supported); C = 512 uses the
 * lw_leff1_fwd + lw_leff2_fwd pair.  resid / out may be bf16 or fp32 (fp32 residual-stream mode) and may be column slices of a
 * wider buffer (row strides in elements: skip-concat fusion, model.py:1288-1300).  out must not alias x (halo reads). */
typedef struct lw_leff_args {
  const void* x;           /* bf16 (B*H*W rows, C) with row stride x_stride */
  void* out;               /* (B*H*W rows, C) bf16 or fp32, row stride out_stride */
  const void* resid;       /* same shape, bf16 or fp32, row stride resid_stride; or NULL */
  const void* w1_img;      /* packed bf16 [hidden/64][KB][64 rows x SW bytes], SW = 2*min(C,64), swizzle SW */
  const float* b1f;        /* (hidden) */
  const float* cs;         /* (hidden) */
  const void* taps;        /* [hidden/SL][10][SL] fp16 per hidden slice: 9 depthwise taps (tap = ky*3+kx), then the conv bias */
  const void* w2_img;      /* packed FP16 [hidden/SL][C rows x 2*SL bytes] (K-major, swizzle 2*SL): the hidden map is fp16 on chip */
  const float* b2;         /* (C) */
  int32_t B, H, W, C, hidden;
  int32_t x_stride, resid_stride, out_stride;   /* elements */
  int32_t resid_fp32, out_fp32;                 /* 0: bf16, 1: fp32 */
  int32_t has_ln;
  float ln_eps;
  void* out_b;             /* optional bf16 copy of out, contiguous (B*H*W, C): the TMA source of the W-MSA kernel that follows
                              when out is the fp32 residual stream */
} lw_leff_args;
int lw_leff_fwd(const lw_leff_args* a, lw_stream_t stream);
/* 1 if lw_leff_fwd handles (C, hidden), else 0 (use lw_leff1_fwd + lw_leff2_fwd). */
int lw_leff_fused_supported(int C, int hidden);
/* hidden channels per slice the fused kernel walks for this C (the host packer cuts w1_img / taps / w2_img with it). */
int lw_leff_slice(int C);

/* ---- Downsample: Conv2d(k4,s2,p1) on the token map as an implicit GEMM (model.py:739-746) */
typedef struct lw_down_args {
  const void* x;           /* bf16 (B, H, W, Cin); rows may be x_stride elements apart (a column slice of the skip-concat buffer) */
  void* out;               /* bf16 (B, H/2*W/2, Cout) */
  const void* w_img;       /* packed bf16 [16*Cin/64][Cout/nch][nch x 128B], K index = tap*Cin+ci */
  const float* bias;       /* (Cout) */
  int32_t B, H, W, Cin, Cout;
  int32_t x_stride;        /* row stride of x in elements; 0 = Cin (contiguous) */
} lw_down_args;
int lw_downsample_fwd(const lw_down_args* a, lw_stream_t stream);

/* ---- Upsample: ConvTranspose2d(k2,s2) = GEMM (N = 4*Cout) + pixel-shuffle scatter
 * (model.py:765-771).  out may point into a wider concat buffer: row stride out_stride elements. */
typedef struct lw_up_args {
  const void* x;           /* bf16 (B, H*W, Cin) */
  void* out;               /* bf16 (B, 4*H*W, out_stride) — first Cout channels of each row written */
  const void* w_img;       /* packed bf16 [4*Cout/nch][KB][nch x 128B], N index = (dy*2+dx)*Cout+co */
  const float* bias;       /* (Cout) */
  int32_t B, H, W, Cin, Cout, out_stride;
} lw_up_args;
int lw_upsample_fwd(const lw_up_args* a, lw_stream_t stream);

/* ---- caller-side projections, HBM-bound 3x3 convolutions that also change the layout ----
 * Both take the reference's raw fp32 conv weight (no host packing).  The usual widths run on the tensor core (csrc/proj.cuh):
 * InputProj with Cin = 3 and E in {16, 32} (im2col rows built in shared memory; image and weight split into bf16 hi + lo terms,
 * products fp32-accurate), OutputProj with Cin in {32, 64} and Cout <= 3 (one GEMM per TMA-loaded halo'd token tile, then the nine
 * taps; the weight split into two bf16 halves; the output planes leave through TMA tensor stores when W % 4 == 0 and `out` is
 * 16-byte aligned).  Other widths take direct SIMT kernels.  W must be even; tokens 16-byte aligned.
 * InputProj (model.py:781-812): y = LeakyReLU_0.01(conv3x3(x)), NCHW fp32 image -> bf16 tokens. */
int lw_input_proj_fwd(const float* img, const float* w /* (E,Cin,3,3) */, const float* b, void* tokens,
                      int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t E, lw_stream_t stream);
/* OutputProj + global residual (model.py:815-846, :1305): out = img + conv3x3(tokens), NCHW fp32. */
int lw_output_proj_fwd(const void* tokens, const float* w /* (Cout,Cin,3,3) */, const float* b,
                       const float* img /* residual or NULL */, float* out, int32_t B, int32_t Cin,
                       int32_t H, int32_t W, int32_t Cout, lw_stream_t stream);

/* ---- training-step periphery: HBM-bound streaming kernels ----
 * CharbonnierLoss (losses.py:41-52) forward AND backward in one pass over the data:
 *   *loss = mean( sqrt((x-y)^2 + eps^2) );   grad[i] = (x[i]-y[i]) / sqrt((x[i]-y[i])^2 + eps^2) / n   (grad may be NULL)
 * x, y, grad: fp32, n elements, 16-byte aligned.  `partial` is caller-provided scratch of LW_CHARBONNIER_PARTIALS
 * floats (the library never allocates); the reduction order is fixed, so the loss is reproducible run to run. */
#define LW_CHARBONNIER_PARTIALS 1024
int lw_charbonnier_fwd_bwd(const float* x, const float* y, float* grad, float* loss, float* partial, int64_t n, float eps,
                           lw_stream_t stream);

/* AdamW over flat fp32 arenas in ONE launch — the arithmetic of torch.optim.AdamW(lr, betas, eps, weight_decay)
 * as the reference constructs it (train/train_denoise.py:77), preceded by g *= grad_scale (1/world after the NCCL sum
 * all-reduce of the gradient arena) and optionally followed by g = 0 (the next step's zero_grad, train_denoise.py:172):
 *   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) */
typedef struct lw_adamw_args {
  float* p;                /* parameters  (n) fp32, updated in place */
  float* g;                /* gradients   (n) fp32 */
  float* m;                /* exp_avg     (n) fp32 */
  float* v;                /* exp_avg_sq  (n) fp32 */
  int64_t n;
  int32_t step;            /* t >= 1: the step being taken */
  int32_t zero_grad;       /* 1: write zeros to g after reading it */
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
} lw_adamw_args;
int lw_adamw_step(const lw_adamw_args* a, lw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LEWIN_B200_H */
