"""ctypes binding of liblewin_b200.so (C ABI: include/lewin_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain nvcc, no torch headers).
There is no fallback: if the library is missing, or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UFORMER_B200_LIB") or os.path.join(_HERE, "lib", "liblewin_b200.so")    # (override: trace builds, tools/)

LW_ERRORS = {-1: "LW_ERR_BAD_SHAPE", -2: "LW_ERR_NULL", -3: "LW_ERR_CUDA", -4: "LW_ERR_ARCH", -5: "LW_ERR_ALIGN"}


class EngineUnavailable(RuntimeError):
    """The native B200 engine cannot run here (library not built / no sm_100 device)."""


class WmsaArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("resid", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("modulator", C.c_void_p), ("wqkv_img", C.c_void_p), ("bqkv", C.c_void_p), ("wproj_img", C.c_void_p),
                ("bproj", C.c_void_p), ("relpos", C.c_void_p), ("mask", C.c_void_p), ("n_mask_windows", C.c_int32),
                ("n_windows", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("head_dim", C.c_int32),
                ("shift", C.c_int32), ("windowed", C.c_int32), ("ln_eps", C.c_float), ("dbg", C.c_int32), ("trace", C.c_void_p),
                ("x_fp32", C.c_int32), ("out_fp32", C.c_int32), ("out_b", C.c_void_p),
                ("wqkv_fold_img", C.c_void_p), ("bqkv_fold", C.c_void_p), ("cs_qkv", C.c_void_p), ("x_b", C.c_void_p),
                ("wmod_fold_img", C.c_void_p), ("win_size", C.c_int32)]


class Leff1Args(C.Structure):
    _fields_ = [("x", C.c_void_p), ("h1", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p), ("w1_img", C.c_void_p),
                ("b1", C.c_void_p), ("n_tokens", C.c_int32), ("C", C.c_int32), ("hidden", C.c_int32), ("ln_eps", C.c_float)]


class Leff2Args(C.Structure):
    _fields_ = [("h1", C.c_void_p), ("out", C.c_void_p), ("resid", C.c_void_p), ("taps", C.c_void_p),
                ("w2_img", C.c_void_p), ("b2", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("hidden", C.c_int32), ("resid_fp32", C.c_int32), ("out_fp32", C.c_int32)]


class LeffArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("resid", C.c_void_p), ("w1_img", C.c_void_p), ("b1f", C.c_void_p),
                ("cs", C.c_void_p), ("taps", C.c_void_p), ("w2_img", C.c_void_p), ("b2", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("hidden", C.c_int32),
                ("x_stride", C.c_int32), ("resid_stride", C.c_int32), ("out_stride", C.c_int32), ("resid_fp32", C.c_int32),
                ("out_fp32", C.c_int32), ("has_ln", C.c_int32), ("ln_eps", C.c_float), ("out_b", C.c_void_p)]


class DownArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("w_img", C.c_void_p), ("bias", C.c_void_p), ("B", C.c_int32),
                ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32), ("x_stride", C.c_int32)]


class UpArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("w_img", C.c_void_p), ("bias", C.c_void_p), ("B", C.c_int32),
                ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32), ("out_stride", C.c_int32)]


class AdamWArgs(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("step", C.c_int32),
                ("zero_grad", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float)]


CHARBONNIER_PARTIALS = 1024      # LW_CHARBONNIER_PARTIALS

# every symbol include/lewin_b200.h declares
EXPORTS = ["lw_abi_version", "lw_struct_size", "lw_last_cuda_error", "lw_check_device", "lw_nch_ares", "lw_set_max_ctas", "lw_set_pdl", "lw_wmsa_fwd", "lw_wmsa_tma_supported", "lw_wmsa16_supported", "lw_leff1_fwd", "lw_leff2_fwd", "lw_leff_fwd", "lw_leff_fused_supported", "lw_leff_slice",
           "lw_downsample_fwd", "lw_upsample_fwd", "lw_input_proj_fwd", "lw_output_proj_fwd", "lw_charbonnier_fwd_bwd", "lw_adamw_step"]

_lib = None


def csrc_hash() -> str:
    """sha256 (16 hex digits) over the kernel sources (uformer_b200/csrc/*, include/lewin_b200.h): identifies the kernels a
    committed measurement (profiles/r02_kernel_metrics.json) describes.  (The built .so is no such identity: nvcc embeds
    per-build module ids, two builds of the same sources differ.)"""
    import hashlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "uformer_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(csrc)) + [os.path.join(root, "include", "lewin_b200.h")]:
        with open(f if os.path.isabs(f) else os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load():
    """Load the shared library (works without a GPU; compute calls do not)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineUnavailable(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    lib.lw_abi_version.restype = C.c_int
    lib.lw_last_cuda_error.restype = C.c_char_p
    lib.lw_check_device.restype = C.c_int
    lib.lw_struct_size.restype = C.c_int
    lib.lw_struct_size.argtypes = [C.c_int]
    for i, st in enumerate([WmsaArgs, Leff1Args, Leff2Args, LeffArgs, DownArgs, UpArgs, AdamWArgs]):
        if lib.lw_struct_size(i) != C.sizeof(st):
            raise EngineUnavailable(f"{LIB_PATH}: argument struct {st.__name__} is {lib.lw_struct_size(i)} bytes in the library, "
                                    f"{C.sizeof(st)} in the binding (stale build?)")
    lib.lw_nch_ares.restype = C.c_int
    lib.lw_nch_ares.argtypes = [C.c_int, C.c_int]
    lib.lw_leff_fused_supported.restype = C.c_int
    lib.lw_leff_fused_supported.argtypes = [C.c_int, C.c_int]
    lib.lw_set_max_ctas.restype = None
    lib.lw_set_max_ctas.argtypes = [C.c_int]
    lib.lw_set_pdl.restype = None
    lib.lw_set_pdl.argtypes = [C.c_int]
    if os.environ.get("UFORMER_B200_PDL", "0") == "1":          # A/B switch: programmatic dependent launches (off by default)
        lib.lw_set_pdl(1)
    lib.lw_wmsa_tma_supported.restype = C.c_int
    lib.lw_wmsa_tma_supported.argtypes = [C.c_int, C.c_int]
    lib.lw_wmsa16_supported.restype = C.c_int
    lib.lw_wmsa16_supported.argtypes = [C.c_int, C.c_int]
    lib.lw_leff_slice.restype = C.c_int
    lib.lw_leff_slice.argtypes = [C.c_int]
    for name, argt in [("lw_wmsa_fwd", WmsaArgs), ("lw_leff1_fwd", Leff1Args), ("lw_leff2_fwd", Leff2Args), ("lw_leff_fwd", LeffArgs),
                       ("lw_downsample_fwd", DownArgs), ("lw_upsample_fwd", UpArgs)]:
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(argt), C.c_void_p]
    lib.lw_input_proj_fwd.restype = C.c_int
    lib.lw_input_proj_fwd.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_void_p]
    lib.lw_output_proj_fwd.restype = C.c_int
    lib.lw_output_proj_fwd.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 5 + [C.c_void_p]
    lib.lw_charbonnier_fwd_bwd.restype = C.c_int
    lib.lw_charbonnier_fwd_bwd.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_float, C.c_void_p]
    lib.lw_adamw_step.restype = C.c_int
    lib.lw_adamw_step.argtypes = [C.POINTER(AdamWArgs), C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        lib = load()
        detail = lib.lw_last_cuda_error().decode() if rc == -3 else ""
        raise RuntimeError(f"{what} failed: {LW_ERRORS.get(rc, rc)} {detail}")


_device_ok = {}


def require_device(device):
    """Raise unless `device` is a CUDA sm_100 device and the library is loadable."""
    import torch
    if device.type != "cuda":
        raise EngineUnavailable("uformer_b200 runs only on CUDA sm_100 (B200) tensors; there is no CPU fallback "
                                f"(got a tensor on {device})")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        lib = load()
        with torch.cuda.device(idx):
            if lib.lw_check_device() != 0:
                raise EngineUnavailable("current CUDA device is not sm_100 (B200)")
        _device_ok[idx] = True
