"""ctypes binding of libldmae_b200.so (include/ldmae_b200.h).  No CPU fallback: if the CUDA
library is missing or fails, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDMAE_B200_LIB") or os.path.join(HERE, "libldmae_b200.so")

_lib = None


class LdmaeError(RuntimeError):
    pass


class DitConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "depth", "hidden_size", "num_heads", "patch_size", "input_size", "in_channels", "num_embeddings",
        "mlp_hidden", "learn_sigma", "use_qknorm", "use_swiglu", "use_rope", "use_rmsnorm", "wo_shift", "max_batch")]


class VmaeConfig(C.Structure):
    _fields_ = [("img_size", C.c_int32), ("patch_size", C.c_int32), ("latent_dim", C.c_int32),
                ("embed_dim", C.c_int32), ("decoder_embed_dim", C.c_int32), ("decoder_depth", C.c_int32),
                ("decoder_num_heads", C.c_int32), ("mlp_hidden", C.c_int32), ("ln_eps", C.c_float),
                ("max_batch", C.c_int32), ("depth", C.c_int32), ("num_heads", C.c_int32), ("to_latent_dim", C.c_int32)]


# name -> (restype, argtypes): every symbol include/ldmae_b200.h declares
vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "ldmae_last_error": (C.c_char_p, []),
    "ldmae_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ldmae_version": (C.c_int, []),
    "ldmae_dit_create": (C.c_int, [C.POINTER(DitConfig), C.POINTER(vp)]),
    "ldmae_dit_destroy": (None, [vp]),
    "ldmae_dit_load_tensor": (C.c_int, [vp, C.c_char_p, vp, i64, vp]),
    "ldmae_dit_finalize": (C.c_int, [vp, vp]),
    "ldmae_dit_forward": (C.c_int, [vp, vp, vp, f32, vp, vp, i32, i32, vp]),
    "ldmae_dit_train_forward": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
    "ldmae_dit_backward": (C.c_int, [vp, vp, i32, vp]),
    "ldmae_dit_generation": (C.c_longlong, [vp]),
    "ldmae_dit_load_tensors": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i64), i32, vp]),
    "ldmae_dit_grad_read_many": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i64), i32, i32, vp]),
    "ldmae_dit_check_labels": (C.c_int, [vp, vp]),
    "ldmae_dit_grad_read": (C.c_int, [vp, C.c_char_p, vp, i64, vp]),
    "ldmae_dit_grad_accumulate": (C.c_int, [vp, C.c_char_p, vp, i64, vp]),
    "ldmae_adamw_ema_step": (C.c_int, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, f32, vp]),
    "ldmae_flow_prepare": (C.c_int, [vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "ldmae_flow_loss": (C.c_int, [vp, vp, vp, vp, f32, i32, i32, vp]),
    "ldmae_dit_debug_stop": (C.c_int, [vp, i32]),
    "ldmae_dit_debug_poison": (C.c_int, [vp, i32, vp]),
    "ldmae_dit_debug_read": (C.c_int, [vp, C.c_char_p, vp, i64, vp]),
    "ldmae_dit_forward_with_cfg": (C.c_int, [vp, vp, vp, f32, vp, vp, i32, f32, i32, vp]),
    "ldmae_sample_ode": (C.c_int, [vp, vp, vp, i32, i32, f32, f32, C.POINTER(C.c_float), i32, i32, vp, i32, vp]),
    "ldmae_vmae_create": (C.c_int, [C.POINTER(VmaeConfig), C.POINTER(vp)]),
    "ldmae_vmae_destroy": (None, [vp]),
    "ldmae_vmae_load_tensor": (C.c_int, [vp, C.c_char_p, vp, i64, vp]),
    "ldmae_vmae_finalize": (C.c_int, [vp, vp]),
    "ldmae_vmae_decode": (C.c_int, [vp, vp, vp, vp, f32, vp, vp, i32, vp]),
    "ldmae_vmae_encode": (C.c_int, [vp, vp, vp, i32, vp]),
    "ldmae_gemm_bias": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "ldmae_gemm_residual": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "ldmae_attention": (C.c_int, [vp, vp, i32, i32, i32, f32, vp]),
    "ldmae_attention_lse": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "ldmae_attention_wide": (C.c_int, [vp, vp, i32, i32, i32, i32, f32, vp]),
    "ldmae_attention_wide_lse": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "ldmae_attention_wide_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "ldmae_attention_bounded": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, f32, vp]),
    "ldmae_attention_prescaled": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "ldmae_attention_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    "ldmae_gemm_wgrad": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "ldmae_attention_trace": (C.c_int, [vp]),
    "ldmae_gemm_trace": (C.c_int, [vp]),
    "ldmae_f32_to_bf16": (C.c_int, [vp, vp, i64, vp]),
    "ldmae_launch_count": (C.c_longlong, []),
    "ldmae_profile_begin": (C.c_int, []),
    "ldmae_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong), i32]),
}


def lib():
    """The loaded library; raises LdmaeError if it was not built (python -m ldmae_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LdmaeError(f"{LIB_PATH} is missing: build it with `python -m ldmae_b200.build` "
                             "(ldmae_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ldmae_last_error()
        raise LdmaeError(f"{what or 'ldmae call'} failed ({rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "ldmae_b200 expects contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


ODE_COND_ONLY_WHEN_UNGUIDED = 1

PROF_CLASSES = ("qkv_gemm", "attention", "proj_gemm", "w12_swiglu_gemm", "w3_gemm", "adaln_shift_gemms", "final_gemm",
                "cond_embed_update", "vmae_decode", "bwd_dgrad_gemms", "bwd_wgrad_gemms", "bwd_attention", "bwd_hbm_kernels")


def profile_begin():
    check(lib().ldmae_profile_begin(), "profile_begin")


def profile_end():
    n = len(PROF_CLASSES)
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(lib().ldmae_profile_end(ms, cnt, n), "profile_end")
    return {PROF_CLASSES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def launch_count() -> int:
    return int(lib().ldmae_launch_count())
