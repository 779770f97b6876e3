"""Helpers shared by the -m gpu parity tests (call the product through its C ABI / Python mirror)."""
import ctypes as C

import numpy as np
import torch

from ldmae_b200 import _lib


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gemm_bias(a_bf16, w_bf16, bias, out_bf16: bool, act=0, cta_group=2, block_n=256):
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(_lib.lib().ldmae_gemm_bias(_lib.ptr(a_bf16), _lib.ptr(w_bf16), _lib.ptr(bias), _lib.ptr(out), int(out_bf16),
                                          M, N, K, act, cta_group, block_n, _lib.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    return out


def attention(qkv_bf16, B, T, H, scale):
    out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().ldmae_attention(_lib.ptr(qkv_bf16), _lib.ptr(out), B, T, H, float(scale), _lib.stream_ptr()), "attn")
    torch.cuda.synchronize()
    return out


def load_npz(golden_dir, name):
    import os
    return np.load(os.path.join(golden_dir, name))


def gemm_residual(a_bf16, w_bf16, bias, x, gate=None, gnext=None, rows_per_sample=None):
    """x += gate*(a.w^T + bias) in place; returns (x, anext or None, ssq)."""
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    rows_per_sample = rows_per_sample or M
    anext = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if gnext is not None else None
    slots = (N + 127) // 128
    ssq = torch.full((M, slots), float("nan"), device="cuda")
    _lib.check(_lib.lib().ldmae_gemm_residual(_lib.ptr(a_bf16), _lib.ptr(w_bf16), _lib.ptr(bias), _lib.ptr(gate), _lib.ptr(gnext),
                                              _lib.ptr(x), _lib.ptr(anext), _lib.ptr(ssq), M, N, K, rows_per_sample,
                                              _lib.stream_ptr()), "gemm_residual")
    torch.cuda.synchronize()
    return x, anext, ssq
