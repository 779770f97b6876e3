"""-m gpu: building blocks of the training backward (weight-gradient GEMM with MN-major operands + split contraction,
attention backward) through the C ABI, against fp32 PyTorch computations on the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    prev = torch.is_grad_enabled()          # other test modules switch autograd off process-wide
    torch.set_grad_enabled(True)
    yield
    torch.set_grad_enabled(prev)


def _wgrad(p, q, c, alpha=1.0):
    from ldmae_b200 import _lib
    M, N1 = p.shape
    N2 = q.shape[1]
    _lib.check(_lib.lib().ldmae_gemm_wgrad(_lib.ptr(p), _lib.ptr(q), _lib.ptr(c), N1, N2, M, float(alpha), _lib.stream_ptr()),
               "wgrad")
    torch.cuda.synchronize()
    return c


WGRAD_SHAPES = [
    (64, 256, 256),        # one k-block, one cluster tile
    (1024, 256, 256),      # several k-blocks
    (4096, 2304, 768),     # qkv weight gradient of 4 samples (split contraction, many tiles)
    (2048, 768, 2048),     # w3
    (2048, 4096, 768),     # w12
    (1000, 200, 72),       # ragged everything (tails in N1, N2 and M)
    (8192, 16, 768),       # final layer: 16 output rows
    (8192, 768, 16),       # patch embed: 16 output columns
    (24, 768, 768),        # contraction over a small batch (adaLN / shift-vector gradients)
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_wgrad_accumulates(shape):
    from gpu_util import rel_err
    M, N1, N2 = shape
    g = torch.Generator().manual_seed(M + 3 * N1 + 7 * N2)
    p = torch.randn(M, N1, generator=g).to(torch.bfloat16)
    q = torch.randn(M, N2, generator=g).to(torch.bfloat16)
    c0 = torch.randn(N1, N2, generator=g)
    ref = c0.double() + 0.5 * (p.double().t() @ q.double())
    out = _wgrad(p.cuda(), q.cuda(), c0.clone().cuda(), alpha=0.5).cpu()
    err = rel_err(out, ref)
    assert err < 2e-5, f"wgrad {shape}: rel err {err}"


ATTN_SHAPES = [(2, 1024, 3, 0.125), (3, 64, 2, 0.125), (1, 320, 2, 0.25), (2, 256, 1, 0.125), (2, 16, 2, 0.125),
               # more work items than SMs: every persistent CTA walks several items (next-item look-ahead; 16, 4, 1 and a ragged 3 blocks)
               (30, 1024, 1, 0.125), (80, 256, 2, 0.125), (200, 64, 1, 0.125), (100, 136, 1, 0.125)]


@pytest.mark.parametrize("B,T,H,scale", ATTN_SHAPES)
def test_attention_backward(B, T, H, scale):
    from gpu_util import rel_err
    from ldmae_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 100 + T)
    qkv = (torch.randn(B * T, 3 * H * 64, generator=g)).to(torch.bfloat16)
    dout = torch.randn(B * T, H * 64, generator=g).to(torch.bfloat16)
    # fp32 reference with autograd on the rounded operands
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    att = torch.softmax(s, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B * T, H * 64)
    o.backward(dout.float())
    ref_d = x.grad.reshape(B * T, 3, H * 64)
    ref_lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634       # [B, H, T]

    qkv_d, dout_d = qkv.cuda(), dout.cuda()
    out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    lse2 = torch.zeros(B * H * T + 64, device="cuda")
    delta = torch.zeros(2 * (B * H * T + 64), device="cuda")
    dqkv = torch.full((B * T, 3 * H * 64), float("nan"), device="cuda").to(torch.bfloat16)
    st = _lib.stream_ptr()
    _lib.check(L.ldmae_attention_lse(_lib.ptr(qkv_d), _lib.ptr(out), _lib.ptr(lse2), B, T, H, float(scale), st), "attn fwd")
    _lib.check(L.ldmae_attention_bwd(_lib.ptr(qkv_d), _lib.ptr(out), _lib.ptr(dout_d), _lib.ptr(lse2), _lib.ptr(delta),
                                     _lib.ptr(dqkv), B, T, H, float(scale), st), "attn bwd")
    torch.cuda.synchronize()
    assert rel_err(out.float(), o.detach()) < 1e-2
    torch.testing.assert_close(lse2[: B * H * T].cpu().reshape(B, H, T), ref_lse2.detach(), rtol=0, atol=2e-2)
    got = dqkv.float().cpu().reshape(B * T, 3, H * 64)
    for i, name in enumerate(("dq", "dk", "dv")):
        err = rel_err(got[:, i], ref_d[:, i])
        assert err < 2e-2, f"attention backward {name} B{B} T{T} H{H}: rel err {err}"


@pytest.mark.parametrize("B,T,H,hd", [(2, 256, 2, 72), (1, 1024, 2, 72), (2, 136, 1, 128), (1, 64, 2, 80)])
def test_attention_backward_wide_heads(B, T, H, hd):
    """head_dim in (64, 128] (LightningDiT-XL: 72): 128-column head slots in qkv / dqkv, dense [B*T, H*hd] o and do."""
    from gpu_util import rel_err
    from ldmae_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 100 + T + hd)
    x = torch.zeros(B * T, 3, H, 128)
    x[..., :hd] = torch.randn(B * T, 3, H, hd, generator=g)
    qkv = x.reshape(B * T, 3 * H * 128).to(torch.bfloat16)
    dout = torch.randn(B * T, H * hd, generator=g).to(torch.bfloat16)
    scale = hd ** -0.5
    xr = qkv.float().reshape(B * T, 3, H, 128)[..., :hd].clone().requires_grad_(True)
    q, k, v = xr.reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    att = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B * T, H * hd)
    o.backward(dout.float())
    ref = xr.grad                                        # [B*T, 3, H, hd]
    qkv_d, dout_d = qkv.cuda(), dout.cuda()
    out = torch.empty(B * T, H * hd, device="cuda", dtype=torch.bfloat16)
    lse2 = torch.zeros(B * H * T + 64, device="cuda")
    ws = torch.zeros(2 * (B * H * T + 64), device="cuda")
    dqkv = torch.full((B * T, 3 * H * 128), float("nan"), device="cuda").to(torch.bfloat16)
    st = _lib.stream_ptr()
    _lib.check(L.ldmae_attention_wide_lse(_lib.ptr(qkv_d), _lib.ptr(out), _lib.ptr(lse2), B, T, H, hd, float(scale), st), "fwd")
    _lib.check(L.ldmae_attention_wide_bwd(_lib.ptr(qkv_d), _lib.ptr(out), _lib.ptr(dout_d), _lib.ptr(lse2), _lib.ptr(ws), _lib.ptr(dqkv),
                                          B, T, H, hd, float(scale), st), "bwd")
    torch.cuda.synchronize()
    assert rel_err(out.float().cpu(), o.detach()) < 1e-2
    got = dqkv.float().cpu().reshape(B * T, 3, H, 128)
    assert torch.isfinite(got).all()
    assert float(got[..., hd:].abs().max()) == 0.0 if hd < 128 else True        # padding columns are written as zeros
    for i, name in enumerate(("dq", "dk", "dv")):
        err = rel_err(got[:, i, :, :hd], ref[:, i])
        assert err < 2e-2, f"wide attention backward {name} hd {hd}: rel err {err}"
