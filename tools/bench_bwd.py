"""Micro-benchmark of the backward building blocks on one B200: weight-gradient GEMMs and the attention backward,
beside torch (cuBLAS / SDPA autograd) on the same shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
L = _lib.lib()
for name, N1, N2 in (("qkv", 2304, 768), ("proj", 768, 768), ("w12", 4096, 768), ("w3", 768, 2048)):
    P = torch.randn(M, N1, device=dev).to(torch.bfloat16); Q = torch.randn(M, N2, device=dev).to(torch.bfloat16)
    Cm = torch.zeros(N1, N2, device=dev)
    fl = 2.0 * M * N1 * N2
    ms = timeit(lambda: _lib.check(L.ldmae_gemm_wgrad(_lib.ptr(P), _lib.ptr(Q), _lib.ptr(Cm), N1, N2, M, 1.0, _lib.stream_ptr())))
    ms2 = timeit(lambda: torch.matmul(P.t(), Q))
    print(f"wgrad {name:5s} M={M} [{N1}x{N2}]: ours {ms:.3f} ms {fl/ms/1e9:.0f} TF/s | cuBLAS {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s", flush=True)
B, T, H = max(1, M // 1024), 1024, 12
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16)
o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16); do = torch.randn(B * T, H * 64, device=dev).to(torch.bfloat16)
lse = torch.zeros(B * H * T + 64, device=dev); ws = torch.zeros(2 * (B * H * T + 64), device=dev); dqkv = torch.empty_like(qkv)
st = _lib.stream_ptr()
_lib.check(L.ldmae_attention_lse(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), B, T, H, 0.125, st))
fl = 4.0 * B * H * T * T * 64
msf = timeit(lambda: _lib.check(L.ldmae_attention_lse(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), B, T, H, 0.125, st)))
msb = timeit(lambda: _lib.check(L.ldmae_attention_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse), _lib.ptr(ws), _lib.ptr(dqkv), B, T, H, 0.125, st)))
from torch.nn.functional import scaled_dot_product_attention as sdpa
q, k, v = (x.contiguous().requires_grad_(True) for x in qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4))
with torch.enable_grad():
    out = sdpa(q, k, v)
g = torch.randn_like(out)
def tb():
    with torch.enable_grad():
        torch.autograd.grad(out, (q, k, v), g, retain_graph=True)
mst = timeit(tb)
print(f"attention B={B} T={T} H={H}: fwd {msf:.3f} ms {fl/msf/1e9:.0f} TF/s | bwd ours {msb:.3f} ms ({2.5*fl/msb/1e9:.0f} TF/s at 2.5x fwd flops) | torch sdpa bwd {mst:.3f} ms")
