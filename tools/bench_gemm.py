"""GPU micro-benchmark: tcgen05 GEMM variants vs torch.matmul (cuBLAS) on the LightningDiT-B shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, N, K in (("qkv", 2304, 768), ("proj", 768, 768), ("w12", 4096, 768), ("w3", 768, 2048)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    ms = timeit(lambda: torch.matmul(A, W.t(), out=out))
    line = f"{name:5s} M={M} N={N} K={K}: cuBLAS {ms:.3f} ms {fl/ms/1e9:.0f} TF/s"
    for cg, bn in ((2, 256), (1, 256), (1, 128)):
        f = lambda: _lib.check(_lib.lib().ldmae_gemm_bias(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), 1, M, N, K, 0, cg, bn, _lib.stream_ptr()))
        ms = timeit(f)
        line += f" | cg{cg}bn{bn} {ms:.3f} ms {fl/ms/1e9:.0f} TF/s"
    print(line, flush=True)
# attention
from torch.nn.functional import scaled_dot_product_attention as sdpa
B, T, H = max(1, M // 1024), 1024, 12
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16); o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
fl = 4.0 * B * H * T * T * 64
ms = timeit(lambda: _lib.check(_lib.lib().ldmae_attention(_lib.ptr(qkv), _lib.ptr(o), B, T, H, 0.125, _lib.stream_ptr())))
q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
ms2 = timeit(lambda: sdpa(q, k, v))
print(f"attention B={B} T={T} H={H}: ours {ms:.3f} ms {fl/ms/1e9:.0f} TF/s | torch sdpa {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s")
