"""Runs the residual-epilogue GEMM at the proj / w3 shapes (timing + target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
which = sys.argv[2] if len(sys.argv) > 2 else "proj"
N, K = (768, 768) if which == "proj" else (768, 2048)
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device=dev); nb = M // 1024
gate = torch.randn(nb, N, device=dev); gnext = torch.randn(nb, N, device=dev)
x = torch.randn(M, N, device=dev); anext = torch.empty(M, N, device=dev, dtype=torch.bfloat16); ssq = torch.empty(M, 6, device=dev)
L = _lib.lib()
def run():
    _lib.check(L.ldmae_gemm_residual(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gate), _lib.ptr(gnext), _lib.ptr(x), _lib.ptr(anext),
                                     _lib.ptr(ssq), M, N, K, 1024, _lib.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = M * (K * 2 + N * 10)
print(f"{which} M={M}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TF/s  {byt/ms/1e6:.0f} GB/s algorithmic")
