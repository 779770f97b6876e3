"""Phase timing of the residual GEMM epilogue (needs a library built with -DLDMAE_GEMM_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
M = 131072; which = sys.argv[1] if len(sys.argv) > 1 else "proj"
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
for _ in range(2): run()
tr = torch.zeros(256, 8, dtype=torch.int64, device=dev)
_lib.check(L.ldmae_gemm_trace(_lib.ptr(tr)))
run(); torch.cuda.synchronize()
tr = tr.cpu()
names = ["wait_x", "ld_acc", "math_x", "wait_store+pump", "math_a", "issue_store"]
t0 = int(tr[0, 0])
for sidx in list(range(0, 40)):
    st = [int(v) for v in tr[sidx]]
    if st[0] == 0: break
    d = [st[k + 1] - st[k] for k in range(6)]
    nxt = int(tr[sidx + 1, 0]) - st[6]
    print(f"pair {sidx:3d} start={st[0]-t0:8d} " + " ".join(f"{n}={v}" for n, v in zip(names, d)) + f" to_next={nxt}")
