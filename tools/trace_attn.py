"""Phase timing of the attention softmax warps (needs a library built with -DLDMAE_ATTN_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
B, T, H = 64, 1024, 12
dev = torch.device("cuda:0")
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16); o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
tr = torch.zeros(2, 64, 8, dtype=torch.int64, device=dev)
L = _lib.lib()
bounded = len(sys.argv) > 1 and sys.argv[1] == "bounded"
m0 = 48.0       # (random un-normed scores here: only the timing is of interest)
run = (lambda: _lib.check(L.ldmae_attention_bounded(_lib.ptr(qkv), _lib.ptr(o), None, B, T, H, 0.125, m0, _lib.stream_ptr()))) if bounded \
    else (lambda: _lib.check(L.ldmae_attention(_lib.ptr(qkv), _lib.ptr(o), B, T, H, 0.125, _lib.stream_ptr())))
for _ in range(2):
    run()
_lib.check(L.ldmae_attention_trace(_lib.ptr(tr)))
run()
torch.cuda.synchronize()
tr = tr.cpu()
t00 = int(tr[0, 0, 0])
names = ["wait_S", "ld_S+release", "max(+rescale)", "wait_turn", "exp", "wait_o_done", "st_P"]
for t in range(2):
    print(f"CTA {t}: block start (rel. cycles) and phase durations")
    for j in range(8):
        st = [int(v) for v in tr[t, j]]
        d = [st[k + 1] - st[k] for k in range(7)]
        nxt = int(tr[t, j + 1, 0]) - st[7] if j < 7 else 0
        print(f"  j={j} start={st[0]-t00:7d} " + " ".join(f"{n}={v}" for n, v in zip(names, d)) + f" tail={nxt}")
