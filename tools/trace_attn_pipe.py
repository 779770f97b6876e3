"""Phase timing of the software-pipelined softmax warps of attn_fwd_persist_kernel<true,true> (needs a library built
with -DLDMAE_ATTN_TRACE; CTA 0, warp 0 of each tile, the CTA's first 64 key-block steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
B, T, H = 64, 1024, 12
dev = torch.device("cuda:0")
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16); o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
tr = torch.zeros(2, 64, 8, dtype=torch.int64, device=dev)
L = _lib.lib()
run = lambda: _lib.check(L.ldmae_attention_prescaled(_lib.ptr(qkv), _lib.ptr(o), None, B, T, H, 48.0, _lib.stream_ptr()))
for _ in range(2):
    run()
_lib.check(L.ldmae_attention_trace(_lib.ptr(tr)))
run()
torch.cuda.synchronize()
tr = tr.cpu()
t00 = int(tr[0, 0, 0])
names = ["ld/issue", "expA", "release", "expB+o_done+st+expC", "probe+prefetch", "expD+st+p_full"]
for t in range(2):
    print(f"tile {t}: step start (rel. cycles), phase durations, pre flag")
    for g in range(47):
        st = [int(v) for v in tr[t, g]]
        d = [st[k + 1] - st[k] for k in range(6)]
        nxt = int(tr[t, g + 1, 0]) - st[6]
        print(f"  g={g:2d} start={st[0]-t00:7d} total={int(tr[t, g + 1, 0]) - st[0]:5d} pre={st[7]} " + " ".join(f"{n}={v}" for n, v in zip(names, d)) + f" tail={nxt}")
