"""Phase timing of the attention backward (rows 15 / 31, first two columns: row warp 4 saw acc_done / finished the item's epilogue) -- needs a library built with -DLDMAE_ATTN_TRACE."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
B, T, H = 32, 1024, 12
dev = torch.device("cuda:0")
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16)
o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16); do = torch.randn(B * T, H * 64, device=dev).to(torch.bfloat16)
lse = torch.zeros(B * H * T + 64, device=dev); ws = torch.zeros(2 * (B * H * T + 64), device=dev); dqkv = torch.empty_like(qkv)
L = _lib.lib(); st = _lib.stream_ptr()
_lib.check(L.ldmae_attention_lse(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), B, T, H, 0.125, st))
tr = torch.zeros(32, 8, dtype=torch.int64, device=dev)
run = lambda: _lib.check(L.ldmae_attention_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse), _lib.ptr(ws), _lib.ptr(dqkv), B, T, H, 0.125, st))
run(); run()
_lib.check(L.ldmae_attention_trace(_lib.ptr(tr)))
run()
torch.cuda.synchronize()
tr = tr.cpu()          # holds the stamps of the LAST kernel launched (dQ), CTA 0
t0 = int(tr[0, 4])
print("dQ kernel, CTA 0: per column block, cycles relative to the first sd_full")
print(" i | mma: c_full scores_issued pd_full acc_issued | rows: sd_full loaded math_done stored")
for i in range(32):
    print(f"{i:2d} | " + " ".join(f"{int(tr[i, k]) - t0:7d}" for k in range(4)) + " | " + " ".join(f"{int(tr[i, k]) - t0:7d}" for k in range(4, 8)))
