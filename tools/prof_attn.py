"""Runs the attention kernel a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
B, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 1024, 12
dev = torch.device("cuda:0")
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16); o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    _lib.check(_lib.lib().ldmae_attention(_lib.ptr(qkv), _lib.ptr(o), B, T, H, 0.125, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok")
