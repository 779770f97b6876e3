import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
dev = torch.device("cuda:0")
B, T, H = 128, 1024, 12
qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(torch.bfloat16); o = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fl = 4.0 * B * H * T * T * 64
ms = timeit(lambda: _lib.check(_lib.lib().ldmae_attention(_lib.ptr(qkv), _lib.ptr(o), B, T, H, 0.125, _lib.stream_ptr())))
q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)[:, :2]
ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(2 * T, H * 64)
err = float((o[:2 * T].float() - ref).norm() / ref.norm())
print(f"{os.environ.get('LDMAE_B200_LIB','default')}: attention fwd {ms:.3f} ms {fl/ms/1e9:.0f} TF/s rel err {err:.2e}")
