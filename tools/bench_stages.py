"""GPU micro-benchmark: time of the DiT forward's launch groups from the debug-stop hook (differences of prefixes).
    python tools/bench_stages.py [Bf]     # stage 1 conditioning, 2 adaLN GEMM + prep, 3 shift-vector GEMMs, 4 patch embed, 5.. blocks
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200 import _lib
from ldmae_b200.pipeline import build_sampling_models
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m, _ = build_sampling_models(dev)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 16, 32, 32, generator=g).to(dev); t = torch.rand(B, generator=g).to(dev); y = torch.randint(0, 1000, (B,), generator=g).to(dev)
m(x, t, y)
L, h = _lib.lib(), m._handle
def timed(stage, n=10):
    _lib.check(L.ldmae_dit_debug_stop(h, stage))
    for _ in range(2): m(x, t, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m(x, t, y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
names = ["conditioning", "adaLN GEMM + prep", "shift-vector GEMMs + final cvec", "patch embed", "b0.qkv", "b0.attention", "b0.proj", "b0.w12", "b0.w3"]
prev = 0.0
for s, nme in enumerate(names, start=1):
    cur = timed(s)
    print(f"stage {s:2d} {nme:32s} {cur - prev:8.3f} ms (prefix {cur:8.3f} ms)")
    prev = cur
_lib.check(L.ldmae_dit_debug_stop(h, -1))
print(f"whole forward Bf={B}: {timed(-1, 5):.3f} ms")
