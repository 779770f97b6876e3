"""GPU diagnostic: find the first launch group of the DiT forward whose result is not bit-reproducible."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200.pipeline import build_sampling_models
from ldmae_b200 import _lib
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
m, vae = build_sampling_models(dev)
L = _lib.lib()
g = torch.Generator().manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = torch.randn(B, 16, 32, 32, generator=g).to(dev); t = torch.rand(B, generator=g).to(dev); y = torch.randint(0, 1000, (B,), generator=g).to(dev)
m(x, t, y)
h = m._handle
M = B * 1024
bufs = {"xres": (M * 768, torch.float32), "abuf": (M * 768, torch.bfloat16), "qkv": (M * 2304, torch.bfloat16), "obuf": (M * 768, torch.bfloat16),
        "hbuf": (M * 2048, torch.bfloat16), "ssq": (M * 6, torch.float32), "mods": (B * 74 * 768, torch.float32),
        "cvec_qkv": (12 * B * 2304, torch.float32), "cvec_12": (12 * B * 4096, torch.float32), "gmul": (25 * B * 768, torch.float32)}
def snap(stage):
    _lib.check(L.ldmae_dit_debug_stop(h, stage))
    m(x, t, y)
    out = {}
    for k, (n, dt) in bufs.items():
        b = torch.empty(n, device=dev, dtype=dt)
        _lib.check(L.ldmae_dit_debug_read(h, k.encode(), _lib.ptr(b), b.numel() * b.element_size(), _lib.stream_ptr()))
        out[k] = b
    torch.cuda.synchronize()
    return out
names = ["cond", "adaln", "shiftvec", "patch"] + [f"b{i}.{n}" for i in range(12) for n in ("qkv", "attn", "proj", "w12", "w3")]
for stage in range(1, 4 + 60 + 1):
    a = snap(stage)
    bad = {}
    for rep in range(3):
        b = snap(stage)
        for k in a:
            if not torch.equal(a[k].view(torch.uint8), b[k].view(torch.uint8)):
                d = (a[k].float() - b[k].float()).abs()
                bad[k] = (int((d > 0).sum()), float(d.max()))
    print(f"stage {stage:2d} {names[stage-1]:10s} nondeterministic buffers: {bad}", flush=True)
    if bad and stage > 12:
        break
_lib.check(L.ldmae_dit_debug_stop(h, -1))
