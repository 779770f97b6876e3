"""GPU diagnostic: first forward after handle creation vs steady state, stage by stage."""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randn(B, 16, 32, 32, generator=g).to(dev); t = torch.rand(B, generator=g).to(dev); y = torch.randint(0, 1000, (B,), generator=g).to(dev)
o = [m(x, t, y).clone() for _ in range(4)]
print("fresh process: call0==call1", torch.equal(o[0], o[1]), "call1==call2", torch.equal(o[1], o[2]), "call2==call3", torch.equal(o[2], o[3]))
m._release()
o2 = [m(x, t, y).clone() for _ in range(3)]
print("after handle re-create: call0==steady", torch.equal(o2[0], o[3]), "call1==steady", torch.equal(o2[1], o[3]))
M = B * 1024
bufs = {"xres": (M * 768, torch.float32), "abuf": (M * 768, torch.bfloat16), "qkv": (M * 2304, torch.bfloat16), "obuf": (M * 768, torch.bfloat16),
        "hbuf": (M * 2048, torch.bfloat16), "ssq": (M * 6, torch.float32), "mods": (B * 74 * 768, torch.float32), "cvec_c": (B * 768, torch.float32),
        "cvec_qkv": (12 * B * 2304, torch.float32), "cvec_12": (12 * B * 4096, torch.float32), "gmul": (25 * B * 768, torch.float32),
        "sc": (B * 768, torch.bfloat16), "th1": (B * 768, torch.float32), "shift_bf16": (25 * B * 768, torch.bfloat16), "cvec_f": (B * 16, torch.float32)}
def snap(stage, fresh):
    if fresh:
        m._release()
        m._ensure_handle(dev, B)
    h = m._handle
    _lib.check(L.ldmae_dit_debug_stop(h, stage))
    m(x, t, y)
    out = {}
    for k, (n, dt) in bufs.items():
        b = torch.empty(n, device=dev, dtype=dt)
        _lib.check(L.ldmae_dit_debug_read(h, k.encode(), _lib.ptr(b), b.numel() * b.element_size(), _lib.stream_ptr()))
        out[k] = b
    torch.cuda.synchronize()
    _lib.check(L.ldmae_dit_debug_stop(h, -1))
    return out
names = ["cond", "adaln", "shiftvec", "patch"] + [f"b{i}.{n}" for i in range(12) for n in ("qkv", "attn", "proj", "w12", "w3")]
written = {"cond": ["th1", "cvec_c", "sc"], "adaln": ["mods", "gmul", "shift_bf16"], "shiftvec": ["cvec_qkv", "cvec_12", "cvec_f"], "patch": ["xres", "abuf", "ssq"],
           "qkv": ["qkv"], "attn": ["obuf"], "proj": ["xres", "abuf", "ssq"], "w12": ["hbuf"], "w3": ["xres", "abuf", "ssq"]}
for stage in range(1, 16):
    a = snap(stage, True); b = snap(stage, False)
    kind = names[stage - 1].split(".")[-1]
    msg = []
    for k in written[kind]:
        eq = torch.equal(a[k].view(torch.uint8), b[k].view(torch.uint8))
        d = (a[k].float() - b[k].float()).abs()
        msg.append(f"{k}: equal={eq} ndiff={int((d>0).sum())} max={float(d.max()):.3e}")
    print(f"stage {stage:2d} {names[stage-1]:9s} " + "; ".join(msg), flush=True)
