"""GPU diagnostic: poison the workspace with NaN bytes before a forward; any read-before-write shows as NaN / changed output."""
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
for B in (1, 3):
    x = torch.randn(B, 16, 32, 32, generator=g).to(dev); t = torch.rand(B, generator=g).to(dev); y = torch.randint(0, 1000, (B,), generator=g).to(dev)
    steady = [m(x, t, y).clone() for _ in range(3)][-1]
    h = m._handle
    res = {}
    for byte in (0x00, 0xFF, 0x3C):
        _lib.check(L.ldmae_dit_debug_poison(h, byte, _lib.stream_ptr()))
        o = m(x, t, y).clone()
        res[byte] = o
        print(f"B={B} poison 0x{byte:02X}: nan count {int(torch.isnan(o).sum())} equal-to-steady {torch.equal(o, steady)} max diff {float((o-steady).abs().nan_to_num(0).max()):.3e}", flush=True)
    M = B * 1024
    bufs = {"xres": (M * 768, torch.float32), "abuf": (M * 768, torch.bfloat16), "qkv": (M * 2304, torch.bfloat16), "obuf": (M * 768, torch.bfloat16),
            "hbuf": (M * 2048, torch.bfloat16), "ssq": (M * 6, torch.float32), "mods": (B * 74 * 768, torch.float32), "cvec_c": (B * 768, torch.float32),
            "cvec_qkv": (12 * B * 2304, torch.float32), "cvec_12": (12 * B * 4096, torch.float32), "gmul": (25 * B * 768, torch.float32),
            "sc": (B * 768, torch.bfloat16), "th1": (B * 768, torch.float32), "shift_bf16": (25 * B * 768, torch.bfloat16)}
    def snap(stage, byte):
        _lib.check(L.ldmae_dit_debug_poison(h, byte, _lib.stream_ptr()))
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
    written = {"cond": ["th1", "cvec_c", "sc"], "adaln": ["mods", "gmul", "shift_bf16"], "shiftvec": ["cvec_qkv", "cvec_12"], "patch": ["xres", "abuf", "ssq"],
               "qkv": ["qkv"], "attn": ["obuf"], "proj": ["xres", "abuf", "ssq"], "w12": ["hbuf"], "w3": ["xres", "abuf", "ssq"]}
    for stage in range(1, 15):
        a, b = snap(stage, 0x00), snap(stage, 0xFF)
        kind = names[stage - 1].split(".")[-1]
        msg = []
        for k in written[kind]:
            na, nb = int(torch.isnan(a[k].float()).sum()), int(torch.isnan(b[k].float()).sum())
            eq = torch.equal(a[k].view(torch.uint8), b[k].view(torch.uint8))
            msg.append(f"{k}: nan(0x00)={na} nan(0xFF)={nb} equal={eq}")
        print(f"B={B} stage {stage:2d} {names[stage-1]:9s} " + "; ".join(msg), flush=True)
    _lib.check(L.ldmae_dit_debug_stop(h, -1))
