"""GPU stress: repeated full weight upload + ldmae_dit_finalize (RoPE table check, score bounds) on a LightningDiT-B/1 handle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldmae_b200.pipeline import build_sampling_models
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
fails = 0
for rep in range(3):
    m, _ = build_sampling_models(dev, seed=rep)
    x = torch.randn(2, 16, 32, 32, device=dev); t = torch.rand(2, device=dev); y = torch.randint(0, 1000, (2,), device=dev)
    for i in range(n):
        m.mark_weights_dirty()
        try:
            m(x, t, y)
        except Exception as e:
            fails += 1
            print("FAIL", rep, i, str(e)[:200], flush=True)
    del m
print(f"stress_finalize: {3 * n} uploads, {fails} failures")
