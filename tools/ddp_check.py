"""2-GPU check of the drop-in training route (train_accum.py:105,215-246): DistributedDataParallel around the ldmae_b200
LightningDiT, transport.training_losses, loss.backward(), torch.optim.AdamW -- gradients must be the cross-rank mean and the
replicas must stay identical.  Run: torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ldmae_b200.models.lightningdit import LightningDiT
from ldmae_b200.transport import create_transport

def make():
    torch.manual_seed(0)                  # identical initial weights for every replica and for the plain reference copy
    m = LightningDiT(input_size=8, patch_size=1, in_channels=16, hidden_size=128, depth=2, num_heads=2, num_classes=10,
                     use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "adaLN_modulation.1" in k or k.startswith("final_layer.linear"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.cuda().eval()           # eval: no label dropout (deterministic comparison)

tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
g = torch.Generator().manual_seed(100)
B = 4
data = [(torch.randn(B, 16, 8, 8, generator=g), torch.randn(B, 16, 8, 8, generator=g), torch.rand(B, generator=g),
         torch.randint(0, 10, (B,), generator=g)) for _ in range(world)]

def grads_for(model, r):
    x1, x0, t, y = (v.cuda() for v in data[r])
    tr.sample = lambda x1_, *a, **k: (t, x0, x1_)
    model.zero_grad(set_to_none=True)
    loss = tr.training_losses(model, x1, dict(y=y))["loss"].mean()
    loss.backward()
    return loss

ddp = DDP(make(), device_ids=[local])
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
loss = grads_for(ddp, rank)
ddp_grads = {k: p.grad.clone() for k, p in ddp.module.named_parameters() if p.grad is not None}
opt.step()
# reference: every rank's local gradient computed on a plain replica, averaged by hand
plain = make()
acc = None
for r in range(world):
    grads_for(plain, r)
    cur = {k: p.grad.clone() for k, p in plain.named_parameters() if p.grad is not None}
    acc = cur if acc is None else {k: acc[k] + cur[k] for k in acc}
worst = 0.0
for k in acc:
    ref = acc[k] / world
    worst = max(worst, float((ddp_grads[k] - ref).norm() / ref.norm().clamp_min(1e-30)))
# replicas identical after the optimizer step
chk = torch.stack([p.detach().double().sum() for p in ddp.parameters()]).sum().reshape(1)
gath = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(gath, chk)
if rank == 0:
    same = all(torch.equal(gath[0], v) for v in gath)
    print(f"DDP drop-in route on {world} GPUs: {len(ddp_grads)} gradients, worst rel. deviation from the hand-averaged mean {worst:.2e}; "
          f"replicas identical after AdamW step: {same}; loss {float(loss):.4f}")
    assert worst < 1e-3 and same
dist.destroy_process_group()
