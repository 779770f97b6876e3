"""One optimizer step of the reference's training loop (LDMAE/train_accum.py:203-246) as one object.

    x1 (latents), y (labels)  ->  transport draw (t, x0), xt = t*x1 + (1-t)*x0, ut = x1 - x0   (transport.py:136-166, path.py:114-136)
    ->  LightningDiT training forward (label dropout as lightningdit.py:157-160)               (libldmae_b200: keeps activations)
    ->  loss = mean_flat((v - ut)^2).mean()                                                      (transport.py:195, train_accum.py:220)
    ->  backward (all parameter gradients, written into ONE flat fp32 buffer)
    ->  data-parallel: one NCCL all-reduce of that flat buffer (train_accum.py:105,230 -- DDP's bucketed all-reduce)
    ->  fused AdamW + EMA over the flat parameter / moment / EMA buffers                         (train_accum.py:121,240-246,337-347)

The drop-in route for train_accum.py itself is unchanged PyTorch: ``LightningDiT.forward`` is an autograd node, so
``accelerator.backward(loss)``, DDP hooks and ``torch.optim.AdamW`` work on ``model.parameters()`` as with the reference.
This class is the B200-native fast path for the same arithmetic (no per-tensor optimizer launches, no gradient
copies, one collective).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .transport import create_transport


class FlatLayout:
    """Host-side layout of the training state: every trainable parameter becomes a view into ONE flat fp32 buffer
    (16-byte aligned slices, in ``named_parameters`` order), with same-shaped flat buffers for the gradients, the two Adam
    moments and the EMA copy.  Device-agnostic (the layout logic is covered by the CPU tests)."""

    def __init__(self, model):
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        sizes = [(p.numel() + 3) // 4 * 4 for _, p in named]
        total = sum(sizes)
        self.device = dev
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.names, self.slices = [], {}
        off = 0
        with torch.no_grad():
            for (k, p), sz in zip(named, sizes):
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)          # parameters become views of the flat buffer
                self.names.append(k)
                self.slices[k] = (off, n, tuple(p.shape))
                off += sz

    def view(self, buf, name):
        off, n, shape = self.slices[name]
        return buf[off:off + n].view(shape)


def reduce_gradients(grad, group=None):
    """Data-parallel gradient exchange (the role of DDP's bucketed all-reduce at train_accum.py:105,230): ONE sum all-reduce
    of the flat gradient buffer; returns the factor (1 / world size) the optimizer kernel folds into its gradient read."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(grad, group=group)
    return 1.0 / world


class FusedTrainer:
    def __init__(self, model, *, lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, ema_decay=0.9999, transport=None,
                 process_group=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = float(lr), betas, float(eps), float(weight_decay), float(ema_decay)
        self.transport = transport or create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False,
                                                       use_lognorm=True)
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        if next(model.parameters()).device.type != "cuda":
            raise _lib.LdmaeError("FusedTrainer needs the model on a CUDA (B200) device")
        self.layout = FlatLayout(model)
        self.flat, self.grad = self.layout.flat, self.layout.grad
        self.names, self.slices = self.layout.names, self.layout.slices
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        dev = self.layout.device
        self.ema = self.flat.clone()                                   # train_accum.py:92 (deepcopy of the fresh model)
        self.step_count = 0
        self.device = dev

    # -- views ---------------------------------------------------------------------------------
    def grad_of(self, name):
        off, n, shape = self.slices[name]
        return self.grad[off:off + n].view(shape)

    def ema_state_dict(self):
        """state_dict of the EMA model (checkpoint key 'ema', train_accum.py:277): frozen tensors are the model's own."""
        sd = {k: v.clone() for k, v in self.model.state_dict().items()}
        for k, (off, n, shape) in self.slices.items():
            sd[k] = self.ema[off:off + n].view(shape).clone()
        return sd

    # -- one micro-batch: loss + gradients into self.grad ----------------------------------------
    def loss_and_grad(self, x1, y, t=None, x0=None, accumulate=False, loss_scale=1.0):
        """One micro-batch.  accumulate=True adds into the flat gradient buffer instead of overwriting it and loss_scale
        divides the loss (train_accum.py:220-223: loss / gradient_accumulation_steps)."""
        m, L = self.model, _lib.lib()
        B = x1.shape[0]
        if t is None or x0 is None:
            t, x0, x1 = self.transport.sample(x1)                      # reference draws: randn_like + logit-normal t
        t = t.to(x1).float().contiguous()
        tt = t.view(B, 1, 1, 1)
        xt = (tt * x1 + (1 - tt) * x0).float().contiguous()
        ut = x1 - x0
        if m.training and m.y_embedder.dropout_prob > 0:
            y = m.y_embedder.token_drop(y)
        y = y.long().contiguous()
        h = m._ensure_handle(x1.device, B)
        out = torch.empty_like(xt)
        with torch.cuda.device(x1.device):
            st = _lib.stream_ptr()
            _lib.check(L.ldmae_dit_train_forward(h, _lib.ptr(xt), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), B, st), "train_forward")
            diff = out - ut
            loss = (diff * diff).mean(dim=(1, 2, 3))                   # mean_flat
            dout = (diff * (2.0 * loss_scale / (diff[0].numel() * B))).contiguous()   # d (mean(loss) * loss_scale) / d out
            _lib.check(L.ldmae_dit_backward(h, _lib.ptr(dout), B, st), "backward")
            base = self.grad.data_ptr()
            for k in self.names:
                off, n, _ = self.slices[k]
                fetch = L.ldmae_dit_grad_accumulate if accumulate else L.ldmae_dit_grad_read
                _lib.check(fetch(h, k.encode(), C.c_void_p(base + 4 * off), n, st), f"grad {k}")
        return loss, out

    def optimizer_step(self):
        """all-reduce (mean) + AdamW + EMA on the flat buffers; marks the library's bf16 weight copies stale."""
        grad_scale = reduce_gradients(self.grad, self.pg)
        self.step_count += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ldmae_adamw_ema_step(
                _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.ema),
                self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                self.ema_decay, grad_scale, _lib.stream_ptr()), "adamw_ema_step")
        self.model.mark_weights_dirty(self.names)                      # parameters changed behind torch's version counters

    def step(self, x1, y, t=None, x0=None, micro_batches=1):
        """One optimizer step; micro_batches > 1 splits the batch and accumulates gradients (train_accum.py's
        gradient_accumulation_steps) -- one all-reduce per optimizer step, not per micro-batch."""
        if micro_batches > 1:
            losses = []
            for i, (xs, ys) in enumerate(zip(x1.chunk(micro_batches), y.chunk(micro_batches))):
                ts = t.chunk(micro_batches)[i] if t is not None else None
                x0s = x0.chunk(micro_batches)[i] if x0 is not None else None
                l, _ = self.loss_and_grad(xs, ys, ts, x0s, accumulate=i > 0, loss_scale=1.0 / micro_batches)
                losses.append(l)
            loss = torch.cat(losses)
        else:
            loss, _ = self.loss_and_grad(x1, y, t, x0)
        self.optimizer_step()
        return loss
