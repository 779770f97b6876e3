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


def adamw_state_dict(model, layout, exp_avg, exp_avg_sq, step, *, lr, betas, eps, weight_decay):
    """The flat moment buffers in ``torch.optim.AdamW(model.parameters()).state_dict()`` form -- the 'opt' entry of the
    reference's checkpoints (train_accum.py:121,278): parameter indices follow ``model.parameters()`` (frozen tensors such
    as pos_embed keep their index and have no state)."""
    index = {k: i for i, (k, _) in enumerate(model.named_parameters())}
    state = {}
    for k in layout.names:
        state[index[k]] = {"step": torch.tensor(float(step)), "exp_avg": layout.view(exp_avg, k).detach().clone().cpu(),
                           "exp_avg_sq": layout.view(exp_avg_sq, k).detach().clone().cpu()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": True,
             "params": list(range(len(index)))}
    return {"state": state if step > 0 else {}, "param_groups": [group]}


def load_adamw_state_dict(model, layout, exp_avg, exp_avg_sq, opt_state):
    """Inverse of ``adamw_state_dict`` (also accepts the state of a real ``torch.optim.AdamW`` over ``model.parameters()``);
    returns the step count.  Parameters without state (fresh optimizer, frozen tensors) keep zero moments."""
    names = [k for k, _ in model.named_parameters()]
    steps = 0
    with torch.no_grad():
        exp_avg.zero_(); exp_avg_sq.zero_()
        for idx, st in opt_state.get("state", {}).items():
            k = names[int(idx)]
            if k not in layout.slices:
                continue
            layout.view(exp_avg, k).copy_(st["exp_avg"].to(exp_avg.device))
            layout.view(exp_avg_sq, k).copy_(st["exp_avg_sq"].to(exp_avg.device))
            steps = max(steps, int(float(st["step"])))
    return steps


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


def clip_flat_gradient_(grad, max_norm, grad_scale=1.0):
    """``accelerator.clip_grad_norm_(model.parameters(), max_grad_norm)`` of train_accum.py:235-238 (torch.nn.utils.clip_grad_norm_,
    2-norm over all gradients, coefficient ``max_norm / (total + 1e-6)`` clamped to 1) on the flat gradient buffer, in place and
    without a host sync.  ``grad`` holds the SUM over ranks and ``grad_scale`` the 1 / world the optimizer kernel applies when it
    reads it, so the norm of the averaged gradient is ``grad_scale * |grad|`` (the 16-byte padding between slices is zero).
    Returns the total norm (0-d tensor) like the reference call."""
    total = torch.linalg.vector_norm(grad.float(), 2) * grad_scale
    grad.mul_(torch.clamp(float(max_norm) / (total + 1e-6), max=1.0))
    return total


def cosine_loss_terms(out, ut, loss_scale=1.0):
    """The optional cosine term of the reference's loss (transport.py:196-197, train_accum.py:216-223 with use_cosine_loss):
    ``cos_loss = mean_flat(1 - cosine_similarity(out, ut, dim=1))`` per sample and the gradient of ``cos_loss.mean() *
    loss_scale`` with respect to ``out`` (added to the MSE term's ``dout``).  No shipped config enables it, so it is host-side
    PyTorch on the model output rather than part of the fused loss kernel."""
    o = out.detach().requires_grad_(True)
    with torch.enable_grad():
        cos = torch.mean(1 - torch.nn.functional.cosine_similarity(o, ut, dim=1), dim=list(range(1, o.dim() - 1)))
        g, = torch.autograd.grad(cos.mean() * loss_scale, o)
    return cos.detach(), g


class FusedTrainer:
    def __init__(self, model, *, lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, ema_decay=0.9999, transport=None,
                 process_group=None, max_grad_norm=None):
        self.model = model
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)   # train_accum.py:235 (config: optimizer.max_grad_norm)
        self.last_grad_norm = None
        self.last_cos_loss = None
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
        # (name, pointer, numel) arrays for the batched ABI calls: gradients land in / weights are re-packed from the flat buffers
        n = len(self.names)
        self._c_names = (C.c_char_p * n)(*[k.encode() for k in self.names])
        self._c_numels = (C.c_int64 * n)(*[self.slices[k][1] for k in self.names])
        self._c_grad_ptrs = (C.c_void_p * n)(*[self.grad.data_ptr() + 4 * self.slices[k][0] for k in self.names])
        self._c_param_ptrs = (C.c_void_p * n)(*[self.flat.data_ptr() + 4 * self.slices[k][0] for k in self.names])

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

    # -- checkpoint contract (train_accum.py:273-284,170-187) ------------------------------------------
    def opt_state_dict(self):
        return adamw_state_dict(self.model, self.layout, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr,
                                betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)

    def load_opt_state_dict(self, opt_state):
        self.step_count = load_adamw_state_dict(self.model, self.layout, self.exp_avg, self.exp_avg_sq, opt_state)
        g = (opt_state.get("param_groups") or [{}])[0]
        self.lr = float(g.get("lr", self.lr)); self.betas = tuple(g.get("betas", self.betas))
        self.eps = float(g.get("eps", self.eps)); self.weight_decay = float(g.get("weight_decay", self.weight_decay))

    def load_ema_state_dict(self, ema_state):
        from .checkpoint import strip_module_prefix
        ema_state = strip_module_prefix(ema_state)
        with torch.no_grad():
            for k in self.names:
                self.layout.view(self.ema, k).copy_(ema_state[k].to(self.ema.device))

    def save_checkpoint(self, checkpoint_dir, train_steps, config=None):
        """``{model, ema, opt, config}`` as ``<dir>/<steps:07d>.pt`` -- loadable by the reference's resume / inference code."""
        from .checkpoint import save_checkpoint
        return save_checkpoint(checkpoint_dir, train_steps, self.model.state_dict(), self.ema_state_dict(), self.opt_state_dict(),
                               config or {})

    def resume(self, checkpoint_dir, load_opt=True, map_location="cpu"):
        """Latest checkpoint of ``checkpoint_dir`` into the flat buffers (parameters are views of them, so
        ``load_state_dict`` writes in place); returns the step count parsed from the file name (0: nothing to resume)."""
        from .checkpoint import latest_checkpoint, strip_module_prefix
        path, steps = latest_checkpoint(checkpoint_dir)
        if path is None:
            return 0
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        self.model.load_state_dict(strip_module_prefix(ckpt["model"]))
        self.load_ema_state_dict(ckpt["ema"])
        if load_opt and ckpt.get("opt"):
            self.load_opt_state_dict(ckpt["opt"])
        self.model.mark_weights_dirty(self.names)
        return steps

    # -- trainer input pipeline on the device --------------------------------------------------------
    def prepare(self, x1=None, *, t, x0, moments=None, moments_flip=None, flip=None, eps_post=None, latent_mean=None,
                latent_std=None, latent_multiplier=1.0, want_x1=False):
        """ONE kernel for everything between the stored features and the model input (ldmae_flow_prepare): flip select +
        posterior sample + per-channel normalise + multiplier (datasets/img_latent_dataset.py:76-94, when ``moments``
        [B, 2C, S, S] are given instead of ready latents ``x1``), then xt = t*x1 + (1-t)*x0 and ut = x1 - x0
        (transport.py:136-166, path.py:114-136).  Returns (xt, ut, x1 or None)."""
        L = _lib.lib()
        src = moments if moments is not None else x1
        B, S = src.shape[0], src.shape[-1]
        Cc = src.shape[1] // 2 if moments is not None else src.shape[1]
        dev = src.device
        f = lambda v: None if v is None else v.detach().to(dev).float().contiguous()
        x1c, x0c, tc, mo, mf, ep = f(x1) if moments is None else None, f(x0), f(t), f(moments), f(moments_flip), f(eps_post)
        fl = None if flip is None else flip.detach().to(dev).to(torch.uint8).contiguous()
        mean = None if latent_mean is None else f(latent_mean).reshape(-1)
        std = None if latent_std is None else f(latent_std).reshape(-1)
        xt = torch.empty(B, Cc, S, S, device=dev)
        ut = torch.empty_like(xt)
        x1o = torch.empty_like(xt) if want_x1 else None
        with torch.cuda.device(dev):
            _lib.check(L.ldmae_flow_prepare(_lib.ptr(mo), _lib.ptr(mf), _lib.ptr(fl), _lib.ptr(ep), _lib.ptr(mean), _lib.ptr(std),
                                            float(latent_multiplier), _lib.ptr(x1c), _lib.ptr(x0c), _lib.ptr(tc), _lib.ptr(x1o),
                                            _lib.ptr(xt), _lib.ptr(ut), B, Cc, S * S, _lib.stream_ptr()), "ldmae_flow_prepare")
        return xt, ut, x1o

    # -- one micro-batch: loss + gradients into self.grad ----------------------------------------
    def loss_and_grad(self, x1, y, t=None, x0=None, accumulate=False, loss_scale=1.0, **pipeline):
        """One micro-batch.  accumulate=True adds into the flat gradient buffer instead of overwriting it and loss_scale
        divides the loss (train_accum.py:220-223: loss / gradient_accumulation_steps).  ``pipeline``: keyword arguments of
        ``prepare`` (moments=..., moments_flip=..., flip=..., eps_post=..., latent_mean/std/multiplier) to start from the
        stored posterior moments instead of ready latents (then ``x1`` may be None)."""
        m, L = self.model, _lib.lib()
        src = pipeline.get("moments") if pipeline.get("moments") is not None else x1
        B = src.shape[0]
        if t is None or x0 is None:
            shape_like = src[:, : src.shape[1] // 2] if pipeline.get("moments") is not None else src
            t, x0, _ = self.transport.sample(shape_like)               # reference draws: randn_like + logit-normal t
        xt, ut, _ = self.prepare(x1, t=t, x0=x0, **pipeline)
        t = t.to(xt.device).float().contiguous()
        if m.training and m.y_embedder.dropout_prob > 0:
            y = m.y_embedder.token_drop(y)
        y = y.long().contiguous()
        h = m._ensure_handle(xt.device, B)
        out = torch.empty_like(xt)
        loss = torch.empty(B, device=xt.device)
        dout = torch.empty_like(xt)
        with torch.cuda.device(xt.device):
            st = _lib.stream_ptr()
            _lib.check(L.ldmae_dit_train_forward(h, _lib.ptr(xt), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), B, st), "train_forward")
            # loss = mean_flat((out - ut)^2); dout = d (mean(loss) * loss_scale) / d out  -- one pass (ldmae_flow_loss)
            _lib.check(L.ldmae_flow_loss(_lib.ptr(out), _lib.ptr(ut), _lib.ptr(loss), _lib.ptr(dout), float(loss_scale), B,
                                         out[0].numel(), st), "flow_loss")
            if self.transport.use_cosine_loss:
                # train_accum.py:216-218: loss = cos_loss.mean() + mse.mean(); the per-sample MSE stays what step() returns
                self.last_cos_loss, dcos = cosine_loss_terms(out, ut, loss_scale)
                dout.add_(dcos)
            _lib.check(L.ldmae_dit_backward(h, _lib.ptr(dout), B, st), "backward")
            _lib.check(L.ldmae_dit_grad_read_many(h, self._c_names, self._c_grad_ptrs, self._c_numels, len(self.names),
                                                  1 if accumulate else 0, st), "grad_read_many")
        return loss, out

    def optimizer_step(self):
        """all-reduce (mean) + AdamW + EMA on the flat buffers; marks the library's bf16 weight copies stale."""
        grad_scale = reduce_gradients(self.grad, self.pg)
        if self.max_grad_norm is not None:
            self.last_grad_norm = clip_flat_gradient_(self.grad, self.max_grad_norm, grad_scale)
        self.step_count += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ldmae_adamw_ema_step(
                _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.ema),
                self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                self.ema_decay, grad_scale, _lib.stream_ptr()), "adamw_ema_step")
        # parameters changed behind torch's version counters: re-pack the library's bf16 copies from the flat buffer in one call
        m = self.model
        if m._handle is not None and m._handle_dev == self.device and m._handle_sig is not None:
            with torch.cuda.device(self.device):
                st = _lib.stream_ptr()
                _lib.check(_lib.lib().ldmae_dit_load_tensors(m._handle, self._c_names, self._c_param_ptrs, self._c_numels,
                                                             len(self.names), st), "load_tensors")
                _lib.check(_lib.lib().ldmae_dit_finalize(m._handle, st), "ldmae_dit_finalize")
        else:
            m.mark_weights_dirty(self.names)

    def step_from_moments(self, moments, moments_flip, y, *, latent_mean=None, latent_std=None, latent_multiplier=1.0,
                          flip=None, eps_post=None, t=None, x0=None):
        """One optimizer step straight from the stored features of extract_features.py (posterior moments of each image and
        of its horizontal flip, [B, 2C, S, S]): the dataset's coin flip, posterior sample and normalisation
        (img_latent_dataset.py:76-94) run inside the same kernel that builds xt / ut.  Random draws can be injected."""
        B, dev = moments.shape[0], moments.device
        if flip is None:
            flip = torch.rand(B, device=dev) <= 0.5                     # "latents" if uniform > 0.5 else "latents_flip"
        if eps_post is None:
            eps_post = torch.randn(B, moments.shape[1] // 2, *moments.shape[2:], device=dev)
        loss, _ = self.loss_and_grad(None, y, t, x0, moments=moments, moments_flip=moments_flip, flip=flip, eps_post=eps_post,
                                     latent_mean=latent_mean, latent_std=latent_std, latent_multiplier=latent_multiplier)
        self.optimizer_step()
        return loss

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
