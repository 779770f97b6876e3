"""Drop-in for the reference's ``models/lightningdit.py`` (LightningDiT denoiser).

Same public surface -- ``LightningDiT(**kw)``, ``forward(x, t, y)``, ``forward_with_cfg(...)``,
``unpatchify``, the nine ``LightningDiT_models`` registry entries (reference
models/lightningdit.py:498-531) -- and the same ``state_dict`` keys/shapes, so
checkpoints written by the reference's ``train_accum.py`` load with ``strict=True``.

The sub-modules below only *hold parameters* under the reference's names; no PyTorch op of theirs
runs on the path.  ``forward`` hands device pointers to ``libldmae_b200.so`` (hand-written sm_100a
CUDA: tcgen05 GEMMs with fused adaLN/RMSNorm/RoPE/SwiGLU/residual epilogues, tcgen05 flash
attention).  There is no CPU or eager fallback: without a B200 and the built library it raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

__all__ = ["LightningDiT", "LightningDiT_models", "get_2d_sincos_pos_embed"]


# ----------------------------------------------------------------------------- parameter containers
class _RMSNormParams(nn.Module):
    """reference models/rmsnorm.py:34-50 (weight only; the arithmetic is fused into the GEMM epilogues)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class _PatchEmbedParams(nn.Module):
    """timm PatchEmbed attributes used by the reference: .proj (Conv2d k=stride=p), .num_patches, .patch_size."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)


class _TimestepEmbedderParams(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.frequency_embedding_size = frequency_embedding_size
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size))


class _LabelEmbedderParams(nn.Module):
    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        use_cfg_embedding = dropout_prob > 0
        self.embedding_table = nn.Embedding(num_classes + use_cfg_embedding, hidden_size)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob

    def token_drop(self, labels, force_drop_ids=None):
        """reference lightningdit.py:152-161."""
        if force_drop_ids is None:
            drop_ids = torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob
        else:
            drop_ids = force_drop_ids == 1
        return torch.where(drop_ids, self.num_classes, labels)


class _RopeBuffers(nn.Module):
    """reference models/pos_embed.py:96-133: persistent buffers freqs_cos / freqs_sin [T, head_dim]."""

    def __init__(self, dim, pt_seq_len):
        super().__init__()
        freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        t = torch.arange(pt_seq_len) / pt_seq_len * pt_seq_len
        f = (t[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
        S = pt_seq_len
        full = torch.cat([f[:, None, :].expand(S, S, dim), f[None, :, :].expand(S, S, dim)], dim=-1).reshape(S * S, 2 * dim)
        self.register_buffer("freqs_cos", full.cos().contiguous())
        self.register_buffer("freqs_sin", full.sin().contiguous())


class _AttentionParams(nn.Module):
    def __init__(self, dim, num_heads, qk_norm, use_rmsnorm=True):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        # reference lightningdit.py:57-61: RMSNorm(head_dim) with use_rmsnorm, else nn.LayerNorm(head_dim) (affine, eps 1e-5)
        norm = _RMSNormParams if use_rmsnorm else nn.LayerNorm
        self.q_norm = norm(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm(self.head_dim) if qk_norm else nn.Identity()
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    """timm Mlp parameter names (fc1, fc2); reference lightningdit.py:219-224 uses it with GELU(approximate='tanh')."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features, bias=True)
        self.fc2 = nn.Linear(hidden_features, in_features, bias=True)


class _SwiGLUParams(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.w12 = nn.Linear(in_features, 2 * hidden_features, bias=True)
        self.w3 = nn.Linear(hidden_features, in_features, bias=True)


def _block_norm(hidden_size, use_rmsnorm):
    """reference lightningdit.py:195-200,258-261: RMSNorm, or LayerNorm without affine parameters (eps 1e-6)."""
    return _RMSNormParams(hidden_size) if use_rmsnorm else nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)


class _BlockParams(nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio, use_qknorm, wo_shift, use_swiglu=True, use_rmsnorm=True):
        super().__init__()
        self.norm1 = _block_norm(hidden_size, use_rmsnorm)
        self.norm2 = _block_norm(hidden_size, use_rmsnorm)
        self.attn = _AttentionParams(hidden_size, num_heads, use_qknorm, use_rmsnorm)
        mlp_hidden = int(hidden_size * mlp_ratio)
        self.mlp = _SwiGLUParams(hidden_size, int(2 / 3 * mlp_hidden)) if use_swiglu else _MlpParams(hidden_size, mlp_hidden)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, (4 if wo_shift else 6) * hidden_size))
        self.wo_shift = wo_shift


class _FinalLayerParams(nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels, use_rmsnorm=True):
        super().__init__()
        self.norm_final = _block_norm(hidden_size, use_rmsnorm)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size))


# ----------------------------------------------------------------------------- model
class LightningDiT(nn.Module):
    """Diffusion transformer denoiser; constructor arguments as reference lightningdit.py:279-297."""

    def __init__(self, input_size=32, patch_size=2, in_channels=32, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, use_qknorm=False,
                 use_swiglu=False, use_rope=False, use_rmsnorm=False, wo_shift=False, use_checkpoint=False):
        super().__init__()
        hd = hidden_size // num_heads
        if not (use_swiglu and use_rmsnorm) and hd != 64:
            raise NotImplementedError("the LayerNorm / GELU-Mlp variants (use_rmsnorm=False / use_swiglu=False; reference "
                                      f"lightningdit.py:195-224) are built for head_dim 64, got {hd}")
        if not (hd == 64 or (64 < hd <= 128 and hd % 8 == 0)):
            raise NotImplementedError(f"ldmae_b200 attention is built for head_dim 64 (B, L, 1p0B, 1p6B: tuned kernels) and for "
                                      f"multiples of 8 in (64, 128] (XL, head_dim 72: 128-column head slots); got {hd}")
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels if not learn_sigma else in_channels * 2
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.use_rope = use_rope
        self.use_rmsnorm = use_rmsnorm
        self.use_qknorm = use_qknorm
        self.use_swiglu = use_swiglu
        self.wo_shift = wo_shift
        self.depth = depth
        self.hidden_size = hidden_size
        self.input_size = input_size
        self.use_checkpoint = use_checkpoint     # activations are never stored on the inference path
        self.x_embedder = _PatchEmbedParams(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.t_embedder = _TimestepEmbedderParams(hidden_size)
        self.y_embedder = _LabelEmbedderParams(num_classes, hidden_size, class_dropout_prob)
        num_patches = self.x_embedder.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, hidden_size), requires_grad=False)
        self.feat_rope = _RopeBuffers(hidden_size // num_heads // 2, input_size // patch_size) if use_rope else None
        self.blocks = nn.ModuleList([_BlockParams(hidden_size, num_heads, mlp_ratio, use_qknorm, wo_shift, use_swiglu, use_rmsnorm)
                                     for _ in range(depth)])
        self.final_layer = _FinalLayerParams(hidden_size, patch_size, self.out_channels, use_rmsnorm)
        self.initialize_weights()
        self._handle = None
        self._handle_sig = None
        self._handle_dev = None
        self._sig_items = None
        self._dirty_names = None

    # -- init exactly as the reference (lightningdit.py:340-374)
    def initialize_weights(self):
        def _basic_init(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

        self.apply(_basic_init)
        pe = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.x_embedder.num_patches ** 0.5))
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for block in self.blocks:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    def unpatchify(self, x):
        """reference lightningdit.py:376-389 (index-only; the kernel path scatters directly to NCHW)."""
        c, p = self.out_channels, self.x_embedder.patch_size[0]
        h = w = int(x.shape[1] ** 0.5)
        assert h * w == x.shape[1]
        x = x.reshape(x.shape[0], h, w, p, p, c)
        x = torch.einsum("nhwpqc->nchpwq", x)
        return x.reshape(x.shape[0], c, h * p, h * p)

    # -- C handle management ---------------------------------------------------------------
    def __deepcopy__(self, memo):
        # EMA copies (reference train_accum.py:92) must not share the C handle
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_handle", "_handle_sig", "_handle_dev", "_sig_items", "_dirty_names"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if getattr(self, "_handle", None):
            _lib.lib().ldmae_dit_destroy(self._handle)
            self._handle = None
            self._handle_sig = None

    # -- weight cache ------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float(): parameters move, the library's packed copies are stale
        self._sig_items = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._sig_items = None
        return super().load_state_dict(*a, **k)

    def mark_weights_dirty(self, names=None):
        """Tell the library that parameters changed behind PyTorch's version counters: in-place edits through ``p.data``
        (custom EMA / clipping / init code), raw-pointer writes (``ldmae_adamw_ema_step``), or re-assigned Parameter
        objects.  Ordinary edits (optimizer steps on the parameters, ``load_state_dict``, ``.to()``) are picked up
        automatically.  ``names``: state_dict keys to re-upload (default: everything)."""
        if names is None:
            self._sig_items = None
            self._handle_sig = None
        else:
            self._dirty_names = set(getattr(self, "_dirty_names", None) or ()) | set(names)

    refresh_weights = mark_weights_dirty

    def _signature_items(self):
        """(key, tensor) of everything the library holds a copy of; the list is cached (Parameter objects are stable
        across .to() and load_state_dict), the per-call check only reads data_ptr / _version of each."""
        items = getattr(self, "_sig_items", None)
        if items is None:
            items = list(self.state_dict(keep_vars=True).items())
            self._sig_items = items
        return items

    def _ensure_handle(self, device, batch):
        """Create the library handle on first use and re-upload the weights that changed since the last call
        (load_state_dict, optimizer / EMA step, .to(), mark_weights_dirty)."""
        if device.type != "cuda":
            raise _lib.LdmaeError("ldmae_b200.LightningDiT runs on a CUDA (B200) device only; there is no CPU path")
        L = _lib.lib()
        items = self._signature_items()
        if self._handle is None or self._handle_dev != device:
            self._release()
            cfg = _lib.DitConfig(
                depth=self.depth, hidden_size=self.hidden_size, num_heads=self.num_heads, patch_size=self.patch_size,
                input_size=self.input_size, in_channels=self.in_channels,
                num_embeddings=self.y_embedder.embedding_table.weight.shape[0],
                mlp_hidden=self.blocks[0].mlp.w3.weight.shape[1] if self.use_swiglu else self.blocks[0].mlp.fc1.weight.shape[0],
                learn_sigma=int(self.learn_sigma),
                use_qknorm=int(self.use_qknorm), use_swiglu=int(self.use_swiglu), use_rope=int(self.use_rope),
                use_rmsnorm=int(self.use_rmsnorm), wo_shift=int(self.wo_shift), max_batch=max(1, batch))
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(L.ldmae_dit_create(C.byref(cfg), C.byref(h)), "ldmae_dit_create")
            self._handle, self._handle_dev, self._handle_sig = h, device, None
        sig = {k: (t.data_ptr(), t._version) for k, t in items}
        old = self._handle_sig
        dirty = getattr(self, "_dirty_names", None)
        if old != sig or dirty:
            changed = [(k, t) for k, t in items if old is None or old.get(k) != sig[k] or (dirty and k in dirty)]
            with torch.cuda.device(device):          # pack kernels must run on the handle's device and its current stream
                st = _lib.stream_ptr()
                keep = []
                for k, v in changed:
                    t = v.detach()
                    if t.device != device:
                        raise _lib.LdmaeError(f"parameter {k} is on {t.device}, expected {device}")
                    if t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.float().contiguous()
                        keep.append(t)
                    _lib.check(L.ldmae_dit_load_tensor(self._handle, k.encode(), _lib.ptr(t), t.numel(), st), f"load {k}")
                _lib.check(L.ldmae_dit_finalize(self._handle, st), "ldmae_dit_finalize")
                if keep:
                    torch.cuda.current_stream().synchronize()      # the temporaries above may be freed after this
            self._handle_sig = sig
            self._dirty_names = None
        return self._handle

    @staticmethod
    def _prep(x, t, y):
        x = x.detach().float().contiguous()
        t = t.detach().float().contiguous()
        y = y.detach().long().contiguous()
        return x, t, y

    # -- forward (reference lightningdit.py:391-418) ------------------------------------------
    def forward(self, x, t=None, y=None):
        if self.training and self.y_embedder.dropout_prob > 0:
            y = self.y_embedder.token_drop(y)             # lightningdit.py:165-167
        x, t, y = self._prep(x, t, y)
        B = x.shape[0]
        h = self._ensure_handle(x.device, B)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not (self.use_swiglu and self.use_rmsnorm):
                raise NotImplementedError("training (autograd through LightningDiT.forward) is built for the shipped recipe "
                                          "(use_rmsnorm=True, use_swiglu=True); the LayerNorm / GELU-Mlp variants are "
                                          "inference-only -- call under torch.no_grad()")
            # differentiable path (train_accum.py:215-230): the library keeps the activations, autograd sees one node whose
            # inputs are the trainable parameters, so .grad / DDP hooks / torch optimizers work as with the reference
            named = [(k, p) for k, p in self.named_parameters() if p.requires_grad]
            return _DitTrainFunction.apply(self, h, x, t, y, tuple(k for k, _ in named), *[p for _, p in named])
        out = torch.empty(B, self.in_channels, x.shape[2], x.shape[3], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ldmae_dit_forward(h, _lib.ptr(x), _lib.ptr(t), 0.0, _lib.ptr(y), _lib.ptr(out), B, B,
                                                    _lib.stream_ptr()), "ldmae_dit_forward")
        return out

    def forward_with_cfg(self, x, t, y, cfg_scale, cfg_interval=None, cfg_interval_start=None):
        """reference lightningdit.py:420-442: model on cat[half, half]; guidance on channels [:3] only;
        below cfg_interval_start the guided channels are the conditional prediction."""
        x, t, y = self._prep(x, t, y)
        n = x.shape[0] // 2
        use_guidance = 1
        if cfg_interval is True:
            if float(t[0]) < cfg_interval_start:          # same host sync as the reference (:437-438)
                use_guidance = 0
        h = self._ensure_handle(x.device, 2 * n)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ldmae_dit_forward_with_cfg(h, _lib.ptr(x), _lib.ptr(t), 0.0, _lib.ptr(y), _lib.ptr(out), n,
                                                             float(cfg_scale), use_guidance, _lib.stream_ptr()),
                       "ldmae_dit_forward_with_cfg")
        return out

    # -- fused sampler entry used by transport.Sampler when model_fn is one of our bound methods
    def _sample_ode(self, x, y, n, use_cfg, cfg_scale, cfg_interval_start, tgrid, method, keep_trajectory, flags=0):
        x = x.detach().float().contiguous().clone()
        y = y.detach().long().contiguous()
        h = self._ensure_handle(x.device, x.shape[0])
        npts = len(tgrid)
        grid = (C.c_float * npts)(*[float(v) for v in tgrid])
        traj = torch.empty((npts,) + tuple(x.shape), device=x.device, dtype=torch.float32) if keep_trajectory else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ldmae_sample_ode(h, _lib.ptr(x), _lib.ptr(y), n, int(use_cfg), float(cfg_scale),
                                                   float(cfg_interval_start), grid, npts, int(method), _lib.ptr(traj),
                                                   int(flags), _lib.stream_ptr()), "ldmae_sample_ode")
        return x, traj


class _DitTrainFunction(torch.autograd.Function):
    """LightningDiT.forward as one autograd node: ldmae_dit_train_forward / ldmae_dit_backward (sm_100a kernels).
    Gradients w.r.t. the latent input are not produced (transport.training_losses never asks for them)."""

    @staticmethod
    def forward(ctx, model, h, x, t, y, names, *params):
        B = x.shape[0]
        out = torch.empty(B, model.in_channels, x.shape[2], x.shape[3], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ldmae_dit_train_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), B,
                                                          _lib.stream_ptr()), "ldmae_dit_train_forward")
        ctx.h, ctx.B, ctx.names, ctx.dev = h, B, names, x.device
        ctx.shapes = [tuple(p.shape) for p in params]
        # the library keeps ONE set of activations per handle: remember which forward they belong to
        ctx.gen = int(_lib.lib().ldmae_dit_generation(h))
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        now = int(L.ldmae_dit_generation(ctx.h))
        if now != ctx.gen:
            raise _lib.LdmaeError(
                "LightningDiT backward: the model ran another forward (training, no_grad, forward_with_cfg or a sampler) after "
                f"the forward this backward belongs to (workspace generation {ctx.gen}, now {now}).  ldmae_b200 keeps the "
                "activations of ONE training forward per model: call backward() before the next forward; expressions like "
                "(model(a) + model(b)).backward() and activation checkpointing are not supported -- use gradient accumulation")
        dout = dout.detach().float().contiguous()
        grads = []
        with torch.cuda.device(ctx.dev):
            st = _lib.stream_ptr()
            _lib.check(L.ldmae_dit_backward(ctx.h, _lib.ptr(dout), ctx.B, st), "ldmae_dit_backward")
            for name, shape in zip(ctx.names, ctx.shapes):
                g = torch.empty(shape, device=ctx.dev, dtype=torch.float32)
                _lib.check(L.ldmae_dit_grad_read(ctx.h, name.encode(), _lib.ptr(g), g.numel(), st), f"grad {name}")
                grads.append(g)
        return (None, None, None, None, None, None, *grads)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False, extra_tokens=0):
    """reference lightningdit.py:444-491 (float64 omega; w-coordinate in the first half of the channels)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])

    def _1d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float64)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    pe = np.concatenate([_1d(embed_dim // 2, grid[0]), _1d(embed_dim // 2, grid[1])], axis=1)
    if cls_token and extra_tokens > 0:
        pe = np.concatenate([np.zeros([extra_tokens, embed_dim]), pe], axis=0)
    return pe


# ----------------------------------------------------------------------------- registry (reference :498-531)
def _cfg(depth, hidden_size, patch_size, num_heads):
    def make(**kwargs):
        return LightningDiT(depth=depth, hidden_size=hidden_size, patch_size=patch_size, num_heads=num_heads, **kwargs)
    return make


LightningDiT_models = {
    "LightningDiT-B/1": _cfg(12, 768, 1, 12), "LightningDiT-B/2": _cfg(12, 768, 2, 12),
    "LightningDiT-L/2": _cfg(24, 1024, 2, 16),
    "LightningDiT-XL/1": _cfg(28, 1152, 1, 16), "LightningDiT-XL/2": _cfg(28, 1152, 2, 16),
    "LightningDiT-1p0B/1": _cfg(24, 1536, 1, 24), "LightningDiT-1p0B/2": _cfg(24, 1536, 2, 24),
    "LightningDiT-1p6B/1": _cfg(28, 1792, 1, 28), "LightningDiT-1p6B/2": _cfg(28, 1792, 2, 28),
}
