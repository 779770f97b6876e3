"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the LDMAE hot path.

Plain, functional, fp32 PyTorch-on-CPU restatement of the reference algorithm for the path
BASELINE.json's ``north_star`` names.  Every function cites the reference file:line it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the CPU baseline -- never as
the product path (``ldmae_b200`` raises if its CUDA library is missing; it never falls back here).

Pinning: the reference has no tests, golden vectors or checkpoints of its own,
so the oracle is pinned against outputs of the reference itself, run in the authoring container
through ``oracle/refshim.py`` by ``oracle/make_golden.py``; the fixtures live in ``tests/golden/``
and ``tests/test_oracle_golden.py`` asserts this file reproduces them.  The torchdiffeq boundary
(``transport/integrators.py:118``) is un-vendored and un-pinned upstream: fixed-grid Euler/Heun are
restated from torchdiffeq's published algorithm -- "parity unpinned" for that single call.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# ----------------------------------------------------------------------------- specs
# registry: models/lightningdit.py:498-531
_DIT_REGISTRY = {
    "LightningDiT-B/1": (12, 768, 1, 12), "LightningDiT-B/2": (12, 768, 2, 12),
    "LightningDiT-L/2": (24, 1024, 2, 16),
    "LightningDiT-XL/1": (28, 1152, 1, 16), "LightningDiT-XL/2": (28, 1152, 2, 16),
    "LightningDiT-1p0B/1": (24, 1536, 1, 24), "LightningDiT-1p0B/2": (24, 1536, 2, 24),
    "LightningDiT-1p6B/1": (28, 1792, 1, 28), "LightningDiT-1p6B/2": (28, 1792, 2, 28),
}


@dataclass
class DiTSpec:
    """Constructor arguments of LightningDiT (models/lightningdit.py:279-297)."""
    depth: int = 12
    hidden_size: int = 768
    patch_size: int = 1
    num_heads: int = 12
    input_size: int = 32
    in_channels: int = 16
    mlp_ratio: float = 4.0
    class_dropout_prob: float = 0.1
    num_classes: int = 1000
    learn_sigma: bool = False
    use_qknorm: bool = True
    use_swiglu: bool = True
    use_rope: bool = True
    use_rmsnorm: bool = True
    wo_shift: bool = False

    @staticmethod
    def named(name: str, **kw) -> "DiTSpec":
        d, h, p, nh = _DIT_REGISTRY[name]
        return DiTSpec(depth=d, hidden_size=h, patch_size=p, num_heads=nh, **kw)

    @property
    def grid(self) -> int:
        return self.input_size // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def out_channels(self) -> int:
        return self.in_channels * (2 if self.learn_sigma else 1)

    @property
    def mlp_hidden(self) -> int:  # lightningdit.py:213-217
        h = int(self.hidden_size * self.mlp_ratio)
        return int(2 / 3 * h) if self.use_swiglu else h


@dataclass
class VMAESpec:
    """mae_for_ldmae_f8d16_prev (tokenizer/models_mae.py:992-997) with the kwargs of inference.py:133."""
    img_size: int = 256
    patch_size: int = 8
    in_chans: int = 3
    embed_dim: int = 192
    depth: int = 12
    num_heads: int = 12
    decoder_embed_dim: int = 192
    decoder_depth: int = 12
    decoder_num_heads: int = 12
    mlp_ratio: float = 4.0
    latent_dim: int = 16
    ln_eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size


# ----------------------------------------------------------------------------- constant tables
def sincos_1d(embed_dim: int, pos: np.ndarray, dtype) -> np.ndarray:
    """models/lightningdit.py:473-491 (float64 omega) / tokenizer/util/pos_embed.py:48-67 (float32)."""
    omega = np.arange(embed_dim // 2, dtype=dtype)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, grid_size: int, dtype=np.float64) -> Tensor:
    """models/lightningdit.py:444-470: meshgrid with w first; first half of the channels encodes
    grid[0] (= the w coordinate), second half grid[1] (= h).  DiT uses float64 omega
    (lightningdit.py:480), VMAE float32 (tokenizer/util/pos_embed.py:56)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)
    emb = np.concatenate([sincos_1d(embed_dim // 2, grid[0], dtype),
                          sincos_1d(embed_dim // 2, grid[1], dtype)], axis=1)
    return torch.from_numpy(emb).float()


def rope_tables(half_head_dim: int, pt_seq_len: int):
    """models/pos_embed.py:96-133 with the call at lightningdit.py:317-323 (freqs_for='lang',
    theta 1e4, ft_seq_len = pt_seq_len).  Returns (cos, sin) of shape [S*S, 2*half_head_dim]:
    first half of the columns rotates with the row index, second half with the column index; each
    frequency is repeated for the adjacent pair (2i, 2i+1)."""
    dim = half_head_dim
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(pt_seq_len) / pt_seq_len * pt_seq_len
    f = t[:, None] * freqs[None, :]                  # [S, dim/2]
    f = f.repeat_interleave(2, dim=-1)               # [S, dim]   '... n -> ... (n r)', r=2
    S = pt_seq_len
    full = torch.cat([f[:, None, :].expand(S, S, dim), f[None, :, :].expand(S, S, dim)], dim=-1)
    full = full.reshape(S * S, 2 * dim)
    return full.cos(), full.sin()


# ----------------------------------------------------------------------------- parameter shapes + synthetic weights
def dit_param_shapes(s: DiTSpec) -> Dict[str, tuple]:
    """state_dict keys/shapes of LightningDiT (probe-listed in SURVEY.md section 8b)."""
    D, p, C = s.hidden_size, s.patch_size, s.in_channels
    T, hd, H = s.tokens, s.head_dim, s.mlp_hidden
    sh: Dict[str, tuple] = {
        "pos_embed": (1, T, D),
        "x_embedder.proj.weight": (D, C, p, p), "x_embedder.proj.bias": (D,),
        "t_embedder.mlp.0.weight": (D, 256), "t_embedder.mlp.0.bias": (D,),
        "t_embedder.mlp.2.weight": (D, D), "t_embedder.mlp.2.bias": (D,),
        "y_embedder.embedding_table.weight": (s.num_classes + (1 if s.class_dropout_prob > 0 else 0), D),
    }
    if s.use_rope:
        sh["feat_rope.freqs_cos"] = (T, hd)
        sh["feat_rope.freqs_sin"] = (T, hd)
    nmod = 4 if s.wo_shift else 6
    for i in range(s.depth):
        b = f"blocks.{i}."
        if s.use_rmsnorm:
            sh[b + "norm1.weight"] = (D,)
            sh[b + "norm2.weight"] = (D,)
        sh[b + "attn.qkv.weight"] = (3 * D, D); sh[b + "attn.qkv.bias"] = (3 * D,)
        if s.use_qknorm:
            sh[b + "attn.q_norm.weight"] = (hd,)
            sh[b + "attn.k_norm.weight"] = (hd,)
            if not s.use_rmsnorm:
                sh[b + "attn.q_norm.bias"] = (hd,)
                sh[b + "attn.k_norm.bias"] = (hd,)
        sh[b + "attn.proj.weight"] = (D, D); sh[b + "attn.proj.bias"] = (D,)
        if s.use_swiglu:
            sh[b + "mlp.w12.weight"] = (2 * H, D); sh[b + "mlp.w12.bias"] = (2 * H,)
            sh[b + "mlp.w3.weight"] = (D, H); sh[b + "mlp.w3.bias"] = (D,)
        else:
            sh[b + "mlp.fc1.weight"] = (H, D); sh[b + "mlp.fc1.bias"] = (H,)
            sh[b + "mlp.fc2.weight"] = (D, H); sh[b + "mlp.fc2.bias"] = (D,)
        sh[b + "adaLN_modulation.1.weight"] = (nmod * D, D); sh[b + "adaLN_modulation.1.bias"] = (nmod * D,)
    if s.use_rmsnorm:
        sh["final_layer.norm_final.weight"] = (D,)
    sh["final_layer.linear.weight"] = (p * p * s.out_channels, D)
    sh["final_layer.linear.bias"] = (p * p * s.out_channels,)
    sh["final_layer.adaLN_modulation.1.weight"] = (2 * D, D)
    sh["final_layer.adaLN_modulation.1.bias"] = (2 * D,)
    return sh


def vmae_decoder_param_shapes(s: VMAESpec) -> Dict[str, tuple]:
    """Decoder-side keys of MaskedAutoencoderViT (tokenizer/models_mae.py:306,357-390)."""
    D, E, L = s.decoder_embed_dim, s.embed_dim, s.grid * s.grid
    Hm = int(D * s.mlp_ratio)
    pp = s.patch_size ** 2 * s.in_chans
    sh: Dict[str, tuple] = {
        "from_latent.weight": (D, s.latent_dim), "from_latent.bias": (D,),
        "decoder_embed.weight": (D, E), "decoder_embed.bias": (D,),
        "decoder_pos_embed": (1, L, D),
    }
    for i in range(s.decoder_depth):
        b = f"decoder_blocks.{i}."
        sh[b + "norm1.weight"] = (D,); sh[b + "norm1.bias"] = (D,)
        sh[b + "attn.qkv.weight"] = (3 * D, D); sh[b + "attn.qkv.bias"] = (3 * D,)
        sh[b + "attn.proj.weight"] = (D, D); sh[b + "attn.proj.bias"] = (D,)
        sh[b + "norm2.weight"] = (D,); sh[b + "norm2.bias"] = (D,)
        sh[b + "mlp.fc1.weight"] = (Hm, D); sh[b + "mlp.fc1.bias"] = (Hm,)
        sh[b + "mlp.fc2.weight"] = (D, Hm); sh[b + "mlp.fc2.bias"] = (D,)
    sh["decoder_norm.weight"] = (D,); sh["decoder_norm.bias"] = (D,)
    sh["decoder_pred.linear_pred.weight"] = (pp, D); sh["decoder_pred.linear_pred.bias"] = (pp,)
    sh["decoder_pred.conv_smoother.weight"] = (s.in_chans, s.in_chans, 3, 3)
    sh["decoder_pred.conv_smoother.bias"] = (s.in_chans,)
    return sh


def vmae_encoder_param_shapes(s: VMAESpec) -> Dict[str, tuple]:
    """Encoder-side keys (tokenizer/models_mae.py:330-352,305) for kl_loss_weight != None."""
    E, L = s.embed_dim, s.grid * s.grid
    Hm = int(E * s.mlp_ratio)
    sh: Dict[str, tuple] = {
        "pos_embed": (1, L, E),
        "patch_embed.proj.weight": (E, s.in_chans, s.patch_size, s.patch_size), "patch_embed.proj.bias": (E,),
    }
    for i in range(s.depth):
        b = f"blocks.{i}."
        sh[b + "norm1.weight"] = (E,); sh[b + "norm1.bias"] = (E,)
        sh[b + "attn.qkv.weight"] = (3 * E, E); sh[b + "attn.qkv.bias"] = (3 * E,)
        sh[b + "attn.proj.weight"] = (E, E); sh[b + "attn.proj.bias"] = (E,)
        sh[b + "norm2.weight"] = (E,); sh[b + "norm2.bias"] = (E,)
        sh[b + "mlp.fc1.weight"] = (Hm, E); sh[b + "mlp.fc1.bias"] = (Hm,)
        sh[b + "mlp.fc2.weight"] = (E, Hm); sh[b + "mlp.fc2.bias"] = (E,)
    sh["norm.weight"] = (E,); sh["norm.bias"] = (E,)
    sh["to_latent.weight"] = (2 * s.latent_dim, E); sh["to_latent.bias"] = (2 * s.latent_dim,)
    return sh


def synth_state(shapes: Dict[str, tuple], seed: int, fixed: Optional[SD] = None) -> SD:
    """Deterministic synthetic weights (CPU generator, key-sorted order) used by the golden
    fixtures, the parity tests, smoke() and bench.py: every tensor is non-trivial (the reference's
    own init leaves final_layer.linear and every adaLN_modulation[-1] at zero --
    lightningdit.py:365-374 -- which would make parity vacuous).  ``fixed`` supplies the constant
    tables (pos_embed, RoPE)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    fixed = fixed or {}
    sd: SD = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k in fixed:
            assert tuple(fixed[k].shape) == tuple(shp), (k, fixed[k].shape, shp)
            sd[k] = fixed[k].clone().float()
            continue
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if len(shp) == 1:
            if k.endswith("weight"):      # norm gains
                sd[k] = 1.0 + 0.1 * r
            else:
                sd[k] = 0.02 * r
        elif "embedding_table" in k:
            sd[k] = 0.02 * r
        else:
            fan_out = shp[0]
            fan_in = int(np.prod(shp[1:]))
            sd[k] = r * math.sqrt(2.0 / (fan_in + fan_out))
    return sd


def synth_dit_state(s: DiTSpec, seed: int) -> SD:
    fixed = {"pos_embed": sincos_2d(s.hidden_size, s.grid, np.float64)[None]}
    if s.use_rope:
        c, sn = rope_tables(s.head_dim // 2, s.grid)
        fixed["feat_rope.freqs_cos"], fixed["feat_rope.freqs_sin"] = c, sn
    return synth_state(dit_param_shapes(s), seed, fixed)


def synth_vmae_state(s: VMAESpec, seed: int, encoder: bool = False) -> SD:
    shapes = dict(vmae_decoder_param_shapes(s))
    fixed = {"decoder_pos_embed": sincos_2d(s.decoder_embed_dim, s.grid, np.float32)[None]}
    if encoder:
        shapes.update(vmae_encoder_param_shapes(s))
        fixed["pos_embed"] = sincos_2d(s.embed_dim, s.grid, np.float32)[None]
    return synth_state(shapes, seed, fixed)


def state_checksum(sd: SD) -> float:
    """Order-independent digest used by fixtures to detect RNG drift between torch builds."""
    return float(sum(v.double().abs().sum().item() * (1 + (len(k) % 7)) for k, v in sd.items()))


# ----------------------------------------------------------------------------- DiT pieces
def rmsnorm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    """models/rmsnorm.py:52-77."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * w


def modulate(x: Tensor, shift: Optional[Tensor], scale: Tensor) -> Tensor:
    """models/lightningdit.py:26-30."""
    y = x * (1 + scale[:, None, :])
    return y if shift is None else y + shift[:, None, :]


def rotate_pairs(x: Tensor) -> Tensor:
    """models/pos_embed.py:38-42: (x0, x1) -> (-x1, x0) on adjacent pairs."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    return torch.stack((-x[..., 1], x[..., 0]), dim=-1).flatten(-2)


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """models/pos_embed.py:135."""
    return x * cos + rotate_pairs(x) * sin


def timestep_embedding(t: Tensor, dim: int = 256, max_period: float = 10000.0) -> Tensor:
    """models/lightningdit.py:108-131 (t is NOT scaled by 1000)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def dit_conditioning(sd: SD, s: DiTSpec, t: Tensor, y: Tensor) -> Tensor:
    """c = t_embedder(t) + y_embedder(y) (lightningdit.py:133-137,163-169,403-405); label dropout
    must already be applied to ``y`` by the caller (lightningdit.py:152-161)."""
    e = timestep_embedding(t)
    h = F.silu(F.linear(e, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"]))
    temb = F.linear(h, sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    return temb + sd["y_embedder.embedding_table.weight"][y.long()]


def patch_embed(sd: SD, s: DiTSpec, x: Tensor) -> Tensor:
    """timm PatchEmbed (Conv2d k=stride=p, flatten(2).transpose(1,2)) + pos_embed
    (lightningdit.py:309,402): token index = h*grid + w."""
    p = s.patch_size
    tok = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    return tok.flatten(2).transpose(1, 2) + sd["pos_embed"]


def dit_attention(sd: SD, s: DiTSpec, pre: str, x: Tensor) -> Tensor:
    """models/lightningdit.py:66-91."""
    B, N, C = x.shape
    nh, hd = s.num_heads, s.head_dim
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if s.use_qknorm:
        if s.use_rmsnorm:
            q, k = rmsnorm(q, sd[pre + "q_norm.weight"]), rmsnorm(k, sd[pre + "k_norm.weight"])
        else:
            q = F.layer_norm(q, (hd,), sd[pre + "q_norm.weight"], sd[pre + "q_norm.bias"], 1e-5)
            k = F.layer_norm(k, (hd,), sd[pre + "k_norm.weight"], sd[pre + "k_norm.bias"], 1e-5)
    if s.use_rope:
        cos, sin = sd["feat_rope.freqs_cos"], sd["feat_rope.freqs_sin"]
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    att = torch.softmax((q @ k.transpose(-2, -1)) * (hd ** -0.5), dim=-1)   # == SDPA, no mask, :77
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def dit_mlp(sd: SD, s: DiTSpec, pre: str, x: Tensor) -> Tensor:
    """models/swiglu_ffn.py:31-36 / timm Mlp with tanh-GELU (lightningdit.py:214-224)."""
    if s.use_swiglu:
        x12 = F.linear(x, sd[pre + "w12.weight"], sd[pre + "w12.bias"])
        x1, x2 = x12.chunk(2, dim=-1)
        return F.linear(F.silu(x1) * x2, sd[pre + "w3.weight"], sd[pre + "w3.bias"])
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]), approximate="tanh")
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def _norm(sd: SD, s: DiTSpec, key: str, x: Tensor) -> Tensor:
    if s.use_rmsnorm:
        return rmsnorm(x, sd[key])
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)   # lightningdit.py:196-197,259


def dit_block(sd: SD, s: DiTSpec, i: int, x: Tensor, c: Tensor) -> Tensor:
    """models/lightningdit.py:239-250."""
    b = f"blocks.{i}."
    mod = F.linear(F.silu(c), sd[b + "adaLN_modulation.1.weight"], sd[b + "adaLN_modulation.1.bias"])
    if s.wo_shift:
        sc_a, g_a, sc_m, g_m = mod.chunk(4, dim=1)
        sh_a = sh_m = None
    else:
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
    x = x + g_a[:, None] * dit_attention(sd, s, b + "attn.", modulate(_norm(sd, s, b + "norm1.weight", x), sh_a, sc_a))
    x = x + g_m[:, None] * dit_mlp(sd, s, b + "mlp.", modulate(_norm(sd, s, b + "norm2.weight", x), sh_m, sc_m))
    return x


def unpatchify(x: Tensor, p: int, c: int) -> Tensor:
    """models/lightningdit.py:376-389: [N,T,p*p*c] -> [N,c,h*p,w*p] via 'nhwpqc->nchpwq'."""
    h = w = int(round(x.shape[1] ** 0.5))
    assert h * w == x.shape[1]
    x = x.reshape(x.shape[0], h, w, p, p, c).permute(0, 5, 1, 3, 2, 4)
    return x.reshape(x.shape[0], c, h * p, w * p)


def dit_forward(sd: SD, s: DiTSpec, x: Tensor, t: Tensor, y: Tensor) -> Tensor:
    """LightningDiT.forward in eval mode (models/lightningdit.py:391-418)."""
    h = patch_embed(sd, s, x)
    c = dit_conditioning(sd, s, t, y)
    for i in range(s.depth):
        h = dit_block(sd, s, i, h, c)
    sh, sc = F.linear(F.silu(c), sd["final_layer.adaLN_modulation.1.weight"],
                      sd["final_layer.adaLN_modulation.1.bias"]).chunk(2, dim=1)      # :267-272
    h = modulate(_norm(sd, s, "final_layer.norm_final.weight", h), sh, sc)
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    out = unpatchify(h, s.patch_size, s.out_channels)
    if s.learn_sigma:
        out = out.chunk(2, dim=1)[0]
    return out


def cfg_combine(model_out: Tensor, t0: float, cfg_scale: float, cfg_interval, cfg_interval_start) -> Tensor:
    """The arithmetic after the forward in forward_with_cfg (lightningdit.py:432-442): guidance on
    channels [:3] only, passthrough for the rest; below cfg_interval_start the guided channels are
    the conditional prediction."""
    eps, rest = model_out[:, :3], model_out[:, 3:]
    n = eps.shape[0] // 2
    cond, uncond = eps[:n], eps[n:]
    half = uncond + cfg_scale * (cond - uncond)
    if cfg_interval is True and t0 < cfg_interval_start:
        half = cond
    return torch.cat([torch.cat([half, half], dim=0), rest], dim=1)


def dit_forward_with_cfg(sd: SD, s: DiTSpec, x: Tensor, t: Tensor, y: Tensor, cfg_scale: float,
                         cfg_interval=None, cfg_interval_start=None) -> Tensor:
    """models/lightningdit.py:420-442."""
    half = x[: len(x) // 2]
    out = dit_forward(sd, s, torch.cat([half, half], dim=0), t, y)
    return cfg_combine(out, float(t[0]), cfg_scale, cfg_interval, cfg_interval_start)


# ----------------------------------------------------------------------------- transport: sampler + loss
def ode_time_grid(num_steps: int, timestep_shift: float = 0.0, t0: float = 0.0, t1: float = 1.0) -> Tensor:
    """transport/integrators.py:93-101: linspace(t0,t1,N) then t <- s*t/(1+(s-1)*t), evaluated per
    element on 0-d fp32 tensors exactly as the reference's list comprehension does."""
    t = torch.linspace(t0, t1, num_steps)
    if timestep_shift > 0:
        t = torch.tensor([(timestep_shift * tn) / (1 + (timestep_shift - 1) * tn) for tn in t])
    return t


def fixed_grid_odeint(fn: Callable[[Tensor, Tensor], Tensor], x: Tensor, t: Tensor, method: str = "euler") -> Tensor:
    """torchdiffeq.odeint with a fixed-grid method, as called at transport/integrators.py:118-125
    (restated from torchdiffeq's FixedGridODESolver: the grid is ``t`` itself, dt = t[k+1]-t[k] in
    fp32, every grid state is returned).  'heun' is accepted as an alias of 'heun2'."""
    out = [x]
    for k in range(len(t) - 1):
        ta, tb = t[k], t[k + 1]
        dt = tb - ta
        if method == "euler":
            x = x + dt * fn(ta, x)
        elif method in ("heun", "heun2"):
            k1 = fn(ta, x)
            k2 = fn(ta + dt, x + dt * k1)
            x = x + dt * (0.5 * k1 + 0.5 * k2)
        else:
            raise NotImplementedError(method)
        out.append(x)
    return torch.stack(out, 0)


def sample_ode(model_fn: Callable, x: Tensor, *, sampling_method: str = "euler", num_steps: int = 250,
               timestep_shift: float = 0.0, **model_kwargs) -> Tensor:
    """Sampler.sample_ode(...) -> fn(x, model, **kw) for path_type Linear / prediction velocity
    (transport/transport.py:398-443 with check_interval :84-111 giving (t0,t1)=(0,1);
    integrators.py:107-125; drift = velocity_ode = the model output, transport.py:234-236)."""
    t = ode_time_grid(num_steps, timestep_shift)

    def _fn(tk, xk):
        tv = torch.ones(xk.size(0), device=xk.device) * tk          # integrators.py:111
        out = model_fn(xk, tv, **model_kwargs)
        assert out.shape == xk.shape              # transport.py:247
        return out

    return fixed_grid_odeint(_fn, x, t, sampling_method)


def training_losses(model_fn: Callable, x1: Tensor, t: Tensor, x0: Tensor, **model_kwargs):
    """Transport.training_losses for Linear path + velocity prediction with the random draws
    (t, x0) injected (transport/transport.py:169-196; path.py:114-136):
    xt = t*x1 + (1-t)*x0, ut = x1 - x0, loss = mean_flat((model(xt,t)-ut)^2)."""
    tt = t.view(-1, *([1] * (x1.dim() - 1)))
    xt = tt * x1 + (1 - tt) * x0
    ut = x1 - x0
    pred = model_fn(xt, t, **model_kwargs)
    loss = ((pred - ut) ** 2).mean(dim=list(range(1, x1.dim())))
    return {"loss": loss, "pred": pred, "xt": xt, "ut": ut}


# ----------------------------------------------------------------------------- VMAE
def vit_block(sd: SD, pre: str, x: Tensor, num_heads: int, eps: float) -> Tensor:
    """tokenizer/models_mae.py:176-187 with Attention :130-147 (explicit softmax, scale hd^-0.5)
    and timm Mlp with exact-erf GELU."""
    B, N, C = x.shape
    hd = C // num_heads
    h = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    att = torch.softmax((qkv[0] @ qkv[1].transpose(-2, -1)) * (hd ** -0.5), dim=-1)
    o = (att @ qkv[2]).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    h = F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def vmae_unpatchify(x: Tensor, p: int) -> Tensor:
    """tokenizer/models_mae.py:458-470."""
    h = w = int(round(x.shape[1] ** 0.5))
    x = x.reshape(x.shape[0], h, w, p, p, 3).permute(0, 5, 1, 3, 2, 4)
    return x.reshape(x.shape[0], 3, h * p, w * p)


def vmae_decode(sd: SD, s: VMAESpec, z: Tensor) -> Tensor:
    """MaskedAutoencoderViT.decode (tokenizer/models_mae.py:865-887) with the smooth_output head
    conv_decoder_pred.forward, pred_with_conv=False branch (:271-279)."""
    B = z.shape[0]
    x = z.flatten(2).transpose(1, 2)                                          # 'b c h w -> b (h w) c'
    x = F.linear(x, sd["from_latent.weight"], sd["from_latent.bias"])
    x = F.linear(x, sd["decoder_embed.weight"], sd["decoder_embed.bias"]) + sd["decoder_pos_embed"]
    for i in range(s.decoder_depth):
        x = vit_block(sd, f"decoder_blocks.{i}.", x, s.decoder_num_heads, s.ln_eps)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], s.ln_eps)
    x = F.linear(x, sd["decoder_pred.linear_pred.weight"], sd["decoder_pred.linear_pred.bias"])
    img = vmae_unpatchify(x, s.patch_size)
    img = F.conv2d(img, sd["decoder_pred.conv_smoother.weight"], sd["decoder_pred.conv_smoother.bias"], padding=1)
    # the reference re-patchifies (:276-278) and decode() unpatchifies again (:883): identity.
    return img


def vmae_encode_moments(sd: SD, s: VMAESpec, img: Tensor) -> Tensor:
    """MaskedAutoencoderViT._encode (tokenizer/models_mae.py:819-836): [B,3,H,W] -> moments
    [B, 2*latent, g, g] (mean || logvar)."""
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=s.patch_size)
    x = x.flatten(2).transpose(1, 2) + sd["pos_embed"]
    for i in range(s.depth):
        x = vit_block(sd, f"blocks.{i}.", x, s.num_heads, s.ln_eps)
    x = F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], s.ln_eps)
    x = F.linear(x, sd["to_latent.weight"], sd["to_latent.bias"])
    return x.transpose(1, 2).reshape(x.shape[0], -1, s.grid, s.grid)


def images_to_uint8(img: Tensor) -> np.ndarray:
    """decode_to_images tail (tokenizer/models_mae.py:972): clamp(127.5*x+128, 0, 255) -> NHWC ->
    truncating cast to uint8."""
    return torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8).numpy()


def denormalize_latents(x: Tensor, mean: Tensor, std: Tensor, multiplier: float) -> Tensor:
    """inference.py:291."""
    return (x * std) / multiplier + mean


# ----------------------------------------------------------------------------- whole job (BASELINE config 1 shape)
def sample_images(dit_sd: SD, ds: DiTSpec, vae_sd: SD, vs: VMAESpec, z: Tensor, y: Tensor, *,
                  num_steps: int, cfg_scale: float, cfg_interval_start: float, timestep_shift: float,
                  method: str = "euler", latent_mean=None, latent_std=None, latent_multiplier: float = 1.0):
    """inference.py:264-292 for one batch: CFG doubling with the null class (= num_classes),
    ODE, keep the first half, de-normalise, decode, uint8."""
    n = z.shape[0]
    if cfg_scale > 1.0:
        zz = torch.cat([z, z], 0)
        yy = torch.cat([y, torch.full((n,), ds.num_classes, dtype=y.dtype, device=y.device)], 0)
        fn = lambda x, t, **kw: dit_forward_with_cfg(dit_sd, ds, x, t, **kw)
        kw = dict(y=yy, cfg_scale=cfg_scale, cfg_interval=True, cfg_interval_start=cfg_interval_start)
    else:
        zz, fn, kw = z, (lambda x, t, **k: dit_forward(dit_sd, ds, x, t, **k)), dict(y=y)
    lat = sample_ode(fn, zz, sampling_method=method, num_steps=num_steps, timestep_shift=timestep_shift, **kw)[-1]
    if cfg_scale > 1.0:
        lat = lat.chunk(2, dim=0)[0]
    mean = torch.zeros(1, ds.in_channels, 1, 1) if latent_mean is None else latent_mean
    std = torch.ones(1, ds.in_channels, 1, 1) if latent_std is None else latent_std
    img = vmae_decode(vae_sd, vs, denormalize_latents(lat, mean, std, latent_multiplier))
    return lat, img, images_to_uint8(img)
