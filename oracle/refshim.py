"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Lets the reference's hot-path files (``/root/reference/LDMAE/models/lightningdit.py``,
``transport/*``, ``tokenizer/models_mae.py``) import verbatim in this container, where
``timm``, ``fairscale``, ``torchdiffeq``, ``diffusers`` and ``taming`` are not installed.
It is used by ``oracle/make_golden.py`` (here, on CPU) to produce the committed fixtures
under ``tests/golden/`` and by ``bench.py``'s reference / cpu_baseline legs.  ``/root/reference`` is absent on
the GPU box; ``oracle/build_ref.py`` stages the files of the path under ``oracle/_ref/`` (git-ignored, travels with
the snapshot) and the shim falls back to that copy.  Nothing in the product imports it.

What is restated (third-party arithmetic that is not under /root/reference):

* ``timm==1.0.12`` (reference ``requirements.txt:8``): ``PatchEmbed`` = ``Conv2d(k=stride=patch)``
  -> ``flatten(2).transpose(1, 2)``; ``Mlp`` = ``fc1 -> act -> fc2``; ``DropPath`` identity in eval.
* ``torchdiffeq`` (un-pinned, reference ``requirements.txt:11``): the fixed-grid solvers used by
  ``transport/integrators.py:118``: ``euler`` (y += dt*f(t0,y)) and ``heun2``
  (y += dt/2*(k1 + f(t1, y + dt*k1))), ``midpoint`` and ``rk4`` (3/8 rule, as torchdiffeq).  The
  grid is exactly the ``t`` passed in; every grid state is returned, stacked on dim 0.
  Parity at this boundary is UNPINNED by the reference (no version pin, no tests).
* ``fairscale``, ``taming``, ``diffusers``: name-only stubs (never executed on the path).
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "LDMAE")     # oracle/build_ref.py


def _default_root():
    """The reference tree itself where it exists (authoring container), else the copy staged by oracle/build_ref.py
    (git-ignored; travels to the GPU box)."""
    if os.path.isdir("/root/reference/LDMAE/models"):
        return "/root/reference/LDMAE"
    return _STAGED


REF_ROOT = os.environ.get("LDMAE_REFERENCE_ROOT") or _default_root()


# --------------------------------------------------------------------------- timm
class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, bias=True, **_):
        super().__init__()
        to2 = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.img_size = to2(img_size)
        self.patch_size = to2(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not (self.training and self.drop_prob > 0), "shim: DropPath only as identity"
        return x


# --------------------------------------------------------------------------- torchdiffeq
def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, **_):
    """Fixed-grid restatement (see module docstring)."""
    assert torch.is_tensor(y0), "shim: tensor state only"
    method = method or "dopri5"
    sol = [y0]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        dt = t1 - t0
        if method == "euler":
            dy = dt * func(t0, y)
        elif method in ("heun2", "heun"):
            k1 = func(t0, y)
            k2 = func(t0 + dt, y + dt * k1)
            dy = dt * (0.5 * k1 + 0.5 * k2)
        elif method == "midpoint":
            half = 0.5 * dt
            dy = dt * func(t0 + half, y + half * func(t0, y))
        elif method == "rk4":
            k1 = func(t0, y)
            k2 = func(t0 + dt / 3, y + dt * k1 / 3)
            k3 = func(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = func(t1, y + dt * (k1 - k2 + k3))
            dy = (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise NotImplementedError(f"shim odeint: method {method!r}")
        y = y + dy
        sol.append(y)
    return torch.stack(sol, 0)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Register the stubs and put the reference on sys.path. Idempotent."""
    if getattr(install, "_done", False):
        return
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")  # @torch.compile decorators -> eager
    _mod("timm"); _mod("timm.models")
    _mod("timm.models.vision_transformer", PatchEmbed=PatchEmbed, Mlp=Mlp, DropPath=DropPath)
    _mod("torchdiffeq", odeint=odeint)

    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    fs = _mod("fairscale"); _mod("fairscale.nn"); _mod("fairscale.nn.model_parallel")
    _mod("fairscale.nn.model_parallel.initialize", get_model_parallel_world_size=lambda: 1,
         get_model_parallel_rank=lambda: 0, initialize_model_parallel=lambda *a, **k: None,
         model_parallel_is_initialized=lambda: False)
    _mod("fairscale.nn.model_parallel.layers", ColumnParallelLinear=_Stub, ParallelEmbedding=_Stub,
         RowParallelLinear=_Stub, VocabParallelEmbedding=_Stub)
    del fs

    class BaseOutput(dict):
        """dataclass-style outputs only need attribute access on this path."""

    _mod("diffusers", ConfigMixin=object, ModelMixin=object)
    _mod("diffusers.utils", BaseOutput=BaseOutput)
    _mod("taming"); _mod("taming.modules"); _mod("taming.modules.losses")
    _mod("taming.modules.losses.lpips", LPIPS=_Stub)
    try:
        import torchvision  # noqa: F401
    except Exception:  # pragma: no cover
        tv = _mod("torchvision")
        tv.transforms = _mod("torchvision.transforms")
    for p in (REF_ROOT, os.path.join(REF_ROOT, "tokenizer")):
        if p not in sys.path:
            sys.path.insert(0, p)
    install._done = True


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))
