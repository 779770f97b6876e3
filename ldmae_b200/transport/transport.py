"""Flow-matching "transport" for the shipped recipe: Linear path + velocity prediction.

Mirrors the reference's API (``create_transport`` -- transport/__init__.py:3-72; ``Transport.training_losses``
-- transport/transport.py:169-215; ``Sampler.sample_ode`` -- transport.py:398-443 with the fixed-grid
solver the reference delegates to torchdiffeq at integrators.py:118).  When the model callable is the bound
``forward`` / ``forward_with_cfg`` of an ``ldmae_b200`` LightningDiT, the whole ODE loop (model evaluations,
CFG combine, Euler/Heun update) runs inside libldmae_b200.so with no per-step host sync; any other callable
runs through the generic fixed-grid loop below.  VP/GVP plans, score/noise prediction, SDE samplers and the
likelihood ODE are outside the hot path and raise NotImplementedError.
"""
from __future__ import annotations

import enum
import os

import numpy as np
import torch as th


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


def mean_flat(x):
    """reference transport/utils.py:12-16."""
    return th.mean(x, dim=list(range(1, len(x.size()))))


def ode_time_grid(num_steps, timestep_shift=0.0, t0=0.0, t1=1.0):
    """reference integrators.py:93-101, evaluated element-wise on 0-d fp32 tensors like the reference."""
    t = th.linspace(t0, t1, num_steps)
    if timestep_shift > 0:
        t = th.tensor([(timestep_shift * tn) / (1 + (timestep_shift - 1) * tn) for tn in t])
    return t


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     use_cosine_loss=None, use_lognorm=None, partitial_train=None, partial_ratio=1.0, shift_lg=False):
    """Same signature and defaults as the reference (transport/__init__.py:3-72)."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    if ptype != PathType.LINEAR or model_type != ModelType.VELOCITY:
        raise NotImplementedError("ldmae_b200 builds the shipped recipe only: path_type='Linear', prediction='velocity' "
                                  "(reference configs/*/lightningdit_b_vmae_f8d16_cfg.yaml:52-59)")
    # velocity & LINEAR is stable everywhere (transport/__init__.py:55-57)
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=0, sample_eps=0,
                     use_cosine_loss=use_cosine_loss, use_lognorm=use_lognorm, partitial_train=partitial_train,
                     partial_ratio=partial_ratio, shift_lg=shift_lg)


class Transport:
    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps, use_cosine_loss=False,
                 use_lognorm=False, partitial_train=None, partial_ratio=1.0, shift_lg=False):
        self.loss_type, self.model_type, self.path_type = loss_type, model_type, path_type
        self.train_eps, self.sample_eps = train_eps, sample_eps
        self.use_cosine_loss, self.use_lognorm = use_cosine_loss, use_lognorm
        self.partitial_train, self.partial_ratio, self.shift_lg = partitial_train, partial_ratio, shift_lg

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """reference transport.py:84-111 for ICPlan + velocity without SDE: always (0, 1)."""
        if sde:
            raise NotImplementedError("SDE sampling is outside the hot path")
        t0, t1 = 0, 1
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1

    def sample_logit_normal(self, mu, sigma, size=1):
        """reference transport.py:113-123 (scipy.stats.norm.rvs == numpy's global normal stream)."""
        samples = np.random.normal(loc=mu, scale=sigma, size=size)
        return th.tensor(1 / (1 + np.exp(-samples)), dtype=th.float32)

    def sample(self, x1, sp_timesteps=None, shifted_mu=0):
        """reference transport.py:136-166."""
        x0 = th.randn_like(x1)
        t0, t1 = self.check_interval(self.train_eps, self.sample_eps)
        if self.partitial_train is not None:
            raise NotImplementedError("partitial_train is not used by any shipped config")
        if not self.use_lognorm:
            t = th.rand((x1.shape[0],)) * (t1 - t0) + t0
        else:
            t = self.sample_logit_normal(shifted_mu if self.shift_lg else 0, 1, size=x1.shape[0]) * (t1 - t0) + t0
        if sp_timesteps is not None:
            t = th.rand((x1.shape[0],)) * (sp_timesteps[1] - sp_timesteps[0]) + sp_timesteps[0]
        return t.to(x1), x0, x1

    def training_losses(self, model, x1, model_kwargs=None, sp_timesteps=None, shifted_mu=0):
        """reference transport.py:169-215 with ICPlan.plan (path.py:114-136): xt = t*x1 + (1-t)*x0, ut = x1 - x0."""
        model_kwargs = model_kwargs or {}
        t, x0, x1 = self.sample(x1, sp_timesteps, shifted_mu)
        tt = t.view(t.size(0), *([1] * (x1.dim() - 1)))
        xt = tt * x1 + (1 - tt) * x0
        ut = x1 - x0
        model_output = model(xt, t, **model_kwargs)
        assert model_output.size() == xt.size()
        terms = {"pred": model_output, "loss": mean_flat((model_output - ut) ** 2)}
        if self.use_cosine_loss:
            terms["cos_loss"] = mean_flat(1 - th.nn.functional.cosine_similarity(model_output, ut, dim=1))
        return terms

    def get_drift(self):
        """reference transport.py:218-250 (velocity_ode)."""
        def body_fn(x, t, model, **model_kwargs):
            out = model(x, t, **model_kwargs)
            assert out.shape == x.shape, "Output shape from ODE solver must match input shape"
            return out
        return body_fn


class _Trajectory:
    """What ``sample_fn(...)`` returns on the fused path: indexable like torchdiffeq's stacked solution
    (``[-1]`` is the final state, ``len`` = number of grid points).  Intermediate states are only
    materialised when the sampler was built with ``keep_trajectory=True`` (or LDMAE_KEEP_TRAJECTORY=1);
    the reference's caller (inference.py:287) reads ``[-1]`` only."""

    def __init__(self, final, traj, npts):
        self._final, self._traj, self._n = final, traj, npts

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if self._traj is not None:
            return self._traj[i]
        if isinstance(i, int) and (i == -1 or i == self._n - 1):
            return self._final
        raise IndexError("intermediate ODE states were not kept (build the sampler with keep_trajectory=True)")


_METHODS = {"euler": 0, "heun2": 1, "heun": 1}


def _fixed_grid_odeint(fn, x, t, method):
    """Generic fallback for arbitrary callables: torchdiffeq's fixed-grid Euler / Heun on the grid ``t``."""
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
            raise NotImplementedError(f"ODE method {method!r}: fixed-grid 'euler' and 'heun2' are built")
        out.append(x)
    return th.stack(out, 0)


class Sampler:
    """reference transport.py:270-283."""

    def __init__(self, transport):
        self.transport = transport
        self.drift = transport.get_drift()

    def sample_sde(self, **_):
        raise NotImplementedError("SDE sampling is outside the hot path (no shipped config uses it)")

    def sample_ode_likelihood(self, **_):
        raise NotImplementedError("likelihood ODE is outside the hot path")

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False,
                   timestep_shift=0.0, keep_trajectory=None, cond_only_when_unguided=False):
        """Returns ``fn(x, model, **model_kwargs)`` like the reference (transport.py:398-443).

        ``cond_only_when_unguided`` (extension, off by default): with ``forward_with_cfg`` and a guidance interval, steps with
        ``t < cfg_interval_start`` evaluate only the conditional half -- its guided velocity is its own prediction
        (lightningdit.py:436-439).  The first half of the returned state is identical; the second half is not advanced, so
        only callers that keep ``chunk(2)[0]`` (inference.py:289) may use it."""
        if reverse:
            raise NotImplementedError("reverse-time ODE is outside the hot path")
        if sampling_method not in _METHODS:
            raise NotImplementedError(f"sampling_method {sampling_method!r}: fixed-grid 'euler' and 'heun2' are built "
                                      "(adaptive dopri5 is a next-round item)")
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False,
                                               eval=True, reverse=reverse, last_step_size=0.0)
        tgrid = ode_time_grid(num_steps, timestep_shift, t0, t1)
        if keep_trajectory is None:
            keep_trajectory = os.environ.get("LDMAE_KEEP_TRAJECTORY", "0") == "1"
        drift = self.drift

        def _sample(x, model, **model_kwargs):
            from ..models.lightningdit import LightningDiT
            owner = getattr(model, "__self__", None)
            name = getattr(model, "__name__", "")
            if isinstance(owner, LightningDiT) and name in ("forward", "forward_with_cfg") and x.is_cuda \
                    and (not owner.training):
                grid = [float(v) for v in tgrid]
                if name == "forward_with_cfg":
                    y = model_kwargs["y"]
                    interval = model_kwargs.get("cfg_interval", None)
                    start = model_kwargs.get("cfg_interval_start", None)
                    start = float(start) if (interval is True and start is not None) else -1.0
                    final, traj = owner._sample_ode(x, y, x.shape[0] // 2, True, model_kwargs["cfg_scale"], start, grid,
                                                    _METHODS[sampling_method], keep_trajectory,
                                                    flags=1 if (cond_only_when_unguided and start >= 0) else 0)
                else:
                    final, traj = owner._sample_ode(x, model_kwargs["y"], x.shape[0], False, 1.0, -1.0, grid,
                                                    _METHODS[sampling_method], keep_trajectory)
                return _Trajectory(final, traj, len(grid))
            device = x.device

            def _fn(t, xx):                                   # reference integrators.py:110-113
                tv = th.ones(xx.size(0)).to(device) * t
                return drift(xx, tv, model, **model_kwargs)

            return _fixed_grid_odeint(_fn, x, tgrid.to(device), sampling_method)

        _sample.t = tgrid
        return _sample
