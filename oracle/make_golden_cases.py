"""TEST INFRASTRUCTURE ONLY -- case tables / helpers shared by oracle/make_golden.py (which needs /root/reference) and the tests."""
import numpy as np
import torch

# (sampling_method, diffusion_form, diffusion_norm, last_step, last_step_size) of the SDE sampler goldens (tests/golden/sde_tiny.npz)
SDE_CASES = (("Euler", "sigma", 1.0, "Mean", 0.04), ("Heun", "sigma", 1.0, "Tweedie", 0.04), ("Euler", "linear", 0.5, "Euler", 0.1),
             ("Heun", "inccreasing-decreasing", 0.7, "Mean", 0.04), ("Euler", "decreasing", 1.0, "Mean", 0.04))


def state_fingerprint(sd):
    """Per-key (sum, sum of |.|) in float64 -- a few hundred numbers instead of megabytes of weights."""
    keys = sorted(k for k, v in sd.items() if torch.is_floating_point(v))
    return keys, np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])
