"""TEST INFRASTRUCTURE ONLY -- case tables shared by oracle/make_golden.py (which needs /root/reference) and the tests."""

# (sampling_method, diffusion_form, diffusion_norm, last_step, last_step_size) of the SDE sampler goldens (tests/golden/sde_tiny.npz)
SDE_CASES = (("Euler", "sigma", 1.0, "Mean", 0.04), ("Heun", "sigma", 1.0, "Tweedie", 0.04), ("Euler", "linear", 0.5, "Euler", 0.1),
             ("Heun", "inccreasing-decreasing", 0.7, "Mean", 0.04), ("Euler", "decreasing", 1.0, "Mean", 0.04))
