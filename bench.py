#!/usr/bin/env python
"""Headline benchmark: LightningDiT-B/1 ImageNet-256 class-conditional sampling (BASELINE.json configs[1]).

One step = one sampling job of `--batch` images per GPU: 250-point shifted Euler grid (249 model evaluations, each on
the CFG-doubled batch), cfg_scale 10 with cfg_interval_start 0.10, then latent de-normalisation + VMAE decode to uint8
(the body of the reference's LDMAE/inference.py:264-292 loop).  Synthetic latents / labels, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo (sm_100a kernels)
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, weak scaling (batch/GPU fixed)
    python bench.py --impl reference ...                                # the reference's own CPU path on the host cores
    python bench.py --impl torch-gpu ...                                # the reference algorithm in PyTorch on the same GPU
                                                                        # (comparator; never imports ldmae_b200)

Prints ONE JSON line (rank 0).  Every step is ONE call of ldmae_b200.pipeline.SamplingJob.run_host (pinned host latents
and labels in, uint8 images back on the host) with four CUDA events on the launching stream: `e2e` is the outer pair
(host-to-device copy ... device-to-host copy), `value` the inner pair (inputs resident in HBM ... images decoded) of the
SAME K steps.  A wall-clock budget (`--time-budget`, default 810 s from process start: the driver's scaling harness allows
870 s per N) first drops the optional extras (training step, decoder, XL, cond-only objects), then warm-up steps beyond 3,
then timed steps -- whatever it dropped is listed under `budget` in the line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
import zlib

T_START = time.time()
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LightningDiT-B sampled img/s"
UNIT = "img/s"


def elapsed():
    return time.time() - T_START


# ----------------------------------------------------------------------------- algorithmic work
def dit_flops_per_sample_forward(depth=12, D=768, T=1024, H=2048, C=16):
    per_block = {
        "qkv_gemm": 2 * T * D * 3 * D,
        "attention": 4 * T * T * D,
        "proj_gemm": 2 * T * D * D,
        "w12_swiglu_gemm": 2 * T * D * 2 * H,
        "w3_gemm": 2 * T * H * D,
    }
    adaln = 2 * D * 6 * D
    total = depth * (sum(per_block.values()) + adaln) + 2 * T * C * D * 2 + 2 * D * 2 * D
    return per_block, total


def vmae_decode_flops_per_image(L=1024, D=192, depth=12, heads=12, Hm=768, E=192, PP=192, latent=16):
    """20.70 GFLOP / image (GEMMs + attention matmuls; tokenizer/models_mae.py:865-887)."""
    blk = L * (2 * D * 3 * D + 2 * D * D + 2 * D * Hm + 2 * Hm * D) + 4 * L * L * D
    return depth * blk + L * (2 * latent * E + 2 * E * D + 2 * D * PP) + L * 64 * 3 * 27 * 2


MODEL_GEOMETRY = {   # registry lightningdit.py:498-531: (depth, width, heads, patch)
    "LightningDiT-B/1": (12, 768, 12, 1), "LightningDiT-L/2": (24, 1024, 16, 2), "LightningDiT-XL/1": (28, 1152, 16, 1),
    "LightningDiT-XL/2": (28, 1152, 16, 2), "LightningDiT-1p0B/1": (24, 1536, 24, 1), "LightningDiT-1p6B/1": (28, 1792, 28, 1),
}


def model_flops(model, input_size):
    depth, D, _, patch = MODEL_GEOMETRY[model]
    T = (input_size // patch) ** 2
    H = int(2 / 3 * 4 * D)
    return dit_flops_per_sample_forward(depth=depth, D=D, T=T, H=H, C=16 * patch * patch)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "tflops_burst": float(d.get("bf16_tflops", 0.0)),
                "hbm_gbs": float(d["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md: 1.59 PF burst / ~1.4 PF sustained)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- the reference algorithm in plain PyTorch
def reference_path(device, model_name="LightningDiT-B/1", input_size=32, img_size=256):
    """The reference's path as callables (dit(x, t, **kw) for forward_with_cfg, decode(lat) -> uint8 images, train_objs).
    kind "reference": the reference's OWN modules (LDMAE/models/lightningdit.py, transport/*, tokenizer/models_mae.py)
    imported from the copy oracle/build_ref.py staged under oracle/_ref/ (or /root/reference where it exists) through
    oracle/refshim.py; kind "port": the oracle restatement, when no staged copy travelled.  bench.py's reference /
    cpu_baseline / comparator legs are the places allowed to execute oracle/ (never the product path)."""
    import torch
    from oracle import ldmae_oracle as O
    from oracle import refshim
    ds = O.DiTSpec.named(model_name, input_size=input_size, in_channels=16)
    vs = O.VMAESpec(img_size=img_size)
    dsd, vsd = O.synth_dit_state(ds, 0), O.synth_vmae_state(vs, 1, encoder=True)
    if refshim.reference_available():
        refshim.install()
        from models.lightningdit import LightningDiT_models            # reference
        from transport import Sampler, create_transport               # reference
        import tokenizer.models_mae as ref_mae                        # reference
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            dit = LightningDiT_models[model_name](input_size=input_size, in_channels=16, use_qknorm=True, use_swiglu=True,
                                                  use_rope=True, use_rmsnorm=True)
            vae = ref_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True,
                                                   img_size=img_size)
        dit.load_state_dict(dsd, strict=True)
        vae.load_state_dict(vsd, strict=False)
        dit, vae = dit.to(device).eval(), vae.to(device).eval()
        transport = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)

        def sample(zz, num_points, **kw):
            fn = Sampler(transport).sample_ode(sampling_method="euler", num_steps=num_points, atol=1e-6, rtol=1e-3,
                                               reverse=False, timestep_shift=0.3)
            return fn(zz, dit.forward_with_cfg, **kw)[-1]

        def decode(lat):
            img = vae.decode(lat, return_dict=False)[0]
            return torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8)      # models_mae.py:972
        return {"kind": "reference", "sample": sample, "decode": decode, "dit": dit, "vae": vae, "transport": transport,
                "null_class": 1000}
    dsd = {k: v.to(device) for k, v in dsd.items()}
    vsd = {k: v.to(device) for k, v in vsd.items()}

    def sample(zz, num_points, **kw):
        fn = lambda x, t, **k: O.dit_forward_with_cfg(dsd, ds, x, t, **k)
        return O.sample_ode(fn, zz, sampling_method="euler", num_steps=num_points, timestep_shift=0.3, **kw)[-1]

    def decode(lat):
        img = O.vmae_decode(vsd, vs, lat)
        return torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    return {"kind": "port", "sample": sample, "decode": decode, "dit": None, "vae": None, "transport": None,
            "null_class": ds.num_classes}


_CPU_PATH = None


def cpu_job(n, num_points, threads):
    """Times the reference's CPU path on a bounded sample of the workload: n images, CFG, `num_points`-point shifted Euler
    grid, then VMAE decode of the n latents.  Returns (seconds per image of the full 249-evaluation job, ODE seconds,
    decode seconds, kind)."""
    import torch
    global _CPU_PATH
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    if _CPU_PATH is None:
        _CPU_PATH = reference_path(torch.device("cpu"))
    P = _CPU_PATH
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n, 16, 32, 32, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    zz = torch.cat([z, z], 0)
    yy = torch.cat([y, torch.full((n,), P["null_class"], dtype=y.dtype)], 0)
    t0 = time.perf_counter()
    lat = P["sample"](zz, num_points, y=yy, cfg_scale=10.0, cfg_interval=True, cfg_interval_start=0.10).chunk(2, dim=0)[0]
    t1 = time.perf_counter()
    P["decode"](lat)
    t2 = time.perf_counter()
    evals = num_points - 1
    per_img = (t1 - t0) / (n * evals) * 249 + (t2 - t1) / n
    return per_img, (t1 - t0), (t2 - t1), P["kind"]


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (all of them), each step a
    bounded sample of the workload; `value` scales the sample's per-image-evaluation time to the 249-evaluation job and
    `ms_per_step` is the wall time one sample step really took."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, pts = args.cpu_images, args.cpu_points
    for _ in range(args.warmup):
        cpu_job(1, 2, threads)                         # warm-up: one evaluation of one image (page in weights / threads)
    per, wall, kind = [], [], "port"
    for _ in range(args.steps):
        t0 = time.perf_counter()
        per_img, _, _, kind = cpu_job(n, pts, threads)
        wall.append(time.perf_counter() - t0)
        per.append(per_img)
    per_img = sum(per) / len(per)
    val = 1.0 / per_img
    sample = (f"{n} images x {pts - 1} CFG evaluations of LightningDiT-B/1 (fp32, {threads} threads) + VMAE decode of {n} images per "
              f"step; per-image time scaled linearly to 249 evaluations")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sum(wall) / len(wall) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "sample_images_per_step": n, "sample_evaluations_per_step": pts - 1,
            "ms_per_full_step_scaled": per_img * args.batch * 1e3,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"{args.model} ImageNet-{8 * args.input_size} class-conditional sampling: {args.num_steps}-point Euler ODE "
                        f"({args.num_steps - 1} evaluations on the CFG-doubled batch), cfg_scale 10, cfg_interval_start 0.10, "
                        f"timestep_shift 0.3, VMAE f8d16 decode to uint8; batch {args.batch}/GPU",
            "batch_per_gpu": args.batch, "num_steps": args.num_steps, "latent": f"{args.input_size}x{args.input_size}x16",
            "image": f"{8 * args.input_size}x{8 * args.input_size}x3",
            "parallelism": f"batch-sharded x{args.gpus} (no collective in the loop)",
            "l2": "inputs larger than L2: each evaluation streams >6 GB of activations (L2 = 126 MB)"}


# ----------------------------------------------------------------------------- GPU comparator
def run_torch_gpu(args, rank, world):
    """--impl torch-gpu: the reference algorithm in plain PyTorch on the same B200 -- the bar the kernels must beat.
    Sampling in the reference's own numerics (fp32 storage + TF32 matmuls, eager; inference.py:79) and under bf16 autocast;
    the training step under bf16 autocast with torch.optim.AdamW + the EMA loop (train_accum.py:215-246,337-347); each
    eager and (when it compiles inside the time limit) with torch.compile.  Per-evaluation time at steady clocks over
    `--cmp-evals` evaluations of the CFG-doubled batch, scaled to the 249-evaluation job (labelled as such).  Imports
    nothing from ldmae_b200."""
    if rank != 0:
        return
    import torch
    assert "ldmae_b200" not in sys.modules
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = True            # inference.py:79-80
    torch.backends.cudnn.allow_tf32 = True
    P = reference_path(dev, args.model, args.input_size, 8 * args.input_size)
    n, S = args.batch, args.input_size
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n, 16, S, S, generator=g).to(dev)
    y = torch.randint(0, 1000, (n,), generator=g).to(dev)
    zz = torch.cat([z, z], 0)
    yy = torch.cat([y, torch.full((n,), P["null_class"], device=dev)], 0)
    kw = dict(y=yy, cfg_scale=10.0, cfg_interval=True, cfg_interval_start=0.10)
    E = args.cmp_evals
    out = {"impl": "torch-gpu", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "kind": P["kind"], "config": workload_config(args),
           "evaluations_timed": E, "note": "img/s = batch / (249 x mean evaluation time + decode time): scaled, not a full job"}

    def time_sampling(tag, ctx, sample_fn):
        try:
            with torch.no_grad(), ctx():
                sample_fn(zz, 4, **kw)                                      # warm-up: 3 evaluations
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                lat = sample_fn(zz, E + 1, **kw).chunk(2, dim=0)[0]
                e1.record()
                P["decode"](lat.float())
                e2.record()
                torch.cuda.synchronize()
            t_eval, t_dec = e0.elapsed_time(e1) / E, e1.elapsed_time(e2)
            job_ms = 249 * t_eval + t_dec
            out[tag] = {"img_per_s": n / (job_ms / 1e3), "ms_per_evaluation": t_eval, "decode_ms": t_dec,
                        "job_ms_scaled": job_ms}
        except Exception as ex:                                             # noqa: BLE001 -- comparator legs are best-effort
            out[tag] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        print(f"[torch-gpu] {tag}: {out[tag]}", file=sys.stderr, flush=True)

    import contextlib
    time_sampling("sampling_fp32_tf32_eager", contextlib.nullcontext, P["sample"])
    time_sampling("sampling_bf16_autocast_eager", lambda: torch.autocast("cuda", dtype=torch.bfloat16), P["sample"])
    dit = P["dit"]
    if dit is not None and not args.no_compile and elapsed() < args.time_budget * 0.5:
        os.environ.pop("TORCHDYNAMO_DISABLE", None)
        try:
            import torch._dynamo
            torch._dynamo.config.disable = False
            cdit = torch.compile(dit)
            from transport import Sampler

            def sample_c(zz_, num_points, **k):
                fn = Sampler(P["transport"]).sample_ode(sampling_method="euler", num_steps=num_points, atol=1e-6, rtol=1e-3,
                                                        reverse=False, timestep_shift=0.3)
                return fn(zz_, lambda x, t, **kk: type(dit).forward_with_cfg(cdit, x, t, **kk), **k)[-1]
            time_sampling("sampling_fp32_tf32_compiled", contextlib.nullcontext, sample_c)
            time_sampling("sampling_bf16_autocast_compiled", lambda: torch.autocast("cuda", dtype=torch.bfloat16), sample_c)
        except Exception as ex:                                             # noqa: BLE001
            out["compile_error"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    # ---- training step: bf16 autocast, AdamW, EMA (train_accum.py) ------------------------------------------------
    if dit is not None and not args.no_train:
        import copy
        B = args.train_batch
        dit.train()
        for p_ in dit.parameters():
            p_.requires_grad_(True)
        dit.pos_embed.requires_grad_(False)
        ema = copy.deepcopy(dit).requires_grad_(False)
        opt = torch.optim.AdamW(dit.parameters(), lr=2e-4, weight_decay=0.0, betas=(0.9, 0.95))
        x1 = torch.randn(B, 16, S, S, generator=g).to(dev)
        yl = torch.randint(0, 1000, (B,), generator=g).to(dev)

        def train_step(model):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = P["transport"].training_losses(model, x1, dict(y=yl))["loss"].mean()
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            with torch.no_grad():                                           # update_ema, train_accum.py:337-347
                for pe, pm in zip(ema.parameters(), dit.parameters()):
                    pe.mul_(0.9999).add_(pm.detach(), alpha=1 - 0.9999)
            return loss

        def time_train(tag, model):
            try:
                for _ in range(3):
                    train_step(model)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.train_steps):
                    loss = train_step(model)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.train_steps
                out[tag] = {"samples_per_s": B / (ms / 1e3), "ms_per_step": ms, "batch": B, "final_loss": float(loss)}
            except Exception as ex:                                         # noqa: BLE001
                out[tag] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            print(f"[torch-gpu] {tag}: {out[tag]}", file=sys.stderr, flush=True)

        time_train("train_bf16_autocast_eager", dit)
        if not args.no_compile and elapsed() < args.time_budget * 0.8:
            try:
                time_train("train_bf16_autocast_compiled", torch.compile(dit))
            except Exception as ex:                                         # noqa: BLE001
                out["train_compile_error"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    best = max((v["img_per_s"] for k, v in out.items() if isinstance(v, dict) and "img_per_s" in v), default=None)
    out["value"] = best
    out["wall_s"] = round(elapsed(), 1)
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- training step (BASELINE.json configs[2])
def measure_train(args, dev, rank, world, model=None):
    """train samples/s: one step = one optimizer step of train_accum.py:203-246 on `--train-batch` synthetic latents per GPU
    (flow-matching loss, forward + backward, gradient all-reduce across ranks, fused AdamW + EMA, weight re-pack).
    Inputs come from pinned host memory every step and the per-sample losses are read back (end-to-end by construction)."""
    import torch
    import torch.distributed as dist
    from ldmae_b200 import _lib
    from ldmae_b200.pipeline import build_sampling_models
    from ldmae_b200.training import FusedTrainer
    B = args.train_batch
    if model is None:
        model, _ = build_sampling_models(dev, seed=0, model_name=args.model, input_size=args.input_size, img_size=8 * args.input_size)
    model.train()
    trainer = FusedTrainer(model, lr=2e-4, betas=(0.9, 0.95), weight_decay=0.0, ema_decay=0.9999)
    g = torch.Generator().manual_seed(100 + rank)
    S = args.input_size
    y_host = torch.randint(0, 1000, (B,), generator=g).pin_memory()
    loss_host = torch.empty(B).pin_memory()
    if args.train_input == "moments":
        # synthetic "extracted features" -- posterior moments of every image and of its flip (extract_features.py:
        # 168-181); coin flip, posterior sample and per-channel normalisation (img_latent_dataset.py:76-94) run on the device
        # inside the kernel that builds xt / ut
        mom_host = (torch.randn(B, 32, S, S, generator=g) * 0.5).pin_memory()
        momf_host = (torch.randn(B, 32, S, S, generator=g) * 0.5).pin_memory()
        lat_mean = (torch.randn(1, 16, 1, 1, generator=g) * 0.1).to(dev)
        lat_std = (1.0 + 0.1 * torch.rand(1, 16, 1, 1, generator=g)).to(dev)
        h2d = (mom_host.numel() + momf_host.numel()) * 4 + y_host.numel() * 8

        def one_step():
            mom = mom_host.to(dev, non_blocking=True)
            momf = momf_host.to(dev, non_blocking=True)
            y = y_host.to(dev, non_blocking=True)
            loss = trainer.step_from_moments(mom, momf, y, latent_mean=lat_mean, latent_std=lat_std, latent_multiplier=1.0)
            loss_host.copy_(loss, non_blocking=True)
    else:
        x_host = torch.randn(B, 16, S, S, generator=g).pin_memory()
        h2d = x_host.numel() * 4 + y_host.numel() * 8

        def one_step():
            x = x_host.to(dev, non_blocking=True)
            y = y_host.to(dev, non_blocking=True)
            loss = trainer.step(x, y)
            loss_host.copy_(loss, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(3):
            one_step()
        barrier()
        launches0 = _lib.launch_count()
        _lib.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.train_steps):
            one_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = _lib.profile_end()
    launches = _lib.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    ms_step = ms / args.train_steps
    _, fwd = model_flops(args.model, args.input_size)
    pk = peaks()
    tfl = 3 * fwd * B / (ms_step / 1e3) / 1e12
    final_loss = float(loss_host.mean())
    del trainer
    torch.cuda.empty_cache()
    return {"metric": ("LightningDiT-B" if args.model == "LightningDiT-B/1" else args.model) + " train samples/s", "value": world * B / (ms_step / 1e3), "unit": "samples/s",
            "batch_per_gpu": B, "global_batch": world * B, "steps": args.train_steps, "ms_per_step": ms_step,
            "tflops_per_gpu": tfl, "frac_of_bf16_peak": tfl / pk["tflops"], "flops_per_sample": 3 * fwd,
            "gpu_launches": launches, "final_loss": final_loss,
            "class_ms_per_step": {k: round(v[0] / args.train_steps, 3) for k, v in prof.items() if v[1] > 0},
            "input": ("posterior moments of each image and of its flip [B,32,S,S] (extract_features.py shard format); flip select, "
                      "posterior sample, normalisation on the device" if args.train_input == "moments" else "ready latents [B,16,S,S]"),
            "includes": "H2D of the stored features/labels, input pipeline, label dropout, transport draws, forward, loss, backward, "
                        "gradient all-reduce (N>1), fused AdamW+EMA, bf16 weight re-pack, D2H of per-sample losses",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 4}


def measure_decode(args, dev, vae, world):
    """BASELINE.json configs[3]: VMAE f8d16 decode of `--batch` synthetic 32x32x16 latents to 256x256x3 uint8, end to end
    (pinned host latents in, uint8 images back on the host)."""
    import torch
    from ldmae_b200 import _lib
    n, S = args.batch, args.input_size
    z_host = torch.randn(n, 16, S, S, generator=torch.Generator().manual_seed(5)).pin_memory()
    out_host = torch.empty(n, 8 * S, 8 * S, 3, dtype=torch.uint8).pin_memory()
    mean = torch.zeros(16, device=dev); std = torch.ones(16, device=dev)

    def step():
        z = z_host.to(dev, non_blocking=True)
        _, u8 = vae._decode(z, False, True, mean, std, 1.0)
        out_host.copy_(u8, non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    K = 10
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    fl = vmae_decode_flops_per_image(L=S * S)
    pk = peaks()
    tfl = fl * n / (ms / 1e3) / 1e12
    return {"metric": "VMAE f8d16 decode img/s", "value": world * n / (ms / 1e3), "unit": UNIT, "batch_per_gpu": n, "steps": K,
            "ms_per_step": ms, "gflop_per_image": fl / 1e9, "tflops_per_gpu": tfl, "frac_of_bf16_peak": tfl / pk["tflops"],
            "gpu_launches": _lib.launch_count() - launches0, "h2d_bytes_per_step": z_host.numel() * 4,
            "d2h_bytes_per_step": out_host.numel(), "note": "rank 0's time x world (independent images, no collective)"}


def measure_xl(args, dev):
    """BASELINE.json configs[4]: LightningDiT-XL/1 at 512 px (64x64x16 latents, 4096 tokens, head_dim 72), sampling only, a
    short grid (`--xl-points`) on `--xl-batch` images: TFLOP/s of executed DiT work and the img/s it scales to."""
    import torch
    from ldmae_b200 import _lib
    from ldmae_b200.pipeline import SamplingJob, build_sampling_models
    model, vae = build_sampling_models(dev, seed=0, model_name="LightningDiT-XL/1", input_size=64, img_size=512)
    n, pts = args.xl_batch, args.xl_points
    job = SamplingJob(model, vae, num_steps=pts, cfg_scale=10.0, cfg_interval_start=0.10, timestep_shift=0.3)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n, 16, 64, 64, generator=g).to(dev)
    y = torch.randint(0, 1000, (n,), generator=g).to(dev)
    job.sample_latents(z, y)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lat = job.sample_latents(z, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    # one more, untimed, pass with the per-class event timers on (they serialise the launch groups)
    _lib.profile_begin()
    job.sample_latents(z, y)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    _, fwd = model_flops("LightningDiT-XL/1", 64)
    pk = peaks()
    tfl = fwd * 2 * n * (pts - 1) / (ms / 1e3) / 1e12
    ok = bool(torch.isfinite(lat).all())
    train = None
    if args.xl_train:
        # training step of the same model (BASELINE configs[4] names both): FusedTrainer on `--xl-train-batch` samples per GPU
        import argparse as _ap
        del job
        targs = _ap.Namespace(**vars(args))
        targs.model, targs.input_size, targs.train_batch, targs.train_steps = "LightningDiT-XL/1", 64, args.xl_train_batch, 3
        train = measure_train(targs, dev, 0, 1, model=model)
        job = None
    del job, model, vae
    torch.cuda.empty_cache()
    return {"metric": "LightningDiT-XL/1 @512px sampling (short grid)", "train": train, "batch": n, "evaluations": pts - 1, "ms": ms,
            "tflops_per_gpu": tfl, "frac_of_bf16_peak": tfl / pk["tflops"], "frac_of_bf16_burst_peak": tfl / pk["tflops_burst"] if pk["tflops_burst"] else None,
            "img_per_s_scaled_to_250_points": n / (ms / 1e3 / (pts - 1) * 249), "latents_finite": ok,
            "class_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0}, "gpu_launches": launches}


# ----------------------------------------------------------------------------- GPU arm
def plan_steps(warmup, steps, t_elapsed, t_step, reserve, budget):
    """Budget guard: how many warm-up / timed steps still fit after the first warm-up step took `t_step` seconds and
    `t_elapsed` seconds have passed since process start.  Warm-up is cut first (never below min(warmup, 3)), then timed
    steps (never below 1).  Returns (warmup, steps, notes)."""
    notes = []
    if warmup <= 0 or t_elapsed + (warmup - 1 + steps) * t_step + reserve <= budget:
        return warmup, steps, notes
    fit = int((budget - reserve - t_elapsed) // t_step)              # steps (remaining warm-up + timed) that still fit
    w_new = max(min(warmup, 3), min(warmup, fit - steps + 1))
    k_new = max(1, min(steps, fit - (w_new - 1)))
    if w_new != warmup:
        notes.append(f"{warmup - w_new} warm-up steps (kept {w_new})")
    if k_new != steps:
        notes.append(f"{steps - k_new} timed steps (kept {k_new})")
    return w_new, k_new, notes


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ldmae_b200 import _lib
    from ldmae_b200.pipeline import SamplingJob, build_sampling_models

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.train_only:
        tr = measure_train(args, dev, rank, world)
        if rank == 0:
            print(json.dumps(tr), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(vals):
        """max over ranks of a few host floats (so every rank takes the same budget decisions)."""
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    model, vae = build_sampling_models(dev, seed=0, model_name=args.model, input_size=args.input_size, img_size=8 * args.input_size)
    job = SamplingJob(model, vae, num_steps=args.num_steps, cfg_scale=10.0, cfg_interval_start=0.10, timestep_shift=0.3)
    n = args.batch
    g = torch.Generator().manual_seed(0 * world + rank)                  # inference.py:87
    S = args.input_size
    z_host = torch.randn(n, 16, S, S, generator=g).pin_memory()
    y_host = torch.randint(0, 1000, (n,), generator=g).pin_memory()
    out_host = torch.empty(n, 8 * S, 8 * S, 3, dtype=torch.uint8).pin_memory()
    h2d_bytes, d2h_bytes = z_host.numel() * 4 + y_host.numel() * 8, out_host.numel()

    # ---- budget: the first warm-up step (includes handle creation + weight upload) bounds the step time ---------------
    budget = {"limit_s": args.time_budget, "steps_requested": args.steps, "warmup_requested": args.warmup, "skipped": []}
    W, K = args.warmup, args.steps
    reserve = (30.0 if (world == 1 and not args.no_cpu_baseline) else 0.0) + 20.0
    t0 = time.time()
    if W > 0:
        job.run_host(z_host, y_host, out_host)
    t_first, t_el = agree([time.time() - t0, elapsed()])
    W, K, notes = plan_steps(W, K, t_el, t_first, reserve, args.time_budget)
    budget["skipped"].extend(notes)
    for _ in range(max(0, W - 1)):
        job.run_host(z_host, y_host, out_host)
    barrier()

    # ---- timed region: K steps, each ONE run_host call; inner event pair = device-resident job, outer pair = end to end
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    launches0 = _lib.launch_count()
    _lib.profile_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        job.run_host(z_host, y_host, out_host, events=evs[i])
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    ms_bracket = ev0.elapsed_time(ev1)                                  # K steps, copies included, bracketed as the contract says
    ms_device = sum(e[1].elapsed_time(e[2]) for e in evs)               # inputs resident in HBM ... images decoded on the device
    ms_e2e = sum(e[0].elapsed_time(e[3]) for e in evs)
    prof = _lib.profile_end()
    launches = _lib.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None

    # ---- validate what was timed: the last step's images -------------------------------------------------------------
    u8 = out_host.numpy()
    first8 = u8[:8]
    check = {"std": float(u8.std()), "mean": float(u8.mean()), "frac_saturated": float(((u8 == 0) | (u8 == 255)).mean()),
             "crc32_first8": int(zlib.crc32(first8.tobytes())), "sum_first8": int(first8.astype("int64").sum())}
    valid = check["std"] > 1.0 and check["frac_saturated"] < 0.9
    ms_device, ms_e2e, ms_bracket, ok_all = agree([ms_device, ms_e2e, ms_bracket, 0.0 if valid else 1.0])
    if ok_all != 0.0:
        raise SystemExit(f"bench.py: the timed job produced degenerate images on some rank (rank {rank}: {check}); no number reported")

    # ---- extras, cheapest first, while the budget lasts (estimates are conservative) --------------------------------------
    t_job = ms_e2e / K / 1e3
    extras = {}

    def fits(name, est_s, enabled=True):
        if not enabled:
            return False
        t_el, = agree([elapsed()])
        if t_el + est_s + 10.0 > args.time_budget:
            budget["skipped"].append(name)
            return False
        return True

    def guarded(fn, *a, **k):
        """The headline is already measured and validated at this point: on one GPU a failing extra is recorded in the line as
        {"error": ...} instead of costing it.  Under torchrun exceptions propagate (a rank failing alone would leave the others
        inside a collective; torchrun then ends the whole job)."""
        if world > 1:
            return fn(*a, **k)
        try:
            return fn(*a, **k)
        except Exception as ex:                                             # noqa: BLE001
            print(f"bench.py: extra object {getattr(fn, '__name__', fn)} failed: {type(ex).__name__}: {ex}", file=sys.stderr, flush=True)
            return {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    if fits("vmae_decode object", 5.0, not args.no_decode_extra):
        extras["vmae_decode"] = guarded(measure_decode, args, dev, vae, world)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # required leg: runs even when the budget is tight (reserved above), on a bounded sample
        def cpu_leg():
            threads = os.cpu_count() or 1
            per_img, t_ode, t_dec, kind = cpu_job(args.cpu_images, args.cpu_points, threads)
            return {"value": 1.0 / per_img, "unit": UNIT, "cores": threads, "kind": kind,
                    "sample": f"{args.cpu_images} images x {args.cpu_points - 1} CFG evaluations ({'the reference modules' if kind == 'reference' else 'fp32 oracle port of the reference'}, "
                              f"{t_ode:.1f} s) + VMAE decode ({t_dec:.1f} s), scaled to 249 evaluations"}
        cpu = guarded(cpu_leg)

    # the cond-only extension is one more full job and comes last (it is never the headline)
    def cond_only_leg():
        job2 = SamplingJob(model, vae, num_steps=args.num_steps, cfg_scale=10.0, cfg_interval_start=0.10, timestep_shift=0.3,
                           cond_only_when_unguided=True)
        z_dev, y_dev = z_host.to(dev), y_host.to(dev)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        job2.run_device(z_dev, y_dev)
        s1.record()
        barrier()
        ms_skip = s0.elapsed_time(s1)
        return {"value": world * n / (ms_skip / 1e3), "unit": UNIT, "ms_per_step": ms_skip, "steps": 1,
                "sample_forwards_per_image": job2.sample_forwards_per_image,
                "note": "extension, NOT the headline: steps with t < cfg_interval_start evaluate only the conditional half (its guided "
                        "velocity is its own prediction, lightningdit.py:436-439); images identical, 13.7% fewer FLOPs; rank 0's time"}

    if fits("cond_only_when_unguided object", t_job + 3.0 + 25.0 + 40.0, not args.no_cond_only_extra):
        extras["cond_only_when_unguided"] = guarded(cond_only_leg)
    del job
    if fits("train object", 25.0, not args.no_train):
        extras["train"] = guarded(measure_train, args, dev, rank, world, model=model)
    del model, vae
    torch.cuda.empty_cache()
    if fits("xl_512 object", 40.0, not args.no_xl_extra and world == 1 and args.model == "LightningDiT-B/1"):
        extras["xl_512"] = guarded(measure_xl, args, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_device / K
    value = world * n / (ms_per_step / 1e3)
    e2e_val = world * n / (ms_e2e / K / 1e3)

    # ---- roofline of the dominant kernel class ----------------------------------------------------------------------
    pk = peaks()
    per_block, fwd_flops = model_flops(args.model, args.input_size)
    Bf = 2 * n
    gemm_classes = {k: v for k, v in prof.items() if k in per_block and v[1] > 0}
    dom = max(gemm_classes, key=lambda k: gemm_classes[k][0]) if gemm_classes else None
    roof = None
    class_rates = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    ncu_traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    for k, (ms, cnt) in gemm_classes.items():
        class_rates[k] = round(per_block[k] * Bf / (ms / cnt / 1e3) / 1e12, 1)
    if dom:
        ms, cnt = gemm_classes[dom]
        flops_per_launch = per_block[dom] * Bf
        achieved = flops_per_launch / (ms / cnt / 1e3) / 1e12
        ent = ncu_traffic.get(dom)
        traffic = ent["bytes_per_sample_forward"] * Bf if ent else None     # measured DRAM bytes per launch (ncu capture)
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": pk["source"],
                "flops_per_launch": flops_per_launch, "avg_launch_ms": ms / cnt, "launches_timed": cnt}
    class_ms = {k: round(v[0] / K, 3) for k, v in prof.items() if v[1] > 0}
    job_flops = fwd_flops * Bf * (args.num_steps - 1)
    line = {"metric": METRIC if args.model == "LightningDiT-B/1" else f"{args.model} sampled img/s", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args), "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / K, "bracket_ms_per_step": ms_bracket / K, "host_wall_ms_per_step": wall_ms / K},
            "gpu_launches": launches, "roofline": roof, "output_check": check,
            "dit_tflops_per_gpu": job_flops / (ms_per_step / 1e3) / 1e12,
            "dit_frac_of_bf16_peak": job_flops / (ms_per_step / 1e3) / 1e12 / pk["tflops"],
            "dit_frac_of_bf16_burst_peak": job_flops / (ms_per_step / 1e3) / 1e12 / pk["tflops_burst"] if pk["tflops_burst"] else None,
            "class_ms_per_step": class_ms, "class_tflops": class_rates, "budget": budget}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    line.update(extras)
    budget["wall_s_at_print"] = round(elapsed(), 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ldmae_b200", choices=["ldmae_b200", "reference", "torch-gpu"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (BASELINE configs[1]: 256)")
    ap.add_argument("--num-steps", type=int, default=250, help="ODE grid points (250 = 249 evaluations)")
    ap.add_argument("--model", default="LightningDiT-B/1", choices=sorted(MODEL_GEOMETRY),
                    help="registry entry (headline: B/1; XL/1 with --input-size 64 is BASELINE configs[4], sampling only)")
    ap.add_argument("--input-size", type=int, default=32, help="latent side (32 = 256 px, 64 = 512 px)")
    ap.add_argument("--time-budget", type=float, default=float(os.environ.get("LDMAE_BENCH_BUDGET_S", "810")),
                    help="wall-clock seconds from process start the whole run must fit in (driver limit: 870 s per N)")
    ap.add_argument("--cpu-images", type=int, default=2)
    ap.add_argument("--cpu-points", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-batch", type=int, default=128, help="training samples per GPU per optimizer step")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--train-input", default="moments", choices=["moments", "latents"],
                    help="training-step input: the stored posterior moments (SURVEY 8d C3; on-device input pipeline) or ready latents")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement (the `train` object)")
    ap.add_argument("--train-only", action="store_true", help="measure only the training step and print its object")
    ap.add_argument("--no-decode-extra", action="store_true", help="skip the VMAE decode object (BASELINE configs[3])")
    ap.add_argument("--no-cond-only-extra", action="store_true", help="skip the cond-only-below-interval extension (one more job)")
    ap.add_argument("--no-xl-extra", action="store_true", help="skip the short XL/1 @512px sampling run (BASELINE configs[4])")
    ap.add_argument("--xl-train", action="store_true", help="also time XL/1 @512px training steps inside the xl_512 object (not in the default run)")
    ap.add_argument("--xl-train-batch", type=int, default=8)
    ap.add_argument("--xl-batch", type=int, default=16)
    ap.add_argument("--xl-points", type=int, default=6)
    ap.add_argument("--cmp-evals", type=int, default=10, help="--impl torch-gpu: evaluations timed per variant")
    ap.add_argument("--no-compile", action="store_true", help="--impl torch-gpu: skip the torch.compile variants")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "torch-gpu":
        run_torch_gpu(args, rank, world)
        return
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE (launch with torchrun for N>1)", file=sys.stderr)
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
