#!/usr/bin/env python
"""Headline benchmark: LightningDiT-B/1 ImageNet-256 class-conditional sampling (BASELINE.json configs[1]).

One step = one sampling job of `--batch` images per GPU: 250-point shifted Euler grid (249 model evaluations, each on
the CFG-doubled batch), cfg_scale 10 with cfg_interval_start 0.10, then latent de-normalisation + VMAE decode to uint8
(the body of the reference's LDMAE/inference.py:264-292 loop).  Synthetic latents / labels, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo (sm_100a kernels)
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, weak scaling (batch/GPU fixed)
    python bench.py --impl reference ...                                # the reference algorithm's CPU path on host cores

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same job through
ldmae_b200.pipeline.SamplingJob.run_host with pinned-host inputs and the uint8 images read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LightningDiT-B sampled img/s"
UNIT = "img/s"


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


MODEL_GEOMETRY = {   # registry lightningdit.py:498-531: (depth, width, heads, patch)
    "LightningDiT-B/1": (12, 768, 12, 1), "LightningDiT-L/2": (24, 1024, 16, 2), "LightningDiT-XL/1": (28, 1152, 16, 1),
    "LightningDiT-XL/2": (28, 1152, 16, 2), "LightningDiT-1p0B/1": (24, 1536, 24, 1), "LightningDiT-1p6B/1": (28, 1792, 28, 1),
}


def model_flops(args):
    depth, D, _, patch = MODEL_GEOMETRY[args.model]
    T = (args.input_size // patch) ** 2
    H = int(2 / 3 * 4 * D)
    return dit_flops_per_sample_forward(depth=depth, D=D, T=T, H=H, C=16 * patch * patch)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "hbm_gbs": float(d["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: 1.59 PF burst / ~1.4 PF sustained)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ----------------------------------------------------------------------------- CPU path (oracle port of the reference algorithm)
def cpu_job(n, num_points, threads):
    """Times the reference algorithm's CPU restatement (oracle/) on a bounded sample of the workload: n images, CFG,
    `num_points`-point shifted Euler grid, then VMAE decode of the n latents.  Returns per-image seconds for the full
    249-evaluation job (DiT seconds per image-evaluation x 249 + decode seconds per image)."""
    import torch
    from oracle import ldmae_oracle as O       # bench.py's cpu_baseline / reference legs are allowed to run the oracle
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    ds = O.DiTSpec.named("LightningDiT-B/1", input_size=32, in_channels=16)
    vs = O.VMAESpec(img_size=256)
    dsd, vsd = O.synth_dit_state(ds, 0), O.synth_vmae_state(vs, 1)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n, 16, 32, 32, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    zz = torch.cat([z, z], 0)
    yy = torch.cat([y, torch.full((n,), ds.num_classes, dtype=y.dtype)], 0)
    fn = lambda x, t, **kw: O.dit_forward_with_cfg(dsd, ds, x, t, **kw)
    t0 = time.perf_counter()
    lat = O.sample_ode(fn, zz, sampling_method="euler", num_steps=num_points, timestep_shift=0.3, y=yy, cfg_scale=10.0,
                       cfg_interval=True, cfg_interval_start=0.10)[-1].chunk(2, dim=0)[0]
    t1 = time.perf_counter()
    img = O.vmae_decode(vsd, vs, lat)
    O.images_to_uint8(img)
    t2 = time.perf_counter()
    evals = num_points - 1
    per_img = (t1 - t0) / (n * evals) * 249 + (t2 - t1) / n
    return per_img, (t1 - t0), (t2 - t1)


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host CPU.  /root/reference is pure Python with un-installed
    dependencies and does not travel to the GPU box, so this leg runs the oracle port (pinned to reference goldens)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, pts = args.cpu_images, args.cpu_points
    for _ in range(args.warmup):
        cpu_job(1, 2, threads)                         # warm-up: one evaluation of one image (page in weights / threads)
    per = []
    for _ in range(args.steps):
        per_img, _, _ = cpu_job(n, pts, threads)
        per.append(per_img)
    per_img = sum(per) / len(per)
    val = 1.0 / per_img
    sample = (f"{n} images x {pts - 1} CFG evaluations of LightningDiT-B/1 (fp32, {threads} threads) + VMAE decode of {n} images per "
              f"step; per-image time extrapolated linearly to 249 evaluations")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_img * args.batch * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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


# ----------------------------------------------------------------------------- training step (BASELINE.json configs[2])
def measure_train(args, dev, rank, world):
    """train samples/s: one step = one optimizer step of train_accum.py:203-246 on `--train-batch` synthetic latents per GPU
    (flow-matching loss, forward + backward, gradient all-reduce across ranks, fused AdamW + EMA, weight re-pack).
    Inputs come from pinned host memory every step and the per-sample losses are read back (end-to-end by construction)."""
    import torch
    import torch.distributed as dist
    from ldmae_b200 import _lib
    from ldmae_b200.pipeline import build_sampling_models
    from ldmae_b200.training import FusedTrainer
    B = args.train_batch
    model, _ = build_sampling_models(dev, seed=0, model_name=args.model, input_size=args.input_size, img_size=8 * args.input_size)
    model.train()
    trainer = FusedTrainer(model, lr=2e-4, betas=(0.9, 0.95), weight_decay=0.0, ema_decay=0.9999)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(B, 16, args.input_size, args.input_size, generator=g).pin_memory()
    y_host = torch.randint(0, 1000, (B,), generator=g).pin_memory()
    loss_host = torch.empty(B).pin_memory()

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
        for _ in range(max(3, args.warmup)):
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
    _, fwd = model_flops(args)
    pk = peaks()
    tfl = 3 * fwd * B / (ms_step / 1e3) / 1e12
    del trainer, model
    torch.cuda.empty_cache()
    return {"metric": ("LightningDiT-B" if args.model == "LightningDiT-B/1" else args.model) + " train samples/s", "value": world * B / (ms_step / 1e3), "unit": "samples/s",
            "batch_per_gpu": B, "global_batch": world * B, "steps": args.train_steps, "ms_per_step": ms_step,
            "tflops_per_gpu": tfl, "frac_of_bf16_peak": tfl / pk["tflops"], "flops_per_sample": 3 * fwd,
            "gpu_launches": launches, "final_loss": float(loss_host.mean()),
            "class_ms_per_step": {k: round(v[0] / args.train_steps, 3) for k, v in prof.items() if v[1] > 0},
            "includes": "H2D of latents/labels, label dropout, forward, loss, backward, gradient all-reduce (N>1), fused AdamW+EMA, "
                        "bf16 weight re-pack, D2H of per-sample losses",
            "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": B * 4}


# ----------------------------------------------------------------------------- GPU arm
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

    model, vae = build_sampling_models(dev, seed=0, model_name=args.model, input_size=args.input_size, img_size=8 * args.input_size)
    job = SamplingJob(model, vae, num_steps=args.num_steps, cfg_scale=10.0, cfg_interval_start=0.10, timestep_shift=0.3)
    n = args.batch
    g = torch.Generator().manual_seed(0 * world + rank)                  # inference.py:87
    S = args.input_size
    z_host = torch.randn(n, 16, S, S, generator=g).pin_memory()
    y_host = torch.randint(0, 1000, (n,), generator=g).pin_memory()
    out_host = torch.empty(n, 8 * S, 8 * S, 3, dtype=torch.uint8).pin_memory()
    z_dev, y_dev = z_host.to(dev), y_host.to(dev)

    for _ in range(args.warmup):
        job.run_device(z_dev, y_dev)
    barrier()

    # ---- timed region A: inputs resident in HBM ---------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    _lib.profile_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        u8 = job.run_device(z_dev, y_dev)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    prof = _lib.profile_end()
    launches = _lib.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None

    # ---- timed region B: end to end from pinned host memory, images read back ------------------------------------
    job.run_host(z_host, y_host, out_host)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        job.run_host(z_host, y_host, out_host)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3              # host wall clock around the same K steps (sanity record)

    # ---- extra (not the headline): the same job with the conditional-only shortcut below the guidance interval -----
    ms_skip = None
    if not args.no_cond_only_extra:
        job2 = SamplingJob(model, vae, num_steps=args.num_steps, cfg_scale=10.0, cfg_interval_start=0.10, timestep_shift=0.3,
                           cond_only_when_unguided=True)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        job2.run_device(z_dev, y_dev)
        s1.record()
        barrier()
        ms_skip = s0.elapsed_time(s1)

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])
    train = None
    if not args.no_train:
        del job, model, vae
        torch.cuda.empty_cache()
        train = measure_train(args, dev, rank, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step / 1e3)
    e2e_val = world * n / (ms_e2e / args.steps / 1e3)

    # ---- roofline of the dominant kernel class ----------------------------------------------------------------------
    pk = peaks()
    per_block, fwd_flops = model_flops(args)
    Bf = 2 * n
    gemm_classes = {k: v for k, v in prof.items() if k in per_block and v[1] > 0}
    dom = max(gemm_classes, key=lambda k: gemm_classes[k][0]) if gemm_classes else None
    roof = None
    if dom:
        ms, cnt = gemm_classes[dom]
        flops_per_launch = per_block[dom] * Bf
        achieved = flops_per_launch / (ms / cnt / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            ent = json.load(open(tp)).get(dom)
            if ent:
                traffic = ent["bytes_per_sample_forward"] * Bf          # measured DRAM bytes per launch (ncu capture)
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": pk["source"],
                "flops_per_launch": flops_per_launch, "avg_launch_ms": ms / cnt, "launches_timed": cnt}
    class_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    job_flops = fwd_flops * Bf * (args.num_steps - 1)
    line = {"metric": METRIC if args.model == "LightningDiT-B/1" else f"{args.model} sampled img/s", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args), "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": z_host.numel() * 4 + y_host.numel() * 8,
                    "d2h_bytes_per_step": out_host.numel(), "ms_per_step": ms_e2e / args.steps,
                    "host_wall_ms_per_step": wall_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof,
            "dit_tflops_per_gpu": job_flops / (ms_per_step / 1e3) / 1e12,
            "dit_frac_of_bf16_peak": job_flops / (ms_per_step / 1e3) / 1e12 / pk["tflops"],
            "class_ms_per_step": class_ms}
    if train is not None:
        line["train"] = train
    if ms_skip is not None:
        line["cond_only_when_unguided"] = {
            "value": world * n / (ms_skip / 1e3), "unit": UNIT, "ms_per_step": ms_skip, "steps": 1,
            "sample_forwards_per_image": job2.sample_forwards_per_image,
            "note": "extension, NOT the headline: steps with t < cfg_interval_start evaluate only the conditional half (its guided "
                    "velocity is its own prediction, lightningdit.py:436-439); images identical, 13.7% fewer FLOPs; rank 0's time"}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        per_img, t_ode, t_dec = cpu_job(args.cpu_images, args.cpu_points, threads)
        line["cpu_baseline"] = {"value": 1.0 / per_img, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_images} images x {args.cpu_points - 1} CFG evaluations (fp32 oracle port of the "
                                          f"reference, {t_ode:.1f} s) + VMAE decode ({t_dec:.1f} s), extrapolated to 249 evaluations"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ldmae_b200", choices=["ldmae_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (BASELINE configs[1]: 256)")
    ap.add_argument("--num-steps", type=int, default=250, help="ODE grid points (250 = 249 evaluations)")
    ap.add_argument("--model", default="LightningDiT-B/1", choices=sorted(MODEL_GEOMETRY),
                    help="registry entry (headline: B/1; XL/1 with --input-size 64 is BASELINE configs[4], sampling only)")
    ap.add_argument("--input-size", type=int, default=32, help="latent side (32 = 256 px, 64 = 512 px)")
    ap.add_argument("--cpu-images", type=int, default=2)
    ap.add_argument("--cpu-points", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-batch", type=int, default=128, help="training samples per GPU per optimizer step")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement (the `train` object)")
    ap.add_argument("--train-only", action="store_true", help="measure only the training step and print its object")
    ap.add_argument("--no-cond-only-extra", action="store_true", help="skip the extra cond-only-below-interval measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE (launch with torchrun for N>1)", file=sys.stderr)
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
