"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run here (authoring container, CPU), where /root/reference exists:

    python oracle/make_golden.py

It imports the reference's own files through oracle/refshim.py, loads deterministic synthetic
weights (oracle.ldmae_oracle.synth_*; the seed is stored, not the weights) into the reference's
own nn.Modules and records inputs + outputs.  tests/test_oracle_golden.py then asserts the oracle
restatement reproduces every output.  The GPU box has no /root/reference; it only reads the npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()

from oracle import ldmae_oracle as O  # noqa: E402

from models.lightningdit import LightningDiT, LightningDiT_models  # noqa: E402  (reference)
from transport import create_transport, Sampler  # noqa: E402               (reference)
import tokenizer.models_mae as ref_mae  # noqa: E402                         (reference)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


def load_into(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return missing


def tiny_dit(patch, **flags):
    spec = O.DiTSpec(depth=2, hidden_size=128, patch_size=patch, num_heads=2, input_size=8, in_channels=16,
                     num_classes=10, **flags)
    ref = LightningDiT(input_size=8, patch_size=patch, in_channels=16, hidden_size=128, depth=2, num_heads=2,
                       num_classes=10, use_qknorm=spec.use_qknorm, use_swiglu=spec.use_swiglu,
                       use_rope=spec.use_rope, use_rmsnorm=spec.use_rmsnorm, wo_shift=spec.wo_shift, learn_sigma=spec.learn_sigma)
    # the reference's own state_dict must have exactly the keys/shapes the oracle lists
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert ref_shapes == O.dit_param_shapes(spec), set(ref_shapes) ^ set(O.dit_param_shapes(spec))
    sd = O.synth_dit_state(spec, seed=11 + patch)
    # constant tables must equal the reference's own
    assert torch.equal(sd["pos_embed"], ref.state_dict()["pos_embed"])
    if spec.use_rope:
        assert torch.equal(sd["feat_rope.freqs_cos"], ref.state_dict()["feat_rope.freqs_cos"])
        assert torch.equal(sd["feat_rope.freqs_sin"], ref.state_dict()["feat_rope.freqs_sin"])
    assert not load_into(ref, sd)
    return spec, ref.eval(), sd


def gen_dit_tiny():
    for patch in (1, 2):
        spec, ref, sd = tiny_dit(patch)
        g = torch.Generator().manual_seed(100 + patch)
        n = 3
        x = torch.randn(2 * n, 16, 8, 8, generator=g)
        t = torch.rand(2 * n, generator=g)
        y = torch.randint(0, 10, (2 * n,), generator=g)
        out = ref(x, t, y)
        ycfg = torch.cat([y[:n], torch.full((n,), 10)])
        tc = torch.full((2 * n,), 0.37)
        cfg_hi = ref.forward_with_cfg(x, tc, ycfg, 4.0, cfg_interval=True, cfg_interval_start=0.10)
        tl = torch.full((2 * n,), 0.05)
        cfg_lo = ref.forward_with_cfg(x, tl, ycfg, 4.0, cfg_interval=True, cfg_interval_start=0.10)
        # sampler through the reference's own transport (Euler + shift, and Heun)
        tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
        smp = Sampler(tr)
        z = torch.cat([x[:n], x[:n]], 0)
        kw = dict(y=ycfg, cfg_scale=4.0, cfg_interval=True, cfg_interval_start=0.10)
        fn_e = smp.sample_ode(sampling_method="euler", num_steps=7, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.3)
        traj_e = fn_e(z, ref.forward_with_cfg, **kw)
        fn_h = smp.sample_ode(sampling_method="heun2", num_steps=4, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.0)
        traj_h = fn_h(z, ref.forward_with_cfg, **kw)
        fn_n = smp.sample_ode(sampling_method="euler", num_steps=5, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.3)
        traj_n = fn_n(x[:n], ref.forward, y=y[:n])
        # training loss with the reference's random draws captured (not replaced)
        torch.manual_seed(7); np.random.seed(7)
        cap = {}
        orig = tr.sample

        def spy(x1, *a, **k):
            r = orig(x1, *a, **k)
            cap["t"], cap["x0"] = r[0].clone(), r[1].clone()
            return r

        tr.sample = spy
        terms = tr.training_losses(ref, x, dict(y=y))       # eval mode: no label dropout
        np.savez_compressed(
            os.path.join(OUT, f"dit_tiny_p{patch}.npz"),
            seed=11 + patch, checksum=O.state_checksum(sd), x=x.numpy(), t=t.numpy(), y=y.numpy(), out=out.numpy(),
            ycfg=ycfg.numpy(), cfg_hi=cfg_hi.numpy(), cfg_lo=cfg_lo.numpy(),
            traj_euler=traj_e.numpy(), traj_heun=traj_h.numpy(), traj_nocfg=traj_n.numpy(),
            grid_euler=fn_e.__self__.t.numpy(), loss=terms["loss"].numpy(), pred=terms["pred"].numpy(),
            loss_t=cap["t"].numpy(), loss_x0=cap["x0"].numpy())
        print("dit tiny patch", patch, "out absmax", out.abs().max().item(), "loss", terms["loss"].mean().item())


def gen_dit_grads():
    """Gradients of the reference's own training loss (transport.training_losses -> loss.mean().backward(),
    train_accum.py:215-230) for the tiny models, with the random draws (t, x0) fixed; eval mode (no label dropout)."""
    for patch in (1, 2):
        spec, ref, sd = tiny_dit(patch)
        g = torch.Generator().manual_seed(300 + patch)
        B = 4
        x1 = torch.randn(B, 16, 8, 8, generator=g)
        x0 = torch.randn(B, 16, 8, 8, generator=g)
        t = torch.rand(B, generator=g)
        y = torch.randint(0, 10, (B,), generator=g)
        tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
        tr.sample = lambda x1_, *a, **k: (t, x0, x1_)
        with torch.enable_grad():
            for p_ in ref.parameters():
                p_.grad = None
            terms = tr.training_losses(ref, x1, dict(y=y))
            terms["loss"].mean().backward()
        grads = {"grad." + k: p_.grad.numpy() for k, p_ in ref.named_parameters() if p_.grad is not None}
        np.savez_compressed(os.path.join(OUT, f"dit_tiny_grads_p{patch}.npz"), seed=11 + patch, checksum=O.state_checksum(sd),
                            x1=x1.numpy(), x0=x0.numpy(), t=t.numpy(), y=y.numpy(), loss=terms["loss"].detach().numpy(), **grads)
        print("dit tiny grads patch", patch, "loss", terms["loss"].mean().item(), len(grads), "tensors")


def gen_dit_variants():
    """Config flags the API must accept (SURVEY section 8a tail): celeba (no qk-norm, 1 class -> no cfg
    embedding is NOT the case: class_dropout_prob stays 0.1 unless num_classes==1), wo_shift."""
    for tag, flags in (("noqk", dict(use_qknorm=False)), ("woshift", dict(wo_shift=True)),
                       # fallbacks of lightningdit.py:195-224,257-261: LayerNorm block / head / final norms, timm Mlp + tanh-GELU
                       ("ln_gelu", dict(use_rmsnorm=False, use_swiglu=False)), ("ln_swiglu", dict(use_rmsnorm=False)),
                       ("rms_gelu", dict(use_swiglu=False)), ("ln_gelu_noqk", dict(use_rmsnorm=False, use_swiglu=False, use_qknorm=False)),
                       # lightningdit.py:298,415-417: twice the output channels, the second half dropped by forward; no RoPE (:317-323)
                       ("learnsigma", dict(learn_sigma=True)), ("learnsigma_p2", dict(learn_sigma=True, patch=2)), ("norope", dict(use_rope=False))):
        flags = dict(flags)
        patch = flags.pop("patch", 1)
        spec, ref, sd = tiny_dit(patch, **flags)
        g = torch.Generator().manual_seed(55)
        x = torch.randn(2, 16, 8, 8, generator=g); t = torch.rand(2, generator=g); y = torch.randint(0, 10, (2,), generator=g)
        out = ref(x, t, y)
        np.savez_compressed(os.path.join(OUT, f"dit_tiny_{tag}.npz"), seed=11 + patch, checksum=O.state_checksum(sd),
                            x=x.numpy(), t=t.numpy(), y=y.numpy(), out=out.numpy())
        print("dit variant", tag, out.abs().max().item())


def gen_dit_b1():
    spec = O.DiTSpec.named("LightningDiT-B/1", input_size=32, in_channels=16)
    ref = LightningDiT_models["LightningDiT-B/1"](input_size=32, in_channels=16, use_qknorm=True, use_swiglu=True,
                                                  use_rope=True, use_rmsnorm=True)
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert ref_shapes == O.dit_param_shapes(spec)
    sd = O.synth_dit_state(spec, seed=1234)
    assert torch.equal(sd["pos_embed"], ref.state_dict()["pos_embed"])
    assert torch.equal(sd["feat_rope.freqs_cos"], ref.state_dict()["feat_rope.freqs_cos"])
    assert not load_into(ref, sd)
    ref.eval()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 32, 32, generator=g)
    t = torch.tensor([0.0316, 0.7312])
    y = torch.tensor([207, 1000])
    out = ref(x, t, y)
    np.savez_compressed(os.path.join(OUT, "dit_b1_forward.npz"), seed=1234, checksum=O.state_checksum(sd),
                        x=x.numpy(), t=t.numpy(), y=y.numpy(), out=out.numpy())
    # key list: the drop-in contract for checkpoints
    with open(os.path.join(OUT, "dit_b1_keys.txt"), "w") as f:
        for k, v in ref.state_dict().items():
            f.write(f"{k} {list(v.shape)}\n")
    print("dit B/1 out absmax", out.abs().max().item(), "std", out.std().item())


def gen_vmae():
    for img_size, bs, tag in ((32, 2, "small"), (256, 1, "full")):
        spec = O.VMAESpec(img_size=img_size)
        ref = ref_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True,
                                               img_size=img_size).eval()
        sd = O.synth_vmae_state(spec, seed=77, encoder=True)
        ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        mine = dict(O.vmae_decoder_param_shapes(spec)); mine.update(O.vmae_encoder_param_shapes(spec))
        assert ref_shapes == mine, set(ref_shapes) ^ set(mine)
        assert torch.allclose(sd["decoder_pos_embed"], ref.state_dict()["decoder_pos_embed"])
        assert torch.equal(sd["decoder_pos_embed"], ref.state_dict()["decoder_pos_embed"])
        assert not load_into(ref, sd)
        g = torch.Generator().manual_seed(5)
        z = torch.randn(bs, 16, spec.grid, spec.grid, generator=g)
        img = ref.decode(z, return_dict=False)[0]
        u8 = torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8).numpy()  # models_mae.py:972
        pix = torch.randn(bs, 3, img_size, img_size, generator=g) * 0.5
        mom = ref._encode(pix)
        np.savez_compressed(os.path.join(OUT, f"vmae_{tag}.npz"), seed=77, checksum=O.state_checksum(sd), z=z.numpy(),
                            img=img.numpy(), u8=u8, pix=pix.numpy(), moments=mom.numpy())
        if tag == "full":
            with open(os.path.join(OUT, "vmae_keys.txt"), "w") as f:
                for k, v in ref.state_dict().items():
                    f.write(f"{k} {list(v.shape)}\n")
        print("vmae", tag, "img absmax", img.abs().max().item(), "u8 mean", u8.mean())


def gen_config1():
    """BASELINE.json configs[0]: LightningDiT-B/1, batch 8 (16 with CFG), 10-point shifted Euler grid (9 evaluations),
    cfg_scale 10 with cfg_interval_start 0.10, then VMAE decode -- the body of inference.py:264-292 through the
    reference's own modules (about 100 s of CPU).  The benchmarked job (configs[1]) is this recipe with 250 points."""
    spec = O.DiTSpec.named("LightningDiT-B/1", input_size=32, in_channels=16)
    ref = LightningDiT_models["LightningDiT-B/1"](input_size=32, in_channels=16, use_qknorm=True, use_swiglu=True,
                                                  use_rope=True, use_rmsnorm=True)
    sd = O.synth_dit_state(spec, seed=1234)
    assert not load_into(ref, sd)
    ref.eval()
    vspec = O.VMAESpec(img_size=256)
    vae = ref_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True,
                                           img_size=256).eval()
    vsd = O.synth_vmae_state(vspec, seed=77, encoder=True)
    assert not load_into(vae, vsd)
    g = torch.Generator().manual_seed(0)
    n = 8
    z = torch.randn(n, 16, 32, 32, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    zz = torch.cat([z, z], 0)                                               # inference.py:278-282
    yy = torch.cat([y, torch.full((n,), 1000)], 0)
    tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
    fn = Sampler(tr).sample_ode(sampling_method="euler", num_steps=10, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.3)
    traj = fn(zz, ref.forward_with_cfg, y=yy, cfg_scale=10.0, cfg_interval=True, cfg_interval_start=0.10)
    lat = traj[-1].chunk(2, dim=0)[0]                                       # inference.py:287-289
    mean = torch.linspace(-0.25, 0.25, 16).view(1, 16, 1, 1)                # synthetic latents_stats (inference.py:291)
    std = torch.linspace(0.75, 1.25, 16).view(1, 16, 1, 1)
    mult = 1.0
    img = vae.decode((lat * std) / mult + mean, return_dict=False)[0]
    u8 = torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8).numpy()
    np.savez_compressed(os.path.join(OUT, "config1_b1_cfg10.npz"), dit_seed=1234, vmae_seed=77,
                        dit_checksum=O.state_checksum(sd), vmae_checksum=O.state_checksum(vsd), z=z.numpy(), y=y.numpy(),
                        grid=fn.__self__.t.numpy(), latents=lat.numpy(), traj_norm=traj.flatten(1).norm(dim=1).numpy(),
                        mid_state=traj[5][:2].numpy(), latent_mean=mean.numpy(), latent_std=std.numpy(), latent_multiplier=mult,
                        img_first=img[:1].numpy(), u8=u8[:4])
    print("config 1: latent absmax", lat.abs().max().item(), "std", lat.std().item(), "u8 mean", u8.mean())


from oracle.make_golden_cases import SDE_CASES, state_fingerprint  # noqa: E402


def gen_sde():
    """The reference's OWN SDE samplers (transport.py:285-396, integrators.py:8-75 -- no torchdiffeq involved) on the tiny
    reference LightningDiT: every state of a 6-point run per (method, diffusion form, norm, last step, last step size), with
    the global torch RNG seeded right before the call (the noise draws `th.randn(x.size())` come from it in call order).
    Not covered: diffusion_form "constant" (the reference's `th.sqrt(2 * diffusion)` raises TypeError on the Python float that
    form returns, integrators.py:33,44) and the default "SBDM" (starts at t = sample_eps = 0 where 1/t is infinite)."""
    spec, ref, sd = tiny_dit(1)
    g = torch.Generator().manual_seed(321)
    x = torch.randn(3, 16, 8, 8, generator=g)
    y = torch.randint(0, 10, (3,), generator=g)
    tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
    smp = Sampler(tr)
    out = {}
    for i, (method, form, norm, last, lsz) in enumerate(SDE_CASES):
        fn = smp.sample_sde(sampling_method=method, diffusion_form=form, diffusion_norm=norm, last_step=last, last_step_size=lsz,
                            num_steps=6)
        torch.manual_seed(1000 + i)
        xs = fn(x, ref.forward, y=y)
        out[f"traj{i}"] = torch.stack(xs).numpy()
        print("sde", method, form, last, "final absmax", xs[-1].abs().max().item())
    np.savez_compressed(os.path.join(OUT, "sde_tiny.npz"), seed=12, checksum=O.state_checksum(sd), x=x.numpy(), y=y.numpy(), **out)


def gen_dataset():
    """The trainer's input format: two safetensors shards in the layout of extract_features.py:168-181 are written
    under tests/golden/latent_shards/ and read back by the reference's OWN ImgLatentDataset (datasets/img_latent_dataset.py,
    loaded by file path: a HuggingFace `datasets` package may shadow the name) with seeded global RNGs -- the cached statistics
    and every item, for (latent_norm, sample, multiplier) = (True, True, 1.3) and (False, False, 1.0)."""
    import importlib.util
    import shutil
    import tempfile
    from safetensors.torch import save_file
    spec_ = importlib.util.spec_from_file_location("ref_img_latent_dataset", os.path.join("/root/reference/LDMAE", "datasets", "img_latent_dataset.py"))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    shard_dir = os.path.join(OUT, "latent_shards")
    shutil.rmtree(shard_dir, ignore_errors=True)
    os.makedirs(shard_dir)
    g = torch.Generator().manual_seed(77)
    for shard, n in enumerate((3, 2)):
        lat = torch.randn(n, 32, 4, 4, generator=g) * 0.5
        flip = torch.randn(n, 32, 4, 4, generator=g) * 0.5
        labels = torch.randint(0, 1000, (n,), generator=g)
        save_file({"latents": lat, "latents_flip": flip, "labels": labels},                     # extract_features.py:168-181
                  os.path.join(shard_dir, f"latents_rank00_shard{shard:03d}.safetensors"),
                  metadata={"total_size": f"{n}", "dtype": f"{lat.dtype}", "device": f"{lat.device}"})
    out = {}
    for tag, (norm, sample, mult) in (("a", (True, True, 1.3)), ("b", (False, False, 1.0))):
        with tempfile.TemporaryDirectory() as tmp:
            for f in os.listdir(shard_dir):
                shutil.copy(os.path.join(shard_dir, f), tmp)
            np.random.seed(5); torch.manual_seed(5)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                ds = mod.ImgLatentDataset(tmp, latent_norm=norm, latent_multiplier=mult, sample=sample)
            items = [ds[i] for i in range(len(ds))]
            out[f"{tag}_features"] = torch.stack([f for f, _ in items]).numpy()
            out[f"{tag}_labels"] = torch.stack([l for _, l in items]).numpy()
            if norm:
                out[f"{tag}_mean"], out[f"{tag}_std"] = ds._latent_mean.numpy(), ds._latent_std.numpy()
    np.savez_compressed(os.path.join(OUT, "dataset_items.npz"), **out)
    print("dataset goldens:", {k: v.shape for k, v in out.items()})


def _ref_function(path, name):
    """One top-level function of a reference driver script that cannot be imported as a whole (train_accum.py needs accelerate
    and CUDA): its source segment is compiled on its own with torch in scope."""
    import ast
    from collections import OrderedDict
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "OrderedDict": OrderedDict}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def gen_host_helpers():
    """Host-side helpers pinned to the reference's own code: the image pre-processing of extract_features.py
    (tokenizer/models_mae.py:85-103 center_crop_arr, :935-950 img_transform) on synthetic PIL images, and the trainer's
    fine-tuning initialisation (train_accum.py:308-334 load_weights_with_shape_check) on tiny reference models."""
    from PIL import Image
    rng = np.random.RandomState(3)
    out = {}
    vae = ref_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True, img_size=64)
    for i, (h, w) in enumerate(((75, 101), (300, 170), (64, 64))):          # plain resize, two BOX halvings first, identity
        arr = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        out[f"img{i}"] = arr
        out[f"crop{i}"] = np.array(ref_mae.center_crop_arr(Image.fromarray(arr), 64))
        out[f"tensor{i}"] = vae.img_transform(p_hflip=0)(Image.fromarray(arr)).numpy()
    # fine-tuning init: 16-channel checkpoint into a 32-channel model, one mismatched and one unknown tensor
    load_ref = _ref_function("/root/reference/LDMAE/train_accum.py", "load_weights_with_shape_check")
    mk = lambda c: LightningDiT(input_size=8, patch_size=1, in_channels=c, hidden_size=128, depth=2, num_heads=2, num_classes=10,
                                use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    spec16 = O.DiTSpec(depth=2, hidden_size=128, patch_size=1, num_heads=2, input_size=8, in_channels=16, num_classes=10)
    spec32 = O.DiTSpec(depth=2, hidden_size=128, patch_size=1, num_heads=2, input_size=8, in_channels=32, num_classes=10)
    target = mk(32)
    assert not load_into(target, O.synth_dit_state(spec32, 21))
    ckpt = {"model": dict(O.synth_dit_state(spec16, 22))}
    ckpt["model"]["blocks.0.attn.q_norm.weight"] = torch.ones(32)           # shape mismatch -> skipped
    ckpt["model"]["not_a_parameter"] = torch.ones(3)                        # unknown -> skipped
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        load_ref(target, ckpt, rank=0)
    keys, fp = state_fingerprint(target.state_dict())
    out["load_keys"], out["load_fp"] = np.array(keys), fp
    out["load_proj"] = target.state_dict()["x_embedder.proj.weight"].numpy()
    np.savez_compressed(os.path.join(OUT, "host_helpers.npz"), **out)
    print("host helper goldens:", {k: getattr(v, "shape", None) for k, v in out.items()})


def gen_dit_hd72():
    """The XL head geometry (head_dim 72 = hidden 1152 / 16 heads, lightningdit.py:509-515) at depth 2 (hidden 1152, 16 heads, 256 tokens),
    so the RoPE tables ([T, 72], 18 frequencies per axis) and the per-head RMSNorm(72) of the reference pin the oracle for the
    wide-head path too (the GPU tests of the XL kernels compare with the oracle)."""
    spec = O.DiTSpec(depth=2, hidden_size=1152, patch_size=1, num_heads=16, input_size=16, in_channels=16, num_classes=10)
    ref = LightningDiT(input_size=16, patch_size=1, in_channels=16, hidden_size=1152, depth=2, num_heads=16, num_classes=10,
                       use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert ref_shapes == O.dit_param_shapes(spec), set(ref_shapes) ^ set(O.dit_param_shapes(spec))
    sd = O.synth_dit_state(spec, seed=72)
    for k in ("pos_embed", "feat_rope.freqs_cos", "feat_rope.freqs_sin"):
        assert torch.equal(sd[k], ref.state_dict()[k]), k
    assert not load_into(ref, sd)
    ref.eval()
    g = torch.Generator().manual_seed(172)
    n = 2
    x = torch.randn(2 * n, 16, 16, 16, generator=g); t = torch.rand(2 * n, generator=g); y = torch.randint(0, 10, (2 * n,), generator=g)
    out = ref(x, t, y)
    ycfg = torch.cat([y[:n], torch.full((n,), 10)])
    tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
    fn = Sampler(tr).sample_ode(sampling_method="euler", num_steps=5, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.3)
    traj = fn(torch.cat([x[:n], x[:n]], 0), ref.forward_with_cfg, y=ycfg, cfg_scale=4.0, cfg_interval=True, cfg_interval_start=0.10)
    np.savez_compressed(os.path.join(OUT, "dit_tiny_hd72.npz"), seed=72, checksum=O.state_checksum(sd), x=x.numpy(), t=t.numpy(),
                        y=y.numpy(), ycfg=ycfg.numpy(), out=out.numpy(), traj_last=traj[-1].numpy(),
                        rope_cos=ref.state_dict()["feat_rope.freqs_cos"].numpy())
    print("dit hd72: out absmax", out.abs().max().item(), "rope table", tuple(ref.state_dict()["feat_rope.freqs_cos"].shape))


if __name__ == "__main__":
    if "--config1-only" in sys.argv:
        gen_config1()
        sys.exit(0)
    if "--variants-only" in sys.argv:
        gen_dit_variants()
        sys.exit(0)
    if "--grads-only" in sys.argv:
        gen_dit_grads()
        sys.exit(0)
    if "--sde-only" in sys.argv:
        gen_sde()
        sys.exit(0)
    if "--dataset-only" in sys.argv:
        gen_dataset()
        sys.exit(0)
    if "--helpers-only" in sys.argv:
        gen_host_helpers()
        sys.exit(0)
    if "--hd72-only" in sys.argv:
        gen_dit_hd72()
        sys.exit(0)
    gen_dit_tiny()
    gen_dit_grads()
    gen_dit_variants()
    gen_dit_b1()
    gen_vmae()
    gen_config1()
    gen_sde()
    gen_dataset()
    gen_host_helpers()
    gen_dit_hd72()
    print("golden fixtures written to", OUT)
