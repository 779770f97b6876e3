"""Checkpoint contract of the reference's trainer and sampler (LDMAE/train_accum.py:94-103,172-187,273-284,308-334,345;
LDMAE/inference.py:100-103).

A training checkpoint is ``{"model": state_dict, "ema": state_dict, "opt": AdamW.state_dict(), "config": train_config}``
written as ``<checkpoint_dir>/<train_steps:07d>.pt``; state_dict keys may carry DDP's ``module.`` prefix; fine-tuning
initialises from a checkpoint with a shape-checked copy that zero-pads ``x_embedder.proj.weight`` when the latent channel
count grew; the sampler loads ``checkpoint["ema"]`` when present.  Everything here is host-side state_dict plumbing -- the
tensors land in the ``nn.Parameter`` containers of ``ldmae_b200.models.lightningdit.LightningDiT`` (same keys as the
reference) and reach the library through the usual weight upload.
"""
from __future__ import annotations

import os
from glob import glob

import torch


def strip_module_prefix(state_dict):
    """train_accum.py:98 / :345 -- ``k.replace('module.', '')`` on every key (DDP-wrapped models)."""
    return {k.replace("module.", ""): v for k, v in state_dict.items()}


def load_weights_with_shape_check(model, checkpoint, rank=0, verbose=True):
    """train_accum.py:308-334: copy every ``checkpoint['model']`` tensor whose name and shape match; ``x_embedder.proj.weight``
    with a different channel count becomes zeros with the first 16 input channels copied (``weight[:, :16] = param[:, :16]``);
    other mismatches and unknown names are skipped with a message on rank 0.  Returns the model (as the reference does)."""
    model_state_dict = model.state_dict()
    skipped = []
    for name, param in checkpoint["model"].items():
        if name in model_state_dict:
            if param.shape == model_state_dict[name].shape:
                model_state_dict[name].copy_(param)
            elif name == "x_embedder.proj.weight":
                weight = torch.zeros_like(model_state_dict[name])
                weight[:, :16] = param[:, :16]
                model_state_dict[name] = weight
            else:
                skipped.append(name)
                if rank == 0 and verbose:
                    print(f"Skipping loading parameter '{name}' due to shape mismatch: "
                          f"checkpoint shape {param.shape}, model shape {model_state_dict[name].shape}")
        else:
            skipped.append(name)
            if rank == 0 and verbose:
                print(f"Parameter '{name}' not found in model, skipping.")
    model.load_state_dict(model_state_dict, strict=False)
    model._ldmae_skipped_keys = skipped
    return model


def init_from_pretrained(model, ema, path, rank=0, map_location="cpu"):
    """train_accum.py:94-102 (``train.weight_init``): both the model and its EMA copy start from ``checkpoint['model']``."""
    checkpoint = torch.load(path, map_location=map_location, weights_only=False)
    checkpoint["model"] = strip_module_prefix(checkpoint["model"])
    load_weights_with_shape_check(model, checkpoint, rank=rank)
    if ema is not None:
        load_weights_with_shape_check(ema, checkpoint, rank=rank)
    return model, ema


def checkpoint_path(checkpoint_dir, train_steps):
    """train_accum.py:281."""
    return f"{checkpoint_dir}/{train_steps:07d}.pt"


def save_checkpoint(checkpoint_dir, train_steps, model_state, ema_state, opt_state, config):
    """train_accum.py:273-284: the four-key layout, file name = zero-padded step count."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    path = checkpoint_path(checkpoint_dir, train_steps)
    to_cpu = lambda sd: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}
    torch.save({"model": to_cpu(model_state), "ema": to_cpu(ema_state), "opt": opt_state, "config": config}, path)
    return path


def latest_checkpoint(checkpoint_dir):
    """train_accum.py:172-178: the reference sorts ``*.pt`` by FILE SIZE and takes the last one (checkpoints of one run have
    equal sizes, so this is the glob order of the largest); the step count is parsed from the file name (:182).  Files are
    additionally ordered by name first, so that equal sizes resolve to the highest step.  Returns (path, train_steps) or
    (None, 0)."""
    files = sorted(glob(f"{checkpoint_dir}/*.pt"))
    if not files:
        return None, 0
    files.sort(key=lambda p: os.path.getsize(p))            # stable: equal sizes keep the name order
    path = files[-1]
    return path, int(path.split("/")[-1].split(".")[0])


def resume(checkpoint_dir, model, ema, opt=None, map_location="cpu", load_opt=False):
    """train_accum.py:170-187: load 'model' and 'ema' (strict) from the latest checkpoint; the reference leaves the optimizer
    state untouched (its ``opt.load_state_dict`` line is commented out, :180) -- ``load_opt=True`` restores it as well.
    ``model`` may be DDP-wrapped or plain: the ``module.`` prefix is reconciled either way.  Returns train_steps (0 when
    there is nothing to resume from)."""
    path, steps = latest_checkpoint(checkpoint_dir)
    if path is None:
        return 0
    ckpt = torch.load(path, map_location=map_location, weights_only=False)

    def fit(sd, target):
        wants_prefix = any(k.startswith("module.") for k in target.state_dict().keys())
        sd = strip_module_prefix(sd)
        return {("module." + k) if wants_prefix else k: v for k, v in sd.items()}

    model.load_state_dict(fit(ckpt["model"], model))
    if ema is not None:
        ema.load_state_dict(fit(ckpt["ema"], ema))
    if load_opt and opt is not None and ckpt.get("opt") is not None:
        opt.load_state_dict(ckpt["opt"])
    return steps


def load_for_sampling(model, path, map_location="cpu"):
    """inference.py:100-103: ``checkpoint['ema']`` when the file is a training checkpoint, else the bare state_dict."""
    checkpoint = torch.load(path, map_location=map_location, weights_only=False)
    if "ema" in checkpoint:
        checkpoint = checkpoint["ema"]
    model.load_state_dict(strip_module_prefix(checkpoint))
    return model
