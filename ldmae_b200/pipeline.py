"""Batch sampling job: the body of the reference's ``do_sample`` loop (LDMAE/inference.py:264-292) as one call.

    z, y (host or device)  ->  CFG doubling with the null class  ->  transport ODE (``Sampler.sample_ode``)
    ->  keep the conditional half  ->  latent de-normalisation  ->  VMAE ``decode_to_images``  ->  uint8 NHWC

It only composes the reference-shaped public API of this package (``LightningDiT.forward_with_cfg``,
``create_transport`` / ``Sampler.sample_ode``, ``MaskedAutoencoderViT.decode_to_images``); every FLOP runs in
libldmae_b200.so.  ``bench.py`` times this call (device-resident inputs for ``value``; pinned host inputs and
the uint8 image read-back for ``e2e``); multi-GPU sampling shards the (z, y) batch by rank exactly like
inference.py:87,266-274 -- one process per GPU, no collective inside the loop.
"""
from __future__ import annotations

import torch

from .transport import Sampler, create_transport

__all__ = ["SamplingJob", "FeatureExtractionJob", "build_sampling_models", "rank_seed", "shard_plan", "output_indices"]


# ----------------------------------------------------------------------------- multi-GPU sampling bookkeeping (no collective)
def rank_seed(global_seed, world, rank):
    """inference.py:87: every process seeds its own noise / label stream."""
    return int(global_seed) * int(world) + int(rank)


def shard_plan(num_samples, per_proc_batch, world):
    """inference.py:194-205: a bit more than ``num_samples`` so that it divides evenly; returns (total_samples, iterations per
    rank).  The batch is sharded by rank with no data-path collective: rank r produces images r, r + W, r + 2W, ... of every
    global batch."""
    import math
    global_batch = int(per_proc_batch) * int(world)
    total = int(math.ceil(num_samples / global_batch) * global_batch)
    per_rank = total // world
    assert total % world == 0 and per_rank % per_proc_batch == 0
    return total, per_rank // per_proc_batch


def output_indices(n, rank, world, total_so_far):
    """inference.py:295-296: file index of the i-th image of this rank's batch = i * W + rank + total."""
    return [i * world + rank + total_so_far for i in range(n)]


class SamplingJob:
    """One rank's sampler + decoder with persistent device / pinned buffers for a fixed batch size."""

    def __init__(self, model, vae, *, num_steps=250, sampling_method="euler", cfg_scale=10.0, cfg_interval_start=0.10,
                 timestep_shift=0.3, latent_mean=None, latent_std=None, latent_multiplier=1.0, device=None,
                 cond_only_when_unguided=False):
        self.model, self.vae = model, vae
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.cfg_scale = float(cfg_scale)
        self.use_cfg = self.cfg_scale > 1.0                      # inference.py:173,278-285
        self.cfg_interval_start = cfg_interval_start
        self.null_class = model.y_embedder.num_classes           # inference.py:279 (hard-coded 1000 there)
        if self.use_cfg and model.y_embedder.embedding_table.weight.shape[0] <= self.null_class:
            # class_dropout_prob == 0 builds the table without the null row (lightningdit.py:146-148); the reference would
            # hit nn.Embedding's device assert on label num_classes -- fail here, before any kernel reads out of bounds
            raise ValueError(f"classifier-free guidance (cfg_scale {self.cfg_scale}) needs the null-class row {self.null_class} "
                             f"of y_embedder.embedding_table, but the model was built with class_dropout_prob=0 "
                             f"({model.y_embedder.embedding_table.weight.shape[0]} rows)")
        transport = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
        self.sample_fn = Sampler(transport).sample_ode(sampling_method=sampling_method, num_steps=num_steps, atol=1e-6,
                                                       rtol=1e-3, reverse=False, timestep_shift=timestep_shift,
                                                       cond_only_when_unguided=cond_only_when_unguided)
        grid = self.sample_fn.t
        self.cond_only_when_unguided = bool(cond_only_when_unguided and self.use_cfg and cfg_interval_start is not None)
        # sample-forwards per image actually executed (for FLOP accounting)
        if self.cond_only_when_unguided:
            below = int((grid[:-1] < float(cfg_interval_start)).sum())
            self.sample_forwards_per_image = 2 * (num_steps - 1 - below) + below
        else:
            self.sample_forwards_per_image = (2 if self.use_cfg else 1) * (num_steps - 1)
        C = model.in_channels
        self.latent_mean = latent_mean if latent_mean is not None else torch.zeros(1, C, 1, 1)
        self.latent_std = latent_std if latent_std is not None else torch.ones(1, C, 1, 1)
        self.latent_mean = self.latent_mean.to(self.device)
        self.latent_std = self.latent_std.to(self.device)
        self.latent_multiplier = float(latent_multiplier)
        self.model_evals = num_steps - 1                         # N grid points = N-1 model evaluations (integrators.py:94)

    # -- device-resident path ------------------------------------------------------------------
    def sample_latents(self, z, y):
        """z [n,C,S,S] fp32, y [n] int64 on the device -> final normalised latents [n,C,S,S] (inference.py:278-289)."""
        n = z.shape[0]
        if self.use_cfg:
            zz = torch.cat([z, z], 0)
            yy = torch.cat([y, torch.full((n,), self.null_class, device=y.device, dtype=y.dtype)], 0)
            kw = dict(y=yy, cfg_scale=self.cfg_scale)
            if self.cfg_interval_start is not None:
                kw.update(cfg_interval=True, cfg_interval_start=self.cfg_interval_start)
            out = self.sample_fn(zz, self.model.forward_with_cfg, **kw)[-1]
            return out.chunk(2, dim=0)[0]
        return self.sample_fn(z, self.model.forward, y=y)[-1]

    def decode_u8(self, latents):
        """normalised latents -> uint8 [n,H,W,3] on the device (inference.py:291-292 without the host copy)."""
        _, u8 = self.vae._decode(latents, False, True, self.latent_mean, self.latent_std, self.latent_multiplier)
        return u8

    def run_device(self, z, y):
        return self.decode_u8(self.sample_latents(z, y))

    def sample_shard(self, *, rank, world, global_seed, num_samples, per_proc_batch, num_classes, on_images, latent_size=None,
                     done_samples=0):
        """This rank's part of a ``num_samples`` job, as the reference's loop does it (inference.py:87,194-205,264-298): seed
        ``global_seed * W + rank``, ``iterations`` batches of ``per_proc_batch`` images with z and y drawn ON THE DEVICE from
        the seeded default generator, ``on_images(indices, uint8[n,H,W,3] numpy)`` per batch with the reference's file indices
        ``i * W + rank + total``.  ``done_samples`` (images already on disk) skips finished iterations like the reference's
        resume count (:204).  No collective: ranks never exchange data.  Returns the number of images produced."""
        total_samples, iterations = shard_plan(num_samples, per_proc_batch, world)
        torch.manual_seed(rank_seed(global_seed, world, rank))
        S = latent_size if latent_size is not None else self.model.input_size
        C, n = self.model.in_channels, per_proc_batch
        done_iterations = (done_samples // world) // n
        total, made = 0, 0
        for it in range(iterations):
            z = torch.randn(n, C, S, S, device=self.device)
            y = torch.randint(0, num_classes, (n,), device=self.device)
            if it >= done_iterations:                         # the draws above keep the RNG stream aligned with a fresh run
                u8 = self.run_device(z, y)
                on_images(output_indices(n, rank, world, total), u8.cpu().numpy())
                made += n
            total += n * world
        return made

    # -- host-to-host path (what inference.py does per iteration) --------------------------------
    def run_host(self, z_host, y_host, out_host=None, events=None):
        """Pinned host z / y in, uint8 images out on the host (pinned ``out_host`` is reused when given).

        ``events``: optional 4 ``torch.cuda.Event(enable_timing=True)`` recorded on the current stream -- before the
        host-to-device copies, after them (= start of the device-resident job), after the decode (= its end) and after
        the device-to-host copy of the images; bench.py derives ``value`` (inner pair) and ``e2e`` (outer pair) from the
        same call."""
        if events is not None:
            events[0].record()
        z = z_host.to(self.device, non_blocking=True)
        y = y_host.to(self.device, non_blocking=True)
        if events is not None:
            events[1].record()
        u8 = self.run_device(z, y)
        if events is not None:
            events[2].record()
        if out_host is None:
            out_host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
        out_host.copy_(u8, non_blocking=True)
        if events is not None:
            events[3].record()
        torch.cuda.current_stream(self.device).synchronize()
        return out_host


class FeatureExtractionJob:
    """One rank's body of the reference's feature-extraction loop (LDMAE/extract_features.py:140-188): images in [-1, 1]
    -> VMAE encoder -> posterior moments of every image AND of its horizontal flip -> safetensors shards
    ``latents_rankRR_shardSSS.safetensors`` of ``shard_images // batch_size`` batches each (10000 images in the reference),
    the remainder in a last shard -- the files ``ImgLatentDataset`` / ``FusedTrainer.step_from_moments`` read.

    The reference runs two DataLoaders (RandomHorizontalFlip p=0 and p=1) and two encoder calls per batch; here the flipped
    batch is ``x.flip(-1)`` on the device (the flip commutes with ToTensor / Normalize, and the centre crop precedes it,
    models_mae.py:935-950) and both halves go through ONE encoder call on the doubled batch.  ``sample=True`` stores the
    moments ``tokenizer._encode(x)`` (configs with a ``data.sample`` key, extract_features.py:150-151), ``sample=False`` the
    posterior mode (:153).  ``vae`` is any object with the tokenizer mirror's ``_encode`` / ``encode``."""

    def __init__(self, vae, output_dir, *, rank=0, batch_size=256, shard_images=10000, sample=True, device=None):
        if batch_size <= 0 or shard_images < batch_size:
            raise ValueError(f"shard_images ({shard_images}) must hold at least one batch of {batch_size}")
        self.vae, self.output_dir, self.rank = vae, output_dir, int(rank)
        self.batches_per_shard = shard_images // batch_size            # extract_features.py:164
        self.sample = bool(sample)
        self.device = torch.device(device) if device is not None else None
        self.latents, self.latents_flip, self.labels = [], [], []
        self.saved_files, self.run_images, self.paths = 0, 0, []

    def _encode(self, x):
        if self.sample:
            return self.vae._encode(x)
        return self.vae.encode(x).latent_dist.mode().detach()

    def add_batch(self, x, y, x_flip=None):
        """x [B,3,H,W] in [-1,1] (host or device), y [B] labels; ``x_flip``: the separately loaded flipped batch of the
        reference's second loader (default: flipped on the device).  Returns the shard path when this batch completed one."""
        dev = self.device if self.device is not None else x.device
        x = x.to(dev, non_blocking=True)
        xf = x.flip(-1) if x_flip is None else x_flip.to(dev, non_blocking=True)
        with torch.no_grad():
            z = self._encode(torch.cat([x, xf], 0))
        B = x.shape[0]
        self.latents.append(z[:B]); self.latents_flip.append(z[B:]); self.labels.append(y.detach().cpu())
        self.run_images += B
        if len(self.latents) == self.batches_per_shard:
            return self._flush()
        return None

    def finish(self):
        """Writes the remainder (extract_features.py:189-206); returns the list of all shard paths."""
        if self.latents:
            self._flush()
        return self.paths

    def compute_stats(self):
        """extract_features.py:213-216 (rank 0, after every rank has finished): constructing the dataset computes and caches
        ``latents_stats.pt``.  Returns (mean, std) of shape [1, C', 1, 1]."""
        from .datasets import ImgLatentDataset
        ds = ImgLatentDataset(self.output_dir, latent_norm=True, sample=self.sample)
        return ds._latent_mean, ds._latent_std

    def _flush(self):
        from .datasets.img_latent_dataset import write_latent_shard
        path = write_latent_shard(self.output_dir, self.rank, self.saved_files, torch.cat(self.latents, 0),
                                  torch.cat(self.latents_flip, 0), torch.cat(self.labels, 0))
        self.latents, self.latents_flip, self.labels = [], [], []
        self.saved_files += 1
        self.paths.append(path)
        return path


def build_sampling_models(device, *, model_name="LightningDiT-B/1", input_size=32, in_channels=16, img_size=256, seed=0,
                          num_classes=1000):
    """Random-init LightningDiT + VMAE decoder of the shipped recipe (configs/imagenet/lightningdit_b_vmae_f8d16_cfg.yaml).
    The reference zero-initialises final_layer.linear and every adaLN_modulation[-1] (lightningdit.py:365-374), which
    makes a freshly constructed model output exactly 0; those tensors are re-drawn N(0, 0.02) so the synthetic
    benchmark exercises non-trivial numerics."""
    from .models.lightningdit import LightningDiT_models
    from .tokenizer import models_mae
    torch.manual_seed(seed)
    model = LightningDiT_models[model_name](input_size=input_size, num_classes=num_classes, use_qknorm=True, use_swiglu=True,
                                            use_rope=True, use_rmsnorm=True, wo_shift=False, in_channels=in_channels)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "adaLN_modulation.1" in name or name.startswith("final_layer.linear"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    vae = models_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True,
                                              img_size=img_size)
    return model.to(device).eval(), vae.to(device).eval()
