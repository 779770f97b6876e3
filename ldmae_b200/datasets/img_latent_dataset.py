"""Drop-in for the reference's ``datasets/img_latent_dataset.py`` (the trainer's input format).

On-disk format (written by the reference's extract_features.py:160-181 and by ``write_latent_shard`` here): safetensors
shards ``latents_rankRR_shardSSS.safetensors`` holding ``latents`` and ``latents_flip`` -- the VMAE posterior moments
``[N, 2*C, g, g]`` of each image and of its horizontal flip (``MaskedAutoencoderViT._encode``) -- and ``labels`` ``[N]``,
plus a cached ``latents_stats.pt`` (``{'mean','std'}`` of shape ``[1, C, 1, 1]``).  ``__getitem__`` follows reference
lines 76-94: pick flip / no-flip with probability 1/2, slice one row out of the shard, optionally sample the posterior,
normalise per channel and scale.  This is host I/O (DataLoader workers); the GPU path starts at the collated batch.
"""
from __future__ import annotations

import os
from glob import glob

import numpy as np
import torch
from torch.utils.data import Dataset

from ..tokenizer.models_mae import DiagonalGaussianDistribution


def write_latent_shard(output_dir, rank, shard, latents, latents_flip, labels):
    """One shard in the layout of extract_features.py:168-181 (tensors are moved to the CPU and made contiguous)."""
    from safetensors.torch import save_file
    os.makedirs(output_dir, exist_ok=True)
    d = {"latents": latents, "latents_flip": latents_flip, "labels": labels}
    d = {k: v.detach().contiguous().cpu() for k, v in d.items()}
    path = os.path.join(output_dir, f"latents_rank{rank:02d}_shard{shard:03d}.safetensors")
    save_file(d, path, metadata={"total_size": f"{latents.shape[0]}", "dtype": f"{latents.dtype}", "device": f"{latents.device}"})
    return path


class ImgLatentDataset(Dataset):
    def __init__(self, data_dir, latent_norm=True, latent_multiplier=1.0, sample=False):
        self.data_dir = data_dir
        self.latent_norm = latent_norm
        self.latent_multiplier = latent_multiplier
        self.sample = sample
        self.files = sorted(glob(os.path.join(data_dir, "*.safetensors")))
        self.img_to_file_map = self.get_img_to_safefile_map()
        if latent_norm:
            self._latent_mean, self._latent_std = self.get_latent_stats()

    def get_img_to_safefile_map(self):
        from safetensors import safe_open
        img_to_file = {}
        for safe_file in self.files:
            with safe_open(safe_file, framework="pt", device="cpu") as f:
                num_imgs = f.get_slice("labels").get_shape()[0]
            base = len(img_to_file)
            for i in range(num_imgs):
                img_to_file[base + i] = {"safe_file": safe_file, "idx_in_file": i}
        return img_to_file

    def get_latent_stats(self):
        cache = os.path.join(self.data_dir, "latents_stats.pt")
        if not os.path.exists(cache):
            stats = self.compute_latent_stats()
            torch.save(stats, cache)
        else:
            stats = torch.load(cache)
        return stats["mean"], stats["std"]

    def compute_latent_stats(self):
        from safetensors import safe_open
        n = min(10000, len(self.img_to_file_map))
        picks = np.random.choice(len(self.img_to_file_map), n, replace=False)
        rows = []
        for idx in picks:
            info = self.img_to_file_map[int(idx)]
            with safe_open(info["safe_file"], framework="pt", device="cpu") as f:
                feature = f.get_slice("latents")[info["idx_in_file"]:info["idx_in_file"] + 1]
            if self.sample:
                feature = DiagonalGaussianDistribution(feature).sample()
            rows.append(feature)
        lat = torch.cat(rows, dim=0)
        return {"mean": lat.mean(dim=[0, 2, 3], keepdim=True), "std": lat.std(dim=[0, 2, 3], keepdim=True)}

    def __len__(self):
        return len(self.img_to_file_map)

    def __getitem__(self, idx):
        from safetensors import safe_open
        info = self.img_to_file_map[idx]
        i = info["idx_in_file"]
        with safe_open(info["safe_file"], framework="pt", device="cpu") as f:
            key = "latents" if np.random.uniform(0, 1) > 0.5 else "latents_flip"
            feature = f.get_slice(key)[i:i + 1]
            label = f.get_slice("labels")[i:i + 1]
        if self.sample:
            feature = DiagonalGaussianDistribution(feature).sample()
        if self.latent_norm:
            feature = (feature - self._latent_mean) / self._latent_std
        feature = feature * self.latent_multiplier
        return feature.squeeze(0), label.squeeze(0)
