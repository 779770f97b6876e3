"""Drop-in for the reference's ``tokenizer/models_mae.py`` (VMAE f8d16 tokenizer), inference-time surface.

Public surface kept: ``mae_for_ldmae_f8d16_prev(**kw)`` (reference models_mae.py:992-997),
``MaskedAutoencoderViT.decode(z, return_dict)`` (:865-887), ``decode_to_images(z)`` (:963-973),
``unpatchify`` (:458-470), ``_encode(x)`` (:819-836), ``encode(x, return_dict)`` (:838-863), ``encode_images`` (:952-961),
``load_state_dict(ckpt['model'], strict=False)`` with the reference's key names.
Decoder (from_latent, decoder_embed, 12 ViT blocks, decoder_norm, linear_pred + 3x3 RGB conv, uint8 pack) and encoder
(patch embed, 12 ViT blocks, norm, to_latent -> posterior moments) run in libldmae_b200.so; the masked pre-training
paths (models_mae.py:472-815) are out of scope (tokenizer training).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import partial
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib


class Config:
    def __init__(self, scaling_factor):
        self.scaling_factor = scaling_factor


@dataclass
class DecoderOutput:
    sample: torch.Tensor
    commit_loss: Optional[torch.Tensor] = None


class DiagonalGaussianDistribution:
    """reference tokenizer/util/misc.py:74-128 (posterior over latents from the encoder's moments)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        dims = list(range(1, self.mean.dim()))
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=dims)
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar
                               + other.logvar, dim=dims)

    def nll(self, sample, dims=(1, 2, 3)):
        if self.deterministic:
            return torch.Tensor([0.0])
        return 0.5 * torch.sum(np.log(2.0 * np.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=list(dims))

    def mode(self):
        return self.mean


@dataclass
class EncoderOutput:
    latent: torch.Tensor

    def sample(self):
        return self.latent


@dataclass
class MAEOutput:
    latent_dist: object


class _PatchEmbedParams(nn.Module):
    """timm PatchEmbed as the reference uses it (models_mae.py:330): .proj = Conv2d(k = stride = patch)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)


def _sincos_2d_f32(embed_dim, grid_size):
    """reference tokenizer/util/pos_embed.py:20-67 (float32 omega)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])

    def _1d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    return np.concatenate([_1d(embed_dim // 2, grid[0]), _1d(embed_dim // 2, grid[1])], axis=1)


class _AttnParams(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _BlockParams(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _AttnParams(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = _MlpParams(dim, int(dim * mlp_ratio))


class _ConvDecoderPredParams(nn.Module):
    """reference conv_decoder_pred (models_mae.py:244-255), pred_with_conv=False branch."""

    def __init__(self, decoder_embed_dim, patch_size, in_chans):
        super().__init__()
        self.p = patch_size
        self.linear_pred = nn.Linear(decoder_embed_dim, patch_size ** 2 * in_chans, bias=True)
        self.conv_smoother = nn.Conv2d(in_chans, in_chans, 3, 1, 1)


class MaskedAutoencoderViT(nn.Module):
    """Decoder half of the reference class (models_mae.py:283-887) under the same parameter names."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4.0, norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, latent_dim=32, ldmae_mode=False, scaling_factor=0.9654248952865601, no_cls=True,
                 gradual_resol=False, finetune_downsample_layer=None, down_nonlinear=False, kl_loss_weight=None,
                 smooth_output=False, pred_with_conv=False, perceptual_loss=None):
        super().__init__()
        if not (ldmae_mode and no_cls and smooth_output) or gradual_resol or down_nonlinear or pred_with_conv:
            raise NotImplementedError("ldmae_b200 builds the tokenizer as inference.py:133 constructs it: ldmae_mode=True, "
                                      "no_cls=True, smooth_output=True (linear_pred + RGB conv), no gradual_resol")
        if in_chans != 3:
            raise NotImplementedError("in_chans must be 3")
        self.config = Config(scaling_factor=scaling_factor)
        self.img_size, self.patch_size, self.latent_dim = img_size, patch_size, latent_dim
        self.latent_resolution = img_size // patch_size
        self.embed_dim, self.decoder_embed_dim = embed_dim, decoder_embed_dim
        self.decoder_depth, self.decoder_num_heads, self.mlp_ratio = decoder_depth, decoder_num_heads, mlp_ratio
        self.kl_loss_weight = kl_loss_weight
        self.no_cls, self.ldmae_mode, self.smooth_output = no_cls, ldmae_mode, smooth_output
        ln = norm_layer(decoder_embed_dim)
        self.ln_eps = float(getattr(ln, "eps", 1e-5))
        num_patches = self.latent_resolution ** 2
        # encoder (models_mae.py:330-352): parameters under the reference's names, arithmetic in the library
        self.depth, self.num_heads = depth, num_heads
        self.patch_embed = _PatchEmbedParams(img_size, patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([_BlockParams(embed_dim, num_heads, mlp_ratio, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.to_latent = nn.Linear(embed_dim, latent_dim * 2 if kl_loss_weight is not None else latent_dim)
        self.from_latent = nn.Linear(latent_dim, decoder_embed_dim)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([_BlockParams(decoder_embed_dim, decoder_num_heads, mlp_ratio, norm_layer)
                                             for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = _ConvDecoderPredParams(decoder_embed_dim, patch_size, in_chans)
        self.initialize_weights()
        self._handle = None
        self._handle_sig = None
        self._handle_dev = None

    def initialize_weights(self):
        """reference models_mae.py:398-435 (decoder side)."""
        pe = _sincos_2d_f32(self.decoder_pos_embed.shape[-1], self.latent_resolution)
        self.decoder_pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        pe = _sincos_2d_f32(self.pos_embed.shape[-1], self.latent_resolution)
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))

        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        self.apply(_init)
        w = self.patch_embed.proj.weight.data                       # models_mae.py:420-421
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))

    def unpatchify(self, x):
        """reference models_mae.py:458-470 (index-only)."""
        p = self.patch_size
        h = w = int(x.shape[1] ** 0.5)
        assert h * w == x.shape[1]
        x = x.reshape(x.shape[0], h, w, p, p, 3)
        x = torch.einsum("nhwpqc->nchpwq", x)
        return x.reshape(x.shape[0], 3, h * p, h * p)

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    @property
    def dtype(self):
        return torch.float32

    # -- C handle -------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ldmae_vmae_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _ensure_handle(self, device, batch):
        if device.type != "cuda":
            raise _lib.LdmaeError("ldmae_b200 VMAE decoder runs on a CUDA (B200) device only; there is no CPU path")
        L = _lib.lib()
        sd = self.state_dict()
        if self._handle is None or self._handle_dev != device:
            if self._handle:
                L.ldmae_vmae_destroy(self._handle)
            cfg = _lib.VmaeConfig(img_size=self.img_size, patch_size=self.patch_size, latent_dim=self.latent_dim,
                                  embed_dim=self.embed_dim, decoder_embed_dim=self.decoder_embed_dim,
                                  decoder_depth=self.decoder_depth, decoder_num_heads=self.decoder_num_heads,
                                  mlp_hidden=int(self.decoder_embed_dim * self.mlp_ratio), ln_eps=self.ln_eps,
                                  max_batch=max(1, batch), depth=self.depth, num_heads=self.num_heads,
                                  to_latent_dim=self.to_latent.weight.shape[0])
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(L.ldmae_vmae_create(C.byref(cfg), C.byref(h)), "ldmae_vmae_create")
            self._handle, self._handle_dev, self._handle_sig = h, device, None
        sig = tuple((t.data_ptr(), t._version) for t in sd.values())
        if sig != self._handle_sig:
            with torch.cuda.device(device):          # pack kernels must run on the handle's device and its current stream
                st = _lib.stream_ptr()
                for k, v in sd.items():
                    t = v.detach()
                    if t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.float().contiguous()
                    if t.device != device:
                        raise _lib.LdmaeError(f"parameter {k} is on {t.device}, expected {device}")
                    _lib.check(L.ldmae_vmae_load_tensor(self._handle, k.encode(), _lib.ptr(t), t.numel(), st), f"load {k}")
                _lib.check(L.ldmae_vmae_finalize(self._handle, st), "ldmae_vmae_finalize")
                torch.cuda.current_stream().synchronize()
            self._handle_sig = sig
        return self._handle

    def mark_weights_dirty(self):
        """Re-upload the weights on the next call (needed after in-place edits through ``p.data`` or raw pointers, which
        PyTorch's version counters do not see)."""
        self._handle_sig = None

    def _decode(self, z, want_f32, want_u8, mean=None, std=None, multiplier=1.0):
        z = z.detach().float().contiguous()
        B = z.shape[0]
        h = self._ensure_handle(z.device, B)
        f32 = torch.empty(B, 3, self.img_size, self.img_size, device=z.device) if want_f32 else None
        u8 = torch.empty(B, self.img_size, self.img_size, 3, device=z.device, dtype=torch.uint8) if want_u8 else None
        m = mean.detach().float().reshape(-1).contiguous() if mean is not None else None
        s = std.detach().float().reshape(-1).contiguous() if std is not None else None
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().ldmae_vmae_decode(h, _lib.ptr(z), _lib.ptr(m), _lib.ptr(s), float(multiplier),
                                                    _lib.ptr(f32), _lib.ptr(u8), B, _lib.stream_ptr()), "ldmae_vmae_decode")
        return f32, u8

    # -- reference API ----------------------------------------------------------------------------
    def decode(self, z, return_dict=True, generator=None):
        img, _ = self._decode(z, True, False)
        return DecoderOutput(sample=img) if return_dict else (img,)

    def decode_to_images(self, z, latent_mean=None, latent_std=None, latent_multiplier=1.0):
        """reference models_mae.py:963-973; the optional statistics fuse inference.py:291's de-normalisation
        (z*std/multiplier + mean) into the decoder's first kernel."""
        with torch.no_grad():
            _, u8 = self._decode(z.cuda(), False, True, latent_mean, latent_std, latent_multiplier)
            return u8.cpu().numpy()

    def _encode(self, x):
        """reference models_mae.py:819-836: images [B,3,H,W] in [-1,1] -> moments [B, 2*latent, H/p, W/p]."""
        x = x.detach().float().contiguous()
        B = x.shape[0]
        if x.shape[1:] != (3, self.img_size, self.img_size):
            raise ValueError(f"expected images of shape [B, 3, {self.img_size}, {self.img_size}], got {tuple(x.shape)}")
        h = self._ensure_handle(x.device, B)
        out = torch.empty(B, self.to_latent.weight.shape[0], self.latent_resolution, self.latent_resolution, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ldmae_vmae_encode(h, _lib.ptr(x), _lib.ptr(out), B, _lib.stream_ptr()), "ldmae_vmae_encode")
        return out

    def encode(self, x, return_dict=True):
        """reference models_mae.py:838-863."""
        m = self._encode(x)
        p = DiagonalGaussianDistribution(m) if self.kl_loss_weight is not None else EncoderOutput(m)
        return MAEOutput(latent_dist=p) if return_dict else (p,)

    def encode_images(self, images):
        """reference models_mae.py:952-961."""
        with torch.no_grad():
            return self.encode(images.cuda(), return_dict=False)[0].sample()

    def img_transform(self, p_hflip=0, img_size=None):
        """reference models_mae.py:935-950 (torchvision pipeline: centre crop, flip, ToTensor, Normalize(0.5, 0.5))."""
        from torchvision import transforms
        img_size = img_size if img_size is not None else self.img_size
        return transforms.Compose([transforms.Lambda(lambda im: center_crop_arr(im, img_size)),
                                   transforms.RandomHorizontalFlip(p=p_hflip), transforms.ToTensor(),
                                   transforms.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5], inplace=True)])


def center_crop_arr(pil_image, image_size):
    """reference models_mae.py:85-103 (ADM centre crop)."""
    from PIL import Image
    while min(*pil_image.size) >= 2 * image_size:
        pil_image = pil_image.resize(tuple(x // 2 for x in pil_image.size), resample=Image.BOX)
    scale = image_size / min(*pil_image.size)
    pil_image = pil_image.resize(tuple(round(x * scale) for x in pil_image.size), resample=Image.BICUBIC)
    arr = np.array(pil_image)
    crop_y = (arr.shape[0] - image_size) // 2
    crop_x = (arr.shape[1] - image_size) // 2
    return Image.fromarray(arr[crop_y: crop_y + image_size, crop_x: crop_x + image_size])


def mae_for_ldmae_f8d16_prev(**kwargs):
    """reference models_mae.py:992-997."""
    return MaskedAutoencoderViT(patch_size=8, embed_dim=192, depth=12, num_heads=12, decoder_embed_dim=192,
                                decoder_depth=12, decoder_num_heads=12, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), latent_dim=16, **kwargs)
