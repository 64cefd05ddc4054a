"""GPU-side mirror of the reference's NoisyDataLoader.__iter__ (src/dataloader.py:88-287).

The reference augments every batch on the host: Gaussian noise on magnitudes / spectra scaled by their errors, uniform noise
on the host-galaxy images scaled by the batch's standard deviation, and a per-image Python loop of RandomRotation by k*90
degrees.  Here the raw batch is uploaded once (images may stay 8-bit: the PNGs are 8-bit and the reference's fp32/255 copy is 4x
the bytes) and the three operations run as single-pass kernels; the rotation is an index permutation.

    aug = DeviceAugment(combinations, max_noise_intensity=0.1, noise_level_mag=1.0)
    batch9 = aug(raw_batch)            # raw_batch: what the reference's TensorDataset yields for `combinations`, on the GPU
    loss = step(batch9)                # same 9-tuple LightCurveImageCLIP.training_step takes

Random numbers: drawn in-kernel (counter-based, seeded per call) unless given explicitly (`noise=...`), in which case the
result is bit-identical to the reference expressions -- that is what the parity tests check.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import ops
from ._lib import check, lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def augment_seq(x: torch.Tensor, err: torch.Tensor, level: float, noise: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """x + randn_like(x) * err * level  (src/dataloader.py:125,136)."""
    x = ops._req(x, "x"); err = ops._req(err, "err")
    if x.shape != err.shape:
        raise ValueError(f"augment_seq: x {tuple(x.shape)} and err {tuple(err.shape)} differ")
    if noise is not None:
        noise = ops._req(noise, "noise")
    out = torch.empty_like(x)
    check(lib().mvn_augment_seq(_p(x), _p(err), _p(noise), float(level), x.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(out), ops._stream()), "augment_seq")
    ops._count(1)
    return out


def image_noise_range(imgs: torch.Tensor, max_noise_intensity: float) -> torch.Tensor:
    """max_noise_intensity * torch.std(host_imgs) as a 1-element device tensor (no host read); imgs fp32 or uint8."""
    L = lib()
    if not imgs.is_cuda:
        raise RuntimeError("maven_b200: images are on the CPU; CUDA only (no CPU fallback)")
    if imgs.dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"maven_b200: images must be float32 or uint8, got {imgs.dtype}")
    imgs = imgs.contiguous()
    wsb = L.mvn_image_noise_range_workspace_bytes()
    ws = torch.empty(wsb, dtype=torch.uint8, device=imgs.device)
    rng = torch.empty(1, dtype=torch.float32, device=imgs.device)
    check(L.mvn_image_noise_range(_p(imgs), 1 if imgs.dtype == torch.uint8 else 0, imgs.numel(), float(max_noise_intensity), _p(rng), _p(ws), wsb,
                                  ops._stream()), "image_noise_range")
    ops._count(2)
    return rng


def augment_images(imgs: torch.Tensor, noise_range: Optional[torch.Tensor], rot_k: Optional[torch.Tensor] = None,
                   noise_u: Optional[torch.Tensor] = None, seed: int = 0, layout: str = "bchw",
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rot90^k(imgs + (2*rand - 1) * noise_range) -> fp32 (B,C,H,W)  (src/dataloader.py:93-112).  imgs: fp32 (B,C,H,W), uint8
    (B,C,H,W) or, with layout='bhwc', the decoder's uint8 (B,H,W,C); rot_k int32 (B,) in 0..3 (counter-clockwise quarter turns).
    noise_range None: no noise (a pure uint8 -> float32/255 conversion / rotation pass); out: write into an existing buffer, e.g.
    the static image input of a GraphedTrainStep."""
    if not imgs.is_cuda:
        raise RuntimeError("maven_b200: images are on the CPU; CUDA only (no CPU fallback)")
    imgs = imgs.contiguous()
    if imgs.dim() != 4:
        raise ValueError(f"augment_images: expected 4 dimensions, got {tuple(imgs.shape)}")
    if layout == "bhwc":
        if imgs.dtype != torch.uint8:
            raise TypeError("augment_images: layout='bhwc' is the raw 8-bit decoder layout")
        B, H, W, C = imgs.shape
        mode = 2
    else:
        B, C, H, W = imgs.shape
        if imgs.dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"maven_b200: images must be float32 or uint8, got {imgs.dtype}")
        mode = 1 if imgs.dtype == torch.uint8 else 0
    if rot_k is not None:
        rot_k = rot_k.to(device=imgs.device, dtype=torch.int32).contiguous()
    if noise_u is not None:
        noise_u = ops._req(noise_u, "noise_u")
    if out is None:
        out = torch.empty(B, C, H, W, dtype=torch.float32, device=imgs.device)
    elif tuple(out.shape) != (B, C, H, W) or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
        raise ValueError(f"augment_images: `out` must be a contiguous float32 CUDA tensor of shape {(B, C, H, W)}")
    check(lib().mvn_augment_images(_p(imgs), mode, _p(noise_u), _p(rot_k), _p(noise_range), int(seed) & 0xFFFFFFFFFFFFFFFF, B, C, H, W, _p(out),
                                   ops._stream()), "augment_images")
    ops._count(1)
    return out


# field order of the reference's TensorDataset per combination set (src/dataloader.py:90-287)
_FIELDS = {
    frozenset(["host_galaxy"]): ("img", "redshift", "classification"),
    frozenset(["lightcurve"]): ("mag", "time", "mask", "magerr", "redshift", "classification"),
    frozenset(["spectral"]): ("spec", "freq", "maskspec", "specerr", "redshift", "classification"),
    frozenset(["host_galaxy", "lightcurve"]): ("img", "mag", "time", "mask", "magerr", "redshift", "classification"),
    frozenset(["host_galaxy", "spectral"]): ("img", "spec", "freq", "maskspec", "specerr", "redshift", "classification"),
    frozenset(["spectral", "lightcurve"]): ("mag", "time", "mask", "magerr", "spec", "freq", "maskspec", "specerr", "redshift", "classification"),
    frozenset(["host_galaxy", "spectral", "lightcurve"]): ("img", "mag", "time", "mask", "magerr", "spec", "freq", "maskspec", "specerr",
                                                             "redshift", "classification"),
}


class DeviceAugment:
    """NoisyDataLoader.__iter__'s per-batch work on the GPU: raw dataset tuple in, the 9-tuple of training_step out."""

    def __init__(self, combinations: Sequence[str], max_noise_intensity: float, noise_level_mag: float, seed: int = 0):
        key = frozenset(c for c in combinations if c != "meta")     # the reference strips 'meta' before dispatch (src/dataloader.py:72-75)
        if key not in _FIELDS:
            raise ValueError(f"DeviceAugment: unsupported combination set {sorted(key)}")
        self.fields = _FIELDS[key]
        self.max_noise_intensity, self.noise_level_mag = float(max_noise_intensity), float(noise_level_mag)
        self.seed, self.calls = int(seed), 0

    def __call__(self, batch: Sequence[torch.Tensor], noise: Optional[dict] = None):
        """noise (optional, for reproducibility / parity): {'img_u': U[0,1) like the images, 'rot_k': int (B,), 'mag': N(0,1)
        like mag, 'spec': N(0,1) like spec}; whatever is missing is drawn on the device."""
        if len(batch) != len(self.fields):
            raise ValueError(f"DeviceAugment: expected {len(self.fields)} tensors {self.fields}, got {len(batch)}")
        f = dict(zip(self.fields, batch))
        noise = noise or {}
        self.calls += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.calls * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        img = mag = spec = None
        if "img" in f:
            x = f["img"]
            rot = noise.get("rot_k")
            if rot is None:
                rot = torch.randint(0, 4, (x.shape[0],), device=x.device, dtype=torch.int32)
            img = augment_images(x, image_noise_range(x, self.max_noise_intensity), rot, noise.get("img_u"), seed)
        if "mag" in f:
            mag = augment_seq(f["mag"], f["magerr"], self.noise_level_mag, noise.get("mag"), seed ^ 0x5851F42D4C957F2D)
        if "spec" in f:
            spec = augment_seq(f["spec"], f["specerr"], self.noise_level_mag, noise.get("spec"), seed ^ 0x14057B7EF767814F)
        return (img, mag, f.get("time"), f.get("mask"), spec, f.get("freq"), f.get("maskspec"), f["redshift"], f["classification"])
