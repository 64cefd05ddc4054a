"""Golden vectors of the reference's NoisyDataLoader.__iter__ (src/dataloader.py:88-287), produced by running the UNMODIFIED
reference class here (build container only; needs /root/reference).  Modules the reference imports at file scope but that
the augmentation path never touches (h5py, astropy, extinction, matplotlib, seaborn, ...) are stubbed.

    python tests/golden/make_golden_noisy.py   ->   tests/golden/noisy_loader.npz

Stored per case: the raw dataset tensors, the augmented 9-tuple the reference yielded, and the random tensors it consumed --
recovered by replaying torch's global generator from the same seed in the order the reference draws them (rand_like on the
images, randn_like on mag, randn_like on spec, randint for the rotations) and verified against the outputs before saving.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MAVEN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    for name in ("h5py", "extinction", "seaborn", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = mod("matplotlib"); mod("matplotlib.pyplot"); mod("matplotlib.ticker", MaxNLocator=object)
        mp.pyplot = sys.modules["matplotlib.pyplot"]; mp.ticker = sys.modules["matplotlib.ticker"]
    try:
        import astropy.cosmology  # noqa: F401
    except Exception:
        mod("astropy"); mod("astropy.cosmology", Planck15=object())
    for name, attrs in (("pytorch_lightning", dict(LightningModule=torch.nn.Module, Callback=object, Trainer=object)),
                        ("pytorch_lightning.callbacks", dict(Callback=object)), ("ruamel", {}), ("ruamel.yaml", dict(YAML=object)),
                        ("torchmetrics", {}), ("torchmetrics.classification", dict(MulticlassFBetaScore=object))):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name, **attrs)


def main():
    _stub()
    sys.path.insert(0, REF)
    from src.dataloader import NoisyDataLoader
    from torch.utils.data import TensorDataset
    g = torch.Generator().manual_seed(123)
    B, T, TS = 4, 20, 24
    img = (torch.randint(0, 256, (B, 3, 60, 60), generator=g).float() / 255.0)
    mag = torch.randn(B, T, generator=g); time = torch.rand(B, T, generator=g); mask = torch.rand(B, T, generator=g) > 0.3
    magerr = torch.rand(B, T, generator=g) * 0.2
    spec = torch.randn(B, TS, generator=g); freq = torch.rand(B, TS, generator=g); maskspec = torch.rand(B, TS, generator=g) > 0.2
    specerr = torch.rand(B, TS, generator=g) * 0.1
    red = torch.rand(B, generator=g); cls = torch.randint(0, 5, (B,), generator=g)
    out = {"img": img, "mag": mag, "time": time, "mask": mask, "magerr": magerr, "spec": spec, "freq": freq, "maskspec": maskspec,
           "specerr": specerr, "redshift": red, "classification": cls, "max_noise_intensity": torch.tensor(0.1), "noise_level_mag": torch.tensor(0.7)}
    cases = {"img": (["host_galaxy"], [img, red, cls]),
             "lc": (["lightcurve"], [mag, time, mask, magerr, red, cls]),
             "img_lc": (["host_galaxy", "lightcurve"], [img, mag, time, mask, magerr, red, cls]),
             "lc_sp": (["spectral", "lightcurve"], [mag, time, mask, magerr, spec, freq, maskspec, specerr, red, cls]),
             "all": (["host_galaxy", "spectral", "lightcurve"], [img, mag, time, mask, magerr, spec, freq, maskspec, specerr, red, cls])}
    for name, (comb, tensors) in cases.items():
        loader = NoisyDataLoader(TensorDataset(*tensors), batch_size=B, noise_level_img=0.1, noise_level_mag=0.7, shuffle=False,
                                 combinations=list(comb))
        torch.manual_seed(77)
        batch = next(iter(loader))
        # replay the generator in the reference's draw order
        torch.manual_seed(77)
        torch.empty((), dtype=torch.int64).random_()          # the DataLoader iterator draws its base seed first
        u = torch.rand_like(img) if "host_galaxy" in comb else None
        n_mag = torch.randn_like(mag) if "lightcurve" in comb else None
        n_sp = torch.randn_like(spec) if "spectral" in comb else None
        rot = torch.randint(0, 4, (B,)) if "host_galaxy" in comb else None
        if u is not None:
            noisy = img + (2 * u - 1) * (0.1 * torch.std(img))
            chk = torch.stack([torch.rot90(noisy[i], int(rot[i]), (1, 2)) for i in range(B)])
            assert torch.equal(chk, batch[0]), name
            out[f"{name}.img_u"] = u; out[f"{name}.rot_k"] = rot
        if n_mag is not None:
            assert torch.equal(mag + n_mag * magerr * 0.7, batch[1]), name
            out[f"{name}.noise_mag"] = n_mag
        if n_sp is not None:
            assert torch.equal(spec + n_sp * specerr * 0.7, batch[4]), name
            out[f"{name}.noise_spec"] = n_sp
        for i, t in enumerate(batch):
            if t is not None:
                out[f"{name}.out{i}"] = t
    np.savez_compressed(os.path.join(HERE, "noisy_loader.npz"), **{k: v.numpy() for k, v in out.items()})
    print("wrote noisy_loader.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
