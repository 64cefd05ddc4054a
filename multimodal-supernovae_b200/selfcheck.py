"""smoke(): one small CLIP training step (light curve + spectra) on cuda:0, checked against the CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synthetic_seq(gen, B, T, nband, tmax, lo, hi, t0=0.0):
    per = T // nband
    mask = torch.zeros(B, T, dtype=torch.bool)
    t = torch.zeros(B, T)
    x = torch.zeros(B, T)
    n = torch.randint(lo, min(hi, per) + 1, (B, nband), generator=gen)
    for b in range(B):
        for k in range(nband):
            nn_ = int(n[b, k])
            tt = torch.sort(torch.rand(nn_, generator=gen) * tmax)[0]
            tt = tt - tt[0] + t0
            sl = slice(k * per, k * per + nn_)
            mask[b, sl] = True
            t[b, sl] = tt
            x[b, sl] = torch.randn(nn_, generator=gen)
    return x, t, mask


def run_smoke():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from maven_b200.models_multimodal import LightCurveImageCLIP
    from oracle import maven_oracle as O
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)
    tk = dict(n_out=32, emb=64, heads=8, depth=2, dropout=0.0, time_norm=20583.37, agg="mean")
    sk = dict(n_out=32, emb=32, heads=2, depth=2, dropout=0.0, time_norm=17945.14, agg="mean")
    model = LightCurveImageCLIP(logit_scale=19.55, nband=2, loss="softmax", transformer_kwargs=tk, transformer_spectral_kwargs=sk,
                                combinations=["lightcurve", "spectral"], lr=1e-3, optimizer_kwargs={"weight_decay": 5.6e-4})
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = 16
    x_lc, t_lc, m_lc = synthetic_seq(gen, B, 40, 2, 300.0, 1, 20)
    x_sp, t_sp, m_sp = synthetic_seq(gen, B, 44, 1, 5500.0, 10, 44, t0=3700.0)
    batch = (None, x_lc, t_lc, m_lc, x_sp, t_sp, m_sp, None, None)
    cfg = dict(combinations=["lightcurve", "spectral"], nband=2, transformer_kwargs=tk, transformer_spectral_kwargs=sk)
    ref = O.training_loss(sd, cfg, batch).item()
    gb = tuple(None if v is None else v.to(dev) for v in batch)
    from maven_b200 import _lib
    from maven_b200.transformer_utils import set_precision
    L = _lib.lib()
    # the tier the benchmark runs: tcgen05 GEMMs + warp-MMA attention + fused feed-forward kernels, 1e-3 bar (north_star);
    # the tier counters prove that those kernels -- not the FFMA ones -- executed the encoder layers
    for tier, tol in (("fp32", 1e-5), ("fused", 1e-3)):
        torch.manual_seed(0)
        m = LightCurveImageCLIP(logit_scale=19.55, nband=2, loss="softmax", transformer_kwargs=tk, transformer_spectral_kwargs=sk,
                                combinations=["lightcurve", "spectral"], lr=1e-3, optimizer_kwargs={"weight_decay": 5.6e-4})
        m.load_state_dict(sd)
        set_precision(m.to(dev).train(), tier)
        opt = m.configure_optimizers()["optimizer"]
        L.mvn_tier_reset()
        loss = m.training_step(gb, 0)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        got = loss.item()
        rel = abs(got - ref) / abs(ref)
        tiers = [L.mvn_tier_count(i) for i in range(4)]
        print(f"smoke[{tier}]: loss gpu={got:.7f} oracle={ref:.7f} rel={rel:.2e} launches by tier (ffma, tcgen05, mma attention, fused)={tiers}")
        assert rel < tol, f"CUDA training step ({tier}) disagrees with the CPU oracle"
        if tier == "fused":
            assert tiers[1] > 0 and tiers[2] == 8 and tiers[3] == 8, "tensor-core / fused kernels did not run"
