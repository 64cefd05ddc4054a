"""GPU parity tests of the round-2 rows: retrieval AUC (N3), "meta" modality and ClipMLP heads (N4), optimizer checkpoint
round trip (torch.optim.RAdam state format), all against fixtures written by the unmodified reference
(tests/golden/make_golden_r2.py)."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden, relerr, split_golden

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-5, 2e-4


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["auc", "auc_small"])
def test_retrieval_auc_golden(name):
    """get_ROC_data / get_AUC (src/utils.py:380-426): the 100-point curve is integer counting -> bit-exact; AUC to 1e-12."""
    from maven_b200.utils import get_AUC, get_ROC_data
    g = load_golden(name)
    thr, frac = get_ROC_data(g["e1"].to(dev()), g["e2"].to(dev()))
    assert np.array_equal(thr, g["thresholds"].numpy())
    assert np.array_equal(frac, g["fraction_correct"].numpy())
    assert abs(get_AUC(g["e1"].to(dev()), g["e2"].to(dev())) - float(g["auc"])) < 1e-12


def test_retrieval_curve_large_vs_oracle():
    from maven_b200.utils import get_ROC_data
    from oracle import maven_oracle as O
    gen = torch.Generator().manual_seed(3)
    e1 = torch.randn(3001, 128, generator=gen)
    e2 = 0.2 * e1 + torch.randn(3001, 128, generator=gen)
    thr, frac = get_ROC_data(e1.to(dev()), e2.to(dev()))
    thr_o, frac_o = O.roc_data(e1.double(), e2.double())
    assert np.array_equal(thr, thr_o)
    assert np.abs(frac - frac_o).max() <= 2.0 / 3001          # fp32 vs fp64 similarity: at most a near-tie or two change rank


def _clip_model(g, combinations, meta=False):
    from maven_b200.models_multimodal import LightCurveImageCLIP
    tk = dict(n_out=32, emb=32, heads=4, depth=2, dropout=0.0, time_norm=20583.37, agg="mean")
    sk = dict(n_out=32, emb=32, heads=2, depth=1, dropout=0.0, time_norm=17945.14, agg="mean")
    kw = dict(meta_kwargs=dict(input_dim=32, hidden_dim=48, num_layers=2, dropout=0.0), n_classes=5) if meta else {}
    return LightCurveImageCLIP(logit_scale=19.545966923442453, lr=1e-3, nband=2, loss="softmax", transformer_kwargs=tk,
                               transformer_spectral_kwargs=sk, optimizer_kwargs={"weight_decay": 5.6e-4}, combinations=combinations, **kw)


def _batch(g):
    return tuple(None if k == "img" else g[k].to(dev()) for k in ("img", "x_lc", "t_lc", "mask_lc", "x_sp", "t_sp", "mask_sp", "redshift", "cls"))


def test_meta_modality_golden():
    """CLIP over [lightcurve, spectral, meta]: 3 pairs, embedding list order fixed (src/models_multimodal.py:260-273)."""
    g = load_golden("model_meta")
    m = _clip_model(g, ["lightcurve", "spectral", "meta"], meta=True)
    sd, grads, _ = split_golden(g)
    for k in ("logit_scale", "logit_bias"):
        sd[k] = g[k]
    m.load_state_dict(sd)
    m = m.to(dev()).train()
    loss = m.training_step(_batch(g), 0)
    assert abs(loss.item() - g["loss"].item()) < TOL * abs(g["loss"].item())
    loss.backward()
    for k, p in m.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert relerr(got, grads[k]) < GTOL or (got.cpu() - grads[k]).abs().max() < 1e-7, k
    m.eval()
    with torch.no_grad():
        out = m(*_batch(g))
    assert len(out) == 3
    for i, o in enumerate(out):
        assert relerr(o, g[f"eval_out{i}"]) < TOL


@pytest.mark.parametrize("name", ["clipmlp_reg", "clipmlp_cls"])
def test_clipmlp_golden(name):
    from maven_b200.models_multimodal import ClipMLP
    g = load_golden(name)
    reg = name.endswith("reg")
    head = ClipMLP(_clip_model(g, ["lightcurve", "spectral"]), dict(hidden_dim=24, num_layers=2, dropout=0.0, output_dim=1 if reg else 5),
                   {"weight_decay": 5.6e-4}, 1e-3, combinations=["lightcurve", "spectral"], regression=reg, classification=not reg, n_classes=5)
    sd, grads, _ = split_golden(g)
    for k in ("clip_model.logit_scale", "clip_model.logit_bias"):
        sd[k] = g[k]
    head.load_state_dict(sd)
    head = head.to(dev()).train()
    loss = head.training_step(_batch(g), 0)
    assert abs(loss.item() - g["loss"].item()) < TOL * abs(g["loss"].item())
    loss.backward()
    for k, p in head.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert relerr(got, grads[k]) < GTOL or (got.cpu() - grads[k]).abs().max() < 1e-7, k
    head.eval()
    with torch.no_grad():
        b = _batch(g)
        y = head(*b[1:7])
    assert relerr(y, g["y_eval"]) < TOL
    opt = head.configure_optimizers()["optimizer"]
    opt.step()                                    # torch.optim.RAdam over the flat-buffer views


def test_optimizer_state_dict_roundtrip_and_reference_format():
    """FusedRAdam checkpoints in torch.optim.RAdam's format: (a) save after 3 steps, resume in a fresh model + optimizer, the 4th
    step equals the uninterrupted run; (b) the saved state loads into torch.optim.RAdam and that optimizer takes the same 4th step."""
    g = load_golden("model_clip2")
    from test_gpu_parity import _build_model
    batch = tuple(g[k].to(dev()) for k in ("img", "x_lc", "t_lc", "mask_lc", "x_sp", "t_sp", "mask_sp", "redshift", "cls"))

    def run(m, opt, n):
        for _ in range(n):
            m.training_step(batch, 0).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)

    m1, _, _ = _build_model("model_clip2", g)
    m1 = m1.to(dev()).train()
    o1 = m1.configure_optimizers()["optimizer"]
    run(m1, o1, 3)
    ck_model, ck_opt = copy.deepcopy(m1.state_dict()), copy.deepcopy(o1.state_dict())
    assert len(ck_opt["state"]) > 0 and all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in ck_opt["state"].values())
    assert all(float(v["step"]) == 3.0 for v in ck_opt["state"].values())
    run(m1, o1, 1)
    # (a) resume with the fused optimizer
    m2, _, _ = _build_model("model_clip2", g)
    m2.load_state_dict(ck_model)
    m2 = m2.to(dev()).train()
    o2 = m2.configure_optimizers()["optimizer"]
    o2.load_state_dict(ck_opt)
    run(m2, o2, 1)
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k
    # (b) the same state in torch's own RAdam
    m3, _, _ = _build_model("model_clip2", g)
    m3.load_state_dict(ck_model)
    m3 = m3.to(dev()).train()
    o3 = torch.optim.RAdam(m3.parameters(), lr=m3.lr, **m3.optimizer_kwargs)
    o3.load_state_dict(ck_opt)
    run(m3, o3, 1)
    for (k, a), (_, b) in zip(m1.named_parameters(), m3.named_parameters()):
        assert (a - b).abs().max() < 2e-6, k


def test_default_constructed_model_fails_fast():
    from maven_b200.models_multimodal import LightCurveImageCLIP
    with pytest.raises(NotImplementedError, match="softmax"):
        LightCurveImageCLIP(combinations=["lightcurve", "spectral"])          # default loss='sigmoid' (reference default) is not built
    m = LightCurveImageCLIP(loss="softmax", combinations=["lightcurve", "spectral"]).to(dev())   # defaults: emb=256, heads=2 -> head dim 128
    x = torch.zeros(2, 8, device=dev()); t = torch.zeros(2, 8, device=dev()); mk = torch.ones(2, 8, dtype=torch.bool, device=dev())
    with pytest.raises(RuntimeError, match="emb=256 not in"):                 # the supported grid is stated in INTEGRATION.md
        m(None, x, t, mk, x, t, mk)


def test_masked_lc_pretraining_golden():
    """N4: MaskedLightCurveEncoder (src/models_pretraining.py:106-226) -- fused encoder with agg="pretraining", last_layer GEMM and the
    masked-MSE kernel against the unmodified reference's training_step: same run selection (random.seed), prediction, loss,
    every parameter gradient."""
    import ast
    import random

    from maven_b200.models_pretraining import MaskedLightCurveEncoder
    g = load_golden("pretrain_lc")
    cfg = ast.literal_eval(g["cfg"])
    sd, grads, _ = split_golden(g)
    m = MaskedLightCurveEncoder(f_mask=0.25, nband=2, transformer_kwargs=dict(cfg), optimizer_kwargs={}, lr_scheduler_kwargs={}, lr=1e-3)
    m.load_state_dict(sd)
    m = m.to(dev()).train()
    x, t, mask = g["x"].to(dev()), g["t"].to(dev()), g["mask"].to(dev())
    random.seed(int(g["random_seed"]))
    x_pred, mask_pred = m.masked_forward(x, t, mask, f_mask=0.25)
    assert torch.equal(mask_pred.cpu(), g["mask_pred"])
    assert relerr(x_pred, g["x_pred"]) < TOL
    random.seed(int(g["random_seed"]))
    loss = m.training_step((t, x, mask), 0)
    assert abs(float(loss.detach()) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    loss.backward()
    for k, p in m.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)          # agg="pretraining" never touches net.projection
        assert relerr(got, grads[k]) < GTOL or (got.cpu() - grads[k]).abs().max() < 1e-6, k
    # the 9-tuple batch of the simulation loader takes the same path (src/models_pretraining.py:223-226)
    random.seed(int(g["random_seed"]))
    loss9 = m.validation_step((None, x, t, mask, None, None, None, None, None), 0)
    assert torch.equal(loss9, loss)
    xs, ps = m.masked_pred(x, t, mask, f_mask=0.25)
    assert xs.shape == ps.shape and xs.ndim == 1


@pytest.mark.parametrize("tier,tol", [("fp32", TOL), ("fused", 1e-3)])
def test_c1_shape_golden(tier, tol):
    """BASELINE.json configs[0] (configs/maven-lite.yaml as shipped: agg="attn" light curves, spectra padded to 1024 tokens) against the
    fixture written by the unmodified reference (tests/golden/make_golden_c1.py): loss, gradients (fp32 tier) and eval embeddings."""
    from maven_b200.models_multimodal import LightCurveImageCLIP
    from maven_b200.transformer_utils import set_precision
    g = load_golden("model_c1")
    tk = dict(n_out=32, emb=64, heads=8, depth=2, dropout=0.0, time_norm=20583.369161312577, agg="attn")
    sk = dict(n_out=32, emb=32, heads=2, depth=3, dropout=0.0, time_norm=17945.142213594805, agg="mean")
    m = LightCurveImageCLIP(logit_scale=19.545966923442453, lr=1e-3, nband=2, loss="softmax", transformer_kwargs=tk, transformer_spectral_kwargs=sk,
                            optimizer_kwargs={"weight_decay": 5.6e-4}, combinations=["lightcurve", "spectral"])
    sd, grads, _ = split_golden(g)
    for k in ("logit_scale", "logit_bias"):
        sd[k] = g[k]
    m.load_state_dict(sd)
    m = set_precision(m.to(dev()).train(), tier)
    loss = m.training_step(_batch(g), 0)
    assert abs(loss.item() - g["loss"].item()) < tol * abs(g["loss"].item())
    loss.backward()
    for k, p in m.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if tier == "fp32":
            assert relerr(got, grads[k]) < GTOL or (got.cpu() - grads[k]).abs().max() < 1e-7, k
        else:
            assert torch.isfinite(got).all(), k
    m.eval()
    with torch.no_grad():
        out = m(*_batch(g))
    for i, o in enumerate(out):
        assert relerr(o, g[f"eval_out{i}"]) < tol
