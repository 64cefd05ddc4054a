"""CPU: the oracle restatement reproduces the reference's outputs (fixtures written by
tests/golden/make_golden.py from the unmodified reference) and its shipped KAT vectors."""
import ast

import pytest
import torch

from conftest import load_golden, relerr, split_golden
from oracle import maven_oracle as O

TOL = 2e-6      # fp32 restatement vs fp32 reference: same ops, same order


def _leafify(sd):
    return {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}


@pytest.mark.parametrize("name", ["enc_lc_mean", "enc_sp_mean", "enc_lc_attn", "enc_lc_max", "enc_lc_pre"])
def test_seq_encoder(name):
    g = load_golden(name)
    cfg = ast.literal_eval(g["cfg"])
    sd, grads, _ = split_golden(g)
    sd["query"] = g["query"] if "query" in g else None
    sd = _leafify({k: v for k, v in sd.items() if v is not None})
    y = O.seq_encoder(sd, "", g["x"][..., None], g["t"], g["mask"], emb=cfg["emb"], heads=cfg["heads"], depth=cfg["depth"],
                      nband=cfg["nband"], agg=cfg["agg"], time_norm=cfg["time_norm"])
    assert relerr(y, g["y"]) < TOL
    (y * g["w"]).sum().backward()
    for k, ref in grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        assert relerr(got, ref) < 5e-5 or (got - ref).abs().max() < 1e-6, k


def test_time_pe_bit_exact():
    g = load_golden("time_pe")
    pe = O.time_positional_encoding(g["t"], 32, 17945.14)
    assert torch.equal(pe, g["pe"])


def test_attention_and_block():
    g = load_golden("attn_block")
    att = {k[4:]: v for k, v in g.items() if k.startswith("att.")}
    blk = {k[4:]: v for k, v in g.items() if k.startswith("blk.")}
    assert relerr(O.self_attention(att, "", g["x"], g["mask"], 2), g["y_att"]) < TOL
    assert relerr(O.transformer_block(blk, "", g["x"], g["mask"], 2), g["y_blk"]) < TOL


def test_convmixer():
    g = load_golden("convmixer")
    sd, grads, after = split_golden(g)
    sd = _leafify(sd)
    stats = {}
    y = O.convmixer(sd, "", g["img"], depth=2, kernel_size=5, patch_size=10, training=True, stats_out=stats)
    assert relerr(y, g["y"]) < 1e-5
    (y * g["w"]).sum().backward()
    for k, ref in grads.items():
        assert relerr(sd[k].grad, ref) < 2e-4, k
    for k, ref in after.items():
        assert relerr(stats[k], ref) < 1e-5, k
    ev = {**{k: v.detach() for k, v in sd.items()}, **stats}
    assert relerr(O.convmixer(ev, "", g["img"], depth=2, kernel_size=5, patch_size=10, training=False), g["y_eval"]) < 1e-5


def test_clip_loss_and_closed_form_backward():
    g = load_golden("clip_loss")
    e = [g[f"e{i}"].clone().requires_grad_() for i in range(3)]
    ls, lb = g["ls"].clone().requires_grad_(), g["lb"].clone().requires_grad_()
    l2 = O.clip_loss(e[0], e[1], ls, lb)
    assert abs(l2.item() - g["loss2"].item()) < 1e-6 * abs(g["loss2"].item())
    d1, d2, dls, dlb = O.clip_loss_grads_closed_form(e[0].detach(), e[1].detach(), ls.detach(), lb.detach())
    assert relerr(d1, g["g2_e0"]) < 1e-5 and relerr(d2, g["g2_e1"]) < 1e-5
    assert abs(dls.item() - g["g2_ls"].item()) < 1e-5 * abs(g["g2_ls"].item()) + 1e-7
    assert abs(dlb.item()) < 1e-6
    l3 = O.clip_loss_multimodal(e, ls, lb)
    assert abs(l3.item() - g["loss3"].item()) < 1e-6 * abs(g["loss3"].item())
    gr = torch.autograd.grad(l3, e + [ls])
    for i in range(3):
        assert relerr(gr[i], g[f"g3_e{i}"]) < 1e-5
    assert abs(gr[3].item() - g["g3_ls"].item()) < 1e-5 * abs(g["g3_ls"].item()) + 1e-7


MODEL_CFG = dict(
    nband=2,
    transformer_kwargs=dict(n_out=32, emb=32, heads=4, depth=2, time_norm=20583.37, agg="mean"),
    transformer_spectral_kwargs=dict(n_out=32, emb=32, heads=2, depth=1, time_norm=17945.14, agg="mean"),
    conv_kwargs=dict(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32),
)
MODEL_CASES = {
    "model_clip3": dict(combinations=["lightcurve", "spectral", "host_galaxy"]),
    "model_clip2": dict(combinations=["lightcurve", "spectral"]),
    "model_cls5": dict(combinations=["lightcurve"], classification=True, n_classes=5),
    "model_reg": dict(combinations=["lightcurve"], regression=True),
}


def golden_batch(g):
    return (g["img"], g["x_lc"], g["t_lc"], g["mask_lc"], g["x_sp"], g["t_sp"], g["mask_sp"], g["redshift"], g["cls"])


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_training_step(name):
    g = load_golden(name)
    cfg = {**MODEL_CFG, **MODEL_CASES[name]}
    sd, grads, after = split_golden(g)
    for k in ("logit_scale", "logit_bias"):
        sd[k] = g[k]
    sd = _leafify(sd)
    stats = {}
    loss = O.training_loss(sd, cfg, golden_batch(g), stats_out=stats)
    assert abs(loss.item() - g["loss"].item()) < 2e-6 * abs(g["loss"].item())
    loss.backward()
    for k, ref in grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        assert relerr(got, ref) < 2e-4 or (got - ref).abs().max() < 1e-7, k
    if after:           # one RAdam step (lr 1e-3, wd 5.6e-4) through the oracle's restatement
        for k, ref in after.items():
            if k not in grads or torch.equal(ref, sd[k].detach()):
                continue        # buffers, and params whose grad was None (torch skips them entirely)
            p = sd[k].detach().clone()
            gk = sd[k].grad if sd[k].grad is not None else torch.zeros_like(p)
            O.radam_step(p, gk.clone(), torch.zeros_like(p), torch.zeros_like(p), 1, 1e-3, weight_decay=5.6e-4)
            assert (p - ref).abs().max() < 1e-6, k


def test_radam_trajectory():
    g = load_golden("radam")
    p = g["params"][0].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(g["grads"].shape[0]):
        O.radam_step(p, g["grads"][s].clone(), m, v, s + 1, float(g["lr"]), weight_decay=float(g["wd"]))
        assert (p - g["params"][s + 1]).abs().max() < 2e-6, s


def test_shipped_kat_lc_reg():
    """The reference's own known-answer vectors: lc-reg checkpoint vs pickled y_pred."""
    g = load_golden("kat_lc_reg")
    cfg = ast.literal_eval(g["cfg"])
    sd = {k: v for k, v in g.items() if torch.is_tensor(v) and ("." in k or k.startswith("logit"))}
    mc = dict(combinations=["lightcurve"], regression=True, nband=2, transformer_kwargs=cfg)
    y = O.model_forward(sd, mc, (None, g["x_lc"], g["t_lc"], g["mask_lc"], None, None, None, None, None), training=False)
    assert (y.flatten() - g["y_pred"]).abs().max() < 2e-6


def test_retrieval_ranks_matches_argsort_loop():
    torch.manual_seed(0)
    e1, e2 = torch.randn(23, 16), torch.randn(23, 16)
    r = O.retrieval_ranks(e1, e2)
    for j in range(23):     # src/utils.py:399-409
        cs = torch.nn.functional.cosine_similarity(e1, e2[j][None], dim=-1)
        order = torch.argsort(cs, descending=True)
        assert int((order == j).nonzero()[0, 0]) == int(r[j])


class _GivenMask(torch.nn.Module):
    """Stands in for nn.Dropout: multiplies by a given keep-factor tensor."""

    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return x * self.scale


def test_convmixer_given_dropout_masks_match_torch_layers():
    """The oracle's drop_scales path == the reference's nn.Sequential layout (src/models_multimodal.py:52-89, rebuilt here
    from stock torch layers) with every nn.Dropout replaced by the same given mask."""
    import torch.nn as nn
    torch.manual_seed(5)
    dim, depth, k, p, B, pd = 16, 2, 5, 10, 6, 0.3
    net = nn.Sequential(nn.Conv2d(3, dim, p, stride=p, bias=False), nn.GELU(), nn.BatchNorm2d(dim))
    masks = {}
    for d in range(depth):
        masks[2 * d + 1] = (torch.rand(B, dim, 6, 6) > pd).double() / (1 - pd)
        masks[2 * d + 2] = (torch.rand(B, dim, 6, 6) > pd).double() / (1 - pd)

        class Res(nn.Module):
            def __init__(self, fn):
                super().__init__()
                self.fn = fn

            def forward(self, x):
                return self.fn(x) + x
        net.append(nn.Sequential(Res(nn.Sequential(nn.Conv2d(dim, dim, k, groups=dim, padding="same"), nn.GELU(), nn.BatchNorm2d(dim),
                                                   _GivenMask(masks[2 * d + 1]))),
                                 nn.Conv2d(dim, dim, 1), nn.GELU(), nn.BatchNorm2d(dim), _GivenMask(masks[2 * d + 2])))
    masks[2 * depth + 1] = (torch.rand(B, 1024) > pd).double() / (1 - pd)
    proj = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(dim, 1024), nn.GELU(), _GivenMask(masks[2 * depth + 1]),
                         nn.Linear(1024, 8))
    full = nn.ModuleDict({"net": net, "projection": proj}).double().train()
    img = torch.rand(B, 3, 60, 60, dtype=torch.float64)
    sd = {k_: v.detach().clone() for k_, v in full.state_dict().items()}
    ref = full["projection"](full["net"](img))
    got = O.convmixer(sd, "", img, depth=depth, kernel_size=k, patch_size=p, training=True, drop_scales=masks)
    assert relerr(got, ref) < 1e-12
    assert relerr(O.convmixer(sd, "", img, depth=depth, kernel_size=k, patch_size=p, training=True), ref) > 1e-3


def test_noisy_loader_golden():
    """N1: the oracle's augmentation == the reference NoisyDataLoader's outputs on the random tensors it consumed."""
    g = load_golden("noisy_loader")
    inten, lvl = float(g["max_noise_intensity"]), float(g["noise_level_mag"])
    for case in ("img", "img_lc", "all"):
        assert torch.equal(O.noisy_images(g["img"], g[f"{case}.img_u"], g[f"{case}.rot_k"], inten), g[f"{case}.out0"]), case
    for case in ("lc", "img_lc", "lc_sp", "all"):
        assert torch.equal(O.noisy_seq(g["mag"], g["magerr"], g[f"{case}.noise_mag"], lvl), g[f"{case}.out1"]), case
    for case in ("lc_sp", "all"):
        assert torch.equal(O.noisy_seq(g["spec"], g["specerr"], g[f"{case}.noise_spec"], lvl), g[f"{case}.out4"]), case


def test_masked_lc_pretraining_objective():
    """N4: MaskedLightCurveEncoder.training_step of the unmodified reference (src/models_pretraining.py:106-226) with the run
    selection it drew: prediction, loss and every parameter gradient."""
    g = load_golden("pretrain_lc")
    cfg = ast.literal_eval(g["cfg"])
    sd, grads, _ = split_golden(g)
    sd = _leafify(sd)
    kw = dict(emb=cfg["emb"], heads=cfg["heads"], depth=cfg["depth"], nband=2, time_norm=cfg["time_norm"])
    x_pred = O.masked_lc_pred(sd, g["x"], g["t"], g["mask"], g["mask_in"], **kw)
    assert relerr(x_pred, g["x_pred"]) < TOL
    loss = O.masked_lc_loss(sd, g["x"], g["t"], g["mask"], g["mask_in"], g["mask_pred"], **kw)
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-6 * abs(float(g["loss"]))
    loss.backward()
    for k, ref in grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        assert relerr(got, ref) < 5e-5 or (got - ref).abs().max() < 1e-6, k


C1_CFG = dict(nband=2, combinations=["lightcurve", "spectral"],
              transformer_kwargs=dict(n_out=32, emb=64, heads=8, depth=2, time_norm=20583.369161312577, agg="attn"),
              transformer_spectral_kwargs=dict(n_out=32, emb=32, heads=2, depth=3, time_norm=17945.142213594805, agg="mean"))


def test_training_step_c1_shape():
    """BASELINE.json configs[0] (configs/maven-lite.yaml as shipped): attention-pooled light curves (nn.MultiheadAttention over the
    zero-padded rows) + spectra padded to 1024 tokens with lengths 214 / 1024 / 517 / 214 -- loss, every gradient and the eval
    embeddings of the unmodified reference (tests/golden/make_golden_c1.py)."""
    g = load_golden("model_c1")
    sd, grads, _ = split_golden(g)
    for k in ("logit_scale", "logit_bias"):
        sd[k] = g[k]
    sd = _leafify(sd)
    batch = (None, g["x_lc"], g["t_lc"], g["mask_lc"], g["x_sp"], g["t_sp"], g["mask_sp"], g["redshift"], g["cls"])
    loss = O.training_loss(sd, C1_CFG, batch)
    assert abs(loss.item() - g["loss"].item()) < 2e-6 * abs(g["loss"].item())
    loss.backward()
    for k, ref in grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        assert relerr(got, ref) < 2e-4 or (got - ref).abs().max() < 1e-7, k
    with torch.no_grad():
        out = O.model_forward({k: v.detach() for k, v in sd.items()}, C1_CFG, batch, training=False)
    for i, o in enumerate(out):
        assert relerr(o, g[f"eval_out{i}"]) < 1e-5
