"""Generate golden input/output vectors by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs small .npz fixtures beside this file.  The GPU box never runs this script; tests read
the committed fixtures.  Framework-only imports of the reference (pytorch_lightning, ruamel,
torchmetrics, matplotlib, seaborn) are stubbed; every arithmetic op is the reference's own.
"""
import os
import pickle
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("MAVEN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    """The unmodified reference modules, framework imports stubbed (oracle/ref_loader.py)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ref_loader
    os.environ.setdefault("MAVEN_REFERENCE", REF)
    return ref_loader.import_reference()


def ragged_seq(gen, B, T, nband, tmax, lo=1):
    """mask/t/x shaped like the reference loaders: per band, first n valid, t sorted from 0."""
    per = T // nband
    mask = torch.zeros(B, T, dtype=torch.bool)
    t = torch.zeros(B, T)
    x = torch.zeros(B, T)
    for b in range(B):
        for k in range(nband):
            n = int(torch.randint(lo, per + 1, (1,), generator=gen))
            tt = torch.sort(torch.rand(n, generator=gen) * tmax)[0]
            tt = tt - tt[0] if nband > 1 else tt + 3700.0
            sl = slice(k * per, k * per + n)
            mask[b, sl] = True
            t[b, sl] = tt
            x[b, sl] = torch.randn(n, generator=gen)
    return x, t, mask


def sd_np(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def grads_np(module):
    return {"grad." + k: (p.grad.detach().cpu().numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
            for k, p in module.named_parameters()}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    rtu, rloss, rmm = import_reference()
    gen = torch.Generator().manual_seed(0)

    # ---- sequence encoders (A1-A7), forward + parameter grads --------------------------------
    enc_cases = {
        "enc_lc_mean": dict(n_out=32, nband=2, agg="mean", time_norm=20583.37, emb=64, heads=8, depth=2, T=40, B=5, tmax=300.0),
        "enc_sp_mean": dict(n_out=32, nband=1, agg="mean", time_norm=17945.14, emb=32, heads=2, depth=3, T=48, B=4, tmax=5500.0),
        "enc_lc_attn": dict(n_out=32, nband=2, agg="attn", time_norm=20583.37, emb=64, heads=8, depth=1, T=24, B=4, tmax=300.0),
        "enc_lc_max": dict(n_out=16, nband=2, agg="max", time_norm=3371.17, emb=32, heads=2, depth=1, T=24, B=4, tmax=300.0),
        "enc_lc_pre": dict(n_out=16, nband=2, agg="pretraining", time_norm=3371.17, emb=32, heads=2, depth=2, T=24, B=3, tmax=300.0),
    }
    for name, c in enc_cases.items():
        torch.manual_seed(1)
        m = rtu.TransformerWithTimeEmbeddings(n_out=c["n_out"], nband=c["nband"], agg=c["agg"], time_norm=c["time_norm"],
                                              emb=c["emb"], heads=c["heads"], depth=c["depth"], dropout=0.0)
        x, t, mask = ragged_seq(gen, c["B"], c["T"], c["nband"], c["tmax"])
        y = m(x[..., None], t, mask)
        w = torch.randn(y.shape, generator=gen)
        (y * w).sum().backward()
        save(name, x=x, t=t, mask=mask, y=y.detach(), w=w, cfg=np.array(repr({k: c[k] for k in ("n_out", "nband", "agg", "time_norm", "emb", "heads", "depth")})),
             **sd_np(m), **grads_np(m))

    # ---- building blocks: PE at large arguments, attention alone, block alone ------------------
    t = torch.rand(3, 16, generator=gen) * 9200.0
    pe = rtu.TimePositionalEncoding(32, 17945.14)(t)
    save("time_pe", t=t, pe=pe)
    torch.manual_seed(2)
    att = rtu.SelfAttention(32, heads=2)
    blk = rtu.TransformerBlock(32, 2, ff_hidden_mult=4)
    xx = torch.randn(3, 12, 32, generator=gen)
    mm = torch.rand(3, 12, generator=gen) > 0.3
    mm[2] = False                      # fully masked row -> uniform softmax (parity trap B-3)
    save("attn_block", x=xx, mask=mm, y_att=att(xx, mm).detach(), y_blk=blk(xx, mm).detach(),
         **{"att." + k: v for k, v in sd_np(att).items()}, **{"blk." + k: v for k, v in sd_np(blk).items()})

    # ---- ConvMixer (A8), train mode: output, grads, running-stat update -------------------------
    torch.manual_seed(3)
    cm = rmm.ConvMixer(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32, dropout_prob=0.0)
    for k, v in cm.named_parameters():     # non-trivial BN affine so its grads are exercised
        if v.dim() == 1 and "net" in k:
            v.data.add_(0.1 * torch.randn(v.shape, generator=gen))
    before = sd_np(cm)
    img = torch.rand(6, 3, 60, 60, generator=gen)
    cm.train()
    y = cm(img)
    w = torch.randn(y.shape, generator=gen)
    (y * w).sum().backward()
    after = {"after." + k: v for k, v in sd_np(cm).items() if "running" in k or "num_batches" in k}
    cm.eval()
    save("convmixer", img=img, y=y.detach(), w=w, y_eval=cm(img).detach(), **before, **after, **grads_np(cm))

    # ---- CLIP loss (A10/A11) forward + grads -------------------------------------------------------
    def unit(n, d):
        z = torch.randn(n, d, generator=gen)
        return (z / z.norm(dim=-1, keepdim=True)).requires_grad_()
    e = [unit(37, 128) for _ in range(3)]
    ls = torch.tensor(float(np.log(19.545966923442453)), requires_grad=True)
    lb = torch.tensor(-10.0, requires_grad=True)
    l2 = rloss.clip_loss(e[0], e[1], ls, lb)
    g2 = torch.autograd.grad(l2, [e[0], e[1], ls, lb], allow_unused=True)
    l3 = rloss.clip_loss_multimodal(e, ls, lb)
    g3 = torch.autograd.grad(l3, e + [ls, lb], allow_unused=True)
    z = lambda g, like: (g if g is not None else torch.zeros_like(like)).detach()
    save("clip_loss", e0=e[0].detach(), e1=e[1].detach(), e2=e[2].detach(), ls=ls.detach(), lb=lb.detach(),
         loss2=l2.detach(), g2_e0=g2[0], g2_e1=g2[1], g2_ls=g2[2], g2_lb=z(g2[3], lb),
         loss3=l3.detach(), g3_e0=g3[0], g3_e1=g3[1], g3_e2=g3[2], g3_ls=g3[3], g3_lb=z(g3[4], lb))

    # ---- full model training_step (A12) : CLIP (lc+sp+img), and C2-style classifier ------------------
    tk = dict(n_out=32, emb=32, heads=4, depth=2, dropout=0.0, time_norm=20583.37, agg="mean")
    sk = dict(n_out=32, emb=32, heads=2, depth=1, dropout=0.0, time_norm=17945.14, agg="mean")
    ck = dict(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32, dropout_prob=0.0)
    B = 6
    x_lc, t_lc, m_lc = ragged_seq(gen, B, 40, 2, 300.0)
    x_sp, t_sp, m_sp = ragged_seq(gen, B, 44, 1, 5500.0, lo=20)
    img = torch.rand(B, 3, 60, 60, generator=gen)
    red = torch.rand(B, generator=gen)
    cls = torch.randint(0, 5, (B,), generator=gen)
    for name, comb, kw in [("model_clip3", ["lightcurve", "spectral", "host_galaxy"], {}),
                           ("model_clip2", ["lightcurve", "spectral"], {}),
                           ("model_cls5", ["lightcurve"], dict(classification=True, n_classes=5)),
                           ("model_reg", ["lightcurve"], dict(regression=True))]:
        torch.manual_seed(4)
        m = rmm.LightCurveImageCLIP(logit_scale=19.545966923442453, lr=1e-3, nband=2, loss="softmax",
                                    transformer_kwargs=dict(tk), transformer_spectral_kwargs=dict(sk), conv_kwargs=dict(ck),
                                    optimizer_kwargs={"weight_decay": 5.6e-4}, combinations=comb, **kw)
        m.train()
        m.y_pred, m.y_true = [], []
        before = sd_np(m)
        batch = (img, x_lc, t_lc, m_lc, x_sp, t_sp, m_sp, red, cls)
        loss = m.training_step(batch, 0)
        loss.backward()
        g = grads_np(m)
        opt = m.configure_optimizers()["optimizer"]
        opt.step()
        after = {"after." + k: v for k, v in sd_np(m).items()} if name in ("model_clip3", "model_cls5") else {}
        m.eval()
        with torch.no_grad():
            out = m(*batch)
        outs = {f"eval_out{i}": o for i, o in enumerate(out)} if isinstance(out, list) else {"eval_out0": out}
        save(name, img=img, x_lc=x_lc, t_lc=t_lc, mask_lc=m_lc, x_sp=x_sp, t_sp=t_sp, mask_sp=m_sp, redshift=red, cls=cls,
             loss=loss.detach(), **before, **g, **after, **outs)

    # ---- RAdam trajectory (A13): torch.optim.RAdam, coupled weight decay, 8 steps ----------------------
    torch.manual_seed(5)
    p = torch.randn(257, requires_grad=True)
    opt = torch.optim.RAdam([p], lr=3.7e-3, weight_decay=5.6e-4)
    traj, gs = [p.detach().clone()], []
    for s in range(8):
        g = torch.randn(257, generator=gen)
        gs.append(g)
        p.grad = g.clone()
        opt.step()
        traj.append(p.detach().clone())
    save("radam", params=torch.stack(traj), grads=torch.stack(gs), lr=3.7e-3, wd=5.6e-4)

    # ---- the reference's own shipped known-answer vectors: lc-reg ------------------------------------
    res = pickle.load(open(os.path.join(REF, "evaluation_metrics/collect_regression_results.pkl"), "rb"))
    ent = [e for e in res if e["Model"] == "lc-reg"][0]
    rid = ent["id"]
    root = os.path.join(REF, "models/lc_reg")
    found = None
    for run in sorted(os.listdir(root)):
        d = os.path.join(root, run)
        if not os.path.isdir(d):
            continue
        ck = sorted([f for f in os.listdir(d) if f.startswith("epoch=")], key=lambda f: int(f.split("=")[1].split("-")[0]))[0]
        sdict = torch.load(os.path.join(d, ck), map_location="cpu", weights_only=False)["state_dict"]
        cfgk = dict(n_out=32, emb=32, heads=2, depth=9, dropout=0.0, time_norm=3371.17, agg="mean")
        import yaml
        cy = yaml.safe_load(open(os.path.join(d, "config.yaml")))
        cfgk.update(n_out=cy["n_out"], emb=cy["emb"], heads=cy["heads"], depth=cy["transformer_depth"], time_norm=cy["time_norm"], agg=cy["agg"])
        m = rmm.LightCurveImageCLIP(logit_scale=cy["logit_scale"], nband=2, loss="softmax", transformer_kwargs=cfgk,
                                    combinations=["lightcurve"], regression=True)
        m.load_state_dict(sdict)
        m.eval()
        lc = ent["lc_data"]
        with torch.no_grad():
            y = m(None, lc["x_lc"], lc["t_lc"], lc["mask_lc"], None, None, None).flatten()
        err = (y - ent["y_pred"]).abs().max().item()
        print(f"  lc-reg id={rid} run={run} ckpt={ck}: max|y-y_pred|={err:.3e}")
        if err < 1e-5:
            found = (run, ck, sdict, cfgk)
            break
    assert found, "no checkpoint reproduced the pickled predictions"
    run, ck, sdict, cfgk = found
    n = 96                                     # subset keeps the fixture small; weights are 0.5 MB
    lc = ent["lc_data"]
    save("kat_lc_reg", x_lc=lc["x_lc"][:n], t_lc=lc["t_lc"][:n], mask_lc=lc["mask_lc"][:n], y_pred=ent["y_pred"][:n],
         cfg=np.array(repr({k: cfgk[k] for k in ("n_out", "emb", "heads", "depth", "time_norm", "agg")})),
         source=np.array(f"models/lc_reg/{run}/{ck} vs evaluation_metrics/collect_regression_results.pkl[Model=lc-reg,id={rid}]"),
         **{k: v.numpy() for k, v in sdict.items()})


if __name__ == "__main__":
    main()
