"""GPU parity tests of the tensor-core tier (prec=1: tcgen05 kind::tf32, fp32 accumulate, fp32 storage).
Tolerance: 1e-3 normwise relative (north_star's reduced-precision tier); references are torch fp64 / the CPU oracle."""
import ctypes
import math

import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL_TC = 1e-3


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def L():
    from maven_b200 import _lib
    return _lib.lib()


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


SHAPES = [(5000, 192, 64), (777, 64, 64), (4097, 256, 64), (3000, 64, 256), (2500, 96, 32), (1300, 32, 32), (2049, 128, 32),
          (1500, 32, 128), (128, 64, 192), (20000, 32, 96), (60001, 192, 64), (60000, 256, 64), (77777, 32, 32)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_tc_linear_fwd(L, M, N, K, act):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K); w = torch.randn(N, K) / math.sqrt(K); b = torch.randn(N)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act:
        ref = torch.relu(ref)
    xg, wg, bg = x.to(dev()), w.to(dev()), b.to(dev())
    y = torch.full((M, N), float("nan"), device=dev())
    assert L.mvn_linear_fwd(P(xg), P(wg), P(bg), P(y), None, M, N, K, act, 1, S()) == 0, L.mvn_last_error()
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    assert relerr(y, ref) < TOL_TC


def test_tc_linear_fwd_device_row_count(L):
    """rows beyond the device-side live count are neither read for output nor written."""
    torch.manual_seed(5)
    M_cap, n, N, K = 4000, 2345, 64, 64
    x = torch.randn(M_cap, K); w = torch.randn(N, K) / 8
    xg, wg = x.to(dev()), w.to(dev())
    y = torch.full((M_cap, N), 7.0, device=dev())
    nrows = torch.tensor([n], dtype=torch.int32, device=dev())
    assert L.mvn_linear_fwd(P(xg), P(wg), None, P(y), P(nrows), M_cap, N, K, 0, 1, S()) == 0
    torch.cuda.synchronize()
    assert relerr(y[:n], x[:n].double() @ w.double().t()) < TOL_TC
    assert (y[n:] == 7.0).all()


# the (60000+, ...) cases give every persistent CTA 3+ tiles: both epilogue groups run and every ring wraps around
@pytest.mark.parametrize("M,N,K", [(3333, 64, 64), (1000, 32, 32), (2000, 64, 256), (900, 32, 128), (60001, 64, 64), (70000, 32, 128), (60000, 64, 256)])
def test_tc_linear_res_ln(L, M, N, K):
    torch.manual_seed(N + K)
    x = torch.randn(M, K); w = torch.randn(N, K) / math.sqrt(K); b = torch.randn(N); r = torch.randn(M, N)
    g = 1 + 0.1 * torch.randn(N); be = 0.1 * torch.randn(N)
    pre = torch.nn.functional.linear(x.double(), w.double(), b.double()) + r.double()
    ref = torch.nn.functional.layer_norm(pre, (N,), g.double(), be.double(), 1e-5)
    mu = pre.mean(1, keepdim=True); var = pre.var(1, unbiased=False, keepdim=True)
    ref_xhat = (pre - mu) / torch.sqrt(var + 1e-5)
    t = [v.to(dev()) for v in (x, w, b, r, g, be)]
    y = torch.empty(M, N, device=dev()); xhat = torch.empty(M, N, device=dev()); rstd = torch.empty(M, device=dev())
    assert L.mvn_linear_res_ln_fwd(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(t[4]), P(t[5]), P(y), P(xhat), P(rstd), None, M, N, K, 1e-5, 1, S()) == 0
    torch.cuda.synchronize()
    assert relerr(y, ref) < TOL_TC
    assert relerr(xhat, ref_xhat) < TOL_TC
    assert relerr(rstd, 1 / torch.sqrt(var + 1e-5).flatten()) < TOL_TC


@pytest.mark.parametrize("M,N,K", [(3000, 192, 64), (2222, 64, 256), (1500, 256, 64), (4000, 96, 32), (1000, 128, 32), (1000, 32, 128),
                                   (60001, 64, 256), (61000, 256, 64), (59999, 192, 64), (80000, 32, 128)])
@pytest.mark.parametrize("mode", ["plain", "addend", "relu_mask"])
def test_tc_linear_bwd_input(L, M, N, K, mode):
    """dX[M,K] = dY[M,N] W[N,K] (+addend) (* relu mask)."""
    torch.manual_seed(N * K)
    dy = torch.randn(M, N); w = torch.randn(N, K) / math.sqrt(N); add = torch.randn(M, K); src = torch.randn(M, K)
    ref = dy.double() @ w.double()
    if mode == "addend":
        ref = ref + add.double()
    if mode == "relu_mask":
        ref = ref * (src.double() > 0)
    dyg, wg, addg, srcg = (v.to(dev()) for v in (dy, w, add, src))
    dx = torch.empty(M, K, device=dev())
    rc = L.mvn_linear_bwd_input(P(dyg), P(wg), P(dx), P(addg) if mode == "addend" else None, P(srcg) if mode == "relu_mask" else None,
                                1 if mode == "relu_mask" else 0, None, M, N, K, 1, S())
    assert rc == 0, L.mvn_last_error()
    torch.cuda.synchronize()
    assert relerr(dx, ref) < TOL_TC


@pytest.mark.parametrize("M,N,K", [(5000, 192, 64), (3000, 64, 256), (4001, 256, 64), (2500, 96, 32), (1300, 32, 128), (999, 128, 32), (700, 64, 64)])
def test_tc_linear_bwd_weight(L, M, N, K):
    torch.manual_seed(M)
    dy = torch.randn(M, N); x = torch.randn(M, K)
    ref_w = dy.double().t() @ x.double(); ref_b = dy.double().sum(0)
    dyg, xg = dy.to(dev()), x.to(dev())
    wsb = L.mvn_linear_bwd_weight_workspace_bytes(M, N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev())
    dw = torch.empty(N, K, device=dev()); db = torch.empty(N, device=dev())
    assert L.mvn_linear_bwd_weight(P(dyg), P(xg), P(dw), P(db), None, M, N, K, 0, P(ws), wsb, 1, S()) == 0, L.mvn_last_error()
    torch.cuda.synchronize()
    assert relerr(dw, ref_w) < TOL_TC
    assert relerr(db, ref_b) < 2e-3          # colsum rides on the same TF32 MMA (against a block of ones)


@pytest.mark.parametrize("case", ["lc", "sp"])
def test_tc_seq_encoder_vs_oracle(case):
    """Whole encoder forward + backward in the tf32 tier vs the fp32 oracle at the C4 shapes."""
    from maven_b200.selfcheck import synthetic_seq
    from maven_b200.transformer_utils import TransformerWithTimeEmbeddings, set_precision
    from oracle import maven_oracle as O
    gen = torch.Generator().manual_seed(11)
    if case == "lc":
        kw = dict(n_out=32, nband=2, agg="mean", time_norm=20583.37, emb=64, heads=8, depth=5)
        x, t, m = synthetic_seq(gen, 48, 200, 2, 300.0, 20, 100, 0.0)
    else:
        kw = dict(n_out=32, nband=1, agg="mean", time_norm=17945.14, emb=32, heads=2, depth=13)
        x, t, m = synthetic_seq(gen, 48, 220, 1, 5500.0, 110, 220, 3700.0)
    torch.manual_seed(0)
    enc = TransformerWithTimeEmbeddings(dropout=0.0, **kw)
    okw = {k: kw[k] for k in ("emb", "heads", "depth", "nband", "agg", "time_norm")}
    sdg = {k: v.detach().float().requires_grad_() for k, v in enc.state_dict().items()}
    yr = O.seq_encoder(sdg, "", x[..., None], t, m, **okw)
    w = torch.randn(yr.shape, generator=gen)
    (yr * w).sum().backward()
    enc = set_precision(enc.to(dev()), "tf32")
    y = enc(x[..., None].to(dev()), t.to(dev()), m.to(dev()))
    (y * w.to(dev())).sum().backward()
    e_fwd = relerr(y, yr)
    worst = max((relerr(p.grad, sdg[k].grad), k) for k, p in enc.named_parameters())
    print(f"tf32 tier {case}: fwd relerr {e_fwd:.3e}, worst grad relerr {worst[0]:.3e} ({worst[1]})")
    assert e_fwd < TOL_TC
    assert worst[0] < 1e-2, worst


@pytest.mark.parametrize("E,H,lens", [(64, 8, [120, 1, 37, 200, 64, 8, 129]), (32, 2, [220, 165, 16, 7, 255, 256]), (32, 2, [300, 1024, 513, 9])])
def test_tc_attention_fwd_bwd(L, E, H, lens):
    """packed-stream attention on tensor cores vs torch fp64 per sequence (softmax(QK^T/sqrt(E)) V and its gradients)."""
    torch.manual_seed(len(lens) * E)
    hd = E // H
    B, M = len(lens), sum(lens)
    qkv = torch.randn(M, 3 * E); dout = torch.randn(M, E)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    ref = qkv.double().requires_grad_()
    outs = []
    for b in range(B):
        r = ref[cu[b]:cu[b + 1]]
        q, k, v = (r[:, i * E:(i + 1) * E].reshape(-1, H, hd).transpose(0, 1) for i in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(E), dim=-1)
        outs.append((p @ v).transpose(0, 1).reshape(-1, E))
    oref = torch.cat(outs)
    oref.backward(dout.double())
    qg, cug, dg = qkv.to(dev()), cu.to(dev()), dout.to(dev())
    out = torch.empty(M, E, device=dev()); lse = torch.empty(M, H, device=dev()); dqkv = torch.full((M, 3 * E), float("nan"), device=dev())
    sc = 1 / math.sqrt(E)
    assert L.mvn_attention_fwd(P(qg), P(cug), None, P(out), P(lse), B, E, H, sc, 1, S()) == 0, L.mvn_last_error()
    assert L.mvn_attention_bwd(P(qg), P(cug), None, P(out), P(lse), P(dg), P(dqkv), B, E, H, sc, 1, S()) == 0, L.mvn_last_error()
    torch.cuda.synchronize()
    assert relerr(out, oref) < TOL_TC
    assert torch.isfinite(dqkv).all()
    assert relerr(dqkv, ref.grad) < 2e-3


@pytest.mark.parametrize("N", [37, 128, 300, 1024, 2500])
def test_tc_clip_loss(N):
    """symmetric InfoNCE with the similarity tiles on tcgen05 (clip_loss_tc.cu) vs the fp64 oracle: 2-modality and 3-modality
    (one op for all pairs), incl. N below one 128-row tile, ragged last tiles and the d logit_scale reduction."""
    from maven_b200 import _lib
    from maven_b200.loss import clip_loss, clip_loss_multimodal
    from oracle import maven_oracle as O
    Lb = _lib.lib()
    gen = torch.Generator().manual_seed(N)
    base = torch.randn(N, 128, generator=gen)
    e = [torch.nn.functional.normalize(base + 0.8 * torch.randn(N, 128, generator=gen), dim=-1) for _ in range(3)]
    ls, lb = torch.tensor(math.log(19.55)), torch.tensor(-10.0)
    er = [t.double().requires_grad_() for t in e]
    lsr, lbr = ls.double().requires_grad_(), lb.double().requires_grad_()
    l2r = O.clip_loss(er[0], er[1], lsr, lbr)
    g2r = torch.autograd.grad(l2r, [er[0], er[1], lsr])
    l3r = O.clip_loss_multimodal(er, lsr, lbr)
    g3r = torch.autograd.grad(l3r, er + [lsr])
    ec = [t.to(dev()).requires_grad_() for t in e]
    lsc, lbc = ls.to(dev()).requires_grad_(), lb.to(dev()).requires_grad_()
    Lb.mvn_tier_reset()
    l2 = clip_loss(ec[0], ec[1], lsc, lbc, prec=1)
    g2 = torch.autograd.grad(l2, [ec[0], ec[1], lsc, lbc])
    assert Lb.mvn_tier_count(1) == 2 and Lb.mvn_tier_count(0) == 0           # forward + backward ran on the tensor-core kernels
    print(f"N={N}: loss tc {l2.item():.6f} ref {l2r.item():.6f}; grad relerr {relerr(g2[0], g2r[0]):.2e} {relerr(g2[1], g2r[1]):.2e}; dls {g2[2].item():.5f} ref {g2r[2].item():.5f}")
    # logits carry ~1e-3 of absolute TF32 noise (scale ~20 x 2^-11-grade products): absolute + relative bound on the loss
    assert abs(l2.item() - l2r.item()) < 2e-4 * abs(l2r.item()) + 3e-4
    assert relerr(g2[0], g2r[0]) < 2e-3 and relerr(g2[1], g2r[1]) < 2e-3
    assert abs(g2[2].item() - g2r[2].item()) < 2e-3 * abs(g2r[2].item()) + 1e-5
    assert g2[3].item() == 0.0
    l3 = clip_loss_multimodal(ec, lsc, lbc, prec=1)
    g3 = torch.autograd.grad(l3, ec + [lsc])
    assert abs(l3.item() - l3r.item()) < 2e-4 * abs(l3r.item()) + 1e-3
    for i in range(3):
        assert relerr(g3[i], g3r[i]) < 2e-3
    assert abs(g3[3].item() - g3r[3].item()) < 2e-3 * abs(g3r[3].item()) + 1e-5
