"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, host-only entry points
(parameter counts, workspace sizes) agree with the Python-side layouts, flat parameter plumbing keeps state_dict
names, and the product path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from maven_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "maven_sm100.h")).read()
    declared = set(re.findall(r"\b(mvn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    raw = ctypes.CDLL(built.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, f"declared in include/maven_sm100.h but not exported: {missing}"
    assert set(built.EXPORTED_SYMBOLS) <= declared, set(built.EXPORTED_SYMBOLS) - declared
    assert built.lib().mvn_abi_version() == 1


def test_error_channel_without_gpu(built):
    L = built.lib()
    rc = L.mvn_l2norm_fwd(None, None, None, 0, 0, None)
    assert rc == -1 and b"l2norm_fwd" in L.mvn_last_error()


@pytest.mark.parametrize("kw", [dict(n_out=32, nband=2, agg="mean", emb=64, heads=8, depth=5),
                                dict(n_out=32, nband=1, agg="mean", emb=32, heads=2, depth=13),
                                dict(n_out=16, nband=2, agg="pretraining", emb=32, heads=2, depth=2)])
def test_seq_param_layout_matches_library(built, kw):
    from maven_b200.transformer_utils import TransformerWithTimeEmbeddings, _AGG
    enc = TransformerWithTimeEmbeddings(time_norm=1e4, dropout=0.0, **kw)
    with_head = kw["agg"] != "pretraining"
    ps = enc.core_params() + (enc.head_params() if with_head else [])
    cfg = enc.make_cfg(_AGG[kw["agg"]], 0, False)
    cfg.B, cfg.T = 4, 200 if kw["nband"] == 2 else 220
    assert built.lib().mvn_seq_param_count(ctypes.byref(cfg)) == sum(p.numel() for p in ps)
    assert built.lib().mvn_seq_workspace_bytes(ctypes.byref(cfg)) > 0
    assert len({id(p) for p in ps}) == len(ps)
    # every parameter except agg-specific ones is covered exactly once
    covered = {id(p) for p in ps}
    rest = [n for n, p in enc.named_parameters() if id(p) not in covered]
    assert all(n.startswith(("projection", "query", "agg_attn")) for n in rest), rest


def test_conv_param_layout_matches_library(built):
    from maven_b200.models_multimodal import ConvMixer
    cm = ConvMixer(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32, dropout_prob=0.0)
    cfg = built.ConvCfg(B=4, C=3, H=60, W=60, dim=32, depth=2, kernel_size=5, patch_size=10, n_out=32, enc_dim=0, hidden=1024,
                        normalize=0, training=1, prec=0, bn_eps=1e-5, bn_momentum=0.1, global_count=4 * 36)
    assert built.lib().mvn_conv_param_count(ctypes.byref(cfg)) == sum(p.numel() for p in cm.core_params()) == sum(p.numel() for p in cm.parameters())
    assert built.lib().mvn_conv_num_bn(ctypes.byref(cfg)) == len(cm.bn_layers()) == 5
    cfg.enc_dim = 128
    assert built.lib().mvn_conv_param_count(ctypes.byref(cfg)) == sum(p.numel() for p in cm.parameters()) + 128 * 32 + 128


def test_state_dict_keys_match_reference_contract():
    """names/shapes the shipped checkpoints carry (checked against a committed golden)."""
    from conftest import load_golden
    from maven_b200.models_multimodal import LightCurveImageCLIP
    g = load_golden("model_clip3")
    ref = {k: tuple(v.shape) for k, v in g.items() if torch.is_tensor(v) and ("." in k or k.startswith("logit"))
           and not k.startswith(("grad.", "after."))}
    m = LightCurveImageCLIP(logit_scale=19.5, nband=2, loss="softmax", combinations=["lightcurve", "spectral", "host_galaxy"],
                            transformer_kwargs=dict(n_out=32, emb=32, heads=4, depth=2, dropout=0.0, time_norm=20583.37, agg="mean"),
                            transformer_spectral_kwargs=dict(n_out=32, emb=32, heads=2, depth=1, dropout=0.0, time_norm=17945.14, agg="mean"),
                            conv_kwargs=dict(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32, dropout_prob=0.0))
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == ref


def test_flat_params_keep_names_and_values():
    from maven_b200.models_multimodal import LightCurveImageCLIP
    torch.manual_seed(0)
    m = LightCurveImageCLIP(logit_scale=19.5, nband=2, loss="softmax", combinations=["lightcurve", "spectral"],
                            transformer_kwargs=dict(n_out=32, emb=32, heads=4, depth=2, dropout=0.0, time_norm=2e4, agg="mean"),
                            transformer_spectral_kwargs=dict(n_out=32, emb=32, heads=2, depth=1, dropout=0.0, time_norm=2e4, agg="mean"))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    g = m.flat_group()
    flat = g.ensure()
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    after = m.state_dict()
    assert list(after) == list(before)
    for k in before:
        assert torch.equal(before[k], after[k]), k
    base = flat.data_ptr()
    for p, o in zip(g.params, g.offsets):
        assert p.data_ptr() == base + 4 * o
    # in-place updates of the flat buffer are visible through the named parameters; load_state_dict keeps the views
    flat.mul_(2.0)
    assert torch.equal(m.logit_bias.detach(), before["logit_bias"] * 2)
    m.load_state_dict(before)
    assert g.ensure() is flat and torch.equal(m.lightcurve_projection.weight.detach(), before["lightcurve_projection.weight"])
    i0, i1 = m._segments["lightcurve"]
    assert g.offsets[i0] % 4 == 0 and g.offsets[m._segments["spectral"][0]] % 4 == 0


def test_product_refuses_cpu_tensors(built):
    from maven_b200.loss import clip_loss
    from maven_b200.transformer_utils import TransformerWithTimeEmbeddings
    enc = TransformerWithTimeEmbeddings(n_out=8, nband=1, emb=16, heads=2, depth=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(2, 6, 1), torch.zeros(2, 6), torch.ones(2, 6, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clip_loss(torch.randn(4, 128), torch.randn(4, 128), torch.tensor(1.0), torch.tensor(0.0))


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "multimodal-supernovae_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py") and f != "selfcheck.py":
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f


def test_graph_and_augment_refuse_cpu(built):
    """The whole-step graph and the device-side augmentation are CUDA-only like everything else: CPU inputs raise."""
    import torch
    from maven_b200.augment import DeviceAugment, augment_images, augment_seq
    from maven_b200.graph import GraphedTrainStep
    from maven_b200.models_multimodal import LightCurveImageCLIP
    m = LightCurveImageCLIP(combinations=["lightcurve"], nband=2, loss="softmax", regression=True,
                            transformer_kwargs=dict(n_out=8, emb=16, heads=2, depth=1, time_norm=100.0, agg="mean"))
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedTrainStep(m, None, [None] * 9)
    with pytest.raises(RuntimeError, match="CUDA|CPU"):
        augment_seq(torch.zeros(3), torch.zeros(3), 1.0)
    with pytest.raises(RuntimeError, match="CUDA|CPU"):
        augment_images(torch.zeros(1, 3, 4, 4), None)
    with pytest.raises(ValueError, match="unsupported combination"):
        DeviceAugment(["meta"], 0.1, 1.0)
    with pytest.raises(ValueError, match="expected 3 tensors"):
        DeviceAugment(["host_galaxy"], 0.1, 1.0)([torch.zeros(1)])


def test_device_augment_field_orders_match_reference_tuple_lengths():
    """NoisyDataLoader asserts the dataset tuple length per combination set (src/dataloader.py:64-87): 3 = images only,
    6 = one sequence modality, 7 = images + one sequence modality, 10 = light curve + spectra, 11 = all three."""
    from maven_b200.augment import _FIELDS
    want = {frozenset(["host_galaxy"]): 3, frozenset(["lightcurve"]): 6, frozenset(["spectral"]): 6,
            frozenset(["host_galaxy", "lightcurve"]): 7, frozenset(["host_galaxy", "spectral"]): 7,
            frozenset(["spectral", "lightcurve"]): 10, frozenset(["host_galaxy", "spectral", "lightcurve"]): 11}
    assert {k: len(v) for k, v in _FIELDS.items()} == want
    for fields in _FIELDS.values():
        assert fields[-2:] == ("redshift", "classification")
        if "mag" in fields:
            i = fields.index("mag")
            assert fields[i:i + 4] == ("mag", "time", "mask", "magerr")
        if "spec" in fields:
            i = fields.index("spec")
            assert fields[i:i + 4] == ("spec", "freq", "maskspec", "specerr")


def test_oracle_auc_matches_reference_fixture():
    """N3 oracle pinned to get_ROC_data / get_AUC outputs of the unmodified reference (tests/golden/make_golden_r2.py)."""
    import numpy as np
    from conftest import load_golden
    from oracle import maven_oracle as O
    for name in ("auc", "auc_small"):
        g = load_golden(name)
        thr, frac = O.roc_data(g["e1"], g["e2"])
        assert np.array_equal(thr, g["thresholds"].numpy()) and np.array_equal(frac, g["fraction_correct"].numpy())
        assert abs(O.auc(g["e1"], g["e2"]) - float(g["auc"])) < 1e-12


def test_device_augment_ignores_meta_and_rejects_unknown():
    from maven_b200.augment import DeviceAugment
    a = DeviceAugment(["lightcurve", "spectral", "meta"], 0.1, 1.0)
    assert a.fields == DeviceAugment(["lightcurve", "spectral"], 0.1, 1.0).fields
    with pytest.raises(ValueError):
        DeviceAugment(["meta"], 0.1, 1.0)


def test_dp_weighted_mean_is_identity_without_group():
    from maven_b200 import ops
    x = torch.tensor(3.0, requires_grad=True)
    assert ops.dp_weighted_mean(x, 5.0) is x


def test_pretraining_masks_match_reference_golden():
    """get_random_mask / get_continous_random_mask (src/models_pretraining.py:17-103) consume torch's / Python's random streams in
    the reference's order: with the fixture's seeds they select exactly the positions the unmodified reference selected."""
    import random

    from conftest import load_golden
    from maven_b200.models_pretraining import get_continous_random_mask, get_random_mask
    g = load_golden("pretrain_masks")
    torch.manual_seed(int(g["torch_seed"]))
    m_in, m_pred = get_random_mask(g["mask"], f_mask=float(g["f_rand"]))
    assert torch.equal(m_in, g["rand_in"]) and torch.equal(m_pred, g["rand_pred"])
    random.seed(int(g["random_seed"]))
    m_in, m_pred = get_continous_random_mask(g["mask"], int(g["nband"]), f_mask=float(g["f_cont"]))
    assert torch.equal(m_in, g["cont_in"]) and torch.equal(m_pred, g["cont_pred"])
    # every selected position is valid, the two masks partition the padding mask
    assert torch.equal(m_in | m_pred, g["mask"]) and not (m_in & m_pred).any()


def test_nvtx_ranges_are_opt_in():
    """MVN_NVTX=1 wraps the library call groups in named NVTX ranges; without it the functions are left untouched (no cost)."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import maven_b200.ops as o, maven_b200.graph as g, maven_b200.optim as r; "
            "print(hasattr(o.SeqEncoderFn.forward, '__wrapped__'), hasattr(o.ClipLossMultiFn.backward, '__wrapped__'), "
            "hasattr(g.GraphedTrainStep.__call__, '__wrapped__'), hasattr(r.FusedRAdam.step, '__wrapped__'))" % ROOT)
    for flag, want in (("1", "True True True True"), ("0", "False False False")):
        env = dict(os.environ, MVN_NVTX=flag)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-400:]
        assert out.stdout.strip().startswith(want), (flag, out.stdout)


def _sass_functions(lib_path, name_filter):
    """{function name: [instruction text, ...]} of the sm_100a SASS in `lib_path` for functions whose mangled name contains `name_filter`."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", lib_path], capture_output=True, text=True, timeout=600).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1) if name_filter in m.group(1) else None
            if cur:
                funcs[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if cur and m:
            funcs[cur].append(m.group(1).strip())
    return funcs


def test_tma_slot_release_is_fenced_after_generic_reads(built):
    """A shared-memory slot that TMA refills may only be released (mbarrier arrive) after the generic-proxy loads that read it have been
    ordered by a proxy fence: in SASS, no `SYNCS.ARRIVE` within a few instructions of a group of `LD.E.128` row loads without a
    `FENCE.VIEW.ASYNC` in between.  Without it the arrive issues with the loads in flight and an early refill is read by the later loads
    (profiles/r02c_aux_ring_experiments.md: 11 / 20 failing runs -> 0 / 20)."""
    lib_path = os.path.join(ROOT, "multimodal-supernovae_b200", "libmaven_sm100.so")
    checked = 0
    for pat in ("tc_gemm_kernel", "tc_ffn_fwd_kernel"):
        for name, ins in _sass_functions(lib_path, pat).items():
            for i, text in enumerate(ins):
                if "SYNCS.ARRIVE" not in text or "A1T0" not in text:
                    continue
                window = ins[max(0, i - 12):i]
                loads = [k for k, t in enumerate(window) if t.split()[0].lstrip("@!P0123456 ").startswith(("LD.E.128", "LDS.128"))]
                if not loads:
                    continue
                checked += 1
                assert any("FENCE.VIEW.ASYNC" in t for t in window[loads[-1]:]), f"{name}: release at instruction {i} right after row loads without a proxy fence"
    assert checked >= 2, "the aux-row / residual-row release sites were not found in the SASS (did the kernels change shape?)"
