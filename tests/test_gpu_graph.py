"""GPU tests of the whole-step CUDA graph (maven_b200.graph.GraphedTrainStep): replays must reproduce the
eager training trajectory, leave no trace of the warm-up, and draw fresh dropout masks every replay."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dev():
    return torch.device("cuda:0")


def _make(wname, dropout, B=48):
    import bench
    from maven_b200.models_multimodal import LightCurveImageCLIP
    from maven_b200.transformer_utils import set_precision
    wl = bench.WORKLOADS[wname]
    torch.manual_seed(0)
    kw = bench.model_kwargs(wl, dropout)
    for k in ("transformer_kwargs", "transformer_spectral_kwargs"):
        if k in kw:
            kw[k] = {**kw[k], "depth": 2}
    model = set_precision(LightCurveImageCLIP(**kw).to(dev()).train(), "tf32")
    opt = model.configure_optimizers()["optimizer"]
    batches = [[None if v is None else v.to(dev()) for v in bench.make_batch(wl, B, seed=50 + i)] for i in range(4)]
    return model, opt, batches


@pytest.mark.parametrize("wname", ["c4", "c5", "c2"])      # c2: classifier, logit_scale / logit_bias get no gradient
def test_graphed_steps_match_eager_trajectory(wname):
    from maven_b200 import ops
    from maven_b200.graph import GraphedTrainStep
    m_e, o_e, batches = _make(wname, 0.0)
    losses_e = []
    for b in batches:
        loss = m_e.training_step(b, 0)
        loss.backward()
        o_e.step()
        o_e.zero_grad(set_to_none=True)
        losses_e.append(loss.item())
    m_g, o_g, _ = _make(wname, 0.0)
    try:
        step = GraphedTrainStep(m_g, o_g, batches[0])
        assert step.launches_per_replay > 20
        losses_g = [step(b).item() for b in batches]
        for a, b in zip(losses_e, losses_g):
            assert abs(a - b) <= 1e-6 * abs(a), (losses_e, losses_g)
        sd_e, sd_g = m_e.state_dict(), m_g.state_dict()
        for k in sd_e:
            if sd_e[k].is_floating_point():
                assert (sd_e[k] - sd_g[k]).abs().max().item() <= 1e-6 * (1 + sd_e[k].abs().max().item()), k
            else:
                assert torch.equal(sd_e[k], sd_g[k]), k           # num_batches_tracked: warm-up rolled back, one tick per replay
        assert o_g._steps == o_e._steps
    finally:
        o_g.disable_device_step()
        assert ops._STEP_COUNTER is None


def test_graph_replays_draw_fresh_dropout_masks():
    from maven_b200.graph import GraphedTrainStep
    m, o, batches = _make("c4", 0.3)
    o.param_groups[0]["lr"] = 0.0                      # freeze the weights: only the dropout masks can change the loss
    o.param_groups[0]["weight_decay"] = 0.0
    try:
        step = GraphedTrainStep(m, o, batches[0])
        vals = [step(None).item() for _ in range(4)]
        assert len({round(v, 6) for v in vals}) == 4, vals
    finally:
        o.disable_device_step()


def test_graph_rejects_shape_change():
    from maven_b200.graph import GraphedTrainStep
    m, o, batches = _make("c4", 0.0)
    try:
        step = GraphedTrainStep(m, o, batches[0])
        bad = [None if v is None else v[:8] for v in batches[1]]
        with pytest.raises(ValueError, match="captured shapes"):
            step(bad)
    finally:
        o.disable_device_step()


@pytest.mark.parametrize("wname", ["c4", "c3"])
def test_prefetched_steps_match_plain_replays(wname):
    """double_buffer=True: batches uploaded from pinned host memory on the copy stream into the alternate input set, replayed by the
    alternate graph -- same losses and parameters as the single-graph replays of the same batches (dropout on: both graphs read the
    one device step counter, so step i draws the same masks either way)."""
    from maven_b200.graph import GraphedTrainStep
    m_a, o_a, batches = _make(wname, 0.1)
    m_b, o_b, _ = _make(wname, 0.1)
    for m in (m_a, m_b):                                   # seeds mix in id(module) and a call counter: pin them so that both
        for i, mod in enumerate(m.modules()):               # models (and both captured graphs) hash the same (seed, site, step)
            mod.dropout_seed = 0x5EED0000 + i
    host = [[None if v is None else v.cpu().pin_memory() for v in b] for b in batches]
    try:
        plain = GraphedTrainStep(m_a, o_a, batches[0])
        losses_a = [plain(b).item() for b in batches + batches]
        o_a.disable_device_step()
        pre = GraphedTrainStep(m_b, o_b, batches[0], double_buffer=True)
        seq = host + host
        losses_b = []
        pre.prefetch(seq[0])
        for i in range(len(seq)):
            loss = pre.step_prefetched()
            if i + 1 < len(seq):
                pre.prefetch(seq[i + 1] if i % 2 else (lambda static, b=seq[i + 1]: b))     # tuple and callable forms
            losses_b.append(loss.item())
        for a, b in zip(losses_a, losses_b):
            assert abs(a - b) <= 1e-6 * abs(a), (losses_a, losses_b)
        sd_a, sd_b = m_a.state_dict(), m_b.state_dict()
        for k in sd_a:
            if sd_a[k].is_floating_point():
                assert (sd_a[k] - sd_b[k]).abs().max().item() <= 1e-6 * (1 + sd_a[k].abs().max().item()), k
            else:
                assert torch.equal(sd_a[k], sd_b[k]), k
        assert o_a._steps == o_b._steps
        with pytest.raises(RuntimeError, match="without a prefetch"):
            pre.step_prefetched()
        assert abs(pre(batches[1]).item()) > 0          # the plain call still works on a double-buffered step
    finally:
        o_a.disable_device_step()
        o_b.disable_device_step()


def test_prefetch_needs_double_buffer():
    from maven_b200.graph import GraphedTrainStep
    m, o, batches = _make("c2", 0.0)
    try:
        step = GraphedTrainStep(m, o, batches[0])
        with pytest.raises(RuntimeError, match="double_buffer"):
            step.prefetch(batches[1])
    finally:
        o.disable_device_step()
