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
