#!/usr/bin/env python
"""Host-side cost of enqueuing one training step: wall time to enqueue (no sync) and a cProfile of where it goes."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from maven_b200.models_multimodal import LightCurveImageCLIP
from maven_b200.transformer_utils import set_precision

wname = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda:0")
wl = bench.WORKLOADS[wname]
torch.manual_seed(0)
model = set_precision(LightCurveImageCLIP(**bench.model_kwargs(wl, 0.0002)).to(dev).train(), "tf32")
opt = model.configure_optimizers()["optimizer"]
batch = [None if v is None else v.to(dev) for v in bench.make_batch(wl, 1024, 1)]

def step():
    loss = model.training_step(batch, 0)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss

for _ in range(5):
    step()
torch.cuda.synchronize()
enq = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); enq.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
print(f"{wname}: enqueue ms/step median {sorted(enq)[len(enq)//2]:.3f}  min {min(enq):.3f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
