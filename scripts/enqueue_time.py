#!/usr/bin/env python
"""CPU time to ENQUEUE one training step (no sync inside) vs its device time -- tells whether the host keeps ahead of the GPU."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from maven_b200.models_multimodal import LightCurveImageCLIP
from maven_b200.transformer_utils import set_precision

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["c4"]
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

for _ in range(3):
    step()
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    l = step()
    t1 = time.perf_counter()
    l.item()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
print(f"MVN_PDL={os.environ.get('MVN_PDL','1')}: enqueue ms/step {sorted(enq)[len(enq)//2]:.2f}  step-to-loss ms {sorted(tot)[len(tot)//2]:.2f}")
