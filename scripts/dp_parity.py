#!/usr/bin/env python
"""torchrun --nproc-per-node R scripts/dp_parity.py : R-GPU data-parallel loss/gradients vs ONE GPU on the concatenated
global batch (the parity definition of DESIGN.md section 5).  Prints relative errors from rank 0; exit code 1 on failure."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from maven_b200 import ops  # noqa: E402
from maven_b200.models_multimodal import LightCurveImageCLIP  # noqa: E402
from maven_b200.transformer_utils import set_precision  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "c4"]
    n = 64
    gbatch = bench.make_batch(wl, n * world, seed=7)                       # the same global batch on every rank
    shard = [None if v is None else v[rank * n:(rank + 1) * n].to(dev) for v in gbatch]

    def build():
        torch.manual_seed(0)
        return set_precision(LightCurveImageCLIP(**bench.model_kwargs(wl, 0.0)).to(dev).train(), prec)

    ops.set_data_parallel_group(dist.group.WORLD)
    m = build()
    loss = m.training_step(shard, 0)
    loss.backward()
    g = m.reduce_gradients().clone()                                       # overlapped per-encoder segments + the rest
    ops.set_data_parallel_group(None)
    ok = True
    if rank == 0:
        m1 = build()
        full = [None if v is None else v.to(dev) for v in gbatch]
        l1 = m1.training_step(full, 0)
        l1.backward()
        g1 = m1.gather_grads()
        el = abs(loss.item() - l1.item()) / abs(l1.item())
        eg = ((g - g1).norm() / g1.norm()).item()
        tol_l, tol_g = (1e-5, 2e-4) if prec == "fp32" else (1e-3, 1e-2)
        ok = el < tol_l and eg < tol_g
        print(f"dp_parity world={world} prec={prec} workload={sys.argv[2] if len(sys.argv) > 2 else 'c4'}: loss dp={loss.item():.7f} single={l1.item():.7f} rel={el:.2e}; flat-grad rel={eg:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
