"""World-size-2 `gloo` test (CPU) of the data-parallel host logic.

The exchange protocol in maven_b200.ops.ClipLossFn -- all-gather embeddings, rank-local LSEs, all-gather LSE vectors,
all-reduce the loss share, then one flat gradient all-reduce -- is driven on two CPU processes.  The two CUDA entry
points it calls per rank (mvn_clip_loss_fwd / _bwd) are replaced BY THE TEST with dense torch restatements of their
header contract (row_offset semantics and all), so what is checked here is ordering, offsets and reductions of the
host code; the kernels themselves are checked on the GPU (test_clip_loss_sharded_rows_match_global).
Parity definition: 2-rank loss and gradients == the single-process oracle on the concatenated batch.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, prec):
    s = ls.exp()
    z_rows = (e2 @ e1_all.T) * s + lb              # local rows i of Z, all columns
    z_cols = (e2_all @ e1.T) * s + lb              # all rows, local columns j of Z
    lse_row = torch.logsumexp(z_rows, dim=1)
    lse_col = torch.logsumexp(z_cols, dim=0)
    diag = z_rows[torch.arange(n), off + torch.arange(n)]
    loss = ((lse_row - diag).sum() + (lse_col - diag).sum()) / (2 * N)
    return loss.reshape(1), torch.stack([lse_row, lse_col])


def _bwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, lse_all, g, prec):
    s = ls.exp()
    idx = off + torch.arange(n)
    z_rows = (e2 @ e1_all.T) * s + lb                                   # (n, N): rows local
    g_rows = (torch.exp(z_rows - lse_all[0][idx][:, None]) + torch.exp(z_rows - lse_all[1][None, :])) / (2 * N)
    g_rows[torch.arange(n), idx] -= 1.0 / N
    z_cols = (e2_all @ e1.T) * s + lb                                   # (N, n): columns local
    g_cols = (torch.exp(z_cols - lse_all[0][:, None]) + torch.exp(z_cols - lse_all[1][idx][None, :])) / (2 * N)
    g_cols[idx, torch.arange(n)] -= 1.0 / N
    d2 = g * s * (g_rows @ e1_all)
    d1 = g * s * (g_cols.T @ e2_all)
    dls = g * (g_rows * (z_rows - lb)).sum()
    return d1, d2, dls.reshape(1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maven_b200 import ops
    from maven_b200.loss import clip_loss_multimodal
    ops._clip_fwd_local, ops._clip_bwd_local = _fwd_local, _bwd_local
    ops.set_data_parallel_group(dist.group.WORLD)
    torch.manual_seed(0)
    N, D, n = 24, 16, 24 // world
    embs = [torch.nn.functional.normalize(torch.randn(N, D, dtype=torch.float64), dim=-1) for _ in range(3)]
    ls = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
    lb = torch.tensor(-2.0, dtype=torch.float64, requires_grad=True)
    local = [e[rank * n:(rank + 1) * n].clone().requires_grad_() for e in embs]
    loss = clip_loss_multimodal(local, ls, lb)
    loss.backward()
    flat = torch.cat([ls.grad.reshape(1), lb.grad.reshape(1)])         # replicated parameters: one flat all-reduce
    dist.all_reduce(flat)
    q.put((rank, loss.item(), [t.grad.tolist() for t in local], flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_clip_loss_matches_single_process_oracle():
    from oracle import maven_oracle as O
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    N, D = 24, 16
    embs = [torch.nn.functional.normalize(torch.randn(N, D, dtype=torch.float64), dim=-1).requires_grad_() for _ in range(3)]
    ls = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
    lb = torch.tensor(-2.0, dtype=torch.float64, requires_grad=True)
    ref = O.clip_loss_multimodal(embs, ls, lb)
    ref.backward()
    n = N // world
    for rank, loss, grads, flat in res:
        assert abs(loss - ref.item()) < 1e-12                          # every rank reports the GLOBAL loss
        for m in range(3):
            assert torch.allclose(torch.tensor(grads[m], dtype=torch.float64), embs[m].grad[rank * n:(rank + 1) * n], atol=1e-12)
        assert abs(flat[0] - ls.grad.item()) < 1e-12
        assert abs(flat[1]) < 1e-12 and abs(lb.grad.item()) < 1e-12      # logit_bias: zero gradient


def _heads_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maven_b200 import ops
    ops.set_data_parallel_group(dist.group.WORLD)
    torch.manual_seed(1)
    n_per = [5, 3]                                                     # unequal shards: the weights, not the rank count, normalise
    logits = torch.randn(sum(n_per), 5, dtype=torch.float64)
    labels = torch.randint(0, 5, (sum(n_per),))
    w = torch.tensor([0.3, 0.08, 1.0, 0.01, 0.2], dtype=torch.float64)
    pred = torch.randn(sum(n_per), dtype=torch.float64); tgt = torch.randn(sum(n_per), dtype=torch.float64)
    lo = sum(n_per[:rank]); hi = lo + n_per[rank]
    lg = logits[lo:hi].clone().requires_grad_()
    local_ce = torch.nn.functional.cross_entropy(lg, labels[lo:hi], weight=w)           # what mvn_weighted_ce_fwd returns per rank
    ce = ops.dp_weighted_mean(local_ce, w[labels[lo:hi]].sum())
    ce.backward()
    pr = pred[lo:hi].clone().requires_grad_()
    mse = ops.dp_weighted_mean(torch.nn.functional.mse_loss(pr, tgt[lo:hi]), float(hi - lo))
    mse.backward()
    q.put((rank, ce.item(), lg.grad.tolist(), mse.item(), pr.grad.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_heads_match_single_process():
    """MSE / weighted-CE heads under data parallelism == the single-process loss on the concatenated batch, and the per-rank
    gradients are the corresponding rows of the global gradient (so the SUM all-reduce of parameter gradients is exact)."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_heads_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(1)
    n_per = [5, 3]
    logits = torch.randn(sum(n_per), 5, dtype=torch.float64, requires_grad=True)
    labels = torch.randint(0, 5, (sum(n_per),))
    w = torch.tensor([0.3, 0.08, 1.0, 0.01, 0.2], dtype=torch.float64)
    pred = torch.randn(sum(n_per), dtype=torch.float64, requires_grad=True); tgt = torch.randn(sum(n_per), dtype=torch.float64)
    ce = torch.nn.functional.cross_entropy(logits, labels, weight=w); ce.backward()
    mse = torch.nn.functional.mse_loss(pred, tgt); mse.backward()
    for rank, ce_r, dlg, mse_r, dpr in res:
        lo = sum(n_per[:rank]); hi = lo + n_per[rank]
        assert abs(ce_r - ce.item()) < 1e-12 and abs(mse_r - mse.item()) < 1e-12
        assert torch.allclose(torch.tensor(dlg, dtype=torch.float64), logits.grad[lo:hi], atol=1e-12)
        assert torch.allclose(torch.tensor(dpr, dtype=torch.float64), pred.grad[lo:hi], atol=1e-12)


def _overlap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maven_b200 import ops
    ops.set_data_parallel_group(dist.group.WORLD)
    torch.manual_seed(rank)
    g = torch.randn(97)
    mine = g.clone()
    ops.begin_grad_overlap(g)
    ops._segment_done(g, 10, 20)          # "encoder A finished its backward": its segment starts reducing now
    ops._segment_done(g, 50, 7)           # encoder B
    ops.finish_grad_reduce(g)             # waits for both, reduces [0,10), [30,50), [57,97)
    q.put((rank, mine.tolist(), g.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_segment_reduce_equals_one_all_reduce():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = torch.tensor(res[0][1]) + torch.tensor(res[1][1])
    for _, _, got in res:
        assert torch.allclose(torch.tensor(got), total, atol=1e-6)
