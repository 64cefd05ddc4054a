#!/usr/bin/env python
"""Per-kernel device timing through the C ABI at the C4 token counts (CUDA events, L2 flushed between launches).
Prints algorithmic GB/s and TFLOP/s per op; used to tune individual kernels.  Not the bench contract (bench.py is)."""
import argparse
import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maven_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes, flops, extra=""):
    print(f"{name:44s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.0f} GB/s  {flops/ms/1e9:8.1f} TFLOP/s {extra}", flush=True)


def gemms(M, E, prec):
    F = 4 * E
    x = torch.randn(M, E, device=dev); h = torch.randn(M, F, device=dev); qkv = torch.randn(M, 3 * E, device=dev)
    wqkv = torch.randn(3 * E, E, device=dev) / 8; wu = torch.randn(E, E, device=dev) / 8
    w1 = torch.randn(F, E, device=dev) / 8; w2 = torch.randn(E, F, device=dev) / 16
    b = torch.randn(F, device=dev); g = torch.ones(E, device=dev)
    y3 = torch.empty(M, 3 * E, device=dev); yE = torch.empty(M, E, device=dev); yF = torch.empty(M, F, device=dev)
    xhat = torch.empty(M, E, device=dev); rstd = torch.empty(M, device=dev)
    tag = f"M={M} E={E} prec={prec}"
    ms = timeit(lambda: L.mvn_linear_fwd(P(x), P(wqkv), None, P(y3), None, M, 3 * E, E, 0, prec, S()))
    report(f"qkv fwd      [{tag}]", ms, M * 4 * (E + 3 * E), 2 * M * E * 3 * E)
    ms = timeit(lambda: L.mvn_linear_res_ln_fwd(P(x), P(wu), P(b), P(yE), P(g), P(g), P(yE), P(xhat), P(rstd), None, M, E, E, 1e-5, prec, S()))
    report(f"unify+res+LN [{tag}]", ms, M * 4 * (E + E + 2 * E), 2 * M * E * E)
    ms = timeit(lambda: L.mvn_linear_fwd(P(x), P(w1), P(b), P(yF), None, M, F, E, 1, prec, S()))
    report(f"ff1+relu     [{tag}]", ms, M * 4 * (E + F), 2 * M * E * F)
    ms = timeit(lambda: L.mvn_linear_res_ln_fwd(P(h), P(w2), P(b), P(x), P(g), P(g), P(yE), P(xhat), P(rstd), None, M, E, F, 1e-5, prec, S()))
    report(f"ff2+res+LN   [{tag}]", ms, M * 4 * (F + E + 2 * E), 2 * M * E * F)
    ms = timeit(lambda: L.mvn_linear_bwd_input(P(x), P(w2), P(yF), None, P(h), 1, None, M, E, F, prec, S()))
    report(f"dgrad ff2 (relu mask) [{tag}]", ms, M * 4 * (E + 2 * F), 2 * M * E * F)
    ms = timeit(lambda: L.mvn_linear_bwd_input(P(h), P(w1), P(yE), P(x), None, 0, None, M, F, E, prec, S()))
    report(f"dgrad ff1 (+res)      [{tag}]", ms, M * 4 * (F + 2 * E), 2 * M * E * F)
    ms = timeit(lambda: L.mvn_linear_bwd_input(P(qkv), P(wqkv), P(yE), P(x), None, 0, None, M, 3 * E, E, prec, S()))
    report(f"dgrad qkv (+res)      [{tag}]", ms, M * 4 * (3 * E + 2 * E), 2 * M * E * 3 * E)
    for (nm, dy, xx, N, K) in (("wgrad qkv", qkv, x, 3 * E, E), ("wgrad ff1", h, x, F, E), ("wgrad ff2", x, h, E, F), ("wgrad unify", x, x, E, E)):
        wsb = L.mvn_linear_bwd_weight_workspace_bytes(M, N, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        dw = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
        ms = timeit(lambda: L.mvn_linear_bwd_weight(P(dy), P(xx), P(dw), P(db), None, M, N, K, 0, P(ws), wsb, prec, S()))
        report(f"{nm:12s} [{tag}]", ms, M * 4 * (N + K), 2 * M * N * K)
        ms = timeit(lambda: L.mvn_linear_bwd_weight(P(dy), P(xx), P(dw), None, None, M, N, K, 0, P(ws), wsb, prec, S()))
        report(f"{nm:12s} nobias [{tag}]", ms, M * 4 * (N + K), 2 * M * N * K)


def attention(B, T, E, H, nmin, nmax, prec):
    gen = torch.Generator().manual_seed(0)
    n = torch.randint(nmin, nmax + 1, (B,), generator=gen)
    cu = torch.cat([torch.zeros(1, dtype=torch.int64), n.cumsum(0)]).int().to(dev)
    M = int(n.sum())
    qkv = torch.randn(M, 3 * E, device=dev); out = torch.empty(M, E, device=dev); lse = torch.empty(M, H, device=dev)
    dout = torch.randn(M, E, device=dev); dqkv = torch.empty(M, 3 * E, device=dev)
    sc = 1 / math.sqrt(E)
    n2 = float((n.double() ** 2).sum())
    tag = f"B={B} E={E} H={H} M={M} prec={prec}"
    ms = timeit(lambda: L.mvn_attention_fwd(P(qkv), P(cu), None, P(out), P(lse), B, E, H, sc, prec, S()))
    report(f"attn fwd [{tag}]", ms, M * 4 * (4 * E + H), 4 * n2 * E)
    ms = timeit(lambda: L.mvn_attention_bwd(P(qkv), P(cu), None, P(out), P(lse), P(dout), P(dqkv), B, E, H, sc, prec, S()))
    report(f"attn bwd [{tag}]", ms, M * 4 * (8 * E + H), 10 * n2 * E)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,attn")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precs", default="0,1")
    a = ap.parse_args()
    B = a.batch
    for prec in [int(p) for p in a.precs.split(",")]:
        if "fixed" in a.what:              # one / two tiles per CTA: the per-launch fixed cost of the persistent kernels
            gemms(128 * 148, 64, prec)
            gemms(2 * 128 * 148, 64, prec)
        if "gemm" in a.what:
            gemms(B * 120, 64, prec)
            gemms(B * 165, 32, prec)
        if "attn" in a.what:
            attention(B, 200, 64, 8, 40, 200, prec)
            attention(B, 220, 32, 2, 110, 220, prec)
