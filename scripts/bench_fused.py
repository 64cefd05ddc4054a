"""Device-time micro-benchmark of the fused block kernels at the C4 token counts (CUDA events, L2 flushed between
launches).  Usage: python scripts/bench_fused.py [ffn|attn]"""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maven_b200 import _lib
L = _lib.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def ffn():
    for (M, E, layers) in [(121563, 64, 5), (168960, 32, 13)]:
        F = 4 * E
        torch.manual_seed(0)
        x = torch.randn(M, E, device=dev); dy = torch.randn(M, E, device=dev)
        w1 = torch.randn(F, E, device=dev) / math.sqrt(E); b1 = torch.randn(F, device=dev); w2 = torch.randn(E, F, device=dev) / math.sqrt(F)
        b2 = torch.randn(E, device=dev); g = torch.ones(E, device=dev); b = torch.zeros(E, device=dev)
        y = torch.empty(M, E, device=dev); xhat = torch.empty(M, E, device=dev); rstd = torch.empty(M, device=dev); dx = torch.empty(M, E, device=dev)
        wsb = L.mvn_ffn_fused_bwd_workspace_bytes(E, 4); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        gr = [torch.empty_like(t) for t in (w1, b1, w2, b2, g, b)]
        fwd = lambda: L.mvn_ffn_fused_fwd(P(x), P(w1), P(b1), P(w2), P(b2), P(g), P(b), P(y), P(xhat), P(rstd), None, M, E, 4, 1e-5, 0.0, 0, 0, S())
        bwd = lambda: L.mvn_ffn_fused_bwd(P(dy), P(xhat), P(rstd), P(x), P(w1), P(b1), P(w2), P(g), P(dx), P(gr[0]), P(gr[1]), P(gr[2]), P(gr[3]), P(gr[4]),
                                          P(gr[5]), None, M, E, 4, 0.0, 0, 0, P(ws), wsb, S())
        assert fwd() == 0 and bwd() == 0, L.mvn_last_error()
        tf, tb = timeit(fwd), timeit(bwd)
        fl_f, fl_b = 4 * M * E * F, 10 * M * E * F
        print(f"ffn E={E} M={M}: fwd {tf*1e3:.1f} us ({fl_f/tf/1e9:.1f} TFLOP/s, {3*M*E*4/tf/1e6:.0f} GB/s)  bwd(+reduce) {tb*1e3:.1f} us ({fl_b/tb/1e9:.1f} TFLOP/s)"
              f"  -> per step x{layers}: fwd {tf*layers:.3f} ms bwd {tb*layers:.3f} ms", flush=True)


def attn():
    """packed-stream attention at the C4 shapes: lengths like bench.py's generator (LC 2 x U{20..100}, SP U{110..220})."""
    g = torch.Generator().manual_seed(0)
    for (name, E, H, lens, layers) in [("lc", 64, 8, (torch.randint(20, 101, (1024, 2), generator=g).sum(1)), 5),
                                       ("sp", 32, 2, torch.randint(110, 221, (1024,), generator=g), 13)]:
        B, M = lens.numel(), int(lens.sum())
        cu = torch.zeros(B + 1, dtype=torch.int32); cu[1:] = lens.cumsum(0)
        cu = cu.to(dev)
        qkv = torch.randn(M, 3 * E, device=dev); dout = torch.randn(M, E, device=dev)
        out = torch.empty(M, E, device=dev); lse = torch.empty(M, H, device=dev); dqkv = torch.empty(M, 3 * E, device=dev)
        sc = 1 / math.sqrt(E)
        fwd = lambda: L.mvn_attention_fwd(P(qkv), P(cu), None, P(out), P(lse), B, E, H, sc, 1, S())
        bwd = lambda: L.mvn_attention_bwd(P(qkv), P(cu), None, P(out), P(lse), P(dout), P(dqkv), B, E, H, sc, 1, S())
        assert fwd() == 0 and bwd() == 0, L.mvn_last_error()
        tf, tb = timeit(fwd), timeit(bwd)
        n2 = float((lens.double() ** 2).sum())
        print(f"attn {name} E={E} H={H} M={M}: fwd {tf*1e3:.1f} us ({4*n2*E/tf/1e9:.1f} TFLOP/s)  bwd {tb*1e3:.1f} us ({10*n2*E/tb/1e9:.1f} TFLOP/s)"
              f"  -> per step x{layers}: fwd {tf*layers:.3f} ms bwd {tb*layers:.3f} ms", flush=True)


def loss():
    """one rank's share of the C5 loss at a global batch of 65 536 (n = 8 192 local rows), per pair: forward (2 directions) and backward"""
    D = 128
    for (n, N) in [(1024, 1024), (8192, 8192), (8192, 65536)]:
        torch.manual_seed(0)
        e1 = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1); e2 = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1)
        ls = torch.tensor([math.log(19.5)], device=dev); lb = torch.tensor([-10.0], device=dev)
        wsb = L.mvn_clip_loss_workspace_bytes(n, N, D); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        loss_ = torch.empty(1, device=dev); lse = torch.empty(2, N, device=dev)
        d1 = torch.empty(n, D, device=dev); d2 = torch.empty(n, D, device=dev); dls = torch.empty(1, device=dev)
        for prec in (0, 1):
            fwd = lambda: L.mvn_clip_loss_fwd(P(e1), P(e2), P(e1), P(e2), n, N, D, 0, P(ls), P(lb), P(loss_), P(lse[0]), P(lse[1]), P(ws), wsb, prec, S())
            bwd = lambda: L.mvn_clip_loss_bwd(P(e1), P(e2), P(e1), P(e2), n, N, D, 0, P(ls), P(lb), P(lse[0]), P(lse[1]), None, P(d1), P(d2), P(dls), P(ws), wsb, prec, S())
            if prec == 0 and n * N > 1 << 27:
                continue
            assert fwd() == 0 and bwd() == 0, L.mvn_last_error()
            tf, tb = timeit(fwd, 5), timeit(bwd, 5)
            print(f"clip loss n={n} N={N} prec={prec}: fwd {tf*1e3:.0f} us ({4*n*N*D/tf/1e9:.0f} TFLOP/s)  bwd {tb*1e3:.0f} us ({8*n*N*D/tb/1e9:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ffn"
    globals()[which]()
