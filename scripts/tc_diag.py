import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maven_b200 import _lib
if os.environ.get('MVN_DIAG_LIB'):
    _lib.LIB_PATH = os.environ['MVN_DIAG_LIB']
L = _lib.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda:0")
REPS = int(os.environ.get("DIAG_REPS", "10"))
for (M, N, K, mode) in [(60001, 64, 256, "addend")] * REPS + [(60001, 64, 256, "relu_mask")] * REPS:
    torch.manual_seed(N * K)
    dy = torch.randn(M, N); w = torch.randn(N, K) / math.sqrt(N); add = torch.randn(M, K); src = torch.randn(M, K)
    ref = dy.double() @ w.double()
    if mode == "addend": ref = ref + add.double()
    if mode == "relu_mask": ref = ref * (src.double() > 0)
    dyg, wg, addg, srcg = (v.to(dev) for v in (dy, w, add, src))
    dx = torch.full((M, K), float("nan"), device=dev)
    rc = L.mvn_linear_bwd_input(P(dyg), P(wg), P(dx), P(addg) if mode == "addend" else None, P(srcg) if mode == "relu_mask" else None,
                                1 if mode == "relu_mask" else 0, None, M, N, K, 1, S())
    torch.cuda.synchronize()
    d = (dx.cpu().double() - ref).abs()
    bad = d > 0.05
    print(M, N, K, mode, "rc", rc, "bad frac", bad.float().mean().item(), "nan", torch.isnan(dx).sum().item(), "marks", (dx.abs() > 1e5).sum().item())
    if bad.any():
        tiles = bad.view(-1)[: (M // 128) * 128 * K].view(M // 128, 128, K // 32, 32).any(dim=3).any(dim=1)   # [tile, chunk]
        bt = tiles.any(dim=1).nonzero().flatten()
        print("  bad tiles:", bt[:24].tolist(), "... count", len(bt), " CTA-seq of bad tiles (tile//148):", sorted(set((bt // 148).tolist())))
        print("  bad chunks in first bad tiles:", [tiles[t].nonzero().flatten().tolist() for t in bt[:6]])
        t0 = int(bt[0]); r = bad[t0 * 128:(t0 + 1) * 128].any(dim=1).nonzero().flatten()
        print("  bad rows in tile", t0, ":", r.tolist(), "count", len(r))
        # which columns of a bad row are wrong (one character per 16-byte chunk of the 128-byte row: x = wrong), and, for the ReLU-mask
        # mode (output = gemm or 0), whether a wrong element is a flipped mask (the aux row was wrong) or neither value (accumulator / box)
        base_full = dy.double() @ w.double()
        ch0 = int(tiles[t0].nonzero().flatten()[0])
        for row in r[:6].tolist():
            gr = t0 * 128 + row
            e = d[gr, ch0 * 32:(ch0 + 1) * 32] > 0.05
            print("    row", row, "chunk", ch0, "wrong 16-byte chunks:", "".join("x" if e[4 * q:4 * q + 4].any() else "." for q in range(8)), "wrong elements", int(e.sum()))
        if mode == "addend":
            # element-wise: do the WRONG elements of a row carry the aux values of a LATER fill of the same ring slot (chunk ch0 + 4k of this
            # tile, or a chunk of the CTA's next tile)?  Then the row was read while / after the slot was being refilled.
            for row in r[:6].tolist():
                gr = t0 * 128 + row
                e = d[gr, ch0 * 32:(ch0 + 1) * 32] > 0.05
                used = dx[gr, ch0 * 32:(ch0 + 1) * 32].cpu().double() - base_full[gr, ch0 * 32:(ch0 + 1) * 32]
                hits = []
                for tt in (t0, t0 + 148):
                    if (tt + 1) * 128 > M: continue
                    for cc in range(K // 32):
                        cand = add[tt * 128 + row, cc * 32:(cc + 1) * 32].double()
                        if ((cand - used).abs()[e] < 1e-3).all(): hits.append((tt - t0, cc))
                print("    row", row, ": the", int(e.sum()), "wrong elements equal aux[(tile offset, chunk)] =", hits if hits else "nothing in this CTA's two first tiles")
        if mode == "relu_mask":
            got = dx.cpu().double()
            flipped_off = (bad & (got.abs() < 1e-6)).sum().item()
            flipped_on = (bad & ((got - base_full).abs() < 0.05) & (ref == 0)).sum().item()
            print("    relu_mask: wrong elements", int(bad.sum()), "of which masked though src > 0:", flipped_off, ", passed though src <= 0:", flipped_on,
                  ", neither 0 nor the gemm value:", int(bad.sum()) - flipped_off - flipped_on)
        if mode == "addend":
            # where does the garbage come from?  Search every (tile, chunk) of the expected OUTPUT and of the aux tensor, same row of the box
            ch0 = int(tiles[t0].nonzero().flatten()[0])
            nt, nc = M // 128, K // 32
            ref4 = ref[: nt * 128].view(nt, 128, nc, 32); add4 = add.double()[: nt * 128].view(nt, 128, nc, 32)
            base4 = (dy.double() @ w.double())[: nt * 128].view(nt, 128, nc, 32)
            for row in r[:4].tolist():
                got = dx[t0 * 128 + row, ch0 * 32:(ch0 + 1) * 32].cpu().double()
                for name, cand in (("output", ref4[:, row]), ("aux", add4[:, row]), ("gemm", base4[:, row]), ("aux-part", add4[:, row] + base4[t0, row, ch0])):
                    e = (cand - got).abs().amax(dim=-1)                       # [tile, chunk]
                    i = int(e.argmin()); print("    row", row, "best match in", name, "-> tile", i // nc, "chunk", i % nc, "err %.2e" % e.view(-1)[i].item())
            ch = int(tiles[t0].nonzero().flatten()[0])
            base = (dy.double() @ w.double())
            for row in r[:6].tolist():
                gr = t0 * 128 + row
                used = dx[gr, ch * 32:(ch + 1) * 32].cpu().double() - base[gr, ch * 32:(ch + 1) * 32]
                # which (tile, chunk) of `add` matches what the kernel added?
                best = None
                for tt in (t0, t0 + 148, t0 + 296):
                    if (tt + 1) * 128 > M: continue
                    for cc in range(K // 32):
                        e = (add[tt * 128 + row, cc * 32:(cc + 1) * 32].double() - used).abs().max().item()
                        if best is None or e < best[0]: best = (e, tt, cc)
                print("    row", row, "chunk", ch, "-> kernel added add[tile %d, chunk %d] (err %.2e)" % (best[1], best[2], best[0]))
