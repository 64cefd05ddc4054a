#!/usr/bin/env python
"""bench.py -- CLIP training-step throughput of the B200-native path (and the CPU reference arm).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (default N=1)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm (oracle port) on host cores

A "step" is one full training step (forward + backward + RAdam) of the reference's LightCurveImageCLIP on one batch
of synthetic data shaped like the simulated-pretrain configuration (light curve T=200 two-band E64 h8
L5 + spectra T=220 E32 h2 L13, 1024 samples per GPU).  One JSON line goes to stdout.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (combinations, lc kwargs, sp kwargs, T_lc, T_sp)
    "c4": dict(desc="C4 maven-pretrain shape: lightcurve(E64,h8,L5,T200,2 bands)+spectral(E32,h2,L13,T220) CLIP",
               combinations=["lightcurve", "spectral"],
               lc=dict(n_out=32, emb=64, heads=8, depth=5, time_norm=20583.369161312577, agg="mean"),
               sp=dict(n_out=32, emb=32, heads=2, depth=13, time_norm=17945.142213594805, agg="mean"),
               T_lc=200, T_sp=220),
    "c3": dict(desc="C3 maven-lite bimodal on ZTFBTS-shaped batches: lightcurve(E64,h8,L5,T200,2 bands)+host_galaxy ConvMixer(dim32,depth2,k5,p10, 3x60x60) CLIP",
               combinations=["lightcurve", "host_galaxy"], img=True, lc_dist="ztfbts",
               lc=dict(n_out=32, emb=64, heads=8, depth=5, time_norm=20583.369161312577, agg="mean"), sp=None, T_lc=200, T_sp=0),
    "c5": dict(desc="C5 trimodal: lightcurve(E64,h8,L5,T200)+spectral(E32,h2,L13,T220)+host_galaxy ConvMixer(dim32,depth2,k5,p10, 3x60x60) CLIP, 3 pairs",
               combinations=["lightcurve", "spectral", "host_galaxy"], img=True,
               lc=dict(n_out=32, emb=64, heads=8, depth=5, time_norm=20583.369161312577, agg="mean"),
               sp=dict(n_out=32, emb=32, heads=2, depth=13, time_norm=17945.142213594805, agg="mean"),
               T_lc=200, T_sp=220),
    # configs/maven-lite.yaml as shipped (BASELINE.json configs[0]): attention-pooled light curves + spectra padded to 1024
    "c1": dict(desc="C1 maven-lite as configured: lightcurve(E64,h8,L5,T200,2 bands,agg=attn)+spectral(E32,h2,L13,T1024) CLIP",
               combinations=["lightcurve", "spectral"], lc_dist="ztfbts", sp_dist="c1",
               lc=dict(n_out=32, emb=64, heads=8, depth=5, time_norm=20583.369161312577, agg="attn"),
               sp=dict(n_out=32, emb=32, heads=2, depth=13, time_norm=17945.142213594805, agg="mean"),
               T_lc=200, T_sp=1024),
    "c2": dict(desc="C2 lc_5way_f1 shape: lightcurve(E32,h2,L9,T200) 5-way classifier",
               combinations=["lightcurve"], classification=True, n_classes=5,
               lc=dict(n_out=32, emb=32, heads=2, depth=9, time_norm=3371.17, agg="mean"), sp=None, T_lc=200, T_sp=0),
}
CONV = dict(dim=32, depth=2, channels=3, kernel_size=5, patch_size=10, n_out=32)      # configs/maven-lite.yaml:12-22
LR, WD, LOGIT_SCALE = 3.716367614864064e-05, 0.000555522900788888, 19.545966923442453


# --------------------------------------------------------------------------------------------------
# synthetic data, generated on the CPU with a fixed seed
# --------------------------------------------------------------------------------------------------
def make_seq(gen, B, T, nband, lo, hi, tmax, t0, lognormal=False, c1_spectra=False):
    per = T // nband
    if c1_spectra:    # SURVEY §8d, C1 spectra: valid length 214 (60 %), the full 1024 (30 %), Uniform{1..1024} (10 %)
        u = torch.rand(B, nband, generator=gen)
        n = torch.where(u < 0.6, torch.full((B, nband), min(214, per)), torch.where(u < 0.9, torch.full((B, nband), per),
                        torch.randint(1, per + 1, (B, nband), generator=gen)))
    elif lognormal:     # ZTFBTS shape: n_obs per band ~ clip(round(LogNormal(ln 8, 0.7)), 1, 100)
        n = torch.exp(math.log(8.0) + 0.7 * torch.randn(B, nband, generator=gen)).round().clamp(1, min(hi, per)).long()
    else:
        n = torch.randint(lo, min(hi, per) + 1, (B, nband), generator=gen)
    pos = torch.arange(per)[None, None, :]
    valid = pos < n[:, :, None]                                             # (B, nband, per)
    t = torch.rand(B, nband, per, generator=gen) * tmax
    t = torch.where(valid, t, torch.full_like(t, float("inf"))).sort(dim=-1)[0]
    t = t - t[:, :, :1] + t0
    t = torch.where(valid, t, torch.zeros_like(t))
    x = torch.where(valid, torch.randn(B, nband, per, generator=gen), torch.zeros(B, nband, per))
    return x.reshape(B, T).contiguous(), t.reshape(B, T).contiguous(), valid.reshape(B, T).contiguous()


def make_batch(wl, B, seed):
    gen = torch.Generator().manual_seed(seed)
    x_lc, t_lc, m_lc = make_seq(gen, B, wl["T_lc"], 2, 20, 100, 300.0, 0.0,           # sim-pretrain shape: Uniform{20..100}/band
                                lognormal=wl.get("lc_dist") == "ztfbts")
    if wl["sp"] is not None:
        x_sp, t_sp, m_sp = make_seq(gen, B, wl["T_sp"], 1, 110, 220, 5500.0, 3700.0,   # valid length Uniform{110..220}
                                    c1_spectra=wl.get("sp_dist") == "c1")
    else:
        x_sp = t_sp = m_sp = None
    cls = torch.randint(0, 5, (B,), generator=gen)
    red = torch.rand(B, generator=gen)
    # host-galaxy cut-outs: 8-bit pixel values / 255 in fp32, exactly what load_images produces from the PNGs (src/dataloader.py:326-331)
    x_img = torch.randint(0, 256, (B, 3, 60, 60), generator=gen).float() / 255.0 if wl.get("img") else None
    return [x_img, x_lc, t_lc, m_lc, x_sp, t_sp, m_sp, red, cls]


def model_kwargs(wl, dropout):
    kw = dict(logit_scale=LOGIT_SCALE, lr=LR, nband=2, loss="softmax", optimizer_kwargs={"weight_decay": WD},
              combinations=wl["combinations"], transformer_kwargs={**wl["lc"], "dropout": dropout})
    if wl["sp"] is not None:
        kw["transformer_spectral_kwargs"] = {**wl["sp"], "dropout": dropout}
    if wl.get("img"):
        kw["conv_kwargs"] = {**CONV, "dropout_prob": dropout}       # script_wandb.py:151: the CNN shares cfg.dropout
    if wl.get("classification"):
        kw.update(classification=True, n_classes=wl["n_classes"])
    return kw


def flops_per_step(wl, batch, precision="tf32"):
    """Algorithmic FLOPs and HBM bytes of one training step, two FLOP accountings: dense over the padded T
    (comparable with the reference) and executed (valid tokens only; recomputation inside fused kernels is NOT credited).
    Returned per kernel class; which class a piece of work lands in depends on the tier (`fused` moves the feed-forward
    half of every block out of the gemm / wgrad / row classes into fused_fwd / fused_bwd)."""
    cls = ("gemm", "wgrad", "attn_fwd", "attn_bwd", "fused_fwd", "fused_bwd", "row", "conv")
    out = {"padded_train": 0.0, **{k: 0.0 for k in cls[:6]}, "bytes": {k: 0.0 for k in cls}}
    if wl.get("img"):
        # ConvMixer: ~1.1 MFLOP fwd per sample, train ~3x; bytes = the fp32 image once forward and once more for the
        # patch-conv weight gradient, plus the 128-B embedding
        Bn = batch[0].shape[0]
        out["padded_train"] += 3.3e6 * Bn
        out["bytes"]["conv"] += Bn * (2 * 43200.0 + 128.0)
    for key, m_idx, T in (("lc", 3, wl["T_lc"]), ("sp", 6, wl["T_sp"])):
        kw = wl[key]
        if kw is None:
            continue
        E, L, H = kw["emb"], kw["depth"], kw["heads"]
        fused = precision == "fused" and E in (32, 64)
        nb = batch[m_idx].sum(dim=1).double()
        B = nb.numel()
        out["padded_train"] += 3.0 * B * L * (24.0 * T * E * E + 4.0 * T * T * E)
        M = nb.sum().item()
        n2 = (nb * nb).sum().item()
        out["attn_fwd"] += L * 4.0 * n2 * E
        out["attn_bwd"] += L * 10.0 * n2 * E
        # FLOPs per token-layer: qkv 6E^2, unify 2E^2, ff1 8E^2, ff2 8E^2 forward; the same again for the input gradients and
        # once more for the weight gradients.
        # HBM bytes per token-layer (fp32 storage, every operand once in / once out -- DESIGN.md):
        #   layer-by-layer: GEMM class fwd qkv 4E, unify+res+LN 4E, ff1 5E, ff2+res+LN 7E; dgrad ff2 9E, ff1 6E, unify 2E, qkv 5E (42E);
        #                   wgrad class ff2 5E, ff1 5E, unify 2E, qkv 4E (16E); attention fwd 4E+H, bwd 8E+H; LayerNorm bwd 2 x 3E
        #   fused FFN     : forward x1 in, x2 + xhat2 out (3E + 1); backward dy, xhat2, x1 in, dx1 out (4E + 1)
        if fused:
            out["gemm"] += L * 16.0 * M * E * E
            out["wgrad"] += L * 8.0 * M * E * E
            out["fused_fwd"] += L * 16.0 * M * E * E
            out["fused_bwd"] += L * 32.0 * M * E * E
            out["bytes"]["gemm"] += 4.0 * L * M * 15 * E
            out["bytes"]["wgrad"] += 4.0 * L * M * 6 * E
            out["bytes"]["fused_fwd"] += 4.0 * L * M * (3 * E + 1)
            out["bytes"]["fused_bwd"] += 4.0 * L * M * (4 * E + 1)
            out["bytes"]["row"] += 4.0 * L * M * 3 * E
        else:
            out["gemm"] += L * 48.0 * M * E * E
            out["wgrad"] += L * 24.0 * M * E * E
            out["bytes"]["gemm"] += 4.0 * L * M * 42 * E
            out["bytes"]["wgrad"] += 4.0 * L * M * 16 * E
            out["bytes"]["row"] += 4.0 * L * M * 6 * E
        out["bytes"]["attn_fwd"] += 4.0 * L * M * (4 * E + H)
        out["bytes"]["attn_bwd"] += 4.0 * L * M * (8 * E + H)
    out["executed_train"] = sum(out[k] for k in cls[:6])
    return out


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in a background thread (nvidia_ml_py), falling
    back to an `nvidia-smi -lms` child process when NVML is not importable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_s=0.05):
        self.index, self.period, self.rows, self.proc, self.thread, self.stop_flag = index, period_s, [], None, None, False
        self.max_mhz, self.mode = None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = index_of_visible(self.index)
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            R = pynvml
            bits = {"hw_slowdown": R.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": R.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": R.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": R.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((mhz, [n for n, b in bits.items() if rs & b]))
                    except Exception:
                        pass
                    time.sleep(self.period)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.mode = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.mode = "nvidia-smi"
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if r and r[0].isdigit():
                if len(r) > 1 and r[1].isdigit():
                    self.max_mhz = max(self.max_mhz or 0, int(r[1]))
                self.rows.append((int(r[0]), [n for i, n in enumerate(names) if len(r) > 3 + i and r[3 + i].lower().startswith("active")]))

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.proc is not None:
            self.proc.terminate()
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({n for r in self.rows for n in r[1]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "source": self.mode}


def index_of_visible(local_index):
    """NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = [p.strip() for p in vis.split(",") if p.strip()]
        if local_index < len(parts) and parts[local_index].isdigit():
            return int(parts[local_index])
    return local_index


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        hbm = float(d["hbm_gbs"])
        tf = float(d.get("bf16_tflops_sustained") or d["bf16_tflops"])
        return hbm, tf, "measured (MEASURED_PEAKS.json: hbm_gbs copy bandwidth / sustained bf16)"
    except Exception:                                          # file absent or in another shape: the recipe's stated fallback
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (staged under oracle/_ref by oracle/build_ref.py; framework imports
# stubbed by oracle/ref_loader.py) running its own LightCurveImageCLIP.training_step + backward + torch.optim.RAdam on the host
# cores.  Falls back to the oracle port (same algorithm restated, oracle/maven_oracle.py) only when no reference tree travelled.
# --------------------------------------------------------------------------------------------------
def reference_kind():
    from oracle import ref_loader
    return "reference" if ref_loader.reference_root() is not None else "port"


def make_reference_stepper(wl, dropout, device="cpu"):
    """-> (kind, step(batch) -> loss tensor) for the reference's training step (src/models_multimodal.py:312-366 + :306-310)."""
    kind = reference_kind()
    if kind == "reference":
        from oracle import ref_loader
        _, _, rmm = ref_loader.import_reference()
        torch.manual_seed(0)
        model = rmm.LightCurveImageCLIP(**model_kwargs(wl, dropout)).to(device).train()
        model.log = lambda *a, **k: None
        opt = model.configure_optimizers()["optimizer"]             # torch.optim.RAdam(self.parameters(), lr, **optimizer_kwargs)

        def step(batch):
            loss = model.training_step(tuple(batch), 0)              # the reference's own 9-tuple batch (src/models_multimodal.py:312-324)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss
        return kind, step
    from oracle import maven_oracle as O
    from maven_b200.models_multimodal import LightCurveImageCLIP
    torch.manual_seed(0)
    ref_model = LightCurveImageCLIP(**model_kwargs(wl, 0.0))        # used only as a weight initialiser (reference default init)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in ref_model.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.RAdam(params, lr=LR, weight_decay=WD)
    cfg = dict(combinations=wl["combinations"], nband=2, transformer_kwargs=wl["lc"], transformer_spectral_kwargs=wl["sp"],
               classification=wl.get("classification", False), n_classes=wl.get("n_classes", 5), conv_kwargs=CONV)

    def step(batch):
        loss = O.training_loss(sd, cfg, batch)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss
    return kind, step


def cpu_reference_steps(wl, sample_B, steps, warmup, threads, dropout=0.0):
    """-> (samples/s, seconds per step, kind) over `steps` timed steps at batch `sample_B` on `threads` host threads."""
    torch.set_num_threads(threads)
    kind, step = make_reference_stepper(wl, dropout)
    batch = make_batch(wl, sample_B, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sample_B * len(times) / sum(times), sum(times) / len(times), kind


def bench_config(wl, args, world):
    """The `config` object of the JSON line: the workload, identical for the B200 arm and the reference arm."""
    return {"workload": wl["desc"], "per_gpu_batch": args.batch, "global_batch": args.batch * world, "parallelism": f"dp{world}",
            "dropout": args.dropout, "data": "SURVEY 8d synthetic generator, seeded on the CPU"}


def run_reference(args):
    """Reference arm: rank 0 alone times the reference's CPU training step on all host cores; each step is a bounded sample
    (a smaller batch) of the B200 arm's workload, sized from a probe step so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    B = args.ref_batch
    if B <= 0:
        probe_B = 64
        sps_probe, _, _ = cpu_reference_steps(wl, probe_B, 1, 1, threads, args.dropout)
        budget_s = 150.0
        B = int(budget_s * sps_probe / max(args.steps + args.warmup, 1))
        B = max(32, min(args.batch, 1 << (B.bit_length() - 1) if B > 0 else 32))      # power of two <= the B200 arm's batch
    sps, sec, kind = cpu_reference_steps(wl, B, args.steps, args.warmup, threads, args.dropout)
    what = ("unmodified reference LightCurveImageCLIP.training_step + backward + torch.optim.RAdam (oracle/_ref, framework imports stubbed)"
            if kind == "reference" else "oracle port of the reference algorithm (no reference tree on this box)")
    line = {"impl": "reference", "metric": "clip_train_samples_per_sec", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(wl, args, args.gpus),
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": kind,
                             "sample": f"{args.steps} training steps at batch {B} (a bounded sample of the per-GPU batch {args.batch}), fp32, "
                                       f"dense over the padded T, dropout {args.dropout}: {what}"},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def reference_eager_on_gpu(wl, dropout, dev, B=256, steps=3):
    """Context figure: what a user of the reference gets on this GPU today -- the unmodified modules run eagerly
    by PyTorch on cuda (cuBLAS/ATen kernels, ~16.6k launches per step), fp32 and with TF32 matmuls allowed."""
    if reference_kind() != "reference":
        return None
    out = {"batch": B, "steps": steps}
    batch = [None if v is None else v.to(dev) for v in make_batch(wl, B, seed=1234)]
    for name, allow in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = allow
        torch.backends.cudnn.allow_tf32 = allow
        _, step = make_reference_stepper(wl, dropout, device=dev)
        for _ in range(2):
            step(batch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(batch)
        torch.cuda.synchronize()
        out[name + "_samples_per_s"] = B * steps / (time.perf_counter() - t0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return out


# --------------------------------------------------------------------------------------------------
def c5_sweep(args, world, rank, dev, L):
    """BASELINE.json configs[4]: the large-global-batch trimodal CLIP (light curve + spectra + image, 3 pairs, all-gathered
    negatives) swept over the global batch 1k-64k.  Every size is a few eager steps (forward + backward + gradient all-reduce +
    RAdam) timed on the device; the CLIP-loss share comes from CUDA events around the loss kernels in a separate pass.
    Runs after the headline measurement and never fails it: an exception is recorded in the result instead."""
    import ctypes
    import torch.distributed as dist
    from maven_b200.models_multimodal import LightCurveImageCLIP
    from maven_b200.transformer_utils import set_precision
    wl = WORKLOADS["c5"]
    out = []
    torch.manual_seed(0)
    model = set_precision(LightCurveImageCLIP(**model_kwargs(wl, args.dropout)).to(dev).train(), args.precision)
    opt = model.configure_optimizers()["optimizer"]

    def step(batch):
        loss = model.training_step(batch, 0)
        loss.backward()
        if world > 1:
            model.reduce_gradients()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for g in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        b = g // world
        if b * world != g or b < 128 or b > args.sweep_max_per_gpu:
            continue
        rec = {"global_batch": g, "per_gpu_batch": b}
        try:
            batch = [None if v is None else v.to(dev) for v in make_batch(wl, b, seed=5000 + rank)]
            # up to 1024 samples per GPU the eager step is bound by its ~350 launches (~6.3 ms of host time): those sizes replay the
            # captured step, like the headline measurement; above it the device time hides the launches and the step is enqueued eagerly
            graphed = None
            if args.graph and b <= 1024:
                from maven_b200.graph import GraphedTrainStep
                try:
                    graphed = GraphedTrainStep(model, opt, batch, group=dist.group.WORLD if world > 1 else None)
                except Exception:
                    graphed = None
                    opt.disable_device_step()
            run = (lambda: graphed(None)) if graphed is not None else (lambda: step(batch))
            rec["launch"] = "graph replay" if graphed is not None else "eager"
            for _ in range(2):
                run()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            n_t = 3
            ev0.record()
            for _ in range(n_t):
                last = run()
            ev1.record()
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1) / n_t], dtype=torch.float64, device=dev)
            last = float(last.detach())
            if graphed is not None:
                opt.disable_device_step()
                opt.zero_grad(set_to_none=True)
                model._gbuf = None
                graphed = run = None
                torch.cuda.empty_cache()
            L.mvn_prof_enable(1 << 5)                          # CLIP-loss kernels only
            step(batch)
            torch.cuda.synchronize()
            L.mvn_prof_enable(0)
            ms, cnt = ctypes.c_double(), ctypes.c_longlong()
            L.mvn_prof_read(5, ctypes.byref(ms), ctypes.byref(cnt))
            tl = torch.tensor([ms.value], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            rec.update(ms_per_step=t.item(), samples_per_s=g / (t.item() / 1e3), loss_kernels_ms=tl.item(), loss_share=tl.item() / t.item(),
                       loss_last=float(last))
            del batch
        except Exception as e:                                 # e.g. out of memory at a size this GPU count cannot hold
            rec["error"] = f"{type(e).__name__}: {str(e)[:160]}"
            out.append(rec)
            break
        out.append(rec)
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import ctypes
    import torch.distributed as dist
    from maven_b200 import _lib, ops
    from maven_b200.models_multimodal import LightCurveImageCLIP
    from maven_b200.transformer_utils import set_precision

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ops.set_data_parallel_group(dist.group.WORLD)
    L = _lib.lib()
    wl = WORKLOADS[args.workload]
    B = args.batch
    torch.manual_seed(0)                                       # identical replicas on every rank
    model = LightCurveImageCLIP(**model_kwargs(wl, args.dropout)).to(dev).train()
    set_precision(model, args.precision)
    opt = model.configure_optimizers()["optimizer"]
    host = make_batch(wl, B, seed=1000 + rank)                 # each rank owns its shard of the global batch
    pinned = [None if v is None else v.pin_memory() for v in host]
    resident = [None if v is None else v.to(dev) for v in host]
    h2d = sum(v.numel() * v.element_size() for v in pinned if v is not None)      # every tensor of the 9-tuple batch is copied per step
    fl = flops_per_step(wl, host, args.precision)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step(batch):
        loss = model.training_step(batch, 0)
        loss.backward()
        if world > 1:
            model.reduce_gradients()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(resident)
    sync()

    # ---- per-class device timing: K steps with every kernel class bracketed by CUDA events on the launching stream ---------
    # (kept out of the throughput region below: event records between launches would defeat the programmatic dependent
    # launches that overlap each tensor-core kernel's prologue with its predecessor's tail)
    names = ["gemm", "wgrad", "attn_fwd", "attn_bwd", "row", "loss", "optim", "conv", "fused_fwd", "fused_bwd"]
    # The class timing runs the modality encoders back to back on one stream (each kernel alone on the GPU, the roofline
    # definition); the throughput region below runs them on one stream each, as the product does by default.
    concurrent = model.concurrent_modalities
    model.concurrent_modalities = False
    L.mvn_prof_enable(0x3FF)
    for _ in range(args.steps):
        flush.fill_(1)
        step(resident)
    torch.cuda.synchronize()
    L.mvn_prof_enable(0)
    model.concurrent_modalities = concurrent
    breakdown, prof_tot = {}, {}
    for c, nme in enumerate(names):
        ms, cnt = ctypes.c_double(), ctypes.c_longlong()
        L.mvn_prof_read(c, ctypes.byref(ms), ctypes.byref(cnt))
        breakdown[nme] = {"ms": round(ms.value / args.steps, 4), "launches": cnt.value // args.steps}
        prof_tot[nme] = (ms.value, cnt.value)
    for _ in range(2):
        step(resident)
    sync()

    # ---- the product's training step: the whole step as ONE CUDA graph (maven_b200.graph), replayed per batch -------------
    graphed, graph_note = None, None
    if args.graph:
        from maven_b200.graph import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(model, opt, resident, group=dist.group.WORLD if world > 1 else None, double_buffer=not args.no_prefetch)
        except Exception as e:                                 # same kernels, launched eagerly: never a different code path
            graph_note = f"eager launches (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            print(f"[bench] rank {rank}: {graph_note}", file=sys.stderr, flush=True)
            opt.disable_device_step()
            opt.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
        if world > 1:                                          # all ranks must take the same path (the collectives have to match)
            ok = torch.tensor([1.0 if graphed is not None else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0.0 and graphed is not None:
                graphed = None
                graph_note = "eager launches (graph capture failed on another rank)"
                opt.disable_device_step()
                opt.zero_grad(set_to_none=True)
    img_u8 = None
    if wl.get("img") and graphed is not None:
        # end-to-end leg of the image workloads: the host keeps the cut-outs as the 8-bit pixels they are; the device converts
        # (maven_b200.augment) straight into the graph's static image input -> 10.8 KB instead of 43.2 KB per sample
        img_u8 = (host[0] * 255.0).round().to(torch.uint8).pin_memory()
        assert torch.equal(img_u8.float() / 255.0, host[0])
        h2d += img_u8.numel() - pinned[0].numel() * pinned[0].element_size()
    if graphed is not None:
        def run_step(batch):                                   # batch None: inputs already resident in the static buffers
            return graphed(batch)
        for _ in range(2):
            run_step(None)
        sync()
    else:
        def run_step(batch):
            return step(resident if batch is None else batch)

    # ---- timed region: K steps, device-resident inputs, per-step CUDA events, L2 flushed between steps ----
    import gc
    gc.collect()
    gc.freeze()                                                # a generation-2 collection inside the timed region stalls the host for tens of ms
    gc.disable()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.mvn_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        run_step(None)
        b.record()
    sync()
    wall = time.perf_counter() - wall0
    launches = L.mvn_launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps     # kernels inside each replayed graph (counted at capture)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(step_ms)
    clocks = sampler.stop()            # sampled during the throughput region only: NVML queries perturb the sync-per-step e2e loop

    # ---- e2e: pinned host batch -> H2D -> step -> D2H loss, everything inside the timed region -----------
    # With the graph's second input set (GraphedTrainStep(double_buffer=True)) the upload of batch i+1 is enqueued on the copy stream
    # before the host reads step i's loss, so it runs under step i's compute -- what DataLoader workers + pinned memory do for the
    # reference.  Every step still uploads its own batch from pinned host memory and reads its own loss back, inside the timed region.
    prefetch = graphed is not None and not args.no_prefetch
    u8_dev = torch.empty_like(img_u8, device=dev) if img_u8 is not None else None

    def host_batch(static):
        if img_u8 is not None:
            from maven_b200.augment import augment_images
            u8_dev.copy_(img_u8, non_blocking=True)
            augment_images(u8_dev, None, out=static[0])                            # uint8 -> fp32/255 on the device
            return [static[0]] + pinned[1:]
        return pinned

    def e2e_step():
        if graphed is not None:
            return graphed(host_batch(graphed.static)).item()  # pinned host batch -> static device buffers -> replay -> D2H loss
        batch = [None if v is None else v.to(dev, non_blocking=True) for v in pinned]
        return step(batch).item()                              # device->host read of the step's loss

    def e2e_loop(n):
        last = None
        if not prefetch:
            for _ in range(n):
                last = e2e_step()
            return last
        graphed.prefetch(host_batch)
        for i in range(n):
            loss_t = graphed.step_prefetched()
            if i + 1 < n:
                graphed.prefetch(host_batch)                   # next batch's upload overlaps this step
            last = loss_t.item()
        return last

    e2e_loop(2)                                                # settle the allocator for the per-step input buffers
    sync()
    t0 = time.perf_counter()
    last = e2e_loop(args.steps)
    sync()
    e2e_wall = time.perf_counter() - t0
    gc.enable()

    t = torch.tensor([dev_ms, e2e_wall * 1e3, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, launches = t[0].item(), t[1].item(), int(t[2].item())
    def finish():
        """Multi-rank runs leave through a hard exit once every rank is done: tearing down a NCCL communicator whose
        collectives were captured into a still-live CUDA graph can block at interpreter exit."""
        sys.stderr.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    sweep = None
    if args.sweep:
        torch.cuda.empty_cache()
        try:
            sweep = c5_sweep(args, world, rank, dev, L)
        except Exception as e:
            sweep = [{"error": f"{type(e).__name__}: {str(e)[:200]}"}]
    if rank != 0:
        finish()
        return
    hbm, tf, src = peaks()
    ms_per_step = dev_ms / args.steps
    value = world * B * args.steps / (dev_ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    # ---- roofline, SURVEY 8(d): the encoder blocks and the similarity are tensor-class (fraction = FLOPs / time / sustained
    # bf16 peak), embed / LayerNorm-backward / pool / ConvMixer are HBM-class (fraction = algorithmic bytes / time / copy peak).
    # Every class reports BOTH fractions; `roofline` is the class with the most device time inside the step.
    tensor_classes = ("gemm", "wgrad", "attn_fwd", "attn_bwd", "fused_fwd", "fused_bwd")
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "traffic_per_launch.json")   # from the committed ncu --set full capture (scripts/ncu_summary.py)
    if os.path.exists(tpath):
        traffic_all = json.load(open(tpath))
    classes = {}
    for k in ("gemm", "wgrad", "attn_fwd", "attn_bwd", "fused_fwd", "fused_bwd", "row", "conv"):
        ms_k, n_k = breakdown[k]["ms"], breakdown[k]["launches"]
        if ms_k <= 0:
            continue
        c = {"ms": ms_k, "launches": n_k, "bound": "tensor" if k in tensor_classes else "hbm",
             "algorithmic_GBps": fl["bytes"][k] / (ms_k / 1e3) / 1e9, "frac_hbm": fl["bytes"][k] / (ms_k / 1e3) / 1e9 / hbm}
        if k in tensor_classes:
            c["executed_TFLOPs"] = fl[k] / (ms_k / 1e3) / 1e12
            c["frac_tensor"] = c["executed_TFLOPs"] / tf
        classes[k] = c
    top = max(classes, key=lambda k: classes[k]["ms"])
    ct = classes[top]
    n_l = max(prof_tot[top][1], 1)
    avg_ms = prof_tot[top][0] / n_l                         # average launch duration of the class inside the timed region
    if ct["bound"] == "tensor":
        per_launch, ach, peak, unit = fl[top] * args.steps / n_l, ct["executed_TFLOPs"], tf, "TFLOP/s"
    else:
        per_launch, ach, peak, unit = fl["bytes"][top] * args.steps / n_l, ct["algorithmic_GBps"], hbm, "GB/s"
    cpu = ref_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb, n_cpu = 128, 6                                      # ~10-30 s of host work: a bounded sample of the workload
        sps, sec, kind = cpu_reference_steps(wl, cb, n_cpu, 1, threads, args.dropout)
        sps1, _, _ = cpu_reference_steps(wl, 16, 2, 1, 1, args.dropout)
        torch.set_num_threads(threads)
        cpu = {"value": sps, "unit": "samples/s", "cores": threads, "kind": kind,
               "sample": f"{n_cpu} training steps at batch {cb} of the same workload (fp32, dense over padded T, dropout {args.dropout}): "
                         + ("unmodified reference training_step + backward + torch RAdam (oracle/_ref)" if kind == "reference" else "oracle port"),
               "one_thread": {"value": sps1, "unit": "samples/s", "cores": 1, "sample": "2 steps at batch 16"}}
        try:
            ref_gpu = reference_eager_on_gpu(wl, args.dropout, dev)
        except Exception as e:                                 # context only: never fails the bench line
            ref_gpu = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    tr = traffic_all.get(args.precision, {})
    step_padded = fl["padded_train"] / (ms_per_step / 1e3) / 1e12
    line = {
        "metric": "clip_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
        "config": bench_config(wl, args, world),
        "runtime": {"precision": args.precision,
                    "streams": "one CUDA stream per modality encoder" if concurrent and len(wl["combinations"]) > 1 else "single stream",
                    "e2e_image_upload": "uint8 pixels, converted on the device (maven_b200.augment)" if img_u8 is not None else None,
                    "launch": "whole step replayed as one CUDA graph (maven_b200.graph.GraphedTrainStep)" if graphed is not None
                              else (graph_note or "eager launches"),
                    "l2": "256 MiB flush written between timed steps; per-step activation working set is GBs (>> 126 MB L2)",
                    "valid_token_fraction": {"lc": float(host[3].float().mean()), "sp": float(host[6].float().mean()) if host[6] is not None else None}},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                "input_prefetch": "batch i+1 uploaded on a copy stream while step i runs (GraphedTrainStep double_buffer)" if (graphed is not None and not args.no_prefetch) else "none"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": ct["bound"], "kernel_class": top, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                     "traffic": tr.get(top), "traffic_source": tr.get("_source"),
                     "algorithmic_per_launch": per_launch, "avg_launch_ms": avg_ms, "peak_source": src,
                     "frac_hbm": ct["frac_hbm"], "launches_timed": prof_tot[top][1], "class_ms_per_step": ct["ms"],
                     "step_frac_of_tensor_peak": {"padded": step_padded / tf, "executed": fl["executed_train"] / (ms_per_step / 1e3) / 1e12 / tf,
                                                  "formula": "samples/s x FLOP/sample / sustained bf16 peak (SURVEY 8d); padded = dense over the padded T, "
                                                             "executed = valid tokens only, recomputation not credited"},
                     "timing": f"CUDA events around every launch of the class over {args.steps} steps run right before the throughput region",
                     "accounting": "class average over its launches in the timed region; executed work (valid tokens only)"},
        "kernel_classes": classes,
        "flops_per_step": {k: v for k, v in fl.items() if k != "bytes"},
        "bytes_per_step": dict(fl["bytes"], total=sum(fl["bytes"].values())),
        "step_tflops": {"padded_equivalent": step_padded, "executed": fl["executed_train"] / (ms_per_step / 1e3) / 1e12,
                        "frac_of_peak_padded": step_padded / tf},
        "kernel_breakdown_ms": breakdown,
        "wall_s_timed_region": wall, "loss_last": last, "step_ms_rank0": [round(x, 3) for x in step_ms],
    }
    if ref_gpu is not None:
        line["reference_eager_b200"] = ref_gpu
    if sweep is not None:
        line["c5_sweep"] = {"workload": WORKLOADS["c5"]["desc"], "launch": "graph replay up to 1024 samples per GPU, eager launches above (see `launch` per size); 3 timed steps per size, max over ranks",
                            "precision": args.precision, "sizes": sweep}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    finish()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (NCCL's version
    banner, library chatter) has been routed to stderr by main()."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                              # C-level writes to stdout (NCCL banner) -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=0, help="samples per CPU reference step (0: sized from a probe step so the run ends within a few minutes)")
    ap.add_argument("--dropout", type=float, default=0.00021844858312997214,
                    help="transformer dropout p (default: pretrain_config/maven_pretrain_config.yaml); the CPU reference arm uses 0")
    ap.add_argument("--precision", default="fused", choices=["fp32", "tf32", "fused"],
                    help="tf32: tcgen05/mma tensor-core tier (fp32 storage, fp32 accumulate; parity 1e-3); fp32: FFMA tier (parity 1e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the C5 large-global-batch sweep that follows the headline measurement")
    ap.add_argument("--sweep-max-per-gpu", type=int, default=8192, help="largest per-GPU batch of the C5 sweep (activation workspace ~5 GB per 1024 samples)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="enqueue every kernel from Python instead of replaying the captured step")
    ap.add_argument("--no-prefetch", action="store_true", help="end-to-end leg: upload every batch right before its own step (no overlap with the previous step)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
