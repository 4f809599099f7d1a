#!/usr/bin/env python
"""bench.py — SAR patches/s of the SNUNet-ECAM training step (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3              # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --steps 3 --warmup 1       # reference arm: CPU port of the reference path
    torchrun ... bench.py --gpus N ...                          # data parallel, one rank per GPU, NCCL all-reduce

A "step" is one full training step on one batch: forward, CE+Dice loss (+argmax), backward, gradient
all-reduce (N>1) and Adam.  `value` = patches/s with inputs resident in HBM; `e2e` = the same step
through the public trainer API from pinned HOST buffers (H2D of both images + mask, D2H of the loss).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "SAR patches/sec (224x224x6ch, bs=64) SNUNet-ECAM train step"
UNIT = "patches/s"
H = W = 224
SNUNET_TRAIN_GFLOP_PER_PATCH = 213.9   # 3 x 71.32 GF forward
# --workload: the headline (default) is BASELINE.json configs[1]; the others are the remaining model families on the path
# (secondary lines, same JSON contract): configs[0] shape (siam-conc, bs=4) and configs[3]'s encoder (FloodViT-B, bs=64).
WORKLOADS = {
    "snunet": dict(metric=METRIC, batch=64, gflop=SNUNET_TRAIN_GFLOP_PER_PATCH, task="cd", method="snunet", lr=1e-3,
                   desc="snunet-ecam (base 32, 12.03M params) train step: fwd + CE+Dice(+argmax) + bwd + allreduce + Adam; "
                        "inputs pre_event_1,post_event of the 3-date x 2-pol 224x224 batch (reference SNUNet takes 2 dates)"),
    "siam-conc": dict(metric="SAR patches/sec (224x224x6ch, bs=4) FC-Siam-conc train step", batch=4, gflop=22.2, task="cd", method="siam-conc",
                      lr=1e-5, desc="siam-conc (1.55M params, Dropout2d p=0.2 on) train step: fwd + CE+Dice(+argmax) + bwd + allreduce + Adam; "
                                    "inputs pre_event_1,post_event (BASELINE.json configs[0] shape)"),
    "changeformer": dict(metric="SAR patches/sec (224x224x6ch, bs=32) ChangeFormerV6 train step", batch=32, gflop=636.0, task="cd",
                         method="changeformer", lr=6e-4,
                         desc="ChangeFormerV6 (embed 256, 41.0M params; stochastic layers ON at the model defaults: Dropout / attention dropout / DropPath 0.1) train step: fwd + CE+Dice(+argmax) on the last "
                              "sigmoid output + bwd + allreduce + SGD(momentum 0.99); inputs pre_event_1,post_event (BASELINE.json configs[2] shape)"),
    "floodvit": dict(metric="SAR patches/sec (224x224x6ch, bs=64) FloodViT-B train step", batch=64, gflop=106.1, task="segmentation",
                     method="finetune", lr=1e-4,
                     desc="FloodViT: ViT-B/16 encoder (6 channels, 86.4M params) + linear FinetunerSegmentation head, train step: fwd + "
                          "CE+Dice(+argmax) + bwd + allreduce + Adam; image = cat(post_event, pre_event_1, pre_event_2)"),
    "floodvit-upernet": dict(metric="SAR patches/sec (224x224x6ch, bs=64) FloodViT-B + UPerNet head train step", batch=64, gflop=118.0,
                             task="segmentation", method="finetune", lr=1e-4,
                             desc="FloodViT (BASELINE.json configs[3]): ViT-B/16 encoder (6 channels) + HF-UperNetHead semantics (hidden 512, pool "
                                  "scales 1/2/3/6, features after blocks 3/6/9/12), train step: fwd + CE+Dice(+argmax) + bwd + allreduce + Adam"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the CPU port of the reference path (oracle/, torch-CPU, all host threads) on a bounded
    sample (bs=4) of the same workload.  /root/reference is Python and does not travel to the GPU box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import snunet_oracle, weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bs = 4
    sd = snunet_oracle.to_torch_state(weights.make_state(999, 2, 3, 32))
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(999, bs, H, W))
    state = {}

    def step():
        loss, _, grads = snunet_oracle.train_step(sd, xA, xB, mask)
        snunet_oracle.adam_step(sd, grads, state, lr=1e-3)
        return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = bs / dt
    sample = f"SNUNet-ECAM fp32 train step (fwd+CE+Dice+bwd+Adam), bs={bs} of the bs=64 workload, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "snunet-ecam train step, 2x[bs,2,224,224] SAR inputs + [bs,224,224] mask", "per_gpu_batch": bs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
def cpu_baseline_leg(seconds_budget: float = 25.0):
    import torch
    from oracle import snunet_oracle, weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bs = 4
    sd = snunet_oracle.to_torch_state(weights.make_state(999, 2, 3, 32))
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(999, bs, H, W))
    state = {}
    times = []
    t_start = time.perf_counter()
    for i in range(4):
        t0 = time.perf_counter()
        _, _, grads = snunet_oracle.train_step(sd, xA, xB, mask)
        snunet_oracle.adam_step(sd, grads, state, lr=1e-3)
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
        if time.perf_counter() - t_start > seconds_budget and times:
            break
    times.sort()
    med = times[len(times) // 2]
    return {"value": bs / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port of the reference SNUNet train step, fp32, bs={bs} (of 64), 1 warm-up + {len(times)} timed steps"}


def library_snunet(base=32):
    """Stock-torch SNUNet-ECAM (nn.Conv2d / BatchNorm2d / ConvTranspose2d -> cuDNN / cuBLAS) with the reference's wiring
    (models/snunet.py:20-29, :41-46, :49-62, :118-153).  Used ONLY as the library comparator of SURVEY.md §8(d): the same step on the
    same B200 through bf16 autocast + channels_last - the bar the hand-written kernels have to beat.  Never on the product path."""
    import torch
    import torch.nn as nn

    class Block(nn.Module):
        def __init__(self, cin, mid, cout):
            super().__init__()
            self.conv1, self.bn1 = nn.Conv2d(cin, mid, 3, padding=1), nn.BatchNorm2d(mid)
            self.conv2, self.bn2 = nn.Conv2d(mid, cout, 3, padding=1), nn.BatchNorm2d(cout)

        def forward(self, x):
            y1 = self.conv1(x)
            h = torch.relu(self.bn1(y1))
            return torch.relu(self.bn2(self.conv2(h)) + y1)

    class CA(nn.Module):
        def __init__(self, c, ratio):
            super().__init__()
            self.fc1, self.fc2 = nn.Conv2d(c, c // ratio, 1, bias=False), nn.Conv2d(c // ratio, c, 1, bias=False)

        def forward(self, x):
            f = lambda t: self.fc2(torch.relu(self.fc1(t)))
            return torch.sigmoid(f(torch.nn.functional.adaptive_avg_pool2d(x, 1)) + f(torch.nn.functional.adaptive_max_pool2d(x, 1)))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            f = [base, base * 2, base * 4, base * 8, base * 16]
            self.enc = nn.ModuleList([Block(2 if l == 0 else f[l - 1], f[l], f[l]) for l in range(5)])
            self.dec = nn.ModuleDict({f"{l}_{j}": Block(f[l] * (j + 1) + f[l + 1], f[l], f[l]) for j in range(1, 5) for l in range(5 - j)})
            self.up = nn.ModuleDict({f"{l}_{j}": nn.ConvTranspose2d(f[l + 1], f[l + 1], 2, stride=2) for j in range(1, 5) for l in range(5 - j)})
            self.ca, self.ca1 = CA(f[0] * 4, 16), CA(f[0], 4)
            self.final = nn.Conv2d(f[0] * 4, 3, 1)
            self.pool = nn.MaxPool2d(2, 2)

        def forward(self, xA, xB):
            X = {}
            for br, x in (("A", xA), ("B", xB)):
                for l in range(5):
                    if l == 4 and br == "A":
                        continue
                    X[(l, br)] = self.enc[l](x if l == 0 else self.pool(X[(l - 1, br)]))
            for (l, j) in [(0, 1), (1, 1), (0, 2), (2, 1), (1, 2), (0, 3), (3, 1), (2, 2), (1, 3), (0, 4)]:
                below = X[(l + 1, "B")] if j == 1 else X[(l + 1, j - 1)]
                cat = [X[(l, "A")], X[(l, "B")]] + [X[(l, k)] for k in range(1, j)] + [self.up[f"{l}_{j}"](below)]
                X[(l, j)] = self.dec[f"{l}_{j}"](torch.cat(cat, 1))
            outs = [X[(0, j)] for j in range(1, 5)]
            out = torch.cat(outs, 1)
            intra = torch.sum(torch.stack(outs), dim=0)
            out = self.ca(out) * (out + self.ca1(intra).repeat(1, 4, 1, 1))
            return self.final(out)

    return Net()


def library_baseline_leg(dev, bs, steps=6, warmup=3):
    """patches/s of the SAME training step (fwd + CE+Dice + bwd + Adam, same batch) on stock torch: cuDNN/cuBLAS under
    torch.autocast(bfloat16), channels_last, fused Adam; eager and - when capture succeeds - as one CUDA graph (the better is reported)."""
    import torch
    import torch.nn.functional as F
    from kurosiwo_b200 import synthetic
    torch.manual_seed(999)
    net = library_snunet().to(dev).to(memory_format=torch.channels_last).train()
    b = synthetic.make_batch(999, bs, H, W, pin=False)
    xA = b[6].to(dev).contiguous(memory_format=torch.channels_last)
    xB = b[2].to(dev).contiguous(memory_format=torch.channels_last)
    mask = b[3].to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True)

    def loss_fn(logits):
        logits = logits.float()
        valid = mask != 3
        t = torch.zeros_like(logits).scatter_(1, (mask * valid).unsqueeze(1), 1.0) + 1e-6
        pr = F.softmax(logits, dim=1)
        dice = torch.mean(1.0 - 2.0 * torch.sum(pr * t, (1, 2, 3)) / (torch.sum(pr + t, (1, 2, 3)) + 1e-6))
        return dice + F.cross_entropy(logits, mask, ignore_index=3)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(xA, xB)
        loss = loss_fn(out)
        loss.backward()
        opt.step()
        return loss

    def timeit(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    res = {"unit": UNIT, "impl": f"torch {torch.__version__} eager nn.Module (cuDNN {torch.backends.cudnn.version()}), autocast bf16, channels_last, "
                                 "fused Adam; same batch and step definition", "per_gpu_batch": bs}
    torch.backends.cudnn.benchmark = True
    ms_eager = timeit(step)
    res["eager_ms_per_step"] = ms_eager
    ms_best, mode = ms_eager, "eager"
    try:
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=False)
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            step()
        torch.cuda.current_stream().wait_stream(s_)
        torch.cuda.synchronize()

        def gstep():
            opt.zero_grad(set_to_none=False)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(xA, xB)
            loss = loss_fn(out)
            loss.backward()
            opt.step()
        with torch.cuda.graph(g):
            gstep()
        ms_graph = timeit(g.replay)
        res["graph_ms_per_step"] = ms_graph
        if ms_graph < ms_best:
            ms_best, mode = ms_graph, "cuda_graph"
    except Exception as e:      # noqa: BLE001 - the comparator is best effort; eager stands
        res["graph_error"] = str(e)[:120]
    res.update(value=bs / (ms_best * 1e-3), ms_per_step=ms_best, mode=mode, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    del net, opt
    torch.cuda.empty_cache()
    return res


def run_library(args):
    """`--impl torch-gpu`: the library comparator alone, as its own JSON line (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    torch.cuda.set_device(0)
    wl = WORKLOADS["snunet"]
    res = library_baseline_leg("cuda:0", args.batch or wl["batch"], steps=args.steps, warmup=max(args.warmup, 3))
    print(json.dumps({"impl": "torch-gpu", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": wl["desc"], "per_gpu_batch": res["per_gpu_batch"], "library": res["impl"], "mode": res["mode"]},
                      "library_baseline": res}))


def conv_roofline(eng, dev_inputs, pk, pk_kind):
    """Per-launch CUDA-event timing of the dominant kernel family (tcgen05 implicit-GEMM conv: fwd + dgrad + wgrad)
    inside one eager training step; achieved = algorithmic FLOPs / summed launch time."""
    import torch
    ops = eng.ops
    recs = []
    orig_conv, orig_wgrad, orig_wgrad_b = ops.conv2d, ops.conv2d_wgrad, ops.conv2d_wgrad_bias

    def timed(fn, kind):
        def wrapper(N, Hh, Ww, ksize, a, *rest, **kw):
            if kind == "conv":
                srcs, dsts = a, rest[2]
                cin, cout = sum(v.C for v in srcs), sum(v.C for v in dsts)
            else:
                cin, cout = sum(v.C for v in a), sum(v.C for v in rest[0])
            fl = 2.0 * N * Hh * Ww * ksize * ksize * cin * cout
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(N, Hh, Ww, ksize, a, *rest, **kw)
            e1.record()
            recs.append((kind, cin, cout, Hh, ksize, fl, e0, e1))
            return r
        return wrapper

    ops.conv2d, ops.conv2d_wgrad, ops.conv2d_wgrad_bias = timed(orig_conv, "conv"), timed(orig_wgrad, "wgrad"), timed(orig_wgrad_b, "wgrad")
    eng._skip_comm = True                  # this leg runs on rank 0 only: no collective may be issued (bucket hooks included)
    passes = []
    try:
        # three eager passes, per launch the MEDIAN of the three: an event pair around an eager launch also sees the host falling
        # behind (tensor-map encodes + launch, ~15 us per call) whenever the GPU runs dry - a property of the box's CPU, not of the kernel
        for _ in range(3):
            recs.clear()
            eng._fwd_loss_bwd(*dev_inputs)
            eng._optimizer()
            torch.cuda.synchronize()
            passes.append([(r[0], r[1], r[2], r[3], r[4], r[5], r[6].elapsed_time(r[7])) for r in recs])
    finally:
        ops.conv2d, ops.conv2d_wgrad, ops.conv2d_wgrad_bias = orig_conv, orig_wgrad, orig_wgrad_b
        eng._skip_comm = False
    assert len({len(p_) for p_ in passes}) == 1
    recs = [pr[0][:6] + (sorted(x[6] for x in pr)[1],) for pr in zip(*passes)]    # (kind, cin, cout, h, k, flop, median ms)
    tot_fl = sum(r[5] for r in recs)
    tot_ms = sum(r[6] for r in recs)
    by_kind = {}
    for kind, cin, cout, hh, ks, fl, ms in recs:
        k = by_kind.setdefault(kind, [0.0, 0.0, 0])
        k[0] += fl; k[1] += ms; k[2] += 1
    ach = tot_fl / (tot_ms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    detail = {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]} for k, v in by_kind.items()}
    layers = [{"kind": r[0], "cin": r[1], "cout": r[2], "h": r[3], "k": r[4], "ms": r[6], "tflops": r[5] / (r[6] * 1e-3) / 1e12} for r in recs]
    traffic = None      # DRAM bytes per launch of this kernel family from the committed ncu launch list (SNUNet bs=64 only)
    tp = ROOT / "profiles" / "r2_conv_traffic.json"
    if tp.exists() and getattr(eng, "f", None) is not None and getattr(eng, "N", 0) == 64:
        try:
            traffic = json.loads(tp.read_text())["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    # Per spatial level: the same launches against BOTH roofs.  Algorithmic bytes of a launch = every operand tensor once (bf16): the
    # full-resolution Cout = 32 / 64 levels of SNUNet sit closer to the HBM roof than to the tensor roof.
    by_level = {}
    for kind, cin, cout, hh, ks, fl, ms in recs:
        lv = by_level.setdefault(str(hh), [0.0, 0.0, 0.0, 0])
        n_ = getattr(eng, "N", 0) or 0
        lv[0] += fl; lv[1] += ms; lv[2] += 2.0 * n_ * hh * hh * (cin + cout); lv[3] += 1
    hbm = pk["hbm_gbs"]
    levels = {h: {"ms": v[1], "launches": v[3], "tflops": v[0] / (v[1] * 1e-3) / 1e12, "frac_tensor": v[0] / (v[1] * 1e-3) / 1e12 / peak,
                  "alg_gbs": v[2] / (v[1] * 1e-3) / 1e9, "frac_hbm": v[2] / (v[1] * 1e-3) / 1e9 / hbm} for h, v in by_level.items()}
    return ({"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
             "kernel": "conv_tc2_kernel + wgrad_tc*_kernel (tcgen05 implicit GEMM: every ks_conv2d / ks_conv2d_wgrad launch of the step)",
             "peak_source": f"{pk_kind} bf16_tflops_sustained", "conv_ms_per_step": tot_ms, "timing": "CUDA events around every launch of three eager steps, per launch the median", "by_kind": detail,
             "by_level": levels, "by_level_note": "key = spatial size of the level; alg_gbs = operand tensors once / time against "
             f"the {pk_kind} HBM peak {hbm:.0f} GB/s"}, layers)


def loss_roofline_leg(dev, N, H, W, pk, pk_kind):
    """The fused CE+Dice kernel against the HBM roof (north star: >= 0.60 at bs=64): 30 calls on 6 rotating buffer sets (so that every
    call reads its logits and labels from HBM, not from the 126 MB L2) captured as ONE CUDA graph and timed with CUDA events."""
    import torch
    from kurosiwo_b200.lib import default_ops
    ops = default_ops()
    sets = []
    for i in range(6):
        g = torch.Generator(device=dev).manual_seed(i)
        sets.append((torch.randn(N, 3, H, W, device=dev, generator=g), torch.randint(0, 4, (N, H, W), device=dev, generator=g),
                     torch.empty(N, 3, H, W, device=dev), torch.empty(N, H, W, dtype=torch.uint8, device=dev)))
    w = torch.ones(3, device=dev); loss3 = torch.zeros(3, device=dev); ws = ops.ce_dice_workspace(N, dev)

    def run(i):
        z, y, dz, pr = sets[i % len(sets)]
        ops.ce_dice(z, y, w, 3, 1.0, loss3, dz, pr, ws)
    reps = 30
    g = torch.cuda.CUDAGraph()
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        for i in range(6):
            run(i)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s_, capture_error_mode="thread_local"):
            for i in range(reps):
                run(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_)
        for _ in range(5):
            g.replay()
        e1.record(s_); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * reps) * 1e3
    alg = N * H * W * 33          # 12 B logits + 8 B int64 label + 12 B gradient + 1 B argmax per pixel
    del sets
    torch.cuda.empty_cache()
    return {"bound": "hbm", "achieved": alg / us / 1e3, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": alg / us / 1e3 / pk["hbm_gbs"],
            "us_per_call": us, "alg_bytes_per_call": alg, "kernel": "ce_dice_resident_kernel (single pass: logits resident on the SMs across "
            "one grid barrier)", "peak_source": f"{pk_kind} hbm_gbs", "how": "30 calls on 6 rotating buffer sets as one CUDA graph, CUDA events"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from kurosiwo_b200 import synthetic
    from kurosiwo_b200.change_detection_trainer import FusedStepper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    pg = None
    if world > 1:
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")      # gradient exchange over NVLink / NVSwitch peer access only
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
        pg = dist.group.WORLD
    wl = WORKLOADS[args.workload]
    bs = args.batch or wl["batch"]
    torch.manual_seed(999)
    configs = {"device": dev, "inputs": ["pre_event_1", "post_event"], "dem": False, "scale_input": "normalize", "num_classes": 3,
               "num_channels": 2, "loss_function": "ce+dice", "class_weights": [1.0, 1.0, 1.0], "method": wl["method"], "epochs": 1,
               "precision": args.precision, "task": wl["task"], "resume_checkpoint": False}
    model_configs = {"method": wl["method"], "optimizer": "adam", "learning_rate": wl["lr"], "lr_schedule": None, "base_channel": 32}
    if wl["method"] == "changeformer":
        model_configs.update({"optimizer": "sgd", "momentum": 0.99, "weight_decay": 1e-5, "embed_dim": 256, "decoder_softmax": True})
    # e2e arm: pinned RAW SAR tiles; clamp / nan_to_num / Normalize (dataset/Dataset.py:162-168, :192-198) run on the device copy
    # (ks_sar_preprocess) inside the timed region.  Device-resident arm: the already-normalised tensors.
    configs.update({"raw_input": True, "data_mean": list(synthetic.DATA_MEAN), "data_std": list(synthetic.DATA_STD), "clamp_input": 0.15})
    host_batches = [synthetic.make_batch(999 + rank + 1000 * i, bs, H, W, pin=True, raw_tiles=True) for i in range(2)]
    b0 = synthetic.make_batch(999 + rank, bs, H, W, pin=False)
    if wl["task"] == "cd":
        from kurosiwo_b200.model_utilities import initialize_cd_model
        model = initialize_cd_model(configs, model_configs).train()
        stepper = FusedStepper(model, configs, model_configs, process_group=pg)
        dev_inputs = (b0[6].to(dev), b0[2].to(dev), b0[3].to(dev))     # pre_event_1, post_event, mask
        h2d = 2 * bs * 2 * H * W * 4 + bs * H * W * 8
    else:
        from kurosiwo_b200.model_utilities import initialize_segmentation_model
        from kurosiwo_b200.segmentation_trainer import FusedSegStepper
        configs.update({"inputs": ["pre_event_1", "pre_event_2", "post_event"], "num_channels": 6, "mlp": False, "decoder": False,
                        "finetuning_patch_size": 16, "linear_eval": False, "encoder": None})
        model_configs["encoder_config"] = {"image_size": 224, "patch_size": 16, "dim": 768, "depth": 12, "heads": 12, "mlp_dim": 3072}
        if args.workload == "floodvit-upernet":
            configs["head"] = "upernet"
        model = initialize_segmentation_model(configs, model_configs).to(dev).train()
        stepper = FusedSegStepper(model, configs, model_configs, process_group=pg)
        dev_inputs = (torch.cat((b0[2], b0[6], b0[9]), 1).to(dev), b0[3].to(dev))   # cat(post, pre1, pre2), mask
        h2d = 3 * bs * 2 * H * W * 4 + bs * H * W * 8
    # ---- device-resident arm -------------------------------------------------------------------
    eng = stepper._engine(dev_inputs[0])
    ops = eng.ops
    if args.tc_mt:
        ops.set_option("tc_mt", args.tc_mt)
    if args.tc_v1:
        ops.set_option("tc_v1", 1)
    use_graph = not args.no_graph
    for _ in range(2):
        eng.train_step(*dev_inputs)
    torch.cuda.synchronize()
    l0 = ops.launches
    m0 = eng.comm_stats["messages"] if world > 1 else 0
    eng.train_step(*dev_inputs)
    calls_per_step = ops.launches - l0
    msgs_per_step = (eng.comm_stats["messages"] - m0) if world > 1 else 0
    if use_graph:
        step = eng.capture(*dev_inputs)
    else:
        step = lambda: eng.train_step(*dev_inputs)
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    value = world * bs / (ms_per_step * 1e-3)
    loss_val = float(eng.loss3[0].item())
    # ---- gradient exchange: bytes, bus bandwidth of the isolated all-reduce, and the time it adds to the step -------------------
    allreduce = None
    if world > 1:
        g = eng.params.grad
        nbytes = 4 * g.numel()
        for _ in range(2):
            dist.all_reduce(g)
        dist.barrier(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(g)
        a1.record(); torch.cuda.synchronize()
        ta = torch.tensor([a0.elapsed_time(a1) / 5], device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        ms_iso = ta.item()
        # the same step WITHOUT its exchange (ranks run free): the difference is what the exchange costs the step (exposed time + skew)
        eng._skip_comm = True
        step_nc = eng.capture(*dev_inputs, warmup=0) if use_graph else (lambda: eng.train_step(*dev_inputs))
        for _ in range(3):
            step_nc()
        dist.barrier(); torch.cuda.synchronize()
        a0.record()
        for _ in range(args.steps):
            step_nc()
        a1.record(); torch.cuda.synchronize()
        tn = torch.tensor([a0.elapsed_time(a1) / args.steps], device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        eng._skip_comm = False
        if use_graph:
            step = eng.capture(*dev_inputs, warmup=0)       # back to the exchanging step for the legs below
        allreduce = {"bytes": nbytes, "ms_isolated": ms_iso, "bus_gbs": nbytes * 2 * (world - 1) / world / (ms_iso * 1e-3) / 1e9,
                     "ms_exposed": ms_per_step - tn.item(), "ms_per_step_without_exchange": tn.item(),
                     "bucket_mb": eng.bucket_bytes / 2 ** 20, "overlapped_with_backward": bool(eng.overlap_comm),
                     "messages_per_step": msgs_per_step, "transport": "NCCL all-reduce, NCCL_P2P_LEVEL=NVL (NVLink/NVSwitch only)"}
    # ---- end-to-end arm: public trainer step from pinned host buffers ---------------------------
    # stepper.step_host: two eager steps, then a CUDA-graph replay per step over static inputs; prefetch() puts the H2D copy of the
    # NEXT step's batch on a second stream.  Every step's inputs are copied from pinned host memory inside the timed region.
    stepper.prefetch(host_batches[0])
    for i in range(4):
        stepper.prefetch(host_batches[(i + 1) % 2])
        l3, _ = stepper.step_host(host_batches[i % 2])
        l3.cpu()
    stepper.step_host(host_batches[0])[0].cpu()       # consumes the last prefetched batch
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    stepper.prefetch(host_batches[0])
    for i in range(args.steps):
        cur, nxt = host_batches[i % 2], host_batches[(i + 1) % 2]
        if i + 1 < args.steps:
            stepper.prefetch(nxt)                     # H2D of step i+1 overlaps step i (what train_change_detection's loop does)
        l3, _ = stepper.step_host(cur)
        _ = l3.cpu()                                  # D2H of the step's loss (blocks: also the per-step sync)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * bs / (t.item() / args.steps * 1e-3)
    # ---- roofline + CPU baseline (rank 0, N==1 only) --------------------------------------------
    pk, pk_kind = peaks()
    roof, layers, cpu, lib, roof_loss = None, None, None, None, None
    if rank == 0:
        roof, layers = conv_roofline(eng, dev_inputs, pk, pk_kind)
        if world == 1:
            try:
                roof_loss = loss_roofline_leg(dev, bs, 224, 224, pk, pk_kind)
            except Exception as e:      # noqa: BLE001
                roof_loss = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline and args.workload == "snunet":
            cpu = cpu_baseline_leg()
        if world == 1 and not args.no_library_baseline and args.workload == "snunet" and args.precision == "bf16":
            try:
                del step
                eng.graph = None
                torch.cuda.empty_cache()
                lib = library_baseline_leg(dev, bs)
            except Exception as e:      # noqa: BLE001
                lib = {"unavailable": str(e)[:200]}
        out_dir = ROOT / "gpurun_out"
        try:
            out_dir.mkdir(exist_ok=True)
            (out_dir / "bench_layers.json").write_text(json.dumps(layers, indent=0))
        except Exception:
            pass
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    step_tflops = world * bs * wl["gflop"] / (ms_per_step * 1e-3) / 1e3 / world
    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"],
                   "per_gpu_batch": bs, "global_batch": bs * world, "parallelism": f"dp{world}", "cuda_graph": use_graph,
                   "l2": "per-step working set (activations kept for the backward: GBs at the BASELINE batch) exceeds the 126 MB L2; no explicit flush",
                   "step_tflops_per_gpu": step_tflops, "gflop_per_patch": wl["gflop"], "final_loss": loss_val},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "input": "pinned RAW float32 SAR tiles; clamp/nan_to_num/normalise on the device (ks_sar_preprocess) inside the timed region"},
        "gpu_launches": calls_per_step * args.steps,
        "gpu_launches_note": f"{calls_per_step} C-ABI calls per step (each >=1 kernel of libkurosiwo_b200.so)",
        "roofline": roof, "roofline_loss": roof_loss, "cpu_baseline": cpu, "library_baseline": lib, "allreduce": allreduce,
    }
    print(json.dumps(line))


def main():
    if os.environ.get("KS_BENCH_WATCHDOG"):      # debugging aid: dump every thread's stack and exit if the run exceeds N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["KS_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"],
                    help="reference = the CPU arm; torch-gpu = the stock-torch (cuDNN/cuBLAS) library comparator on the same GPU")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = the workload's BASELINE.json batch)")
    ap.add_argument("--workload", default="snunet", choices=sorted(WORKLOADS), help="snunet = the headline (BASELINE.json configs[1])")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--tc-v1", type=int, default=0, help="1 = non-persistent v1 conv kernel (A/B comparisons)")
    ap.add_argument("--tc-mt", type=int, default=0, help="output windows per CTA of the tcgen05 conv (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "torch-gpu":
        run_library(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        port = 29500 + (os.getpid() % 1000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(Path(__file__).resolve())] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
