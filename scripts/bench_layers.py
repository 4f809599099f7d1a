"""Layer-level microbenchmark of the tcgen05 conv / wgrad kernels at SNUNet bs=64 shapes (B200 only).
   python scripts/bench_layers.py [out.json]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import IMPL_TC, CudaOps, KsError, View  # noqa: E402

dev, bf = "cuda:0", torch.bfloat16
ops = CudaOps()
N = 64


def buf(H, C):
    return View.alloc(N, H, H, C, bf, dev, zero=False)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(H, cins, cout_views, ks=3, stats=False, acc=False):
    srcs = []
    for c in cins:
        b = buf(H, c); b.base.normal_(); srcs.append(b)
    dsts = []
    for c in cout_views:
        b = buf(H, c); b.base.zero_(); dsts.append(b)
    cin, cout = sum(cins), sum(cout_views)
    w = torch.randn(ks * ks * cout * cin, device=dev).mul_(0.05).to(bf)
    bias = torch.zeros(cout, device=dev)
    st = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if stats else None
    fl = 2.0 * N * H * H * ks * ks * cin * cout
    return (lambda: ops.conv2d(N, H, H, ks, srcs, w, bias, dsts, [acc] * len(dsts), st, IMPL_TC)), fl


def wgrad_case(H, cins, couts, ks=3):
    xs = []
    for c in cins:
        b = buf(H, c); b.base.normal_(); xs.append(b)
    dys = []
    for c in couts:
        b = buf(H, c); b.base.normal_(); dys.append(b)
    cin, cout = sum(cins), sum(couts)
    dw = torch.zeros(ks * ks * cout * cin, device=dev)
    fl = 2.0 * N * H * H * ks * ks * cin * cout
    return (lambda: ops.conv2d_wgrad(N, H, H, ks, xs, dys, dw, False, IMPL_TC)), fl


CONV = {
    "L0 fwd 32->32": (224, [32], [32], 3, True),
    "L0 fwd 128->32 (64+64)": (224, [64, 64], [32], 3, True),
    "L0 fwd 160->32 (96+64)": (224, [96, 64], [32], 3, True),
    "L0 fwd 224->32 (160+64)": (224, [160, 64], [32], 3, True),
    "L0 dgrad 32->224": (224, [32], [64, 96, 64], 3, False),
    "L0 dgrad 32->32": (224, [32], [32], 3, False),
    "L1 fwd 64->64": (112, [64], [64], 3, True),
    "L1 dgrad 64->64": (112, [64], [64], 3, False),
    "L1 fwd 32->64": (112, [32], [64], 3, True),
    "L1 fwd 384->64 (256+128)": (112, [256, 128], [64], 3, True),
    "L1 fwd 256->64 (128+128)": (112, [128, 128], [64], 3, True),
    "L2 fwd 256->128 as 2x64": (56, [256], [128], 3, True),
    "L1 dgrad 64->384": (112, [64], [256, 128], 3, False),
    "L2 fwd 640->128": (56, [384, 256], [128], 3, True),
    "L2 fwd 128->128": (56, [128], [128], 3, True),
    "L3 fwd 1024->256": (28, [512, 512], [256], 3, True),
    "L4 fwd 512->512": (14, [512], [512], 3, True),
    "Up 64->4x64 @112 (1x1)": (112, [64], [256], 1, False),
}
WGRAD = {
    "L0 wgrad 32x32": (224, [32], [32]),
    "L0 wgrad 224x32": (224, [160, 64], [32]),
    "L1 wgrad 64x64": (112, [64], [64]),
    "L1 wgrad 384x64": (112, [256, 128], [64]),
    "L2 wgrad 640x128": (56, [384, 256], [128]),
    "L3 wgrad 1024x256": (28, [512, 512], [256]),
}
VARIANTS = {
    "auto": {},
    "v1": {"tc_v1": 1},
    "nores_mt1": {"tc_no_resident": 1, "tc_mt": 1},
    "nores_mt2": {"tc_no_resident": 1, "tc_mt": 2},
    "nores_mt4": {"tc_no_resident": 1, "tc_mt": 4},
    "res_mt2": {"tc_mt": 2},
    "no_ns3": {"tc_no_ns3": 1},
    "ew8": {"tc_ew": 8},
    "ew16": {"tc_ew": 16},
    "stat_butterfly": {"tc_stat_mode": 1},
    "ns3_all": {"tc_ns3_min_cin": 1},
    "ns3_two": {"tc_ns3_min_cin": 1, "tc_ns3_mode": 1},
    "ns3_one": {"tc_ns3_min_cin": 1, "tc_ns3_mode": 2},
    "ns3_one_mt2": {"tc_ns3_min_cin": 1, "tc_ns3_mode": 2, "tc_mt": 2},
    "ns3_nores_one": {"tc_ns3_min_cin": 1, "tc_ns3_mode": 2, "tc_no_resident": 1},
    "ns3_nores_two": {"tc_ns3_min_cin": 1, "tc_ns3_mode": 1, "tc_no_resident": 1},
}
import os
SEL = [x.strip() for x in os.environ.get("KS_LAYERS", "").split(",") if x.strip()]
VSEL = [x.strip() for x in os.environ.get("KS_VARIANTS", "").split(",") if x.strip()]
REPS = int(os.environ.get("KS_REPS", "5"))
ops.set_option("tc_debug", int(os.environ.get("KS_DEBUG", "0")))   # ablations: 1 no stores, 2 no TMEM loads, 4 no MMAs, 8 no epilogue, 16 no TMA
if SEL:
    CONV = {k: v for k, v in CONV.items() if k in SEL}
    WGRAD = {k: v for k, v in WGRAD.items() if k in SEL}
if VSEL:
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in VSEL}
out = {"conv": {}, "wgrad": {}}
for name, (H, cins, couts, ks, stats) in CONV.items():
    row = {}
    for vn, opts in VARIANTS.items():
        for o in ("tc_v1", "tc_mt", "tc_no_resident", "tc_no_ns3", "tc_ns3_min_cin", "tc_ns3_mode", "tc_ew", "tc_stat_mode"):
            ops.set_option(o, opts.get(o, 0))
        try:
            fn, fl = conv_case(H, cins, couts, ks, stats and vn != "v1")
            ms = timeit(fn, REPS)
            row[vn] = (round(ms, 4), round(fl / ms / 1e9, 1))
        except KsError as e:
            row[vn] = str(e)[:60]
        torch.cuda.empty_cache()
    out["conv"][name] = row
    print(f"{name:28s}", "  ".join(f"{k}={v}" for k, v in row.items()), flush=True)
for o in ("tc_v1", "tc_mt", "tc_no_resident", "tc_no_ns3", "tc_ns3_min_cin", "tc_ns3_mode", "tc_ew", "tc_stat_mode"):
    ops.set_option(o, 0)
for name, (H, cins, couts) in WGRAD.items():
    fn, fl = wgrad_case(H, cins, couts)
    ms = timeit(fn, REPS)
    out["wgrad"][name] = (round(ms, 4), round(fl / ms / 1e9, 1))
    print(f"{name:28s} ms={ms:.4f} TF={fl / ms / 1e9:.1f}", flush=True)
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
