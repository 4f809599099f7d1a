"""Perf experiment: which part of the conv epilogue limits the store-heavy layers? (B200 only)"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import importlib
bl = importlib.import_module("bench_layers_lib") if False else None
from kurosiwo_b200.lib import IMPL_TC, CudaOps, View
ops = CudaOps(); dev = "cuda:0"; bf = torch.bfloat16; N = 64
def buf(H, C):
    v = View.alloc(N, H, H, C, bf, dev, zero=False); v.base.normal_(); return v
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = {
  "Up 64->4x64 @112": (112, [64], [256], 1, True, False),
  "L0 dgrad 32->224": (224, [32], [64, 96, 64], 3, False, False),
  "L0 dgrad 32->224 acc": (224, [32], [64, 96, 64], 3, False, True),
  "L0 fwd 224->32": (224, [160, 64], [32], 3, True, False),
  "L1 fwd 64->64": (112, [64], [64], 3, True, False),
  "L0 fwd 32->32": (224, [32], [32], 3, True, False),
  "L0 dgrad 32->32": (224, [32], [32], 3, False, False),
  "L1 dgrad 64->384": (112, [64], [256, 128], 3, False, False),
}
for name, (H, cins, couts, ks, bias, acc) in cases.items():
    srcs = [buf(H, c) for c in cins]; dsts = [buf(H, c) for c in couts]
    cin, cout = sum(cins), sum(couts)
    w = torch.randn(ks * ks * cout * cin, device=dev).mul_(0.05).to(bf)
    b = torch.zeros(cout, device=dev) if bias else None
    row = {}
    for dbg in (0, 1, 2, 3, 32, 8, 8 + 4, 16, 16 + 8, 16 + 8 + 4, 16 + 8 + 4 + 32):
        ops.set_option("tc_debug", dbg)
        row[dbg] = round(timeit(lambda: ops.conv2d(N, H, H, ks, srcs, w, b, dsts, [acc] * len(dsts), None, IMPL_TC)), 4)
    ops.set_option("tc_debug", 0)
    print(f"{name:24s} full={row[0]} nostore={row[1]} notmemld={row[2]} nostore+notmemld={row[3]} spin={row[32]} noepi={row[8]} noepi+nomma={row[12]} notma={row[16]} notma+noepi={row[24]} onlybarriers={row[28]} onlybarriers+spin={row[60]}", flush=True)
    del srcs, dsts; torch.cuda.empty_cache()
