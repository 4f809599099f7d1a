"""Perf experiment: is the A-ring handshake latency-bound? Sweep pipeline depth / MT (B200 only)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
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
  "L0 fwd 224->32": (224, [160, 64], [32], 3),
  "L0 fwd 128->32": (224, [64, 64], [32], 3),
  "L0 dgrad 32->224": (224, [32], [64, 96, 64], 3),
  "L1 fwd 384->64": (112, [256, 128], [64], 3),
}
for name, (H, cins, couts, ks) in cases.items():
    srcs = [buf(H, c) for c in cins]; dsts = [buf(H, c) for c in couts]
    cin, cout = sum(cins), sum(couts)
    w = torch.randn(ks * ks * cout * cin, device=dev).mul_(0.05).to(bf)
    b = torch.zeros(cout, device=dev)
    for mt in (1, 2):
        row = []
        for sa in (2, 3, 4, 6, 8):
            ops.set_option("tc_mt", mt); ops.set_option("tc_sa", sa)
            r = []
            for dbg in (0, 28):
                ops.set_option("tc_debug", dbg)
                try:
                    r.append(round(timeit(lambda: ops.conv2d(N, H, H, ks, srcs, w, b, dsts, None, None, IMPL_TC)), 3))
                except Exception as e:
                    r.append("err")
            row.append(f"sa{sa}={r[0]}/{r[1]}")
        print(f"{name:20s} mt={mt}  " + "  ".join(row), flush=True)
    ops.set_option("tc_debug", 0); ops.set_option("tc_mt", 0); ops.set_option("tc_sa", 0)
    del srcs, dsts; torch.cuda.empty_cache()
