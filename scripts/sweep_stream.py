"""Sweep (MT, TB, SA, SB) of the streamed-weight conv path on the mid-level SNUNet layers (B200 only)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import IMPL_TC, CudaOps, KsError, View
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
  "L1 64->64": (112, [64], [64]), "L1 384->64": (112, [256, 128], [64]), "L1 dgrad 64->384": (112, [64], [256, 128]),
  "L2 128->128": (56, [128], [128]), "L2 640->128": (56, [384, 256], [128]), "L2 dgrad 128->640": (56, [128], [384, 256]),
  "L3 256->256": (28, [256], [256]), "L3 1024->256": (28, [512, 512], [256]), "L4 512->512": (14, [512], [512]),
}
for name, (H, cins, couts) in cases.items():
    srcs = [buf(H, c) for c in cins]; dsts = [buf(H, c) for c in couts]
    cin, cout = sum(cins), sum(couts)
    w = torch.randn(9 * cout * cin, device=dev).mul_(0.05).to(bf); b = torch.zeros(cout, device=dev)
    st = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if len(couts) == 1 else None
    fl = 2.0 * N * H * H * 9 * cin * cout
    res = []
    for mt in (1, 2, 4):
        for tb in (1, 3):
            for (sa, sb) in ((2, 2), (3, 3), (3, 4), (4, 4), (2, 4)):
                for o, v in (("tc_no_resident", 1), ("tc_mt", mt), ("tc_tb", tb), ("tc_sa", sa), ("tc_sb", sb)):
                    ops.set_option(o, v)
                try:
                    ms = timeit(lambda: ops.conv2d(N, H, H, 3, srcs, w, b, dsts, None, st, IMPL_TC), 4)
                    res.append((ms, f"mt{mt} tb{tb} sa{sa} sb{sb}"))
                except KsError:
                    pass
    for o in ("tc_no_resident", "tc_mt", "tc_tb", "tc_sa", "tc_sb"):
        ops.set_option(o, 0)
    auto = timeit(lambda: ops.conv2d(N, H, H, 3, srcs, w, b, dsts, None, st, IMPL_TC), 4)
    res.sort()
    print(f"{name:20s} auto={auto:.4f} ({fl / auto / 1e9:.0f} TF)  best: " + "  ".join(f"{n}={ms:.4f}" for ms, n in res[:5]) + f"   worst={res[-1][1]}={res[-1][0]:.4f}", flush=True)
    del srcs, dsts; torch.cuda.empty_cache()
