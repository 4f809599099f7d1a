"""Perf experiment (B200 only): HBM throughput of the BatchNorm passes at the SNUNet level-0/1 shapes vs grid size."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import CudaOps, View
ops = CudaOps(); dev = "cuda:0"; bf = torch.bfloat16; N = 64
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (H, C) in ((224, 32), (112, 64)):
    mk = lambda: View.alloc(N, H, H, C, bf, dev, zero=False)
    y, out, dout, dy, res = mk(), mk(), mk(), mk(), mk()
    for v in (y, dout, res): v.base.normal_()
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1; mu = torch.zeros(C, device=dev); rs = torch.ones(C, device=dev)
    gamma = torch.ones(C, device=dev); sums = torch.zeros(2 * C, dtype=torch.float64, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    nbytes = N * H * H * C * 2
    for cap in (0, 2, 4, 6, 12, 16, 24, 32):
        ops.set_option("ew_cap", cap)
        t1 = timeit(lambda: ops.bn_act(y, sc, sh, None, True, out, None))
        t1r = timeit(lambda: ops.bn_act(y, sc, sh, res, True, out, None))
        t2 = timeit(lambda: ops.bn_bwd_reduce(dout, out, y, None, None, mu, rs, sums))
        t2b = timeit(lambda: ops.bn_bwd_reduce(dout, None, y, sc, sh, mu, rs, sums))
        t3 = timeit(lambda: ops.bn_bwd_apply(dout, True, y, None, None, mu, rs, gamma, sums, float(N * H * H), None, dy, dg, db, None, False))
        print(f"H={H} C={C} cap={cap:2d}: bn_act {t1*1e3:7.1f} us {2*nbytes/t1/1e6:6.0f} GB/s | +res {t1r*1e3:7.1f} us {3*nbytes/t1r/1e6:6.0f} | "
              f"reduce(out) {t2*1e3:7.1f} us {3*nbytes/t2/1e6:6.0f} | reduce(y) {t2b*1e3:7.1f} us {2*nbytes/t2b/1e6:6.0f} | apply {t3*1e3:7.1f} us {3*nbytes/t3/1e6:6.0f}", flush=True)
    ops.set_option("ew_cap", 0)
    # reference points: torch copy and read-only sum on the same buffers
    t = timeit(lambda: out.base.copy_(y.base)); print(f"  torch copy_ {t*1e3:7.1f} us {2*nbytes/t/1e6:6.0f} GB/s")
    t = timeit(lambda: torch.relu_(out.base)); print(f"  torch relu_ (in place) {t*1e3:7.1f} us {2*nbytes/t/1e6:6.0f} GB/s")
    t = timeit(lambda: y.base.view(torch.int16).sum(dtype=torch.int64)); print(f"  torch sum {t*1e3:7.1f} us {nbytes/t/1e6:6.0f} GB/s")
