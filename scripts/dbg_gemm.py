"""Perf experiment (B200 only): the conv engine as a plain GEMM on the ViT shapes ([13312 tokens] x K x N), stage by stage ablation."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import IMPL_TC, CudaOps, View
ops = CudaOps(); dev = "cuda:0"; bf = torch.bfloat16; R = 13312
def mat(C):
    t = torch.randn(R * C, device=dev).to(bf)
    return View(t, 0, 1, R // 16, 16, C, R * C, 16 * C, C)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (K, N) in ((768, 768), (768, 2304), (768, 3072), (3072, 768)):
    a, o = mat(K), mat(N)
    w = torch.randn(N * K, device=dev).mul_(0.03).to(bf)
    b = torch.zeros(N, device=dev)
    row = {}
    for dbg in (0, 1, 8, 8 + 4, 16, 16 + 8 + 4):
        ops.set_option("tc_debug", dbg)
        row[dbg] = timeit(lambda: ops.conv2d(1, R // 16, 16, 1, [a], w, b, [o], [False], None, IMPL_TC))
    ops.set_option("tc_debug", 0)
    fl = 2.0 * R * K * N
    print(f"GEMM {R}x{K}x{N}: full {row[0]*1e3:7.1f} us ({fl/row[0]/1e9:6.0f} TF/s) nostore {row[1]*1e3:7.1f} noepi {row[8]*1e3:7.1f} noepi+nomma {row[12]*1e3:7.1f} notma {row[16]*1e3:7.1f} onlybarriers {row[28]*1e3:7.1f}", flush=True)
    for mt in (1, 2):
        ops.set_option("tc_mt", mt)
        t = timeit(lambda: ops.conv2d(1, R // 16, 16, 1, [a], w, b, [o], [False], None, IMPL_TC))
        print(f"   tc_mt={mt}: {t*1e3:7.1f} us ({fl/t/1e9:6.0f} TF/s)")
    ops.set_option("tc_mt", 0)
    # weight gradient of the same layer
    dw = torch.zeros(N * K, device=dev)
    t = timeit(lambda: ops.conv2d_wgrad(1, R // 16, 16, 1, [a], [o], dw, False, IMPL_TC))
    print(f"   wgrad: {t*1e3:7.1f} us ({fl/t/1e9:6.0f} TF/s)")
