"""Perf experiment (B200 only): ring depth of the streamed conv path on the ViT GEMM shapes."""
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
    fl = 2.0 * R * K * N
    out = []
    for (sa, sb) in ((0, 0), (2, 2), (3, 3), (4, 4), (5, 5), (6, 6), (4, 2), (6, 3), (8, 4)):
        ops.set_option("tc_sa", sa); ops.set_option("tc_sb", sb)
        try:
            t = timeit(lambda: ops.conv2d(1, R // 16, 16, 1, [a], w, b, [o], [False], None, IMPL_TC))
            out.append(f"sa{sa}/sb{sb}: {t*1e3:6.1f}us {fl/t/1e9:5.0f}TF")
        except Exception as e:
            out.append(f"sa{sa}/sb{sb}: ERR")
    ops.set_option("tc_sa", 0); ops.set_option("tc_sb", 0)
    print(f"GEMM {R}x{K}x{N}: " + " | ".join(out), flush=True)
