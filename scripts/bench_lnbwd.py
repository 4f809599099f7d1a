"""In-graph time of ks_layernorm_bwd (rows + cols kernels) on token matrices; LN_ROWS sweep of the column pass."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import CudaOps
ops = CudaOps(); dev = "cuda:0"
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (R, C) in ((13312, 768), (64 * 3136, 64), (64 * 784, 128)):
    dy = torch.randn(R, C, device=dev).bfloat16(); x = torch.randn(R, C, device=dev).bfloat16(); dx = torch.zeros_like(x)
    mean = torch.zeros(R, device=dev); rstd = torch.ones(R, device=dev); gamma = torch.ones(C, device=dev)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    for rows in (16, 32, 64, 128):
        ops.set_option("ln_rows", rows)
        both = t(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, True, dg, db))
        rows_only = t(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, True, None, None))
        print(R, C, "ln_rows", rows, f"both {both:.1f} us, rows kernel {rows_only:.1f} us, cols kernel {both - rows_only:.1f} us; traffic {4*R*C*2/1e6:.0f}+{2*R*C*2/1e6:.0f} MB")
