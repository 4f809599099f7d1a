import sys, torch
sys.path.insert(0, "/root/repo")
from kurosiwo_b200.lib import CudaOps, View
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
for (R, C) in ((13312, 768), (13312, 3072), (64 * 3136, 256)):
    m = torch.randn(R, C, device=dev).bfloat16()
    out = torch.zeros(C, device=dev)
    for rows in (8, 16, 32, 64, 128):
        ops.set_option("cs_rows", rows)
        def fn():
            for c0 in range(0, C, 1024):
                c = min(1024, C - c0)
                v = View(m.view(-1), c0, 1, R // 16, 16, c, R * C, 16 * C, C)
                ops.channel_sum(v, out[c0:c0 + c], True)
        print(R, C, "rows/lane", rows, f"{t(fn):.1f} us", f"{R*C*2/t(fn)/1e6:.2f} TB/s")
