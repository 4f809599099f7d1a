"""Layer microbenchmark of the tcgen05 conv / wgrad kernels at ChangeFormer decoder shapes (bs=32)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import IMPL_TC, CudaOps, View
dev, bf = "cuda:0", torch.bfloat16
ops = CudaOps(); N = 32
def buf(H, C): return View.alloc(N, H, H, C, bf, dev, zero=False)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def conv(H, cins, couts, phases=False, acc=False):
    srcs = [buf(H, c) for c in cins]
    for s in srcs: s.base.normal_()
    if phases:
        big = buf(2 * H, couts[0]); dsts = [big.phase(k // 2, k % 2) for k in range(4)]; cout = 4 * couts[0]
    else:
        dsts = [buf(H, c) for c in couts]; cout = sum(couts)
    cin = sum(cins)
    w = torch.randn(9 * cout * cin, device=dev).mul_(0.02).to(bf); bias = torch.zeros(cout, device=dev)
    fl = 2.0 * N * H * H * 9 * cin * cout
    return (lambda: ops.conv2d(N, H, H, 3, srcs, w, bias, dsts, [acc] * len(dsts), None, IMPL_TC)), fl
def wgrad(H, cin, cout, phases=False):
    x = buf(H, cin); x.base.normal_()
    if phases:
        big = buf(2 * H, cout); big.base.normal_(); dys = [big.phase(k // 2, k % 2) for k in range(4)]; ct = 4 * cout
    else:
        d = buf(H, cout); d.base.normal_(); dys = [d]; ct = cout
    dw = torch.zeros(9 * ct * cin, device=dev)
    fl = 2.0 * N * H * H * 9 * cin * ct
    return (lambda: ops.conv2d_wgrad(N, H, H, 3, [x], dys, dw, False, IMPL_TC)), fl
cases = {
  "fwd 256->256 @224": conv(224, [256], [256]), "fwd 256->256 @112": conv(112, [256], [256]),
  "fwd 256->256 @224 acc": conv(224, [256], [256], acc=True),
  "convT fwd 256->4x256 @112 (-> 224)": conv(112, [256], [256], phases=True),
  "fwd 512->256 @56": conv(56, [256, 256], [256]), "fwd 256->16 @224": conv(224, [256], [16]),
  "wgrad 256x256 @224": wgrad(224, 256, 256), "wgrad 256x256 @112": wgrad(112, 256, 256),
  "wgrad convT 256x(4x256) @112": wgrad(112, 256, 256, phases=True), "wgrad 256x16 @224": wgrad(224, 256, 16),
}
for name, (fn, fl) in cases.items():
    ms = timeit(fn); print(f"{name:40s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TF/s", flush=True)
    torch.cuda.empty_cache()
