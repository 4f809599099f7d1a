"""How much of the CUDA-graph step is idle between kernels?  torch.profiler (CUPTI) timeline of graph replays of a bench workload:
prints the span, the sum of kernel durations and the gap histogram.   python scripts/gap_probe.py snunet"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["snunet"])
wl = sys.argv[1]
import runpy
ns = {"__file__": str(Path(__file__).resolve().parent / "ab_step.py"), "__name__": "ab_step_prefix"}
src = (Path(__file__).resolve().parent / "ab_step.py").read_text().split("sets = [{}]")[0]
exec(compile(src, "ab_step_prefix", "exec"), ns)
eng, inputs = ns["eng"], ns["inputs"]
eng.graph = None
eng.capture(*inputs, warmup=2)
for _ in range(3):
    eng.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 0],
             key=lambda e: e.time_range.start)
n = len(evs) // 3
evs = evs[n:2 * n]          # the middle replay
span = evs[-1].time_range.end - evs[0].time_range.start
busy = sum(e.time_range.elapsed_us() for e in evs)
gaps = [evs[i + 1].time_range.start - evs[i].time_range.end for i in range(len(evs) - 1)]
pos = [g for g in gaps if g > 0]
print(f"{wl}: {len(evs)} kernels, span {span / 1e3:.3f} ms, sum of kernel durations {busy / 1e3:.3f} ms, idle {sum(pos) / 1e3:.3f} ms "
      f"({100 * sum(pos) / span:.1f} %), mean gap {sum(pos) / max(1, len(pos)):.2f} us, max gap {max(gaps):.1f} us, overlapped {sum(1 for g in gaps if g < 0)}")
