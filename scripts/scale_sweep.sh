#!/bin/bash
# scripts/scale_sweep.sh N [workloads...]: bench.py at N GPUs for each workload -> gpurun_out/r2_scale_N<N>.jsonl (one JSON line each)
N=$1; shift
WLS=${@:-"snunet changeformer floodvit-upernet"}
mkdir -p gpurun_out
out=gpurun_out/r2_scale_N${N}.jsonl
: > $out
port=29520
for wl in $WLS; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    KS_BENCH_WATCHDOG=400 timeout 500 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-library-baseline 2>gpurun_out/r2_scale_N${N}_$wl.err | tail -1 >> $out
  else
    KS_BENCH_WATCHDOG=400 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_scale_N${N}_$wl.err | tail -1 >> $out
  fi
  echo "$wl N=$N rc=$?"
done
python - <<PY
import json
for ln in open("$out"):
    try:
        d = json.loads(ln)
        print(d["config"]["parallelism"], d["metric"][:60], round(d["value"], 1), "patches/s", round(d["ms_per_step"], 2), "ms", "e2e", round(d["e2e"]["value"], 1), d.get("allreduce"))
    except Exception as e:
        print("bad line:", ln[:200], e)
PY
