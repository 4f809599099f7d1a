set -x
timeout 300 python scripts/prof_ops.py floodvit > gpurun_out/prof_vit50.log 2>&1; head -24 gpurun_out/prof_vit50.log
timeout 300 python scripts/prof_ops.py floodvit-upernet > gpurun_out/prof_up50.log 2>&1; head -30 gpurun_out/prof_up50.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench50.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench50.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e'])"
