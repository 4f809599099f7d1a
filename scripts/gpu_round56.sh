set -x
timeout 900 python -m pytest tests/test_gpu_vit_kernels.py tests/test_gpu_vit.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_56.log 2>&1; tail -5 gpurun_out/pytest_56.log
timeout 300 python scripts/prof_ops.py floodvit-upernet > gpurun_out/prof_up56.log 2>&1; head -32 gpurun_out/prof_up56.log
for wl in floodvit floodvit-upernet; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench56_$wl.log 2>&1; tail -1 gpurun_out/bench56_$wl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
