set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vit.py tests/test_gpu_changeformer.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_57.log 2>&1; tail -3 gpurun_out/pytest_57.log
for wl in snunet floodvit floodvit-upernet changeformer; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench57_$wl.log 2>&1; tail -1 gpurun_out/bench57_$wl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
