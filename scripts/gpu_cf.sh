set -x
timeout 600 python -m pytest tests/test_gpu_cformer_kernels.py tests/test_gpu_changeformer.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_cf.log 2>&1; tail -5 gpurun_out/pytest_cf.log
timeout 300 python scripts/prof_ops.py changeformer > gpurun_out/prof_cf.log 2>&1; head -8 gpurun_out/prof_cf.log
timeout 600 python bench.py --workload changeformer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cf.log 2>&1; tail -1 gpurun_out/bench_cf.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
