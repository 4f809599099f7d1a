set -x
timeout 600 python -m pytest tests/test_gpu_cformer_kernels.py tests/test_gpu_changeformer.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_cf49.log 2>&1; tail -15 gpurun_out/pytest_cf49.log
timeout 600 python bench.py --workload changeformer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench49_cf.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench49_cf.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e'])"
timeout 300 python scripts/prof_ops.py changeformer > gpurun_out/prof_cf49.log 2>&1; head -30 gpurun_out/prof_cf49.log
