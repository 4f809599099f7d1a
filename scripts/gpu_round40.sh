set -x
timeout 900 python -m pytest tests/test_gpu_changeformer.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_cf.log 2>&1; tail -50 gpurun_out/pytest_cf.log
timeout 600 python bench.py --workload changeformer --steps 5 --warmup 3 > gpurun_out/bench40_cf.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench40_cf.log | cut -c1-1500
