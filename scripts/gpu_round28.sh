set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_snunet.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_28.log 2>&1; tail -8 gpurun_out/pytest_28.log
KS_VARIANTS=auto python scripts/bench_layers.py gpurun_out/layers28.json > gpurun_out/layers28.log 2>&1; cat gpurun_out/layers28.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench28.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench28.log | cut -c1-300
