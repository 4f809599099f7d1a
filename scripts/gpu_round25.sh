set -x
timeout 900 python -m pytest tests/test_gpu_vit_kernels.py tests/test_gpu_vit.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_vit.log 2>&1; tail -30 gpurun_out/pytest_vit.log
timeout 600 python scripts/bench_vit.py 64 > gpurun_out/bench_vit.log 2>&1; tail -24 gpurun_out/bench_vit.log
