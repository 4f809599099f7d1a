set -x
timeout 900 python -m pytest tests/test_gpu_siam.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_siam.log 2>&1; tail -40 gpurun_out/pytest_siam.log
