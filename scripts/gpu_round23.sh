set -x
timeout 900 python -m pytest tests/test_gpu_vit_kernels.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_vitk.log 2>&1; tail -60 gpurun_out/pytest_vitk.log
