set -x
timeout 900 python -m pytest tests/test_gpu_vit_kernels.py tests/test_gpu_changeformer.py tests/test_gpu_cformer_kernels.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_60.log 2>&1; tail -3 gpurun_out/pytest_60.log
timeout 300 python scripts/prof_ops.py changeformer > gpurun_out/prof_cf60.log 2>&1; head -24 gpurun_out/prof_cf60.log
