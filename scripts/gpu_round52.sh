set -x
timeout 600 python -m pytest tests/test_gpu_vit_kernels.py tests/test_gpu_vit.py tests/test_gpu_changeformer.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_52.log 2>&1; tail -5 gpurun_out/pytest_52.log
timeout 300 python scripts/prof_ops.py floodvit > gpurun_out/prof_vit52.log 2>&1; head -12 gpurun_out/prof_vit52.log
timeout 300 python scripts/prof_ops.py changeformer > gpurun_out/prof_cf52.log 2>&1; head -12 gpurun_out/prof_cf52.log
