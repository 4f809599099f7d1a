set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vit.py tests/test_gpu_changeformer.py tests/test_gpu_snunet.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_55.log 2>&1; tail -5 gpurun_out/pytest_55.log
timeout 300 python scripts/prof_ops.py floodvit > gpurun_out/prof_vit55.log 2>&1; grep "permute_cast_table\|sum" gpurun_out/prof_vit55.log
timeout 300 python scripts/prof_ops.py changeformer > gpurun_out/prof_cf55.log 2>&1; grep "permute_cast_table\|sum" gpurun_out/prof_cf55.log
