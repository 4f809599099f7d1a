set -x
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench29.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench29.log | cut -c1-400
