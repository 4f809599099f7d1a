set -x
timeout 900 python -m pytest tests/test_gpu_vit.py -q -m gpu --timeout 600 -p no:cacheprovider -s -k "upernet or adaptive" > gpurun_out/pytest_up.log 2>&1; tail -30 gpurun_out/pytest_up.log
timeout 600 python bench.py --workload floodvit-upernet --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench47_up.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench47_up.log | cut -c1-400
