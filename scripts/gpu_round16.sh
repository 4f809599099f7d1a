set -x
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench16.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench16_ref.log 2>&1; tail -1 gpurun_out/bench16_ref.log
