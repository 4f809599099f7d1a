set -x
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench17.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r17.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench17.log 2>&1; echo "ncu rc=$?"
