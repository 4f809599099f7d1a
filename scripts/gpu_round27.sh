set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "ecam" > gpurun_out/pytest_ecam.log 2>&1; tail -30 gpurun_out/pytest_ecam.log
timeout 900 python -m pytest tests/test_gpu_snunet.py -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_snunet.log 2>&1; tail -5 gpurun_out/pytest_snunet.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench27.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench27.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 340 -c 330 --csv --log-file gpurun_out/launches_r27.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench27.log 2>&1; echo "ncu rc=$?"
