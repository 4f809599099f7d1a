set -x
timeout 900 python -m pytest tests/test_gpu_siam.py -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_siam.log 2>&1; tail -3 gpurun_out/pytest_siam.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r21.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench21.log 2>&1; echo "ncu rc=$?"
