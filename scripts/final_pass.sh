set -x
timeout 1400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_final.json | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 2600 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu list rc=$?"
