set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench54.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench54.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench54_ref.log 2>&1; tail -1 gpurun_out/bench54_ref.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke54.log 2>&1; tail -3 gpurun_out/smoke54.log
for wl in floodvit floodvit-upernet changeformer siam-conc; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench54_$wl.log 2>&1; tail -1 gpurun_out/bench54_$wl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 340 -c 330 --csv --log-file gpurun_out/launches_r54.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench54.log 2>&1; echo "ncu rc=$?"
