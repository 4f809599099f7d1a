set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log
for mt in 1 2 4; do
  timeout 600 python bench.py --steps 5 --warmup 3 --tc-mt $mt --no-cpu-baseline > gpurun_out/bench_mt$mt.log 2>&1; echo "bench mt=$mt rc=$?"; tail -1 gpurun_out/bench_mt$mt.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
  cp gpurun_out/bench_layers.json gpurun_out/bench_layers_mt$mt.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
