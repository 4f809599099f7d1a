set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 600 python tests/tc_probe.py gpurun_out/tc_probe3.json > gpurun_out/tc_probe3.log 2>&1; echo "probe rc=$?"; cat gpurun_out/tc_probe3.log | tail -20
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -30 gpurun_out/pytest_gpu.log
for v1 in 0 1; do
  KS_V1=$v1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tc-v1 $v1 > gpurun_out/bench_v1_$v1.log 2>&1; echo "bench v1=$v1 rc=$?"; tail -1 gpurun_out/bench_v1_$v1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
  cp gpurun_out/bench_layers.json gpurun_out/bench_layers_v1_$v1.json
done
