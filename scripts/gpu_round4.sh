set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 600 python tests/tc_probe.py gpurun_out/tc_probe4.json > gpurun_out/tc_probe4.log 2>&1; echo "probe rc=$?"; tail -16 gpurun_out/tc_probe4.log | cut -c1-330
timeout 900 python scripts/bench_layers.py gpurun_out/layers4.json > gpurun_out/layers4.log 2>&1; echo "layers rc=$?"; cat gpurun_out/layers4.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
