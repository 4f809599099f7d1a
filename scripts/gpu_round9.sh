set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 600 python tests/tc_probe.py gpurun_out/tc_probe9.json > gpurun_out/tc_probe9.log 2>&1; echo "probe rc=$?"; grep -c error gpurun_out/tc_probe9.log; tail -6 gpurun_out/tc_probe9.log | cut -c1-200
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
KS_VARIANTS=auto timeout 900 python scripts/bench_layers.py gpurun_out/layers9.json > gpurun_out/layers9.log 2>&1; echo "layers rc=$?"; cat gpurun_out/layers9.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench9.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
