set -x
timeout 300 python tests/tc_probe.py gpurun_out/tc_probe15.json > gpurun_out/tc_probe15.log 2>&1; echo "probe rc=$?"; tail -11 gpurun_out/tc_probe15.log | cut -c1-220
KS_VARIANTS=auto timeout 600 python scripts/bench_layers.py gpurun_out/layers15.json > gpurun_out/layers15.log 2>&1; echo "layers rc=$?"; tail -7 gpurun_out/layers15.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench15.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench15.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
