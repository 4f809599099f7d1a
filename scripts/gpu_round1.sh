set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m kurosiwo_b200.build 2>&1 | tail -2
timeout 600 python tests/tc_probe.py gpurun_out/tc_probe_first.json > gpurun_out/tc_probe_first.log 2>&1; echo "probe rc=$?"
tail -30 gpurun_out/tc_probe_first.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench1.log
