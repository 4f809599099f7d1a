set -x
timeout 600 python -m pytest tests/test_gpu_snunet.py -q -m gpu --timeout 600 -p no:cacheprovider -k "step_host" > gpurun_out/pytest_pl.log 2>&1; tail -15 gpurun_out/pytest_pl.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench48.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench48.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'])"
timeout 600 python bench.py --workload floodvit --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench48_vit.log 2>&1; tail -1 gpurun_out/bench48_vit.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e'])"
