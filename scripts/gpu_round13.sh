set -x
timeout 300 python tests/tc_probe.py gpurun_out/tc_probe13.json > gpurun_out/tc_probe13.log 2>&1; echo "probe rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/tc_probe13.json'))
bad=[(k,kk,vv) for k,v in d['conv'].items() for kk,vv in v.items() if not (isinstance(vv,float) and vv<5e-3)]
print("conv bad:",bad); print("wgrad bad:", {k:v for k,v in d['wgrad'].items() if 'error' in v or v['assign']>5e-3})
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench13.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
