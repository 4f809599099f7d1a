set -x
timeout 300 python tests/tc_probe.py gpurun_out/tc_probe14.json > gpurun_out/tc_probe14.log 2>&1; echo "probe rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/tc_probe14.json'))
bad=[(k,kk,vv) for k,v in d['conv'].items() for kk,vv in v.items() if not (isinstance(vv,float) and vv<5e-3)]
print("conv bad:",bad); print("wgrad bad:", {k:v for k,v in d['wgrad'].items() if 'error' in v or v['assign']>5e-3})
PY
for pl in 0 1; do
KS_PLANAR_SLOTS=$pl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_$pl.log 2>&1; echo "bench planar=$pl rc=$?"; tail -1 gpurun_out/bench14_$pl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
done
KS_PLANAR_SLOTS=1 timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python scripts/dbg_epilogue.py 2>&1 | tail -6 | cut -c1-120
