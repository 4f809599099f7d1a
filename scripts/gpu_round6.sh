set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 600 python tests/tc_probe.py gpurun_out/tc_probe6.json > gpurun_out/tc_probe6.log 2>&1; echo "probe rc=$?"; grep -c error gpurun_out/tc_probe6.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_layers.py gpurun_out/layers6.json > gpurun_out/layers6.log 2>&1; echo "layers rc=$?"; cat gpurun_out/layers6.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kind'], d['roofline']['conv_ms_per_step'])"
# ncu: full sections for the bandwidth-bound BN passes (first launches of the backward = level 0)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_bwd_apply|bn_bwd_reduce|bn_act_kernel' -c 6 -o gpurun_out/prof_bn python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline > gpurun_out/ncu_bn.log 2>&1; echo "ncu bn rc=$?"
KS_LAYERS="L0 fwd 224->32 (160+64),L0 dgrad 32->224,L1 fwd 384->64 (256+128),L0 wgrad 32x32,L0 wgrad 224x32" KS_VARIANTS=auto KS_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|wgrad_tc' -o gpurun_out/prof_conv python scripts/bench_layers.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
