set -x
KS_VARIANTS=auto python scripts/bench_layers.py gpurun_out/layers22.json > gpurun_out/layers22.log 2>&1; cat gpurun_out/layers22.log
KS_VARIANTS=auto KS_REPS=1 KS_LAYERS="L0 fwd 32->32,L0 fwd 224->32 (160+64),L0 dgrad 32->224,L1 fwd 64->64" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -c 8 -o gpurun_out/prof22 python scripts/bench_layers.py > gpurun_out/ncu22.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/
