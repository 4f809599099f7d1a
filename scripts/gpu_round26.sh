set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ecam_final|ecam_pool_kernel|ecam_bwd|stem_" -s 10 -c 8 -o gpurun_out/prof26 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu26.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
