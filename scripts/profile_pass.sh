#!/bin/bash
# scripts/profile_pass.sh: the measurement pass behind profiles/r2_* (run under gpurun on ONE B200).
#   1. default bench.py line (with the CPU port and the stock-torch comparator)              -> gpurun_out/r2_bench_final.json
#   2. ncu launch list of ONE eager SNUNet step (time + DRAM bytes per launch)              -> gpurun_out/r2_launches_raw.csv
#   3. ncu --set full of the kernels added in the second half of round 2                    -> gpurun_out/r2_new_kernels.ncu-rep
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_final.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 2600 --csv \
  --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ce_dice_resident|stem_fwd_mma|stem_wgrad_mma|ecam_final_mma" -c 8 --launch-skip 8 \
  -o gpurun_out/r2_new_kernels -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline > gpurun_out/r2_new_kernels.log 2>&1; echo "ncu full rc=$?"
