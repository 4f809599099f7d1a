# round 2c closing pass (one gpurun call): full -m gpu suite, smoke(), the default bench line, the ChangeFormer per-op table,
# ncu --set full of the three depth-wise conv tile kernels at the stage-1 shape
set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2c_bench_final.json 2> gpurun_out/r2c_bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/r2c_bench_final.json | cut -c1-200
timeout 120 python scripts/prof_ops.py changeformer 32 2>&1 | head -26 > gpurun_out/r2c_prof_changeformer.txt; tail -3 gpurun_out/r2c_prof_changeformer.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dwconv -c 3 -o gpurun_out/r2c_dwconv -f python scripts/bench_dwconv.py 64 > gpurun_out/r2c_ncu_dwconv.log 2>&1; echo "ncu rc=$?"
