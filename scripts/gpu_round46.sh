set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench46.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench46.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench46_ref.log 2>&1; tail -1 gpurun_out/bench46_ref.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke46.log 2>&1; tail -3 gpurun_out/smoke46.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 340 -c 330 --csv --log-file gpurun_out/launches_r46.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench46.log 2>&1; echo "ncu rc=$?"
KS_VARIANTS=auto KS_REPS=1 KS_LAYERS="L0 fwd 32->32,L0 fwd 224->32 (160+64),L1 fwd 384->64 (256+128),L2 fwd 640->128" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -c 8 -o gpurun_out/prof46_conv python scripts/bench_layers.py > gpurun_out/ncu46_conv.log 2>&1; echo "ncu rc=$?"
