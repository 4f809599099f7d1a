set -x
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench35_snunet.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench35_snunet.log | cut -c1-300
timeout 600 python bench.py --workload floodvit --steps 10 --warmup 3 > gpurun_out/bench35_floodvit.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench35_floodvit.log | cut -c1-1200
timeout 600 python bench.py --workload siam-conc --steps 20 --warmup 3 > gpurun_out/bench35_siam.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench35_siam.log | cut -c1-1200
