#!/bin/bash
# tests + bench + ncu launch list + one full capture of the integration kernel
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "Error|assert|passed|failed" | cut -c1-250 | tail -20 | tee gpurun_out/pytest_gpu.log
echo "== bench"
timeout 300 python bench.py --cells 262144 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_small.log
timeout 300 python bench.py --mech LiDryer --cells 262144 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_lidryer.log
timeout 600 python bench.py --mech dodecane_lu --cells 16384 --box-cells 16384 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_dodecane.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:react_kernel -s 3 -c 1 -f -o gpurun_out/prof_react \
   python bench.py --cells 32768 --box-cells 32768 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
