#!/bin/bash
# quick GPU iteration: parity tests + small bench per mechanism
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -${PYTAIL:-15} | tee gpurun_out/pytest_gpu.log
echo "== bench" 
timeout 300 python bench.py --cells 262144 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_small.log
timeout 300 python bench.py --mech LiDryer --cells 262144 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_lidryer.log
timeout 600 python bench.py --mech dodecane_lu --cells 16384 --box-cells 16384 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_dodecane.log
