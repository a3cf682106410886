#!/bin/bash
# first GPU pass of round 2: lane = cell kernels -- A/B against the warp-per-cell kernels, parity suite, short benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 900 python tools/lane_ab.py > gpurun_out/r02a_lane_ab.log 2>&1; echo "ab exit $?"
tail -4 gpurun_out/r02a_lane_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 --cells 1048576 --no-cpu-baseline > gpurun_out/r02a_bench_drm19.json 2> gpurun_out/r02a_bench_drm19.err; echo "bench exit $?"
cat gpurun_out/r02a_bench_drm19.json
timeout 300 python bench.py --mech LiDryer --integrator ReactorRK64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_bench_lidryer_rk64.json 2> gpurun_out/r02a_bench_lidryer_rk64.err
cat gpurun_out/r02a_bench_lidryer_rk64.json
