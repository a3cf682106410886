#!/bin/bash
# One GPU pass: parity tests, smoke, short benches, ncu launch list + one full capture of the integration kernel.
# usage: tools/gpu_pass.sh [tests] [bench] [ncu]   (default: all three)
what="${*:-tests bench ncu}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [[ $what == *tests* ]]; then
  echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
  echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
fi
if [[ $what == *bench* ]]; then
  echo "== bench drm19 2^20"
  timeout 600 python bench.py --cells 1048576 --steps 2 --warmup 3 --cpu-seconds 5 2>&1 | tail -2 | tee gpurun_out/bench_drm19.log
  echo "== bench drm19 PeleLM-like tolerances (rtol 1e-6, order 4)"
  timeout 300 python bench.py --cells 1048576 --rtol 1e-6 --max-order 4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_drm19_loose.log
  echo "== bench LiDryer RK64 / LiDryer cvode / dodecane"
  timeout 300 python bench.py --mech LiDryer --integrator ReactorRK64 --cells 262144 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_lidryer_rk64.log
  timeout 300 python bench.py --mech LiDryer --cells 1048576 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_lidryer.log
  timeout 600 python bench.py --mech dodecane_lu --cells 16384 --box-cells 16384 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | tee gpurun_out/bench_dodecane.log
fi
if [[ $what == *ncu* ]]; then
  echo "== ncu launches"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
     python bench.py --cells 65536 --box-cells 16384 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch.log 2>&1
  echo "== ncu full"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:react_kernel -s 3 -c 1 -f -o gpurun_out/prof_react \
     python bench.py --cells 32768 --box-cells 32768 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
