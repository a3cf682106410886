#!/bin/bash
# Round-end evidence pass: tests, smoke, benches of every headline configuration, launch list, full ncu captures.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/final_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "== bench default (drm19 2^22)"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench_drm19.json; cut -c1-200 gpurun_out/final_bench_drm19.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/final_bench_reference.json; cut -c1-200 gpurun_out/final_bench_reference.json
for cfg in "loose --cells 1048576 --rtol 1e-6 --max-order 4" "lidryer_rk64 --mech LiDryer --integrator ReactorRK64" "lidryer --mech LiDryer --cells 1048576" "dodecane --mech dodecane_lu --cells 65536 --box-cells 16384"; do
  set -- $cfg; tag=$1; shift
  timeout 600 python bench.py "$@" --steps 2 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 > gpurun_out/final_bench_$tag.json; echo "$tag: $(cut -c1-160 gpurun_out/final_bench_$tag.json)"
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --cells 65536 --box-cells 16384 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_ncu_launch.log 2>&1
echo "== ncu full"
bash tools/gpu_prof.sh final_drm19 "--cells 32768 --box-cells 32768" final_dodecane "--mech dodecane_lu --cells 4096 --box-cells 4096"
ls gpurun_out | head -50
