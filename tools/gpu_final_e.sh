#!/bin/bash
# Round-end evidence pass (r01e): smoke, default bench, reference arm, the other headline configurations, launch list, ncu of drm19.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/e_smoke.log
echo "== bench default (drm19 2^22)"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/e_bench_drm19.json; cut -c1-200 gpurun_out/e_bench_drm19.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/e_bench_reference.json; cut -c1-200 gpurun_out/e_bench_reference.json
for cfg in "loose --cells 1048576 --rtol 1e-6 --max-order 4" "lidryer_rk64 --mech LiDryer --integrator ReactorRK64" "dodecane --mech dodecane_lu --cells 65536 --box-cells 16384"; do
  set -- $cfg; tag=$1; shift
  timeout 600 python bench.py "$@" --steps 2 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 > gpurun_out/e_bench_$tag.json; echo "$tag: $(cut -c1-160 gpurun_out/e_bench_$tag.json)"
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/e_launches.csv \
   python bench.py --cells 65536 --box-cells 16384 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/e_ncu_launch.log 2>&1
echo "== ncu full"
bash tools/gpu_prof.sh e_drm19 "--cells 32768 --box-cells 32768"
ls gpurun_out | grep "^e_\|prof_e"
