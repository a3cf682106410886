#!/bin/bash
# Round-end evidence pass, second half (after the Gauss-Jordan change): GPU suite, smoke, default bench, dodecane_lu / loose / LiDryer benches, ncu of dodecane_lu and drm19
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/e2_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/e2_smoke.log
echo "== bench default (drm19 2^22)"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/e2_bench_drm19.json; cut -c1-200 gpurun_out/e2_bench_drm19.json
for cfg in "dodecane --mech dodecane_lu --cells 65536 --box-cells 16384" "loose --cells 1048576 --rtol 1e-6 --max-order 4" "lidryer --mech LiDryer --cells 1048576"; do
  set -- $cfg; tag=$1; shift
  timeout 600 python bench.py "$@" --steps 2 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 > gpurun_out/e2_bench_$tag.json; echo "$tag: $(cut -c1-160 gpurun_out/e2_bench_$tag.json)"
done
echo "== ncu full"
bash tools/gpu_prof.sh e2_dodecane "--mech dodecane_lu --cells 4096 --box-cells 4096" e2_drm19 "--cells 32768 --box-cells 32768"
ls gpurun_out | grep "^e2_\|prof_e2"
