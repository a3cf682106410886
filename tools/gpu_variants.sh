#!/bin/bash
# Development: time library variants (PLM_LIBPATH) on short benches.  usage: tools/gpu_variants.sh "v1 v2 ..." [mechs]
mkdir -p gpurun_out
variants="$1"; mechs="${2:-drm19}"
for v in $variants; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  for m in $mechs; do
    case $m in
      drm19) args="--cells 262144";;
      LiDryer) args="--mech LiDryer --cells 262144";;
      rk64) args="--mech LiDryer --integrator ReactorRK64 --cells 131072";;
      dodecane_lu) args="--mech dodecane_lu --cells 16384 --box-cells 16384";;
      dodecane64k) args="--mech dodecane_lu --cells 65536 --box-cells 16384";;
      loose) args="--cells 1048576 --rtol 1e-6 --max-order 4";;
    esac
    r=$(PLM_LIBPATH=$PWD/$lib timeout 300 python bench.py $args --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1)
    echo "$v $m $(echo "$r" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("%.0f cells/s  %.1f ms  frac %.4f nfe %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["per_cell"]["nfe"]))
except Exception as e: print("FAILED", e)')" | tee -a gpurun_out/variants.log
  done
done
