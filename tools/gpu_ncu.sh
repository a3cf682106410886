#!/bin/bash
mkdir -p gpurun_out
MECH=${MECH:-drm19}; CELLS=${CELLS:-32768}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:react_kernel -s 3 -c 1 -f -o gpurun_out/prof_react_$MECH \
   python bench.py --mech $MECH --cells $CELLS --box-cells $CELLS --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_$MECH.log 2>&1
tail -1 gpurun_out/ncu_full_$MECH.log | cut -c1-200
