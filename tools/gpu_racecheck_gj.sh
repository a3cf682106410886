#!/bin/bash
# racecheck over one dodecane_lu golden case (two pivots per sweep)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 170 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and dodecane_lu and cv_plm" > gpurun_out/racecheck_gj_dodecane.log 2>&1; echo "rc=$?"
grep -c "Race reported" gpurun_out/racecheck_gj_dodecane.log; grep -o "bdf_warp.cuh:[0-9]*\|chem_warp.cuh:[0-9]*\|kernels.cu:[0-9]*" gpurun_out/racecheck_gj_dodecane.log | sort | uniq -c | sort -rn | head -20; tail -3 gpurun_out/racecheck_gj_dodecane.log
