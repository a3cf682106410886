#!/bin/bash
# memcheck over the reworked inverse: the smoke run (one row per lane, pipelined sweep) and one dodecane_lu golden case (two pivots per sweep)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck: dodecane_lu golden"; timeout 110 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and dodecane_lu and cv_plm" > gpurun_out/memcheck_gj_dodecane.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_gj_dodecane.log
echo "== memcheck: smoke"; timeout 80 $S --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_gj_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_gj_smoke.log
