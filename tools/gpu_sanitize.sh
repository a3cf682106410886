#!/bin/bash
# compute-sanitizer passes over a small run of every kernel (memcheck) and of the smoke run (racecheck).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck: smoke"; timeout 900 $S --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_smoke.log
echo "== memcheck: fused / rate / dual RK64 / dodecane golden tests"
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "advance_chemistry or reaction_rate or (golden and dodecane_lu and cv_plm) or (golden and rk_mild) or subbox or edge_sizes" > gpurun_out/memcheck_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/memcheck_tests.log
echo "== memcheck: dual RK64"; PLM_RK64_DUAL=1 timeout 900 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and rk_mild and LiDryer" > gpurun_out/memcheck_dual.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/memcheck_dual.log
echo "== racecheck: smoke"; timeout 1500 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/racecheck_smoke.log
