#!/bin/bash
# eleventh GPU pass: FlameSheet through the adaptor (2-D box bounds fixed), racecheck after the Gauss-Jordan read fixes, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -s -k flamesheet > gpurun_out/r02k_flamesheet.log 2>&1; echo "flamesheet exit $?"
grep -E "FlameSheet 3 steps|passed|failed" gpurun_out/r02k_flamesheet.log | tail -4
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02k_pytest.log
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -6 gpurun_out/r02k_racecheck_smoke.log
timeout 1200 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and dodecane_lu and cv_tight" > gpurun_out/r02k_racecheck_dodecane.log 2>&1; echo "racecheck dodecane rc=$?"; tail -6 gpurun_out/r02k_racecheck_dodecane.log
