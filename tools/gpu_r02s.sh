#!/bin/bash
# nineteenth GPU pass (final of round 2): smoke, the default bench line (with the N = 8-only 256^3 entry forced at a small size to
# exercise its code path), the ncu launch list of a short bench, and a --set full capture of the shipped drm19 kernel
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
PLM_BENCH_256CUBED=1 PLM_BENCH_256CUBED_CELLS=131072 timeout 1500 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/r02s_bench.json; tail -3 gpurun_out/r02s_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 2 --warmup 1 --also-steps 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r02s_bench_under_ncu.log 2>&1; echo "launch list exit $?"; wc -l gpurun_out/r02s_launches.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:react_kernel -c 1 -f -o gpurun_out/r02s_drm19_warp python tools/one_react.py drm19 ReactorCvode 32768 warp > gpurun_out/r02s_ncu.log 2>&1; tail -1 gpurun_out/r02s_ncu.log
