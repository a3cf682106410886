#!/bin/bash
# fifteenth GPU pass: the soft (membership) barrier in front of every right-hand side -- a warp leaves the group during a
# linear-solver setup and when it runs out of cells -- in CTAs of 4, 8, 16 warps (PLM_SOFTALIGN), against the shipped kernel
# (free-running warps).  Throughput, then the instruction-fetch counters of each.
mkdir -p gpurun_out
M=sm__icc_request_hit_rate.pct,smsp__issue_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.per_cycle_active
for v in base sal4 sal8 sal16; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== $v"
  PLM_LIBPATH=$PWD/$lib timeout 120 python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -2
  PLM_LIBPATH=$PWD/$lib timeout 200 ncu --metrics $M --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py drm19 ReactorCvode 32768 warp 2>&1 | grep -E "icc_request|issue_active|stalled|type_instruction|inst_executed|warps_active" | sed 's/  */ /g'
done 2>&1 | tee gpurun_out/r02o_softalign_ab.txt
