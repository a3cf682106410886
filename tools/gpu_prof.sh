#!/bin/bash
# ncu --set full captures of the integration kernel: usage tools/gpu_prof.sh TAG "bench args" [TAG "bench args" ...]
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; args=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:react_kernel -s 3 -c 1 -f -o gpurun_out/prof_$tag \
     python bench.py $args --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
