#!/bin/bash
# configs[3] scaled down: dodecane_lu over 2 GPUs, boxes dealt by the knapsack on measured FC
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mech dodecane_lu --cells 32768 --box-cells 4096 --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/e2_bench_dodecane_n2.json; cut -c1-1200 gpurun_out/e2_bench_dodecane_n2.json
