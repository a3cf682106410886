#!/bin/bash
# configs[4] at 2 GPUs: knapsack on measured FC, box payloads over NCCL both ways
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload tripleflame --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/amr_bench_drm19_n2.json; cut -c1-1500 gpurun_out/amr_bench_drm19_n2.json
