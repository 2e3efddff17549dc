#!/bin/bash
# round 2, call zh (8 GPUs): bench.py --gpus 8 with every leg (c2 split over query groups, row-sharded c4, sharded c3 training, c5)
mkdir -p gpurun_out
timeout -s KILL 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02zs_bench8.json 2> gpurun_out/r02zs_bench8.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r02zs_bench8.json; tail -5 gpurun_out/r02zs_bench8.err
