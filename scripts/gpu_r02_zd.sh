#!/bin/bash
# round 2, call zd (2 GPUs): the tensor scan under torchrun: sharded tests, then bench.py --gpus 2 with every leg
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_tscan.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r02zd_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02zd_tests.log | cut -c1-600
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02zd_bench2.json 2> gpurun_out/r02zd_bench2.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r02zd_bench2.json; tail -5 gpurun_out/r02zd_bench2.err
