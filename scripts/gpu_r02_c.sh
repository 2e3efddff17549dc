# round 2, call C (2 GPUs): the bench under torchrun -- NCCL sharded query, sharded k-means, row-sharded leg
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02c_bench2.json 2> gpurun_out/r02c_bench2.err ) 2>&1 | grep real; echo "bench2 rc=$?"; cut -c1-6000 gpurun_out/r02c_bench2.json; tail -15 gpurun_out/r02c_bench2.err | cut -c1-400
timeout -s KILL 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k two_devices > gpurun_out/r02c_twodev.log 2>&1; echo "twodev rc=$?"; tail -5 gpurun_out/r02c_twodev.log | cut -c1-300
