# 2-GPU check of the sharded scan through bench.py (torchrun, NCCL)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err
cut -c1-2200 gpurun_out/bench_g2.json; tail -5 gpurun_out/bench_g2.err
