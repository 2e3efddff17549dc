mkdir -p gpurun_out
N=${NG:-8}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench$N rc=$?"; grep -v "^NCCL" gpurun_out/bench_g$N.json | cut -c1-400; tail -3 gpurun_out/bench_g$N.err
