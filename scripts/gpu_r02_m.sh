# round 2, call M (8 GPUs): the bench as the driver's SCALE run launches it
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02m_bench8.json 2> gpurun_out/r02m_bench8.err ) 2>&1 | grep real; echo "bench8 rc=$?"; tail -5 gpurun_out/r02m_bench8.err | cut -c1-300
