# round 2, call G: why did the prefilter not speed encode up?  c3 noise check + ncu --set full of the two tc_assign variants
mkdir -p gpurun_out
for i in 1 2; do timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 1; done 2>&1 | tee gpurun_out/r02g_c3.txt
timeout -s KILL 300 python scripts/bench_encode.py 1000000 300 30 1 > gpurun_out/r02g_enc.txt 2>&1; tail -2 gpurun_out/r02g_enc.txt
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_assign_kernel.*unsigned" -s 1 -c 1 -f -o gpurun_out/r02g_tc_encode python scripts/bench_encode.py 1000000 300 30 1 > gpurun_out/r02g_ncu1.log 2>&1; tail -2 gpurun_out/r02g_ncu1.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_assign_kernel" -s 2 -c 1 -f -o gpurun_out/r02g_tc_fused python scripts/bench_train.py 1000000 300 30 3 1 1 > gpurun_out/r02g_ncu2.log 2>&1; tail -2 gpurun_out/r02g_ncu2.log
