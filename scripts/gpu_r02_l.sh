# round 2, call L: ncu evidence -- launch list of the bench command, ncu --set full of the main-stage pruned scan
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --queries 4736 --no-cpu-baseline --no-recall --no-extra-legs --opt pruned_lb_quantizers=15"
timeout -s KILL 300 $CMD > gpurun_out/r02l_plain.json 2> gpurun_out/r02l_plain.err; echo "plain rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_q4736.csv $CMD > gpurun_out/r02l_ncu1.log 2>&1; echo "launch list rc=$?"
# main-stage launch = the 2nd pruned_scan launch of a batch (after the short first stage); skip the warm-up step's launches
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"pruned_scan_kernel" -s 5 -c 1 -f -o gpurun_out/r02_pscan_main $CMD > gpurun_out/r02l_ncu2.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/r02l_ncu2.log
