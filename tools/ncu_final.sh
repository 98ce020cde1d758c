#!/bin/bash
# run under gpurun: ncu evidence of the final round-2 build (launch list of the bench command + full captures of the
# projection and Metropolis kernels), summarised on the box (the .ncu-rep files are too large to travel back)
B="python bench.py --steps 2 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final2.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walker --launch-skip 6 --launch-count 3 -o /tmp/walker $B > gpurun_out/ncu_w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mcmc2 --launch-skip 3 --launch-count 1 -o /tmp/mcmc $B > gpurun_out/ncu_m.log 2>&1
python tools/ncu_summary.py /tmp/walker.ncu-rep > gpurun_out/r2_full_walker_final2.md 2>&1
python tools/ncu_summary.py /tmp/mcmc.ncu-rep > gpurun_out/r2_full_mcmc_final2.md 2>&1
ncu -i /tmp/walker.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/r2_traffic_walker2.csv 2>&1
python tools/ncu_stalls.py /tmp/walker.ncu-rep > gpurun_out/r2_stalls_walker_final2.txt 2>&1
python tools/ncu_codesize.py /tmp/walker.ncu-rep > gpurun_out/r2_codesize_walker_final2.txt 2>&1
grep -E "^###|duration|fp64 % \(act|issue slots|local loads|top stalls|bank conflicts|dram read|dram write" gpurun_out/r2_full_walker_final2.md gpurun_out/r2_full_mcmc_final2.md
