#!/bin/bash
# run under gpurun: ncu --set full of the general-family LRDMC kernels inside the water JAGP step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kw_move|kw_electron|kw_bmm|kw_lrdmc_select|kw_mesh' \
  --launch-skip 400 --launch-count 12 -o /tmp/jagp2 python bench.py --config water_jagp --steps 1 --warmup 1 --no-cpu > gpurun_out/jagp_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/jagp2.ncu-rep > gpurun_out/r2_jagp_wide2.md 2>&1
grep -E "^###|duration|grid|block:|regs|achieved occ|dram read|dram write|issue slots|top stalls|L2 bytes" gpurun_out/r2_jagp_wide2.md | cut -c1-170
