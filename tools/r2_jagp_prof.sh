#!/bin/bash
# run under gpurun: GPU force tests, then ncu --set full of the general-family kernels on the water JAGP LRDMC step
python -m pytest tests/test_gpu_forces.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_forces.log
cat gpurun_out/pytest_forces.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kw_move|kw_electron|kw_mesh|kw_bmm|kw_dgemm|kw_ao_store|kw_lrdmc_select' \
  --launch-skip 200 --launch-count 14 -o /tmp/jagp python bench.py --config water_jagp --steps 1 --warmup 1 --no-cpu > gpurun_out/jagp_ncu.log 2>&1
python tools/ncu_summary.py /tmp/jagp.ncu-rep > gpurun_out/r2_jagp_wide.md 2>&1
python tools/ncu_stalls.py /tmp/jagp.ncu-rep > gpurun_out/r2_jagp_stalls.txt 2>&1
tail -3 gpurun_out/jagp_ncu.log
