#!/bin/bash
# run under gpurun: branching parity after the cumsum change, GFMC_t step time of the final build, bench line
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -x -q -k "branch or gfmc or lrdmc" 2>&1 | tail -4
python tools/time_gfmc_t.py --out gpurun_out/r2_gfmc_t.json > /dev/null 2> gpurun_out/gfmc_t.err
python -c "
import json; d=json.load(open('gpurun_out/r2_gfmc_t.json')); print('gfmc_t', d['ms_per_step'], d['walker_steps_per_s'], d['projections_mean'])"
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tail.json')); print(round(d['value']), d['vmc']['ms_per_step'], d['lrdmc']['ms_per_step'], {k: round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items() if v['share']>0.004})"
