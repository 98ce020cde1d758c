#!/bin/bash
# run under gpurun: register-family parity tests, phase clocks and the bench line after the P3 / Sherman-Morrison changes
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullshape.py tests/test_gpu_variants.py -x -q 2>&1 | tail -6 > gpurun_out/pytest_p3.log
cat gpurun_out/pytest_p3.log
python tools/phase_clocks.py > gpurun_out/r2_phases_v5.json 2> gpurun_out/phases.err
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_p3.json 2> gpurun_out/bench_p3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_p3.json"))
print("value", round(d["value"]), "vmc", d["vmc"]["ms_per_step"], "lrdmc", d["lrdmc"]["ms_per_step"], "e2e", round(d["e2e"]["value"]))
print({k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items() if v["share"] > 0.01})
p = json.load(open("gpurun_out/r2_phases_v5.json"))
print(p["projection"])
PY
