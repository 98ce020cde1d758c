#!/bin/bash
# run under gpurun: bench lines, full GPU suite and smoke of the deterministic (no --split-compile) build
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_det.json 2> gpurun_out/bench_det.err
python bench.py --config water_jagp --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_det_jagp.json 2> gpurun_out/bench_det_jagp.err
python - <<'PY'
import json
for f in ("bench_det", "bench_det_jagp"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items() if v["share"] > 0.02})
    except Exception as e:
        print(f, "failed", e)
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_det.log
cat gpurun_out/pytest_det.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_det.log 2>&1; tail -2 gpurun_out/smoke_det.log
