#!/bin/bash
# quick GPU check: parity tests + bench summary (run under gpurun)
python -m pytest tests -m gpu -x -q ${1:-} 2>&1 | tail -15 > gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/pytest.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", round(d["value"]), "vmc", round(d["vmc"]["value"]), d["vmc"]["ms_per_step"], "lrdmc", round(d["lrdmc"]["value"]), d["lrdmc"]["ms_per_step"], "e2e", round(d["e2e"]["value"]))
    print({k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
    print({k: round(v["frac"], 3) for k, v in d["roofline"]["per_kernel"].items()})
    print(d["check"])
except Exception as e:
    print("bench failed", e)
PY
tail -5 gpurun_out/bench.err
