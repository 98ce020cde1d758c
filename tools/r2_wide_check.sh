#!/bin/bash
# run under gpurun: general-family parity tests + force tests + water_jagp / synthetic bench lines
python -m pytest tests/test_gpu_wide.py tests/test_gpu_forces.py -x -q 2>&1 | tail -12 > gpurun_out/pytest_wide.log
cat gpurun_out/pytest_wide.log
python bench.py --config water_jagp --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_jagp.json 2> gpurun_out/bench_jagp.err
python bench.py --config benzene_sr --no-cpu > gpurun_out/bench_benzene.json 2> gpurun_out/bench_benzene.err
python - <<'PY'
import json
for f in ("bench_jagp", "bench_benzene"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), d["ms_per_step"], {k: d[k] for k in ("vmc", "lrdmc", "sr") if k in d})
        print({k: (round(v["ms_per_launch"], 4), v["launches_per_step"]) for k, v in d["roofline"]["kernels"].items() if v["share"] > 0.02})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_jagp.err gpurun_out/bench_benzene.err
