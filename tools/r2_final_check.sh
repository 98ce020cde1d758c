#!/bin/bash
# run under gpurun: A/B of the final library against the earlier round-2 build, full GPU suite, smoke, bench lines
bash tools/ab_bench.sh "A Z" 2>&1 | tee gpurun_out/ab_final.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_final.log
cat gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_water_jsd.json 2> gpurun_out/bench_water_jsd.err
python bench.py --precision mixed --no-cpu > gpurun_out/bench_water_jsd_mixed.json 2> gpurun_out/bench_mixed.err
python bench.py --config water_jagp --no-cpu > gpurun_out/bench_water_jagp.json 2> gpurun_out/bench_jagp.err
python - <<'PY'
import json
for f in ("bench_water_jsd", "bench_water_jsd_mixed", "bench_water_jagp"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"].get("frac"), d["roofline"].get("frac_executed"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
