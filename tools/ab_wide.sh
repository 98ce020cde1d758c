#!/bin/bash
# A/B of library builds on the general-family configuration (water JAGP LRDMC) + general-family parity tests with the last tag
TAGS=${1:-"Z W"}
cp jqmc_b200/lib/libjqmc_b200.so /tmp/lib_keep.so
for rep in 1 2; do
  for t in $TAGS; do
    cp jqmc_b200/lib/ab/lib_$t.so jqmc_b200/lib/libjqmc_b200.so
    python bench.py --config water_jagp --steps 10 --warmup 3 --no-cpu > gpurun_out/abw_${t}_${rep}.json 2> gpurun_out/abw_${t}_${rep}.err
    python - "$t" "$rep" <<'PY'
import json, sys
t, rep = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/abw_{t}_{rep}.json"))
    k = d["roofline"]["kernels"]
    print(t, rep, "value", round(d["value"]), round(d["ms_per_step"], 3), {n: round(v["ms_per_launch"], 4) for n, v in k.items() if v["share"] > 0.02})
except Exception as e:
    print(t, rep, "failed", e)
PY
  done
done
cp /tmp/lib_keep.so jqmc_b200/lib/libjqmc_b200.so
