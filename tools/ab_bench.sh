#!/bin/bash
# A/B of two builds of the library on the SAME GPU box (run under gpurun): jqmc_b200/lib/ab/lib_<tag>.so for every tag given,
# interleaved (A B A B) so that box-to-box and drift effects cancel.  usage: tools/ab_bench.sh "A B" [bench.py args]
TAGS=${1:-"A B"}; shift
ARGS=${@:-"--steps 60 --warmup 5 --no-cpu"}
cp jqmc_b200/lib/libjqmc_b200.so /tmp/lib_keep.so
for rep in 1 2; do
  for t in $TAGS; do
    cp jqmc_b200/lib/ab/lib_$t.so jqmc_b200/lib/libjqmc_b200.so
    python bench.py $ARGS > gpurun_out/ab_${t}_${rep}.json 2> gpurun_out/ab_${t}_${rep}.err
    python - "$t" "$rep" <<'PY'
import json, sys
t, rep = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/ab_{t}_{rep}.json"))
    k = d["roofline"]["kernels"]
    print(t, rep, "value", round(d["value"]), {n: round(d[n]["ms_per_step"], 4) for n in ("vmc", "lrdmc") if n in d},
          {n: round(v["ms_per_launch"], 4) for n, v in k.items() if v["share"] > 0.02})
except Exception as e:
    print(t, rep, "failed", e)
PY
  done
done
cp /tmp/lib_keep.so jqmc_b200/lib/libjqmc_b200.so
