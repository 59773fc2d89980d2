#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_r1split_tests.log 2>&1
tail -4 gpurun_out/r2b_r1split_tests.log
for sk in 0 1; do
  TTRNN_SPLIT_KEPT=$sk timeout 300 python bench.py --configs 1,2 --headline 1 --no-cpu-baseline > gpurun_out/r2b_r1split_$sk.json 2> gpurun_out/r2b_r1split_$sk.err
  python - $sk <<'PY'
import json, sys
sk = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2b_r1split_%s.json" % sk))
    for c in d["all_configs"]:
        r = c["roofline"]
        print("split_kept=%s id=%d ms=%.3f" % (sk, c["id"], c["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in r["kernels"].items() if v["ms_per_step"] > 0}, r["plan"][1].get("bwd_kernel"), r["plan"][1].get("bwd_rows"), r["plan"][1].get("bwd_rows2"), "frac %.3f" % r["frac"])
except Exception as e:
    print(sk, "failed", e)
PY
done
