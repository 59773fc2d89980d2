#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rank or pad or row_groups or static or alt" > gpurun_out/r2b_d2split_tests.log 2>&1
tail -3 gpurun_out/r2b_d2split_tests.log
for sk in 0 1; do
  TTRNN_SPLIT_KEPT=$sk timeout 300 python bench.py --config 6 --no-cpu-baseline > gpurun_out/r2b_d2split_$sk.json 2> gpurun_out/r2b_d2split_$sk.err
  python - $sk <<'PY'
import json, sys
sk = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2b_d2split_%s.json" % sk))
    c = d["all_configs"][0]; r = c["roofline"]
    print("split_kept=%s ms=%.3f" % (sk, c["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in r["kernels"].items()}, r["plan"][1].get("bwd_kernel"), r["plan"][1].get("bwd_rows"), r["plan"][0])
except Exception as e:
    print(sk, "failed", e)
PY
done
