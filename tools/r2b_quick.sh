#!/bin/bash
# full GPU suite + the headline / d2 configs (quick A/B point for kernel changes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_quick_tests.log 2>&1
tail -2 gpurun_out/r2b_quick_tests.log
timeout 300 python bench.py --configs 1,2,3,6 --no-cpu-baseline > gpurun_out/r2b_quick.json 2> gpurun_out/r2b_quick.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_quick.json"))
    for c in d["all_configs"]:
        r = c["roofline"]
        print("id=%d ms=%.3f dom %s frac %.3f |" % (c["id"], c["ms_per_step"], r["kernel"], r["frac"]),
              " ".join("%s/R%d/%d:%.3f(%.3f)" % (v["kernel"][2:], v["rows_per_cta"], v["ctas"], v["ms_per_launch"], v["frac_of_ffma_peak"]) for v in r["recurrent_variants"]))
except Exception as e:
    print("failed", e)
PY
