#!/bin/bash
mkdir -p gpurun_out
for ov in 1 2; do
  TTRNN_BWD_OVERLAP=$ov timeout 300 python bench.py --configs 3,9 --no-cpu-baseline > gpurun_out/r2b_ov$ov.json 2> gpurun_out/r2b_ov$ov.err
done
python - <<'PY'
import json
for ov in (1, 2):
    try:
        d = json.load(open("gpurun_out/r2b_ov%d.json" % ov))
    except Exception as e:
        print("ov", ov, "failed", e); continue
    for c in d["all_configs"]:
        r = c["roofline"]
        print("ov=%d id=%d ms=%.3f (timing pass %.3f) e2e=%.3g" % (ov, c["id"], c["ms_per_step"], c["ms_per_step_kernel_timing_pass"], c["e2e"]["value"]))
PY
