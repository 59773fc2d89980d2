#!/bin/bash
# row groups: parity tests, then A/B of the multi-layer configs with the split off / on (1 GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_row_groups.py -x -q > gpurun_out/r2b_groups_test.log 2>&1
tail -12 gpurun_out/r2b_groups_test.log
for rg in 0 1; do
  TTRNN_ROW_GROUPS=$rg timeout 300 python bench.py --configs 3,6,9 --no-cpu-baseline > gpurun_out/r2b_ab_rg$rg.json 2> gpurun_out/r2b_ab_rg$rg.err
done
python - <<'PY'
import json
for rg in (0, 1):
    try:
        d = json.load(open("gpurun_out/r2b_ab_rg%d.json" % rg))
    except Exception as e:
        print("rg", rg, "failed", e); continue
    for c in d["all_configs"]:
        r = c["roofline"]
        print("rg=%d id=%d ms=%.3f (timing pass %.3f) e2e=%.3g plan=%s fwd=%.3f bwd=%.3f frac=%.3f" % (
            rg, c["id"], c["ms_per_step"], c["ms_per_step_kernel_timing_pass"], c["e2e"]["value"], r["plan"][0],
            r["kernels"]["k_rnn_fwd"]["ms_per_step"], r["kernels"]["k_rnn_bwd"]["ms_per_step"], r["frac"]))
PY
