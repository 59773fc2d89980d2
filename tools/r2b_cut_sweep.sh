#!/bin/bash
# forced row-group cuts of the cfg3 stack around the planner's choice (444 + 196): does the list-schedule model rank them right?
mkdir -p gpurun_out
for r0 in 1 296 392 444 500 520 592; do
  TTRNN_ROW_GROUPS=$r0 timeout 200 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2b_cut_$r0.json 2> gpurun_out/r2b_cut_$r0.err
  python - $r0 <<'PY'
import json, sys
r0 = sys.argv[1]
try:
    c = json.load(open("gpurun_out/r2b_cut_%s.json" % r0))["all_configs"][0]
    h = c["roofline"]["plan"][0]
    k = c["roofline"]["kernels"]
    print("row_groups=%s cut=%s+%s ms=%.3f  (serial pass %.3f; fwd %.2f bwd %.2f)" % (r0, h.get("group_rows0"), h.get("group_rows1"), c["ms_per_step"],
          c["ms_per_step_kernel_timing_pass"], k["k_rnn_fwd"]["ms_per_step"], k["k_rnn_bwd"]["ms_per_step"]))
except Exception as e:
    print(r0, "failed", e)
PY
done
