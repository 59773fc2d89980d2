#!/bin/bash
# row groups at other batch sizes of the cfg3 stack (strong-scaling shards) and the alternative forced cuts
mkdir -p gpurun_out
run() {  # tag, row_groups value, extra bench args
  TTRNN_ROW_GROUPS=$2 timeout 200 python bench.py --no-cpu-baseline ${@:3} > gpurun_out/r2b_$1.json 2> gpurun_out/r2b_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2b_%s.json" % tag))
    for c in d["all_configs"]:
        h = c["roofline"]["plan"][0]
        print("%s id=%d B=%d ms=%.3f groups=%s rows=%s+%s" % (tag, c["id"], c["config"]["batch_per_gpu"], c["ms_per_step"],
              h.get("row_groups"), h.get("group_rows0"), h.get("group_rows1")))
except Exception as e:
    print(tag, "failed", e)
PY
}
run b320_off 0 --config 3 --batch 320
run b320_auto 1 --config 3 --batch 320
run b320_f196 196 --config 3 --batch 320
run b480_off 0 --config 3 --batch 480
run b480_auto 1 --config 3 --batch 480
run b640_auto 1 --configs 3,6
run b640_alt_f148 148 --config 6
