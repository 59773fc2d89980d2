#!/bin/bash
# compute-sanitizer memcheck over the paths added in the second half of round 2
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9"
( $S python -m pytest tests/test_gpu_row_groups.py -q -x -k "forced or initial or inference or options" ; echo "exit $?" ) > gpurun_out/r2b_san_groups.txt 2>&1
( $S python -m pytest tests/test_gpu_variants.py -q -x -k "loggrads" ; echo "exit $?" ) > gpurun_out/r2b_san_logging.txt 2>&1
( $S python tools/r1split_probe.py 20 16 ; echo "exit $?" ) > gpurun_out/r2b_san_r1split.txt 2>&1
for f in gpurun_out/r2b_san_*.txt; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|exit|ok " $f | tail -6; done
