#!/bin/bash
# full GPU test suite + the default bench line + smoke
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_gputests.log 2>&1
tail -8 gpurun_out/r2b_gputests.log
( time python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2> gpurun_out/r2b_bench_time.txt
cat gpurun_out/r2b_bench_time.txt
python tools/bench_summary.py gpurun_out/r2b_bench.json 2>/dev/null | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
