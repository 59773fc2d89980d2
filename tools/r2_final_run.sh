#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err ) 2> gpurun_out/r2_bench_time.txt
tail -c 600 gpurun_out/r2_bench_final.json | head -c 300; echo
cat gpurun_out/r2_bench_time.txt
bash tools/r2_profile.sh > gpurun_out/r2_profile.log 2>&1
bash tools/r2_tc_ncu.sh > gpurun_out/r2_tc_ncu.txt 2>&1
tools/tc_gemm_test > gpurun_out/r2_tc_test_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out | tail -30
