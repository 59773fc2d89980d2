#!/bin/bash
# usage: tools/prof_on_box.sh <tag> <kernel-regex> <count> <skip> <bench args...>
# Runs one `ncu --set full` capture on the GPU box, writes text summaries under gpurun_out/ and drops the
# (large) .ncu-rep so the result fits gpurun's 64 MiB return limit.
tag=$1; regex=$2; count=$3; skip=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o /tmp/prof_$tag \
    python bench.py "$@" --no-cpu-baseline > gpurun_out/prof_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep > gpurun_out/prof_$tag.summary.txt 2>&1
# per-SASS-instruction metrics (stall samples, executed counts) with the source line of each instruction
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_$tag.sass.csv.gz
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/prof_$tag.cudasass.csv.gz
ls -la /tmp/prof_$tag.ncu-rep gpurun_out/ >> gpurun_out/prof_$tag.log
