#!/bin/bash
# cfg3 (GE2E speaker encoder, the batch-sharded north-star config): launch list + ncu --set full of its four kernel families
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches_cfg3.csv \
    python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_launches_cfg3.log 2>&1
tools/prof_on_box.sh final_cfg3_bwd k_rnn_bwd_s 1 2 --config 3 --steps 1 --warmup 0
tools/prof_on_box.sh final_cfg3_fwd k_rnn_fwd_s 1 1 --config 3 --steps 1 --warmup 0
tools/prof_on_box.sh final_cfg3_red k_gemm_red 1 3 --config 3 --steps 1 --warmup 0
tools/prof_on_box.sh final_cfg3_rows k_gemm_rows 1 1 --config 3 --steps 1 --warmup 0
ls -la gpurun_out | grep final_cfg3
