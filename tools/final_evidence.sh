#!/bin/bash
# Round-end evidence on the GPU box (one B200): launch list of the default bench command, ncu --set full of the
# two recurrent kernels of the default workload, bench lines of all five BASELINE.json configs, the reference arm.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_launches_cfg2.log 2>&1
tools/prof_on_box.sh final_cfg2_bwd k_rnn_bwd_s 1 0 --config 2 --steps 1 --warmup 0
tools/prof_on_box.sh final_cfg2_fwd k_rnn_fwd_s 1 0 --config 2 --steps 1 --warmup 0
python bench.py > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
for c in 1 3 4 5; do
    python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/final_cfg$c.json 2> gpurun_out/final_cfg$c.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference_cfg2.json 2> gpurun_out/final_reference_cfg2.err
python tools/bench_summary.py gpurun_out/final_cfg?.json
