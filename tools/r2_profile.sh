#!/bin/bash
# Round-2 ncu captures (run under gpurun on one B200).  Numbers printed by runs under ncu are never bench values.
set -x
B="python bench.py --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_cfg3.csv $B --config 3 --steps 2 --warmup 1 > gpurun_out/r2_launches_cfg3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv $B --config 2 --steps 2 --warmup 1 > gpurun_out/r2_launches_cfg2.log 2>&1
for k in k_tc_red k_tc_rows; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/r2_prof_cfg3_$k $B --config 3 --steps 1 --warmup 1 > gpurun_out/r2_prof_cfg3_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_rnn_bwd_s -s 3 -c 1 -f -o gpurun_out/r2_prof_cfg3_bwd $B --config 3 --steps 1 --warmup 1 > gpurun_out/r2_prof_cfg3_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rnn_fwd_s -s 3 -c 1 -f -o gpurun_out/r2_prof_cfg3_fwd $B --config 3 --steps 1 --warmup 1 > gpurun_out/r2_prof_cfg3_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rnn_bwd_s -s 1 -c 1 -f -o gpurun_out/r2_prof_cfg2_bwd $B --config 2 --steps 1 --warmup 1 > gpurun_out/r2_prof_cfg2_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
# export what is read on the CPU box, drop the big reports (gpurun_out is capped at 64 MiB)
for f in gpurun_out/r2_prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  python tools/ncu_summary.py $f > $b.summary.txt 2>&1
  ncu -i $f --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $b.cudasass.csv.gz
  rm -f $f
done
ls -la gpurun_out/ | tail -20
