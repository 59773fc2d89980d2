#!/bin/bash
# Final-build evidence (run under gpurun on one B200): default bench line, launch lists, ncu --set full captures of the
# dominant kernels.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2b_bench_final.json 2> gpurun_out/r2b_bench_final.err ) 2> gpurun_out/r2b_bench_final_time.txt
cat gpurun_out/r2b_bench_final_time.txt
python tools/bench_summary.py gpurun_out/r2b_bench_final.json
B="python bench.py --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_cfg3.csv $B --config 3 --steps 2 --warmup 1 > gpurun_out/r2b_launches_cfg3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg2.csv $B --config 2 --steps 2 --warmup 1 > gpurun_out/r2b_launches_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rnn_bwd_s -s 6 -c 1 -f -o gpurun_out/r2b_prof_cfg3_bwd $B --config 3 --steps 1 --warmup 1 > gpurun_out/r2b_prof_cfg3_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rnn_fwd_s -s 6 -c 1 -f -o gpurun_out/r2b_prof_cfg3_fwd $B --config 3 --steps 1 --warmup 1 > gpurun_out/r2b_prof_cfg3_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rnn_bwd_s -s 1 -c 1 -f -o gpurun_out/r2b_prof_cfg2_bwd $B --config 2 --steps 1 --warmup 1 > gpurun_out/r2b_prof_cfg2_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/r2b_prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  python tools/ncu_summary.py $f > $b.summary.txt 2>&1
  ncu -i $f --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $b.cudasass.csv.gz
  rm -f $f
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la gpurun_out/ | grep r2b_ | tail -30
