#!/bin/bash
# ncu --set full on the tcgen05 GEMMs of tools/tc_gemm_test (timing shapes); prints the metrics that say what binds them
for k in k_tc_red k_tc_rows; do
  ncu --set full --clock-control none -k regex:$k -s 30 -c 1 -f -o /tmp/tc_$k tools/tc_gemm_test > /dev/null 2>&1
  echo "=== $k"
  ncu -i /tmp/tc_$k.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; r=rows[2]
for i,k in enumerate(h):
    if any(s in k for s in ('gpu__time_duration.sum','tensor','sm__cycles_active.avg','sm__throughput.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','sm__inst_executed_pipe_alu.avg.pct','sm__inst_executed_pipe_lsu.avg.pct','smsp__issue_active.avg.pct','smsp__cycles_active.avg','sm__cycles_elapsed.avg.per_second','lts__t_sectors.avg.pct','dram__throughput.avg.pct','tmem','smsp__average_warp')):
        print('  %-90s %s %s' % (k, r[i], u[i]))
"
done
