#!/bin/bash
# bwd_overlap on/off at the cfg3 shape for several batch sizes (strong-scaling shards of the 640-utterance batch)
for b in 80 160 320 640; do
  for o in 1 0; do
    TTRNN_BWD_OVERLAP=$o python bench.py --config 3 --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=$b overlap=$o ms/step %.3f' % d['ms_per_step'])"
  done
done
