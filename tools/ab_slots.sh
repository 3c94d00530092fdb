#!/bin/bash
# per-GPU work of the 8-GPU run (2,500 genes x 200,000 cells) on one GPU: wall ms per step for slot / batch-size choices
for mb in 512 1024; do for bs in 1 2 3 4; do
  SCLANE_MIN_BATCH=$mb python bench.py --genes 2500 --steps 5 --cpu-sample 0 --sweep-genes 0 --no-pageable --batch-streams $bs > gpurun_out/ab_${mb}_${bs}.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_${mb}_${bs}.json').read().strip().splitlines()[-1])
print('min_batch $mb streams $bs: wall %.2f ms  seq wall %.2f  events(seq) %.2f  e2e %.2f ms' % (d['ms_per_step'], d['sequential_wall_ms_per_step'], d['device_event_ms_per_step'], d['e2e']['ms_per_step']))
PY
done; done
