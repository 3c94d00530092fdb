#!/bin/bash
# times bench.py for every prebuilt library variant in sclane_b200/lib/variant_*.so:  tools/variant_bench.sh "c5 c2"
cd "$(dirname "$0")/.."
CFGS=${1:-"c5 c2"}
cp sclane_b200/lib/libsclane_b200.so /tmp/lib_backup.so
for f in sclane_b200/lib/variant_*.so; do
  cp "$f" sclane_b200/lib/libsclane_b200.so
  for c in $CFGS; do
    python bench.py --config $c --steps 3 --warmup 2 --cpu-sample 0 --sweep-genes 0 --no-pageable 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$f $c', 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'seq', round(d['sequential_wall_ms_per_step'],2), 'stages', {k: round(v, 2) for k, v in d['stage_ms_per_step'].items() if v > 0.3})
"
  done
done
cp /tmp/lib_backup.so sclane_b200/lib/libsclane_b200.so
