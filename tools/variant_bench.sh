#!/bin/bash
# times bench.py's sweep stage for every prebuilt library variant in sclane_b200/lib/variant_*.so
cd "$(dirname "$0")/.."
cp sclane_b200/lib/libsclane_b200.so /tmp/lib_backup.so
for f in sclane_b200/lib/variant_*.so; do
  cp "$f" sclane_b200/lib/libsclane_b200.so
  echo "== $f"
  timeout 300 python bench.py --steps 3 --warmup 2 --cpu-sample 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value']), 'sweep ms/launch', d['roofline']['ms_per_launch'], 'frac', round(d['roofline']['frac'], 3), 'select', d['roofline']['select_kernel_ms_per_launch'], 'stages', {k: round(v, 2) for k, v in d['stage_ms_per_step'].items()})
"
done
cp /tmp/lib_backup.so sclane_b200/lib/libsclane_b200.so
