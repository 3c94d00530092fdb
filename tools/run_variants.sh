#!/bin/bash
# runs "$@" once per prebuilt library variant (sclane_b200/lib/variant_*.so) and once with the default build
cd "$(dirname "$0")/.."
cp sclane_b200/lib/libsclane_b200.so /tmp/lib_backup.so
echo "== default"; "$@"
for f in sclane_b200/lib/variant_*.so; do
  cp "$f" sclane_b200/lib/libsclane_b200.so
  echo "== $f"; "$@"
done
cp /tmp/lib_backup.so sclane_b200/lib/libsclane_b200.so
