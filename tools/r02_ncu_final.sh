#!/bin/bash
# ncu evidence for the final round-2 build (run on the GPU box; never a bench value): per-launch times of one bench step per
# config (the kernel SHARES must agree with stage_ms_per_step), then --set full of the fit / plan / aggregation kernels.
set -u
O=gpurun_out
mkdir -p $O
P="--steps 1 --warmup 1 --cpu-sample 0 --sweep-genes 0 --no-pageable --batch-streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k|^void k' -c 4000 --csv --log-file $O/r02_launches_c5.csv python bench.py $P > $O/ncu_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k|^void k' -c 4000 --csv --log-file $O/r02_launches_c4.csv python bench.py --config c4 $P > $O/ncu_c4.log 2>&1
python - <<'PY'
import csv, collections, sys
for cfg in ("c5", "c4"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r02_launches_{cfg}.csv", errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1; a[1] += v
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "?"
    tot = sum(v for _, v in agg.values())
    with open(f"gpurun_out/r02_ncu_launches_{cfg}.txt", "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, one bench.py process ({cfg}: 1 warm-up + 1 timed + 3 sequential-profile steps + e2e steps), unit {unit}\n")
        fh.write(f"# per-launch times under ncu are cold-cache and serialised: the SHARES are what is compared with stage_ms_per_step\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k[:70]:72s} launches {n:5d}  total {v/1e6:10.3f} ms  share {100*v/tot:5.1f} %\n")
PY
F="--steps 1 --warmup 1 --cpu-sample 0 --sweep-genes 0 --no-pageable --batch-streams 1"
ncu --set full --clock-control none --import-source on -k regex:'k_comp_stage|k_aggregate|k_gene_bucket_nodes|k_gather_copy' -c 7 -o $O/r02_final_c5 -f python bench.py --genes 4000 $F > $O/ncu_full_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_stage|k_off_aggregate|k_wg_stage' -c 4 -o $O/r02_final_c4 -f python bench.py --config c4 --genes 3000 $F > $O/ncu_full_c4.log 2>&1
ncu --set full --clock-control none -k regex:'k0_' -c 16 -o $O/r02_final_k0 -f python bench.py --genes 1000 $F > $O/ncu_full_k0.log 2>&1
for r in r02_final_c5 r02_final_c4 r02_final_k0; do python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.txt 2>&1; done
rm -f $O/r02_final_k0.ncu-rep $O/r02_final_c4.ncu-rep
ls -la $O | tail -20
