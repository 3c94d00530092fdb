#!/bin/bash
# Round-2 measurement pass (run on the GPU box from the repo root): bench lines of every BASELINE config, the exact-knot mode,
# the reference arm, then (separately, never a bench value) the ncu launch list and --set full captures of the fit kernels.
set -u
O=gpurun_out
mkdir -p $O
WHAT=${1:-all}
if [ "$WHAT" != ncu ]; then
python bench.py                                   > $O/r02_bench_c5.json      2> $O/r02_bench_c5.err
python bench.py --config c2                       > $O/r02_bench_c2.json      2> $O/r02_bench_c2.err
python bench.py --config c4 --cpu-sample 0        > $O/r02_bench_c4.json      2> $O/r02_bench_c4.err
python bench.py --config c3 --cpu-sample 0        > $O/r02_bench_c3.json      2> $O/r02_bench_c3.err
python bench.py --config c2 --approx-knot 0 --cpu-sample 0 --sweep-genes 0 --no-pageable > $O/r02_bench_c2_exact_knots.json 2> $O/r02_bench_c2_exact_knots.err
python bench.py --config c4 --offset-per-cell --cpu-sample 0 --sweep-genes 0 --no-pageable > $O/r02_bench_c4_offset_per_cell.json 2> $O/r02_bench_c4_offset_per_cell.err
python bench.py --impl reference                  > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
tail -c 600 $O/r02_bench_*.err | tail -30
fi
if [ "$WHAT" = bench ]; then exit 0; fi
P="--steps 1 --warmup 1 --cpu-sample 0 --sweep-genes 0 --no-pageable"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 3000 --csv --log-file $O/r02_launches_c5.csv python bench.py $P > $O/ncu_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 3000 --csv --log-file $O/r02_launches_c4.csv python bench.py --config c4 $P > $O/ncu_c4.log 2>&1
ncu --set full --clock-control none -k regex:'k_comp_stage|k_stage|k_wg_stage|k_off_aggregate' -c 10 -o $O/r02_fit_c4 -f python bench.py --config c4 --genes 4000 $P > $O/ncu_full_c4.log 2>&1
ncu --set full --clock-control none -k regex:'k_comp_stage|k_aggregate' -c 6 -o $O/r02_fit_c5 -f python bench.py --genes 4000 $P > $O/ncu_full_c5.log 2>&1
ncu --set full --clock-control none -k regex:'k_exact' -c 2 -o $O/r02_fit_exact -f python bench.py --config c2 --approx-knot 0 $P > $O/ncu_full_exact.log 2>&1
for r in r02_fit_c4 r02_fit_c5 r02_fit_exact; do python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.txt 2>&1; rm -f $O/$r.ncu-rep; done
du -sh $O; ls -la $O | tail -30
