python -m pytest tests/test_parity_r02.py -m gpu -q -x -k "with_offsets" 2>&1 | grep -E "^E  |passed|failed" | head -20
for g in 1000 1250 1700; do python bench.py --cpu-sample 0 --sweep-genes 0 --no-pageable --genes-per-batch $g > gpurun_out/gpb_c5_$g.json 2>&1; python tools/show_bench.py gpurun_out/gpb_c5_$g.json; done
for g in 0 2500 1250; do python bench.py --config c4 --cpu-sample 0 --sweep-genes 0 --no-pageable --genes-per-batch $g > gpurun_out/gpb_c4_$g.json 2>&1; python tools/show_bench.py gpurun_out/gpb_c4_$g.json; done
