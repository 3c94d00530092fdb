python tools/debug_exact.py 4000 32 43 off 2>&1 | tail -12
for c in 1 2 4 8; do echo "== cluster $c"; SCLANE_FINAL_CLUSTER=$c python bench.py --config c4 --cpu-sample 0 --sweep-genes 0 --no-pageable > gpurun_out/c4_cl$c.json 2> gpurun_out/c4_cl$c.err; tail -3 gpurun_out/c4_cl$c.err; python tools/show_bench.py gpurun_out/c4_cl$c.json 2>&1 | tail -12; done
python -m pytest tests -m gpu -q -x -k "offset or lineages or compressed_path or full_size" 2>&1 | tail -8
