#!/bin/bash
cd "$(dirname "$0")/.."
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include"
for cfg in "1 1 1 0 3 8 1" "1 1 1 0 3 8 0" "1 1 1 0 4 8 1" "1 1 1 0 8 4 1" "1 1 1 0 8 4 0" "1 1 1 0 6 4 1" "1 1 1 1 6 4 1" "0 1 1 0 3 8 1"; do
  set -- $cfg
  nvcc $F -DSCLANE_SWEEP_DO_FLUSH=$1 -DSCLANE_SWEEP_DO_BOUNDARY=$2 -DSCLANE_SWEEP_DO_EPILOGUE=$3 -DSCLANE_SWEEP_PREFETCH=$4 -DSCLANE_SWEEP_MINB=$5 -DSCLANE_SWEEP_WARPS=$6 -DSCLANE_SWEEP_CELL2=$7 tools/sweep_ablation.cu -o /tmp/abl 2>/dev/null && /tmp/abl ${1000:-2000} 10000
done
