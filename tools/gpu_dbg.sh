#!/bin/bash
# per-layer times with parts of the GEMM kernels disabled (results are wrong on purpose; timing only)
out=gpurun_out; tag=${1:-dbg}; modes=${2:-"0 1"}
mkdir -p $out
for d in $modes; do
B200OCR_IGEMM_DBG=$d timeout 300 python bench.py --no-cpu-baseline --steps 5 --profile-out $out/${tag}_per_layer_d$d.json > $out/${tag}_bench_d$d.json 2> $out/${tag}_bench_d$d.err
done
echo done
