#!/bin/bash
# per-layer times with parts of the halo kernel disabled (results are wrong on purpose; timing only)
out=gpurun_out; tag=${1:-dbg}
mkdir -p $out
for d in 0 1 2 3; do
B200OCR_IGEMM_DBG=$d timeout 300 python bench.py --no-cpu-baseline --steps 5 --profile-out $out/${tag}_per_layer_d$d.json > $out/${tag}_bench_d$d.json 2> $out/${tag}_bench_d$d.err
done
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
echo done
