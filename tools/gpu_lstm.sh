#!/bin/bash
# LSTM bring-up: phase cycle counters, parity tests, per-layer times
out=gpurun_out; tag=${1:-lstm}
mkdir -p $out
: > $out/${tag}_dbg.log
for v in 0 1; do
B200OCR_LSTM_VAR=$v B200OCR_LSTM_DBG=1 timeout 300 python tools/prof_step.py lstm fp16f8 1 >> $out/${tag}_dbg.log 2>&1
B200OCR_LSTM_VAR=$v timeout 300 python bench.py --no-cpu-baseline --steps 10 --profile-out $out/${tag}_per_layer_v$v.json > $out/${tag}_bench_v$v.json 2> $out/${tag}_bench_v$v.err
done
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
echo done
