#!/bin/bash
out=gpurun_out; tag=${1:-lstm}
mkdir -p $out
: > $out/${tag}_dbg.log
for v in 0 1 2 3; do
B200OCR_LSTM_VAR=$v B200OCR_LSTM_DBG=1 timeout 300 python tools/prof_step.py lstm fp16f8 1 >> $out/${tag}_dbg.log 2>&1
done
echo done
