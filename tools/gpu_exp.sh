#!/bin/bash
# quick experiment pass: engine parity tests + per-layer times (+ optional ncu of one kernel family)
out=gpurun_out; tag=${1:-exp}; kern=${2:-}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err
if [ -n "$kern" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kern" -c 2 -o $out/${tag}_ncu -f python tools/prof_step.py lstm fp16f8 1 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_ncu.ncu-rep --page raw --csv > $out/${tag}_ncu_raw.csv 2>/dev/null
ncu -i $out/${tag}_ncu.ncu-rep --page details > $out/${tag}_ncu_details.txt 2>/dev/null
fi
echo done
