#!/bin/bash
# Transformer-encoder variant (BASELINE config 3 forward): parity tests + per-layer times
out=gpurun_out; tag=${1:-tr}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_decoders.py -m gpu -x -q -k "transformer or recognise" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py --net transformer --no-cpu-baseline --steps 5 --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo done
