#!/bin/bash
# round-2 GPU pass b: parity tests, precision sweep, bench line (with incumbent, configs 3/4)
out=gpurun_out; tag=${1:-r02b}
mkdir -p $out
python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -15 $out/${tag}_pytest_gpu.log
python tools/gpu_precision_sweep.py > $out/${tag}_precision_sweep.json 2> $out/${tag}_precision_sweep.err; echo "sweep rc=$?"
tail -3 $out/${tag}_precision_sweep.err
python bench.py --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -5 $out/${tag}_bench.err
head -c 6000 $out/${tag}_bench.json
