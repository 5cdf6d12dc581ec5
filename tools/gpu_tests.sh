#!/bin/bash
# all GPU parity tests + a short bench (+ the auxiliary measurements when asked)
out=gpurun_out; tag=${1:-t}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err
if [ -n "$2" ]; then timeout 600 python -m tests.aux_bench > $out/${tag}_aux_bench.json 2> $out/${tag}_aux_bench.err; fi
echo done
