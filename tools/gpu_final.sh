#!/bin/bash
# last check of a round: smoke, all GPU parity tests, the default bench command
out=gpurun_out; tag=${1:-final}
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo done
