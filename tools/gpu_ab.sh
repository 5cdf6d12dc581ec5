#!/bin/bash
# A/B two builds of the library in one session (box-to-box clock variance is ~10 %): ab_libA.so vs the in-tree build
out=gpurun_out; tag=${1:-ab}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parsenet.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
for rep in 1 2; do
B200OCR_LIB=$PWD/ab_libA.so timeout 300 python bench.py --no-cpu-baseline --steps 8 --profile-out $out/${tag}_A${rep}.json > $out/${tag}_benchA${rep}.json 2> $out/${tag}_A.err
timeout 300 python bench.py --no-cpu-baseline --steps 8 --profile-out $out/${tag}_B${rep}.json > $out/${tag}_benchB${rep}.json 2> $out/${tag}_B.err
done
echo done
