#!/bin/bash
out=gpurun_out; tag=${1:-beam}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_decoders.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
for lib in ""; do
B200OCR_LIB=$lib timeout 300 python - <<'PY' >> $out/${tag}_time.log 2>&1
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
print(os.environ.get('B200OCR_LIB') or 'in-tree', json.dumps(bench.ctc_decode_times(torch.device('cuda', 0), False)))
PY
done
echo done
