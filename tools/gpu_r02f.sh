#!/bin/bash
# round-2 GPU pass f: parity tests (new conv_first staging, tile scheduler, decoders fix, pages pipeline), bench with
# 1 and 2 replicas, static vs dynamic tiles
out=gpurun_out; tag=${1:-r02f}
mkdir -p $out
python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED|correction mode" $out/${tag}_pytest_gpu.log | tail -12
for r in 1 2; do
  python bench.py --replicas $r --no-cpu-baseline --no-incumbent --no-configs --profile-out $out/${tag}_per_layer_rep$r.json > $out/${tag}_bench_rep$r.json 2> $out/${tag}_bench_rep$r.err; echo "bench rep$r rc=$?"
  tail -3 $out/${tag}_bench_rep$r.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_rep$r.json'))
print('replicas $r: value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'nolog',round(d['e2e']['no_logits']['value']), d.get('precision_variants'))
print(d['roofline']['per_layer_ms'])
PY
done
