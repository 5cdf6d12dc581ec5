#!/bin/bash
# round-2 GPU pass c: all parity tests (no -x), bench line
out=gpurun_out; tag=${1:-r02c}
mkdir -p $out
python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED|frames;|autotune:|correction mode|fp16f8w" $out/${tag}_pytest_gpu.log | tail -30
python bench.py --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -5 $out/${tag}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/'+'%s'%"r02c"+'_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nolog',d['e2e']['no_logits']['value'])
print('host',d['e2e']['host_ms_per_step_rank0'])
print('auto',d['config']['precision_autotune'])
print('inc',{k:v for k,v in d['incumbent_gpu'].items() if k!='variants' and k!='what'})
print('c3',d['config3'].get('value'),d['config3'].get('forward_only'),'c4',d['config4'].get('value'))
PY
