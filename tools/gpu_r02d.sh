#!/bin/bash
# round-2 GPU pass d: parity tests, c119 beam diagnostic, staging A/B (+ ncu captures of the three variants), bench
out=gpurun_out; tag=${1:-r02d}
mkdir -p $out
python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED|frames;|autotune:|correction mode|fp16f8w" $out/${tag}_pytest_gpu.log | tail -30
python tools/gpu_diag_c119.py > $out/${tag}_diag_c119.log 2>&1; tail -6 $out/${tag}_diag_c119.log
python tools/gpu_staging_ab.py > $out/${tag}_staging_ab.json 2> $out/${tag}_staging_ab.err; echo "staging rc=$?"; tail -3 $out/${tag}_staging_ab.err
grep -E "staging|conv_first_ms_mean|identical" $out/${tag}_staging_ab.json | paste - - - | head -8
for v in 0 1 2; do
  B200OCR_CROP_STAGING=$v ncu --set full --clock-control none --import-source on -k regex:conv_first_mma -s 1 -c 1 \
      -o $out/${tag}_conv_first_staging$v -f python tools/prof_step.py lstm fp16f8 2 > $out/${tag}_ncu_staging$v.log 2>&1
  ncu -i $out/${tag}_conv_first_staging$v.ncu-rep --page raw --csv > $out/${tag}_conv_first_staging${v}_raw.csv 2>/dev/null
done
python bench.py --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -5 $out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nolog',d['e2e']['no_logits']['value'])
print('host',d['e2e']['host_ms_per_step_rank0'])
print('inc',{k:v for k,v in d['incumbent_gpu'].items() if k!='variants' and k!='what'})
print('c3',d['config3'].get('value'),d['config3'].get('forward_only'),'c4',d['config4'].get('value'))
PY
