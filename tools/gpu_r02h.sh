#!/bin/bash
# round-2 GPU pass h: parity tests, full bench line (incumbent, configs 3/4), ncu launch list of the bench command and
# one `ncu --set full` capture of a device-resident step in the benched arithmetic (autotune budget 5e-4)
out=gpurun_out; tag=${1:-r02h}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log; tail -2 $out/${tag}_smoke.log
python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED" $out/${tag}_pytest_gpu.log | tail -8
python bench.py --profile-out $out/${tag}_per_layer.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -5 $out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'nolog',round(d['e2e']['no_logits']['value']),'frac',round(d['roofline']['frac'],3))
print('host',d['e2e']['host_ms_per_step_rank0'])
print('inc',{k:v for k,v in d['incumbent_gpu'].items() if k not in ('variants','what','kernels','hosted_as')})
print('c3',d['config3'].get('value'),d['config3'].get('forward_only'),d['config3'].get('error'))
print('c4',d['config4'].get('value'),d['config4'].get('page_by_page'),d['config4'].get('error'))
print(d['roofline']['per_layer_ms'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-incumbent --no-configs > $out/${tag}_ncu_bench.log 2>&1
B200OCR_AUTOTUNE_BUDGET=5e-4 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o $out/${tag}_step_fp16f8 -f python tools/prof_step.py lstm fp16f8 3 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_step_fp16f8.ncu-rep --page raw --csv > $out/${tag}_step_fp16f8_raw.csv 2>/dev/null
ls -la $out | grep ${tag}
