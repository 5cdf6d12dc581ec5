#!/bin/bash
# round-2 pass t: token-loop kernels v2 (parity under all three variants, timing, ncu capture of one position,
# sanitizers) and the recycled page-locked result blocks of process_lines (test + bench e2e)
out=gpurun_out; tag=${1:-r02t}
mkdir -p $out
python -m pytest tests/test_zz_gpu_ar_decoder.py -m gpu -q > $out/${tag}_pytest_ar.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_ar.log
tail -3 $out/${tag}_pytest_ar.log
python -m pytest tests/test_gpu_engine.py -m gpu -q -k "recycled or sparse" > $out/${tag}_pytest_pool.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_pool.log
tail -3 $out/${tag}_pytest_pool.log
python -m tests.gpu_ar_diag $out/${tag}_ar_decoder_diag.json > $out/${tag}_ar_diag.log 2>&1; echo "diag rc=$?" >> $out/${tag}_ar_diag.log
tail -2 $out/${tag}_ar_diag.log
# one decoded position = 25 launches (embed, 2 x 11, class projection, argmax); position 40 of the call
ncu --set full --clock-control none --import-source on -k 'regex:linear_f32|step_attention|layernorm|embed_pe|argmax_alive' \
    -s 1005 -c 25 -o $out/${tag}_ar_step -f python -m tests.prof_ar 128 1 > $out/${tag}_ar_ncu_full.log 2>&1
ncu -i $out/${tag}_ar_step.ncu-rep --page raw --csv > $out/${tag}_ar_step_raw.csv 2>/dev/null
timeout 420 compute-sanitizer --tool memcheck python -m tests.prof_ar 3 1 > $out/${tag}_ar_memcheck.log 2>&1; tail -2 $out/${tag}_ar_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --kernel-name kns=linear_f32 --kernel-name kns=step_attention --kernel-name kns=sum_layernorm \
    python -m tests.prof_ar 3 1 > $out/${tag}_ar_racecheck.log 2>&1; tail -2 $out/${tag}_ar_racecheck.log
python bench.py --no-incumbent --no-configs > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('host_ms_per_step_rank0'), d['e2e'].get('pinned_result_pool'))
print('no_logits', d['e2e']['no_logits']['value'])
PY
echo done
