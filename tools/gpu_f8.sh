#!/bin/bash
# bring-up of the fp16f8 precision: per-layer diagnostics, parity tests, bench lines, ncu captures
out=gpurun_out; tag=${1:-f8a}
mkdir -p $out
timeout 300 python tests/gpu_diag.py lstm fp16f8,fp16x3 > $out/${tag}_diag_lstm.log 2>&1
timeout 300 python tests/gpu_diag.py transformer fp16f8 > $out/${tag}_diag_tr.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py --precision fp16f8 --no-cpu-baseline --profile-out $out/${tag}_per_layer_fp16f8.json > $out/${tag}_bench_fp16f8.json 2> $out/${tag}_bench_fp16f8.err
timeout 300 python bench.py --precision fp16x3 --no-cpu-baseline --profile-out $out/${tag}_per_layer_fp16x3.json > $out/${tag}_bench_fp16x3.json 2> $out/${tag}_bench_fp16x3.err
nl=16
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:igemm|lstm_tc|conv_first|ctc_collapse' -s ${nl} -c ${nl} -o $out/${tag}_step_fp16f8 -f \
    python tools/prof_step.py lstm fp16f8 2 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_step_fp16f8.ncu-rep --page raw --csv > $out/${tag}_step_fp16f8_raw.csv 2>/dev/null
echo done
