#!/bin/bash
# autoregressive Transformer decoder: parity tests under both step-projection kernels, stage-wise diagnostics + timing
# (profiles/*_ar_decoder_diag.json), ncu launch list + full capture of one decoded position, compute-sanitizer over the
# token loop, and the config-4 page pipeline + the reference's GPU eager path (aux_bench)
out=gpurun_out; tag=${1:-ar}
mkdir -p $out
python -m pytest tests/test_zz_gpu_ar_decoder.py -m gpu -q > $out/${tag}_pytest_ar.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_ar.log
B200OCR_AR_LINEAR=0 python -m pytest tests/test_zz_gpu_ar_decoder.py -m gpu -q > $out/${tag}_pytest_ar_tiled.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_ar_tiled.log
python -m tests.gpu_ar_diag $out/${tag}_ar_decoder_diag.json > $out/${tag}_ar_diag.log 2>&1; echo "diag rc=$?" >> $out/${tag}_ar_diag.log
# ncu: launch list of one decode call, then a full capture of the 27 kernels of one decoded position (position 40 of the
# call: 5 encoder LayerNorms + 40 x 27 matching launches are skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_ar_launches.csv \
    python -m tests.prof_ar 64 1 > $out/${tag}_ar_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:linear_f32|step_attention|layernorm|embed_pe|argmax_alive' \
    -s 1085 -c 27 -o $out/${tag}_ar_step -f python -m tests.prof_ar 64 1 > $out/${tag}_ar_ncu_full.log 2>&1
ncu -i $out/${tag}_ar_step.ncu-rep --page raw --csv > $out/${tag}_ar_step_raw.csv 2>/dev/null
# compute-sanitizer over the token loop (3 lines keep the encoder short under instrumentation)
timeout 420 compute-sanitizer --tool memcheck python -m tests.prof_ar 3 1 > $out/${tag}_ar_memcheck.log 2>&1
timeout 420 compute-sanitizer --tool racecheck python -m tests.prof_ar 3 1 > $out/${tag}_ar_racecheck.log 2>&1
echo done
