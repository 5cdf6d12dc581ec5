#!/bin/bash
# autoregressive Transformer decoder: parity tests under both step-projection kernels, stage-wise diagnostics + timing
# (profiles/*_ar_decoder_diag.json), and the config-4 page pipeline (profiles/*_config4_page_pipeline.json)
out=gpurun_out; tag=${1:-ar}
mkdir -p $out
python -m pytest tests/test_zz_gpu_ar_decoder.py -m gpu -q > $out/${tag}_pytest_ar.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_ar.log
B200OCR_AR_LINEAR=0 python -m pytest tests/test_zz_gpu_ar_decoder.py -m gpu -q > $out/${tag}_pytest_ar_tiled.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_ar_tiled.log
python -m tests.gpu_ar_diag $out/${tag}_ar_decoder_diag.json > $out/${tag}_ar_diag.log 2>&1; echo "diag rc=$?" >> $out/${tag}_ar_diag.log
python -m tests.aux_bench config4 > $out/${tag}_config4_page_pipeline.json 2> $out/${tag}_config4.err
echo done
