#!/bin/bash
# One GPU-box pass: parity tests, bench lines (default precision + fp16x3 + the reference arm), the ncu launch list of
# the bench command and one `ncu --set full` capture of a single device-resident step.  Outputs: gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01e'
tag=${1:-r01}
prec=${2:-fp16f8}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
python bench.py --precision $prec --profile-out $out/${tag}_per_layer_${prec}.json > $out/${tag}_bench_${prec}.json 2> $out/${tag}_bench_${prec}.err
python bench.py --precision fp16x3 --no-cpu-baseline --profile-out $out/${tag}_per_layer_fp16x3.json > $out/${tag}_bench_fp16x3.json 2> $out/${tag}_bench_fp16x3.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of the bench command itself (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --precision $prec --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
# full capture of one device-resident step (second step: skip the first step's launches)
nl=16
ncu --set full --clock-control none --import-source on -k 'regex:igemm|lstm_tc|conv_first|ctc_collapse' -s ${nl} -c ${nl} -o $out/${tag}_step_${prec} -f \
    python tools/prof_step.py lstm $prec 2 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_step_${prec}.ncu-rep --page raw --csv > $out/${tag}_step_${prec}_raw.csv 2>/dev/null
timeout 600 python -m tests.aux_bench > $out/${tag}_aux_bench.json 2> $out/${tag}_aux_bench.err
echo done
