#!/bin/bash
# compute-sanitizer passes: memcheck over one small forward of each family; racecheck filtered to the kernels that use
# shared memory across warps without TMA / mbarrier-only hand-offs; racecheck over the AR decoder's token-loop kernels
out=gpurun_out; tag=${1:-r02s}
mkdir -p $out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $out/${tag}_memcheck_small.log 2>&1; tail -4 $out/${tag}_memcheck_small.log
timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=conv_first_mma --kernel-name kns=attention_tc --kernel-name kns=pad_lines \
    --kernel-name kns=sparsify --kernel-name kns=ctc_collapse --kernel-name kns=layernorm \
    python tools/sanitize_small.py > $out/${tag}_racecheck_small.log 2>&1; tail -4 $out/${tag}_racecheck_small.log
timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=linear_f32 --kernel-name kns=step_attention --kernel-name kns=embed_pe \
    --kernel-name kns=argmax_alive --kernel-name kns=layernorm python -m tests.prof_ar 3 1 > $out/${tag}_racecheck_ar.log 2>&1; tail -4 $out/${tag}_racecheck_ar.log
timeout 600 compute-sanitizer --tool synccheck --kernel-name kns=attention_tc --kernel-name kns=conv_first_mma python tools/sanitize_small.py > $out/${tag}_synccheck_small.log 2>&1; tail -4 $out/${tag}_synccheck_small.log
