#!/bin/bash
# round-2 GPU pass e: parity tests, c119 diag, L2 chunk sweep, DRAM traffic of the first two layers with / without chunking
out=gpurun_out; tag=${1:-r02e}
mkdir -p $out
python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
grep -E "passed|failed|FAILED|correction mode" $out/${tag}_pytest_gpu.log | tail -12
python tools/gpu_diag_c119.py > $out/${tag}_diag_c119.log 2>&1; tail -9 $out/${tag}_diag_c119.log
B200OCR_AUTOTUNE_BUDGET=5e-4 python tools/gpu_chunk_sweep.py > $out/${tag}_chunk_sweep.json 2> $out/${tag}_chunk_sweep.err; echo "sweep rc=$?"; tail -3 $out/${tag}_chunk_sweep.err
python - <<PY
import json
for r in json.load(open('gpurun_out/${tag}_chunk_sweep.json'))['rows']:
    print(r)
PY
for c in 0 4; do
  B200OCR_L2_CHUNK=$c B200OCR_AUTOTUNE_BUDGET=5e-4 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
      --clock-control none -k regex:'conv_first_mma|igemm_halo' -c 400 --csv --log-file $out/${tag}_dram_chunk$c.csv \
      python tools/prof_step.py lstm fp16f8 1 > $out/${tag}_ncu_chunk$c.log 2>&1
done
python - <<PY
import csv
for c in (0,4):
    rows=list(csv.reader(open('gpurun_out/${tag}_dram_chunk%d.csv'%c)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID']
    if not hdr: print('no data',c); continue
    h=rows[hdr[0]]; body=rows[hdr[0]+1:]
    ki=h.index('Kernel Name'); mi=h.index('Metric Name'); ui=h.index('Metric Unit'); vi=h.index('Metric Value')
    tot={}
    for r in body:
        if len(r)<=vi: continue
        name='conv_first' if 'conv_first' in r[ki] else ('halo64' if 'igemm_halo_kernel<64' in r[ki] else 'halo128')
        v=float(r[vi].replace(',',''))
        u=r[ui]
        scale={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-6,'us':1e-3,'ms':1,'s':1e3}.get(u,1)
        tot[(name,r[mi])]=tot.get((name,r[mi]),0)+v*scale
    print('chunk',c,{k:round(v/(1e9 if 'bytes' in k[1] else 1),3) for k,v in sorted(tot.items())})
PY
