#!/bin/bash
# round-2 GPU pass k: attention tests, config-3 timing with / without the tcgen05 attention, ncu of the attention kernel
out=gpurun_out; tag=${1:-r02k}
mkdir -p $out
python -m pytest tests/test_gpu_engine.py -m gpu -q -s -k "attention" > $out/${tag}_pytest_attention.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_attention.log
grep -E "passed|failed|FAILED|T = " $out/${tag}_pytest_attention.log | tail -14
python - <<'PY' > $out/${tag}_attention_ab.json
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from pero_ocr_b200 import netdesc, synthetic
from pero_ocr_b200.engine import LineRecognizer
net = synthetic.make_net('transformer', 120, seed=0, out_gain=2.5, layers=2)
layers, _ = netdesc.describe_line_net(net)
rec = LineRecognizer(layers, precision='fp16f8')
crops = torch.zeros((256, 40, 1344, 3), dtype=torch.uint8, device='cuda')
crops[:, :, 32:-32] = torch.from_numpy(synthetic.bench_crops(256, 1280, seed=0)).cuda()
out, res = {}, {}
for flag in (1, 0, 1):
    rec.set_flag(8, flag)
    for _ in range(2):
        rec.forward(crops, want_logits=False, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        rec.forward(crops, want_logits=False, out=out)
    b.record(); torch.cuda.synchronize()
    rec.profile(True)
    rec.forward(crops, want_logits=False, out=out)
    tags, lidx, ms = rec.profile_read()
    rec.profile(False)
    res.setdefault('tcgen05' if flag else 'cuda_core', []).append(
        {'step_ms': a.elapsed_time(b) / 5, 'lines_per_s': 256 * 5 / (a.elapsed_time(b) / 1e3),
         'other_ms (2 attention + 5 layernorm launches + collapse)': float(ms[tags == 3].sum()),
         'largest_other_launch_ms': float(ms[tags == 3].max())})
print(json.dumps(res, indent=1))
PY
cat $out/${tag}_attention_ab.json
ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o $out/${tag}_attention_tc -f \
    python tools/prof_step.py transformer fp16f8 1 > $out/${tag}_ncu_attention.log 2>&1
ncu -i $out/${tag}_attention_tc.ncu-rep --page raw --csv > $out/${tag}_attention_tc_raw.csv 2>/dev/null
tail -3 $out/${tag}_ncu_attention.log
