"""GPU session helper: the first-conv + second-conv pair over L2-sized chunks of lines (debug flag 6) -- step time and
the two layers' time for several chunk sizes (0 = whole batch), CUDA events, 8 steps each.
`python tools/gpu_chunk_sweep.py > gpurun_out/<tag>_chunk_sweep.json`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer       # noqa: E402

net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
layers, _ = netdesc.describe_line_net(net)
rec = LineRecognizer(layers, precision='fp16f8')
if os.environ.get('B200OCR_AUTOTUNE_BUDGET'):
    rec.autotune_precision(budget=float(os.environ['B200OCR_AUTOTUNE_BUDGET']))
crops = torch.zeros((256, 40, 1344, 3), dtype=torch.uint8, device='cuda')
crops[:, :, 32:-32] = torch.from_numpy(synthetic.bench_crops(256, 1280, seed=0)).cuda()
out, ref, rows = {}, None, []
for chunk in (0, 2, 4, 6, 8, 12, 16, 32, 0, 4):
    rec.set_flag(6, chunk)
    lg = rec.forward(crops, want_logits=True, out={})['logits'].clone()
    if ref is None:
        ref = lg
    for _ in range(2):
        rec.forward(crops, want_logits=False, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        rec.forward(crops, want_logits=False, out=out)
    b.record()
    torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / 8
    rec.profile(True)
    for _ in range(3):
        rec.forward(crops, want_logits=False, out=out)
    tags, lidx, ms = rec.profile_read()
    rec.profile(False)
    rows.append({'chunk_lines': chunk, 'step_ms': step_ms, 'lines_per_s': 256 / (step_ms / 1e3),
                 'layer0_ms': float(ms[lidx == 0].sum() / 3), 'layer1_ms': float(ms[lidx == 1].sum() / 3),
                 'launches_per_step': int(len(ms) // 3), 'logits_identical': bool(torch.equal(lg, ref))})
print(json.dumps({'workload': 'config 2, 256 x 40 x 1344, fp16f8', 'rows': rows}, indent=1))
