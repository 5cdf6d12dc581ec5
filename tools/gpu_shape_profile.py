"""GPU session helper: per-launch CUDA-event times of one forward at a given batch shape.
`python tools/gpu_shape_profile.py LINES WIDTH [LINES WIDTH ...]`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer       # noqa: E402

kind = os.environ.get('B200OCR_KIND', 'lstm')
net = synthetic.make_net(kind, 120, seed=0, out_gain=6.0)
layers, _ = netdesc.describe_line_net(net)
rec = LineRecognizer(layers, precision='fp16f8')
rec.autotune_precision(budget=5e-4)
args = [int(a) for a in sys.argv[1:]] or [256, 1344, 60, 2688]
for n, w in zip(args[0::2], args[1::2]):
    crops = torch.randint(0, 256, (n, 40, w, 3), dtype=torch.uint8, device='cuda')
    out = {}
    for _ in range(3):
        rec.forward(crops, want_logits=False, out=out)
    rec.profile(True)
    for _ in range(5):
        rec.forward(crops, want_logits=False, out=out)
    tags, lidx, ms = rec.profile_read()
    rec.profile(False)
    k = len(ms) // 5
    per = ms.reshape(5, k).mean(axis=0)
    print(json.dumps({'lines': n, 'width': w, 'step_ms': round(float(ms.sum() / 5), 3),
                      'per_launch_ms': [round(float(x), 3) for x in per], 'tags': [int(t) for t in tags[:k]],
                      'layers': [int(t) for t in lidx[:k]]}))
