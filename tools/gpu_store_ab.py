"""GPU session helper: A/B of the epilogue store paths of the tensor-core GEMM kernels (flag 9: 0 = 256-bit st.global.v8, the default;
4 = 16-byte stores; the first run of this script, r02u, compared v8 against the then-default 16-byte / shared-memory
staged stores) -- per-layer CUDA-event times inside a device-resident
config-2 step, 10 steps each, interleaved twice; logits must be bit-identical.
`python tools/gpu_store_ab.py > gpurun_out/<tag>_store_ab.json`"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer       # noqa: E402

net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
layers, _ = netdesc.describe_line_net(net)
rec = LineRecognizer(layers, precision='fp16f8')
rec.autotune_precision(budget=float(os.environ.get('B200OCR_AUTOTUNE_BUDGET', '5e-4')))
crops = torch.zeros((256, 40, 1344, 3), dtype=torch.uint8, device='cuda')
crops[:, :, 32:-32] = torch.from_numpy(synthetic.bench_crops(256, 1280, seed=0)).cuda()
out = {}
ref = None
res = {'workload': 'config 2 step, 256 x 40 x 1344, fp16f8 autotuned', 'runs': []}
for rnd in range(2):
    for variant, name in ((0, 'complete records'), (1, "lo' plane unwritten where unread")):
        rec.set_flag(11, variant)
        o = rec.forward(crops, want_logits=True, out={})
        lg = o['logits'].clone()
        if ref is None:
            ref = lg
        same = bool(torch.equal(lg, ref))
        for _ in range(2):
            rec.forward(crops, want_logits=False, out=out)
        rec.profile(True)
        for _ in range(10):
            rec.forward(crops, want_logits=False, out=out)
        tags, lidx, ms = rec.profile_read()
        rec.profile(False)
        n = len(ms) // 10
        per = ms.reshape(10, n).mean(axis=0)
        res['runs'].append({'round': rnd, 'stores': name, 'step_ms': float(ms.sum() / 10),
                            'per_launch_ms': [round(float(x), 4) for x in per],
                            'launch_layer': [int(x) for x in lidx[:n]], 'launch_tag': [int(x) for x in tags[:n]],
                            'logits_identical': same})
rec.set_flag(11, 0)
print(json.dumps(res, indent=1))
