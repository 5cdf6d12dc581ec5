"""GPU session helper: A/B of the three ways the first convolution stages its uint8 crop patch from HBM into shared
memory (conv_first.cu: 0 plain loads, 1 cp.async, 2 TMA) -- CUDA-event time of the kernel inside a device-resident
config-2 step, 10 steps each, interleaved twice.  `python tools/gpu_staging_ab.py > gpurun_out/<tag>_staging_ab.json`"""
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
crops = torch.zeros((256, 40, 1344, 3), dtype=torch.uint8, device='cuda')
crops[:, :, 32:-32] = torch.from_numpy(synthetic.bench_crops(256, 1280, seed=0)).cuda()
out = {}
ref = None
res = {'workload': '256 x 40 x 1344 x 3 uint8 crops -> 64-channel records (conv_first_mma_kernel<64>)', 'runs': []}
for rnd in range(2):
    for variant, name in ((0, 'plain'), (1, 'cp.async'), (2, 'tma'), (3, 'tma + tcgen05')):
        rec.set_flag(4, variant)
        o = rec.forward(crops, want_logits=True, out={})
        lg = o['logits'].clone()
        if ref is None:
            ref = lg
        same = bool(torch.equal(lg, ref))
        dmax = float((lg - ref).abs().max())
        for _ in range(2):
            rec.forward(crops, want_logits=False, out=out)
        rec.profile(True)
        for _ in range(10):
            rec.forward(crops, want_logits=False, out=out)
        tags, lidx, ms = rec.profile_read()
        rec.profile(False)
        first = ms[tags == 0]
        res['runs'].append({'round': rnd, 'staging': name, 'conv_first_ms_mean': float(first.mean()),
                            'conv_first_ms_min': float(first.min()), 'step_ms': float(ms.sum() / 10),
                            'logits_identical_to_plain': same, 'max_abs_logit_diff_to_plain': dmax})
print(json.dumps(res, indent=1))
