"""Short driver for ncu captures: builds the config-2 recogniser and runs a few device-resident steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pero_ocr_b200 import netdesc, synthetic
from pero_ocr_b200.engine import LineRecognizer

kind = sys.argv[1] if len(sys.argv) > 1 else 'lstm'
precision = sys.argv[2] if len(sys.argv) > 2 else 'fp16x3'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 256
net = synthetic.make_net(kind, 120, seed=0, out_gain=6.0 if kind == 'lstm' else 2.5)
layers, _ = netdesc.describe_line_net(net)
rec = LineRecognizer(layers, precision=precision)
if os.environ.get('B200OCR_CROP_STAGING'):          # A/B of the first conv's uint8 staging (0 plain, 1 cp.async, 2 TMA)
    rec.set_flag(4, int(os.environ['B200OCR_CROP_STAGING']))
if os.environ.get('B200OCR_L2_CHUNK'):               # lines per chunk of the first two conv layers (flag 6)
    rec.set_flag(6, int(os.environ['B200OCR_L2_CHUNK']))
if os.environ.get('B200OCR_AUTOTUNE_BUDGET'):
    rec.autotune_precision(budget=float(os.environ['B200OCR_AUTOTUNE_BUDGET']))
crops = torch.zeros((batch, 40, 1344, 3), dtype=torch.uint8, device='cuda')
crops[:, :, 32:-32] = torch.from_numpy(synthetic.bench_crops(batch, 1280, seed=0)).cuda()
out = {}
for _ in range(max(0, steps - 1)):
    rec.forward(crops, want_logits=False, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures exactly the last step
rec.forward(crops, want_logits=False, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done', rec.launch_count)
