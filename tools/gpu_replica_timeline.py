"""GPU session helper: timeline of two engine replicas alternating config-2 steps on two streams -- begin / end of every
launch of both engines relative to one reference event.  `python tools/gpu_replica_timeline.py [steps]`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer       # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
layers, _ = netdesc.describe_line_net(net)
recs = [LineRecognizer(layers, precision='fp16f8') for _ in range(2)]
recs[0].autotune_precision(budget=5e-4)
for i in recs[0].corrections:
    recs[1].set_layer_correction(i, recs[0].corrections[i])
if os.environ.get('B200OCR_LINK_REPLICAS', '1') != '0':
    recs[0].run_after(recs[1])
    recs[1].run_after(recs[0])
crops = torch.randint(0, 256, (256, 40, 1344, 3), dtype=torch.uint8, device='cuda')
streams = [torch.cuda.Stream() for _ in recs]
outs = [{}, {}]
for i in range(6):
    with torch.cuda.stream(streams[i & 1]):
        recs[i & 1].forward(crops, want_logits=False, out=outs[i & 1])
torch.cuda.synchronize()
for r in recs:
    r.profile(True)
ref = torch.cuda.Event(enable_timing=True)
ref.record()
for s in streams:
    s.wait_event(ref)
for i in range(steps):
    with torch.cuda.stream(streams[i & 1]):
        recs[i & 1].forward(crops, want_logits=False, out=outs[i & 1])
torch.cuda.synchronize()
rows = []
for r, rec in enumerate(recs):
    tags, lidx, t0, t1 = rec.profile_read_since(ref)
    rec.profile(False)
    for a, b, c, d in zip(tags, lidx, t0, t1):
        rows.append((float(c), float(d), r, int(b), int(a)))
rows.sort()
print('total', round(max(r[1] for r in rows), 3), 'ms for', steps, 'steps ->', round(max(r[1] for r in rows) / steps, 3), 'ms/step')
for c, d, r, layer, tag in rows:
    print(f'{c:8.3f} {d:8.3f}  {"A" if r == 0 else "    B"}  layer {layer:2d} tag {tag}  ({d - c:.3f})')
