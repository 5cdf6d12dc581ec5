"""Short driver for compute-sanitizer: one small forward of each recogniser family through every product kernel
(first conv with TMA staging, halo / per-tap tcgen05 GEMMs with the tile ring, BiLSTM, CTC head, tcgen05 attention,
pad_lines, sparsification).  `python tools/sanitize_small.py`"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer       # noqa: E402
from pero_ocr_b200.sparse_logits import sparsify_device  # noqa: E402

rng = np.random.default_rng(0)
for kind, kw in (('lstm', {}), ('transformer', {'layers': 1})):
    net = synthetic.make_net(kind, 120, seed=0, out_gain=2.5, **kw)
    layers, _ = netdesc.describe_line_net(net)
    rec = LineRecognizer(layers, precision='fp16f8')
    crops = torch.from_numpy(rng.integers(0, 256, (3, 40, 272, 3), dtype=np.uint8)).cuda()
    out = rec.forward(crops, want_logits=True, want_confidence=True)
    sp = sparsify_device(out['logits'])
    torch.cuda.synchronize()
    print(kind, 'ok', out['lengths'].cpu().tolist(), int(sp.base.cpu()[-1]))
    rec.close()
