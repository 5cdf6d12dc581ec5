"""GPU session helper: device time of b200ocr_sparsify_logits at the config-2 shape (256 x 336 x 120), on flat logits
(every class kept: the bench net's degenerate case) and on peaky ones (a trained recogniser)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200.sparse_logits import sparsify_device          # noqa: E402

rng = np.random.default_rng(0)
for name, scale in (('flat', 0.3), ('peaky', 12.0)):
    x = torch.from_numpy((rng.standard_normal((256, 336, 120)) * scale).astype(np.float32)).cuda()
    sp = sparsify_device(x)
    torch.cuda.synchronize()
    total = int(sp.base[-1])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        sp = sparsify_device(x, out=sp)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    gb = (2 * x.numel() * 4 + total * 8) / 1e9
    print(f'{name}: {ms * 1e3:.1f} us per batch, {total / x.shape[0] / x.shape[1]:.1f} entries per frame, '
          f'{gb / (ms / 1e3):.0f} GB/s of logits read twice + entries written')
