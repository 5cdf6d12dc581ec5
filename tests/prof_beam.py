"""ncu driver: one prefix-beam launch on bench.py's decode workload (256 x 336 x 120 peaky log-probs, k = 16).
Lives under tests/ because the workload generator is the oracle's (oracle/cases.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases                                  # noqa: E402
from pero_ocr_b200.decoders import prefix_beam_device     # noqa: E402

rng = np.random.default_rng(9)
lp = cases.peaky_logprobs(rng, 8, 336, 120, sharp=11.0)
print('classes above the -10 pruning threshold per frame:', float((lp[:, :, :-1] > -10).sum(axis=2).mean()))
x = torch.from_numpy(np.ascontiguousarray(np.tile(lp, (32, 1, 1)))).cuda()
for _ in range(2):
    prefix_beam_device(x, 16)
torch.cuda.synchronize()
print('done')
