"""GPU bring-up diagnostic: why the device prefix beam rejected the c119 case's log-probs as un-normalised."""
import os, sys, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases
from oracle.forward_oracle import full_logprobs
from tests.util import make_case_net, write_engine_json
from pero_ocr_b200.engine import B200EngineLineOCR
from pero_ocr_b200.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder, prefix_beam_device

tmp = tempfile.mkdtemp()
eng = B200EngineLineOCR(write_engine_json(tmp, 'lstm_c119'), torch.device('cuda', 0), batch_size=2, module=make_case_net('lstm_c119'))
lines = cases.engine_lines('lstm_c119')
tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
for i in range(len(lines)):
    lp = full_logprobs(lg[i])[co[i][0]:co[i][1]]
    x64 = lp.astype(np.float64)
    dev_host = np.abs(np.exp(x64).sum(1) - 1).max()
    xd = torch.from_numpy(np.ascontiguousarray(x64[None])).cuda()
    dev_dev = float((xd.exp().sum(2) - 1).abs().max())
    out = {}
    for k in (1, 4, 16):
        labels, lengths, scores, status = prefix_beam_device(xd, k)
        out[k] = int(status.cpu()[0])
    # halves
    t = x64.shape[0]
    parts = {}
    for a, b in ((0, t // 2), (t // 2, t), (0, 8), (0, 64), (0, 72), (0, 74), (1, t)):
        xs = torch.from_numpy(np.ascontiguousarray(x64[None, a:b])).cuda()
        parts[(a, b)] = int(prefix_beam_device(xs, 4)[3].cpu()[0])
    print(i, lp.shape, lp.dtype, 'host dev', dev_host, 'torch dev', dev_dev, 'status by k', out, parts,
          'min', float(x64.min()), 'has -inf', bool(np.isinf(x64).any()))
