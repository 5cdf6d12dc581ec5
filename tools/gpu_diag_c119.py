"""GPU bring-up diagnostic: why the device prefix beam rejected the c119 case's log-probs as un-normalised."""
import os, sys, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases
from oracle.forward_oracle import full_logprobs
from tests.util import make_case_net, write_engine_json, load_golden
from pero_ocr_b200.engine import B200EngineLineOCR
from pero_ocr_b200.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder, GreedyDecoder, prefix_beam_device

gold = load_golden(os.path.join(ROOT, 'tests', 'golden'), 'engine_lstm_c119.npz')
spec = cases.ENGINE_CASES['lstm_c119']
for precision in ('fp16x3', 'fp16f8'):
    tmp = tempfile.mkdtemp()
    eng = B200EngineLineOCR(write_engine_json(tmp, 'lstm_c119'), torch.device('cuda', 0), batch_size=2, precision=precision,
                            module=make_case_net('lstm_c119'))
    lines = cases.engine_lines('lstm_c119')
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    letters = cases.json_characters(spec['json_chars']) + [BLANK_SYMBOL]
    gd, bd = GreedyDecoder(letters), CTCPrefixLogRawNumpyDecoder(letters, 4)
    for i in range(len(lines)):
        lp = full_logprobs(lg[i])[co[i][0]:co[i][1]]
        x64 = lp.astype(np.float64)
        dev_host = np.abs(np.exp(x64).sum(1) - 1).max()
        g = gd(lp).best_hyp()
        try:
            b = bd(x64).best_hyp()
            ok = b == str(gold['decoder_beam4'][i])
        except ValueError as e:
            b, ok = f'ValueError {e}', False
        xd = torch.from_numpy(np.ascontiguousarray(x64[None])).cuda()
        st = int(prefix_beam_device(xd, 4)[3].cpu()[0])
        # per-frame deviation as the kernel computes it, and a look for non-finite values
        print(precision, i, lp.shape, 'host dev %.2e' % dev_host, 'greedy ok', g == str(gold['decoder_greedy'][i]), 'beam', ok,
              b[:40] if not ok else '', 'direct status', st, 'nan', bool(np.isnan(x64).any()), 'min %.1f' % x64.min(),
              'contig', x64.flags['C_CONTIGUOUS'], lp.flags['C_CONTIGUOUS'])
