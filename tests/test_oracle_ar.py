"""CPU: oracle/ar_oracle.py (restatement of the reference's autoregressive Transformer decoding)
against the output of the unmodified TransformerEngineLineOCR.transcribe_batch stored in tests/golden/ar_decoder*.npz
(oracle/make_golden.py: golden_ar_decoder).  This pins the parity reference of the device path
(tests/test_zz_gpu_ar_decoder.py).  Case 'small': 2 decoder layers, 32 classes, every line stops by itself; case
'wide': 3 decoder layers, 122 classes, no line ever emits the sentence boundary, so the loop ends on the length limit
(transformer_ocr_engine.py:79-82)."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.ar_oracle import ar_decoder_state, greedy_transcribe, postprocess_decoded
from oracle.nets import make_net
from tests.util import load_golden


@pytest.mark.parametrize('case', ['small', 'wide'])
def test_ar_greedy_decoding_matches_reference_golden(golden_dir, case):
    spec = cases.AR_CASES[case]
    gold = load_golden(golden_dir, spec['golden'])
    net = make_net('transformer', 120, seed=spec['encoder_seed'], layers=2)
    x = torch.from_numpy(cases.ar_inputs(spec)).float() / 255.0
    with torch.no_grad():                                     # TransformerOCR.encode (transformer.py:548-555)
        y = net.agg_act(net.agg(net.conv(x))).squeeze(2).permute(2, 0, 1)
        memory = net.trans_encoder(net.input_norm(y) + net.pe[:y.size(0)]).numpy()
    sd = ar_decoder_state(seed=spec['decoder_seed'], layers=spec['decoder_layers'], classes=spec['classes'])
    tokens, logits = greedy_transcribe(memory, sd, spec['decoder_layers'], 8, spec['classes'] - 2, spec['width'])
    assert logits.shape == gold['logits'].shape
    assert np.abs(logits - gold['logits']).max() <= 5e-4
    # the greedy choice wherever the reference's own decision is numerically meaningful
    srt = np.sort(gold['logits'], axis=2)
    decided = (srt[..., -1] - srt[..., -2]) > 2e-3
    assert np.array_equal(logits.argmax(axis=2)[decided], gold['logits'].argmax(axis=2)[decided])
    outs = postprocess_decoded(tokens, spec['classes'] - 1, spec['classes'] - 2)
    for i, o in enumerate(outs):
        assert o == list(gold['tokens'][i, :gold['lengths'][i]]), i
    if case == 'wide':                                        # ended by the length limit, not by the alive mask
        assert logits.shape[1] == spec['width'] // 4 + 1 and all(len(o) == spec['width'] // 4 for o in outs)
