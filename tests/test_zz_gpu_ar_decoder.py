"""GPU parity of the autoregressive Transformer decoding behind B200TransformerEngineLineOCR:
encoder walk + b200ocr_ar_transcribe against (1) tokens / logits of the unmodified reference
TransformerEngineLineOCR.transcribe_batch (tests/golden/ar_decoder.npz) and (2) the CPU oracle (torch-CPU encoder +
oracle/ar_oracle.py) on other seeded batches.

Bars: every step's logits within 2e-3 of the reference (the decoder steps run in fp32; the encoder and the memory
K / V projection in the engine's fp16f8 tensor-core precision); greedy tokens identical wherever the reference's
top-2 margin exceeds 2e-3 -- a line is compared up to its first numerically undecidable step, after which two fp32
implementations may legitimately decode different continuations.  (The file name sorts last on purpose: the
parity tests of the CTC hot path run first.)
"""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.ar_oracle import greedy_transcribe, postprocess_decoded
from tests.util import load_golden

pytestmark = pytest.mark.gpu

TOL = 2e-3
MARGIN = 2e-3


def _engine(tmp_path, sd, **kw):
    from pero_ocr_b200.transformer_engine import B200TransformerEngineLineOCR
    js = cases.write_ar_engine_json(tmp_path, max_line_width=kw.pop('max_line_width', None))
    return B200TransformerEngineLineOCR(js, torch.device('cuda', 0), state_dict=sd, **kw)


def _compare(ours_logits, ref_logits):
    """-> (max |diff| over compared steps, per line number of steps compared).  Lines are compared step by step up
    to the first step whose reference margin is below MARGIN and whose greedy choice differs."""
    n = ref_logits.shape[0]
    steps = min(ours_logits.shape[1], ref_logits.shape[1])
    worst, compared = 0.0, []
    for i in range(n):
        k = 0
        for s in range(steps):
            ref, got = ref_logits[i, s], ours_logits[i, s]
            srt = np.sort(ref)
            same = int(got.argmax()) == int(ref.argmax())
            if not same:
                assert srt[-1] - srt[-2] <= MARGIN, (i, s, srt[-1] - srt[-2])
                break
            worst = max(worst, float(np.abs(got - ref).max()))
            k += 1
        compared.append(k)
    return worst, compared


def _oracle_memory(net, nchw_u8):
    x = torch.from_numpy(nchw_u8).float() / 255.0
    with torch.no_grad():                                     # TransformerOCR.encode (transformer.py:548-555)
        y = net.agg_act(net.agg(net.conv(x))).squeeze(2).permute(2, 0, 1)
        return net.trans_encoder(net.input_norm(y) + net.pe[:y.size(0)]).numpy()


@pytest.mark.parametrize('linear_variant', [3, 2, 1, 0])
def test_transcribe_batch_matches_reference_golden(tmp_path, golden_dir, linear_variant):
    gold = load_golden(golden_dir, 'ar_decoder.npz')
    net, dec, sd = cases.ar_state_dict()
    eng = _engine(tmp_path, sd)
    # token-loop kernels: 3 = the launches of a position replayed as a CUDA graph (device-side position counter) on top
    # of 2 = (q|k|v in one launch, K split over CTAs + summing LayerNorm, CTA-per-head step
    # attention), 1 = split-K projections + warp-per-head attention, 0 = single K walker
    eng.net.set_flag(2, linear_variant)
    assert eng.sentence_boundary_ind == cases.AR_CASE['classes'] - 2 and eng.ignore_ind == cases.AR_CASE['classes'] - 1
    launches0 = eng.net.launch_count
    outs, logits = eng.transcribe_batch(cases.ar_inputs(), is_cached=True)
    assert eng.net.launch_count - launches0 > 100          # the CUDA path ran (encoder + token loop)
    assert logits.dtype == np.float32 and logits.shape[0] == gold['logits'].shape[0]
    worst, compared = _compare(logits, gold['logits'])
    assert worst <= TOL, worst
    if min(compared) == gold['logits'].shape[1]:           # no undecidable step met: everything must be identical
        assert logits.shape == gold['logits'].shape
        for i, o in enumerate(outs):
            assert list(o) == list(gold['tokens'][i, :gold['lengths'][i]]), i
    else:
        assert min(compared) >= 20, compared
    # the stop test may run every step or every few steps: same result
    eng.check_every = 1
    outs1, logits1 = eng.transcribe_batch(cases.ar_inputs(), is_cached=True)
    assert logits1.shape == logits.shape and np.array_equal(logits1, logits)
    assert all(np.array_equal(a, b) for a, b in zip(outs, outs1))
    # no_logits: same tokens
    outs2, none = eng.transcribe_batch(cases.ar_inputs(), is_cached=True, no_logits=True)
    assert none is None and all(np.array_equal(a, b) for a, b in zip(outs, outs2))


def test_run_ocr_matches_oracle_on_a_narrow_batch(tmp_path):
    """run_ocr centre-pads batches narrower than 1088 px (transformer_ocr_engine.py:36-40); 5 lines (more than one
    32-row tile boundary is not needed, but n is not a multiple of 4) of ragged content in a 704 px batch."""
    net, dec, sd = cases.ar_state_dict()
    eng = _engine(tmp_path, sd)
    rng = np.random.default_rng(123)
    n, w = 5, 704
    batch = np.zeros((n, 40, w, 3), dtype=np.uint8)
    for i, wi in enumerate([640, 600, 333, 128, 40]):
        batch[i, :, 32:32 + wi] = rng.integers(0, 256, (40, wi, 1), dtype=np.uint8)
    texts, logits = eng.run_ocr(batch)
    padded = np.zeros((n, 3, 40, 1088), dtype=np.uint8)
    s = (1088 - w) // 2
    padded[:, :, :, s:s + w] = batch.transpose(0, 3, 1, 2)
    memory = _oracle_memory(net, padded)
    sb = cases.AR_CASE['classes'] - 2
    tokens, ref_logits = greedy_transcribe(memory, dec, cases.AR_CASE['decoder_layers'], 8, sb, 1088)
    worst, compared = _compare(logits, ref_logits)
    assert worst <= TOL, worst
    if min(compared) == ref_logits.shape[1]:
        assert logits.shape == ref_logits.shape
        ref_out = postprocess_decoded(tokens, sb + 1, sb)
        assert texts == [''.join(eng.characters[c] for c in o) for o in ref_out]
    else:
        assert min(compared) >= 10, compared


def test_process_lines_splits_and_merges_on_the_device_path(tmp_path):
    """process_lines end to end: a line wider than max_line_width is split into overlapping parts, recognised in one
    batch and merged (line_ocr_engine.py:95-119, 180-211); the result equals run_ocr on the same parts + the merge."""
    from pero_ocr_b200.transformer_engine import merge_transcriptions_and_logits
    net, dec, sd = cases.ar_state_dict()
    eng = _engine(tmp_path, sd, max_line_width=512, batch_size=8)
    rng = np.random.default_rng(7)
    lines = [np.repeat(rng.integers(0, 256, (40, w, 1), dtype=np.uint8), 3, axis=2) for w in (900, 300)]
    tr, lg, co = eng.process_lines(lines, sparse_logits=False)
    assert all(isinstance(t, str) for t in tr) and co[0] == [0, len(tr[0])] and co[1] == [0, len(tr[1])]
    # manual composition of the one batch process_lines builds: the 900 px line in parts [0:512], [384:896],
    # [768:900] + the 300 px line, under width min(928, 512 + 64) + 64
    parts = [lines[0][:, 0:512], lines[0][:, 384:896], lines[0][:, 768:1280], lines[1]]
    batch = np.zeros((4, 40, 512 + 64 + 64, 3), dtype=np.uint8)
    for d, p in zip(batch, parts):
        d[:, 32:32 + p.shape[1]] = p
    t_parts, l_parts = eng.run_ocr(batch)
    t_merged, l_merged = merge_transcriptions_and_logits(t_parts[:3], l_parts[:3])
    assert tr[0] == t_merged and np.array_equal(lg[0], l_merged)
    assert tr[1] == t_parts[3] and np.array_equal(lg[1], l_parts[3][:len(t_parts[3])])
    sp_tr, sp_lg, _ = eng.process_lines(lines, sparse_logits=True)
    assert sp_tr == tr and sp_lg[0].shape == lg[0].shape


def test_wide_case_matches_reference_golden(tmp_path, golden_dir):
    """Second golden of the unmodified TransformerEngineLineOCR.transcribe_batch: 3 decoder layers, 122 classes
    (partial 32-wide output tiles, several argmax passes per warp), no line emits the stop symbol, so the loop ends on
    the length limit after W/4 + 1 = 289 steps (transformer_ocr_engine.py:79-82)."""
    from pero_ocr_b200.transformer_engine import B200TransformerEngineLineOCR
    spec = cases.AR_CASES['wide']
    gold = load_golden(golden_dir, spec['golden'])
    _, _, sd = cases.ar_state_dict(spec)
    eng = B200TransformerEngineLineOCR(cases.write_ar_engine_json(tmp_path, spec=spec), torch.device('cuda', 0),
                                       state_dict=sd)
    outs, logits = eng.transcribe_batch(cases.ar_inputs(spec), is_cached=True)
    assert logits.shape[1] == gold['logits'].shape[1] == spec['width'] // 4 + 1
    worst, compared = _compare(logits, gold['logits'])
    assert worst <= TOL, worst
    if min(compared) == gold['logits'].shape[1]:
        for i, o in enumerate(outs):
            assert list(o) == list(gold['tokens'][i, :gold['lengths'][i]]), i
    else:
        assert min(compared) >= 40, compared
