"""CPU: host side of the autoregressive Transformer engine (pero_ocr_b200/transformer_engine.py, netdesc) against
outputs of the unmodified reference stored in tests/golden/ar_host.npz and transformer_ocr_keys.json
(oracle/make_golden.py: golden_ar_host).  No CUDA call is made: run_ocr is the deterministic stand-in of
oracle/cases.py, on both sides."""
import json
import os

import numpy as np
import pytest

from oracle import cases
from tests.util import load_golden


def _host_engine():
    from pero_ocr_b200.transformer_engine import B200TransformerEngineLineOCR
    spec = cases.AR_HOST_CASE
    eng = B200TransformerEngineLineOCR.__new__(B200TransformerEngineLineOCR)     # no device: host logic only
    eng.characters = cases.json_characters(spec['classes'] - 2) + ['​', '']
    eng.line_px_height, eng.line_padding_px, eng.net_subsampling = 40, 32, 4
    eng.max_line_width = spec['max_line_width']
    eng.batch_size = spec['batch_size']
    eng.max_input_horizontal_pixels = 480 * spec['batch_size']
    eng.run_ocr = cases.ar_host_fake_run_ocr(eng.characters[:-2])
    return eng


def test_process_lines_transformer_mode_matches_reference(golden_dir):
    gold = load_golden(golden_dir, 'ar_host.npz')
    eng = _host_engine()
    lines = cases.ar_host_lines()
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    assert tr == list(gold['transcriptions'])
    for i in range(len(lines)):
        assert lg[i].dtype == np.float32 and np.array_equal(lg[i], gold[f'logits_{i}']), i
        assert list(co[i]) == list(gold[f'coords_{i}'])
    tr_t, lg_t, co_t = eng.process_lines([l.copy() for l in lines], sparse_logits=False, tight_crop_logits=True)
    assert tr_t == tr
    for i in range(len(lines)):
        assert co_t[i] == [None, None] and np.array_equal(lg_t[i], gold[f'tight_{i}'])
    tr_s, lg_s, co_s = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    assert tr_s == tr and co_s == co
    for i in range(len(lines)):
        assert np.array_equal(lg_s[i].toarray(), gold[f'sparse_{i}']), i
    tr_n, lg_n, co_n = eng.process_lines([l.copy() for l in lines], no_logits=True)
    assert tr_n == tr and all(x is None for x in lg_n) and all(x is None for x in co_n)


def test_find_best_overlap_and_levenshtein_match_reference(golden_dir):
    from pero_ocr_b200.transformer_engine import find_best_overlap, levenshtein_distance
    gold = load_golden(golden_dir, 'ar_host.npz')
    for (a, b), want in zip(gold['overlap_pairs'], gold['overlaps']):
        assert find_best_overlap(str(a), str(b)) == int(want), (a, b)
    assert levenshtein_distance(list('kitten'), list('sitting')) == 3
    assert levenshtein_distance([], list('abc')) == 3 and levenshtein_distance(list('abc'), []) == 3


@pytest.mark.parametrize('a, b, want', [
    # the reference's known answers for levenshtein_distance (test/test_sequence_alignment.py:9-48)
    ([1], [1], 0), ([1], [2], 1), ([1], [2, 1], 1), ([1, 2], [1], 1), ([1, 2, 3], [1, -1, -2, 3], 2),
    ([1, -1, -2, 3], [1, 2, 3], 2), ([1, 2, 3], [], 3), ([], [1, 2, 3], 3)])
def test_levenshtein_distance_reference_known_answers(a, b, want):
    from pero_ocr_b200.transformer_engine import levenshtein_distance
    assert levenshtein_distance(a, b) == want


def test_postprocess_decoded_matches_oracle():
    from oracle.ar_oracle import postprocess_decoded as oracle_pp
    from pero_ocr_b200.transformer_engine import postprocess_decoded
    rng = np.random.default_rng(4)
    tokens = rng.integers(0, 8, (40, 17))                     # [steps, N]; 6 = sentence boundary, 7 = ignore
    ours = postprocess_decoded(tokens.T, 7, 6)
    assert [list(o) for o in ours] == oracle_pp(tokens, 7, 6)
    assert [list(o) for o in postprocess_decoded(np.zeros((3, 0), dtype=np.int64), 7, 6)] == [[], [], []]


def test_describe_transformer_ocr_reads_the_reference_checkpoint_layout(golden_dir):
    """The synthetic checkpoint has exactly the keys / shapes of the unmodified transformer.build_net state dict, and
    its encoder half turns into the same layer list as the nn.Module walk that the CTC Transformer variant uses."""
    from pero_ocr_b200 import _lib, netdesc
    with open(os.path.join(golden_dir, 'transformer_ocr_keys.json')) as f:
        ref_keys = json.load(f)
    net, dec, sd = cases.ar_state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == ref_keys
    layers, decoder = netdesc.describe_transformer_ocr(sd, json.dumps(cases.AR_NET_CONFIG), 40)
    ref_layers, _ = netdesc.describe_line_net(net)
    assert len(layers) == len(ref_layers) - 1 and ref_layers[-1]['kind'] == _lib.CTC_HEAD
    for a, b in zip(layers, ref_layers):
        assert set(a) == set(b)
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], (k, a[k], b[k])
    assert [(l['pool_h'], l['pool_w']) for l in layers if l['kind'] == _lib.CONV] == \
        [(2, 2), (1, 1), (2, 2), (1, 1), (1, 1), (2, 1), (1, 1), (1, 1), (1, 1)]
    assert decoder['classes'] == cases.AR_CASE['classes'] and len(decoder['layers']) == cases.AR_CASE['decoder_layers']
    for i, ly in enumerate(decoder['layers']):
        assert np.array_equal(ly['cross_in_w'], dec[f'trans_decoder.layers.{i}.multihead_attn.in_proj_weight'])
    desc, keep = netdesc.ar_to_ctypes(decoder)
    assert desc.n_layers == 2 and desc.heads == 8 and desc.dim_ff == 2048 and desc.classes == 32
    _, _, sd_wide = cases.ar_state_dict(cases.AR_CASES['wide'])
    _, dec_wide = netdesc.describe_transformer_ocr(sd_wide, cases.ar_net_config(cases.AR_CASES['wide']), 40)
    assert len(dec_wide['layers']) == 3 and dec_wide['classes'] == 122
    bad = dict(sd)
    bad.pop('trans_decoder.layers.1.norm3.bias')
    with pytest.raises(KeyError):
        netdesc.describe_transformer_ocr(bad, cases.AR_NET_CONFIG, 40)
    with pytest.raises(ValueError):
        netdesc.describe_transformer_ocr(sd, dict(cases.AR_NET_CONFIG, conv_subsampling=[4, 4]), 40)
    with pytest.raises(ValueError):
        netdesc.describe_transformer_ocr(sd, dict(cases.AR_NET_CONFIG, decoder_layers=1), 40)
