"""GPU parity: CTC greedy and prefix-beam kernels against the reference outputs in tests/golden/decoders.npz,
the reference's own known-answer tests (test/test_decoding/test_decoders.py) and the numpy oracle."""
import numpy as np
import pytest

from oracle import cases
from oracle.decoders_oracle import prefix_beam
from tests.util import load_golden

pytestmark = pytest.mark.gpu

BLANK = '<BLANK>'
LETTERS = ['a', 'b', 'c', BLANK]


@pytest.fixture(scope='module')
def gold(golden_dir):
    return load_golden(golden_dir, 'decoders.npz')


def test_config1_greedy_bit_exact(gold):
    from pero_ocr_b200.decoders import GreedyDecoder, greedy_decode_ctc
    raw, lp, letters = cases.config1_logits()
    bags = GreedyDecoder(letters).decode_batch(lp)
    assert [b.best_hyp() for b in bags] == list(gold['config1_greedy'])
    np.testing.assert_allclose([list(b)[0].vis_sc for b in bags], gold['config1_greedy_score'], rtol=1e-5)
    assert greedy_decode_ctc(raw.transpose(0, 2, 1).copy(), letters[:-1] + ['']) == list(gold['config1_greedy'])


def test_greedy_edge_cases_bit_exact(gold):
    from pero_ocr_b200.decoders import greedy_decode_ctc
    for name, (arr, chars) in cases.greedy_edge_cases().items():
        assert greedy_decode_ctc(arr, chars) == list(gold[f'edge_{name}']), name


def test_greedy_ragged_batch_and_single_frame():
    from pero_ocr_b200.decoders import GreedyDecoder
    lp, letters = cases.peaky_cases()['peaky_small']
    dec = GreedyDecoder(letters)
    ragged = [lp[0][:7], lp[1], lp[2][:1]]
    together = [b.best_hyp() for b in dec.decode_batch(ragged)]
    alone = [dec(m).best_hyp() for m in ragged]
    assert together == alone


def test_config1_prefix_beam_matches_reference(gold):
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder
    _, lp, letters = cases.config1_logits()
    dec = CTCPrefixLogRawNumpyDecoder(letters, k=16)
    bags = dec.decode_batch([lp[i].astype(np.float64) for i in range(cases.CONFIG1_BEAM_LINES)])
    for i, bag in enumerate(bags):
        assert bag.best_hyp() == str(gold[f'config1_beam_best_{i}'])
        assert sorted(h.transcript for h in bag) == sorted(gold[f'config1_beam_hyps_{i}'])
        np.testing.assert_allclose(sorted(h.vis_sc for h in bag), sorted(gold[f'config1_beam_scores_{i}']), rtol=1e-9)


@pytest.mark.parametrize('name', ['peaky_small', 'peaky_big'])
@pytest.mark.parametrize('k', [1, 4, 16])
def test_peaky_prefix_beam_matches_reference(gold, name, k):
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder
    lp, letters = cases.peaky_cases()[name]
    bags = CTCPrefixLogRawNumpyDecoder(letters, k=k).decode_batch(list(lp))
    for i, bag in enumerate(bags):
        assert bag.best_hyp() == str(gold[f'{name}_k{k}_best_{i}'])
        assert sorted(h.transcript for h in bag) == sorted(gold[f'{name}_k{k}_hyps_{i}'])
        np.testing.assert_allclose(sorted(h.vis_sc for h in bag), sorted(gold[f'{name}_k{k}_scores_{i}']), rtol=1e-9)


def _lp(rows):
    return np.log(np.asarray(rows, dtype=np.float64))


def _decoders():
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder, GreedyDecoder
    return [GreedyDecoder(LETTERS), CTCPrefixLogRawNumpyDecoder(LETTERS, 1), CTCPrefixLogRawNumpyDecoder(LETTERS, 2)]


def test_reference_known_answer_tests():
    """test/test_decoding/test_decoders.py:23-104, run on both GPU decoders."""
    for dec in _decoders():
        assert dec(_lp([[0.8, 0.1, 0.05, 0.05]])).best_hyp() == 'a'
        assert dec(_lp([[0.8, 0.1, 0.05, 0.05], [0.1, 0.8, 0.05, 0.05], [0.05, 0.1, 0.8, 0.05]])).best_hyp() == 'abc'
        assert dec(_lp([[0.8, 0.1, 0.05, 0.05], [0.8, 0.1, 0.05, 0.05]])).best_hyp() == 'a'
        assert dec(_lp([[0.9, 0.03, 0.03, 0.04], [0.03, 0.03, 0.04, 0.9], [0.9, 0.03, 0.03, 0.04]])).best_hyp() == 'aa'
        assert dec(_lp([[0.9, 0.03, 0.03, 0.04], [0.03, 0.9, 0.03, 0.04]])).best_hyp() == 'ab'
        assert dec(np.log(np.asarray([[1e-3, 1e-3, 1e-3, 1 - 3e-3]]))).best_hyp() == ''
        with pytest.raises(ValueError):
            dec(np.asarray([[-80.0, -80.0, -80.0, -5.0]]))            # not normalised (:97-104)


def test_prefix_joining_and_wide_beam_regressions():
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder
    bag = CTCPrefixLogRawNumpyDecoder(LETTERS, 2)(_lp([[0.5, 1e-30, 1e-30, 0.5], [0.5, 1e-30, 1e-30, 0.5]]))   # :106-121
    d = {h.transcript: h.vis_sc for h in bag}
    assert sorted(d) == ['', 'a']
    assert d['a'] == pytest.approx(np.log(0.75), abs=1e-9) and d[''] == pytest.approx(np.log(0.25), abs=1e-9)
    bag = CTCPrefixLogRawNumpyDecoder(LETTERS, 50)(_lp([[0.4, 0.3, 0.2, 0.1], [0.4, 0.3, 0.2, 0.1], [0.1, 0.2, 0.3, 0.4]]))
    texts = [h.transcript for h in bag]                                                                    # :448-462
    assert len(texts) == len(set(texts))
    assert np.logaddexp.reduce([h.vis_sc for h in bag]) == pytest.approx(0.0, abs=1e-9)


def test_constructor_validation():
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder, GreedyDecoder
    with pytest.raises(ValueError):
        GreedyDecoder(['a', 'a', BLANK])
    with pytest.raises(ValueError):
        GreedyDecoder(['a', BLANK, 'b'])
    with pytest.raises(ValueError):
        CTCPrefixLogRawNumpyDecoder(['a', 'b'], 1)
    with pytest.raises(TypeError):
        CTCPrefixLogRawNumpyDecoder(LETTERS, None)
    with pytest.raises(ValueError):
        CTCPrefixLogRawNumpyDecoder(LETTERS, 0)


def test_prefix_beam_full_size_vs_oracle_and_mass():
    """Config-3-sized input (T=336, C=120, k=16) on peaky data: identical to the numpy oracle; and with an
    effectively unbounded beam on a tiny alphabet the hypothesis mass sums to one."""
    from pero_ocr_b200.decoders import CTCPrefixLogRawNumpyDecoder
    rng = np.random.default_rng(9)
    letters = [chr(0x100 + i) for i in range(119)] + [BLANK]
    lp = cases.peaky_logprobs(rng, 3, 336, 120, sharp=11.0)
    bags = CTCPrefixLogRawNumpyDecoder(letters, 16).decode_batch(list(lp))
    for m, bag in zip(lp, bags):
        want = prefix_beam(m, 16)
        assert [h.transcript for h in bag][0] == ''.join(letters[i] for i in want[0][0])
        np.testing.assert_allclose(sorted(h.vis_sc for h in bag), sorted(s for _, s in want), rtol=1e-9)
    small = cases.peaky_logprobs(rng, 1, 5, 4, sharp=2.0)[0]
    bag = CTCPrefixLogRawNumpyDecoder(LETTERS, 64)(small)
    want = prefix_beam(small, 64)
    np.testing.assert_allclose(sorted(h.vis_sc for h in bag), sorted(s for _, s in want), rtol=1e-9)


@pytest.mark.parametrize('kind', ['transformer', 'lstm'])
def test_recognise_and_beam_decode_on_device_matches_reference_chain(tmp_path, kind):
    """BASELINE config 3 chain (Transformer-encoder variant + CTC prefix beam, k = 16), device-resident:
    B200EngineLineOCR.decode_lines against the reference chain restated by the oracles -- process_lines (sparse
    logits) -> get_full_logprobs -> [logit_coords] slice -> CTCPrefixLogRawNumpyDecoder(k) (page_parser.py:108-142,
    418-430).  Bar: same best hypothesis, vis_sc within rtol 1e-4."""
    import torch
    from oracle.forward_oracle import OracleEngine, full_logprobs
    from pero_ocr_b200.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder
    from pero_ocr_b200.engine import B200EngineLineOCR
    from tests.util import make_case_net, write_engine_json
    spec = cases.ENGINE_CASES[kind]
    net = make_case_net(kind)
    eng = B200EngineLineOCR(write_engine_json(tmp_path, kind), torch.device('cuda', 0), batch_size=4, module=net)
    letters = eng.characters + [BLANK_SYMBOL]
    lines = cases.engine_lines(kind)
    bags = eng.decode_lines([l.copy() for l in lines], CTCPrefixLogRawNumpyDecoder(letters, 16))
    ref = OracleEngine(dict(net.state_dict()), cases.json_characters(spec['classes'] - 2), kind=kind)
    ref.model = lambda x: net(x)
    ref.batch_size = 4
    ref.max_input_horizontal_pixels = 480 * 4
    with torch.no_grad():
        _, sparse_logits, coords = ref.process_lines([l.copy() for l in lines], sparse_logits=True)
    for i in range(len(lines)):
        lp = full_logprobs(sparse_logits[i])[coords[i][0]:coords[i][1]]
        want = prefix_beam(lp.astype(np.float64), 16)
        best = max(want, key=lambda h: h[1])
        got = max(bags[i], key=lambda h: h.vis_sc)
        assert got.transcript == ''.join(letters[j] for j in best[0]), i
        assert abs(got.vis_sc - best[1]) <= 1e-4 * max(1.0, abs(best[1])), i
