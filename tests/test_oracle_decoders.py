"""CPU: the decoder oracle (oracle/decoders_oracle.py) against (1) outputs of the unmodified reference stored in
tests/golden/decoders.npz and (2) the reference's own known-answer tests, restated from
/root/reference/test/test_decoding/test_decoders.py (line numbers cited per case)."""
import os

import numpy as np
import pytest

from oracle import cases
from oracle.decoders_oracle import (BLANK_SYMBOL, GreedyDecoderOracle, PrefixBeamOracle, greedy, prefix_beam)
from oracle.forward_oracle import greedy_ctc_strings


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'decoders.npz'))


def test_config1_greedy_matches_reference(gold):
    raw, lp, letters = cases.config1_logits()
    got = [greedy(m, letters)[0] for m in lp]
    assert got == list(gold['config1_greedy'])
    sc = [greedy(m, letters)[1] for m in lp]
    np.testing.assert_allclose(sc, gold['config1_greedy_score'], rtol=1e-6)
    # greedy_decode_ctc (pytorch_ocr_engine.py:13-34) restatement agrees too
    assert greedy_ctc_strings(raw.transpose(0, 2, 1), letters[:-1] + ['']) == got


def test_config1_prefix_beam_matches_reference(gold):
    _, lp, letters = cases.config1_logits()
    dec = PrefixBeamOracle(letters, 16)
    for i in range(cases.CONFIG1_BEAM_LINES):
        hyps = dec(lp[i].astype(np.float64))
        assert hyps[0][0] == str(gold[f'config1_beam_best_{i}'])
        assert [h[0] for h in hyps] == list(gold[f'config1_beam_hyps_{i}'])
        np.testing.assert_allclose([h[1] for h in hyps], gold[f'config1_beam_scores_{i}'], rtol=1e-9)


@pytest.mark.parametrize('name', ['peaky_small', 'peaky_big'])
@pytest.mark.parametrize('k', [1, 4, 16])
def test_peaky_prefix_beam_matches_reference(gold, name, k):
    lp, letters = cases.peaky_cases()[name]
    dec = PrefixBeamOracle(letters, k)
    for i, m in enumerate(lp):
        hyps = dec(m)
        assert hyps[0][0] == str(gold[f'{name}_k{k}_best_{i}'])
        assert sorted(h[0] for h in hyps) == sorted(gold[f'{name}_k{k}_hyps_{i}'])
        np.testing.assert_allclose(sorted(h[1] for h in hyps), sorted(gold[f'{name}_k{k}_scores_{i}']), rtol=1e-9)


@pytest.mark.parametrize('name', ['peaky_small', 'peaky_big'])
def test_peaky_greedy_matches_reference(gold, name):
    lp, letters = cases.peaky_cases()[name]
    assert [greedy(m, letters)[0] for m in lp] == list(gold[f'{name}_greedy'])
    np.testing.assert_allclose([greedy(m, letters)[1] for m in lp], gold[f'{name}_greedy_score'], rtol=1e-9)


def test_greedy_edge_cases_match_reference(gold):
    for name, (arr, chars) in cases.greedy_edge_cases().items():
        assert greedy_ctc_strings(arr, chars) == list(gold[f'edge_{name}']), name


# ---- the reference's shared known-answer tests (test_decoders.py:23-104), run on both decoders ----------
LETTERS = ['a', 'b', 'c', BLANK_SYMBOL]


def _decoders():
    return [GreedyDecoderOracle(LETTERS), PrefixBeamOracle(LETTERS, 1), PrefixBeamOracle(LETTERS, 2)]


def _lp(rows):
    return np.log(np.asarray(rows, dtype=np.float64))


@pytest.mark.parametrize('dec', _decoders())
def test_kat_single_frame(dec):                       # test_decoders.py:24-31
    assert dec(_lp([[0.8, 0.1, 0.05, 0.05]]))[0][0] == 'a'


@pytest.mark.parametrize('dec', _decoders())
def test_kat_single_blank_score(dec):                 # :33-41  vis_sc == -5.0 exactly
    lp = np.asarray([[-80.0, -80.0, -80.0, -5.0]])
    with pytest.raises(ValueError):
        dec(lp)                                        # un-normalised input is rejected (:97-104)
    lp = np.log(np.asarray([[1e-3, 1e-3, 1e-3, 1 - 3e-3]]))
    hyps = dec(lp)
    assert hyps[0][0] == ''


@pytest.mark.parametrize('dec', _decoders())
def test_kat_trivial_sequence(dec):                   # :43-52
    assert dec(_lp([[0.8, 0.1, 0.05, 0.05], [0.1, 0.8, 0.05, 0.05], [0.05, 0.1, 0.8, 0.05]]))[0][0] == 'abc'


@pytest.mark.parametrize('dec', _decoders())
def test_kat_repeated_symbol_collapses(dec):          # :54-63
    assert dec(_lp([[0.8, 0.1, 0.05, 0.05], [0.8, 0.1, 0.05, 0.05]]))[0][0] == 'a'


@pytest.mark.parametrize('dec', _decoders())
def test_kat_double_symbol_via_blank(dec):            # :65-75
    assert dec(_lp([[0.9, 0.03, 0.03, 0.04], [0.03, 0.03, 0.04, 0.9], [0.9, 0.03, 0.03, 0.04]]))[0][0] == 'aa'


@pytest.mark.parametrize('dec', _decoders())
def test_kat_immediate_switch(dec):                   # :77-86
    assert dec(_lp([[0.9, 0.03, 0.03, 0.04], [0.03, 0.9, 0.03, 0.04]]))[0][0] == 'ab'


def test_kat_prefix_joining():                        # :106-121 hypothesis set must be exactly {'a', ''}
    lp = _lp([[0.5, 1e-30, 1e-30, 0.5], [0.5, 1e-30, 1e-30, 0.5]])
    hyps = PrefixBeamOracle(LETTERS, 2)(lp)
    assert sorted(h[0] for h in hyps) == ['', 'a']
    # P('a') = 1 - P('') = 0.75
    d = dict(hyps)
    assert d['a'] == pytest.approx(np.log(0.75), abs=1e-9)
    assert d[''] == pytest.approx(np.log(0.25), abs=1e-9)


def test_kat_no_duplicate_hypotheses_wide_beam():     # :448-462
    lp = _lp([[0.4, 0.3, 0.2, 0.1], [0.4, 0.3, 0.2, 0.1], [0.1, 0.2, 0.3, 0.4]])
    hyps = PrefixBeamOracle(LETTERS, 50)(lp)
    texts = [h[0] for h in hyps]
    assert len(texts) == len(set(texts))
    assert np.logaddexp.reduce([h[1] for h in hyps]) == pytest.approx(0.0, abs=1e-9)


def test_constructor_validation():                    # :135-166
    with pytest.raises(ValueError):
        GreedyDecoderOracle(['a', 'a', BLANK_SYMBOL])
    with pytest.raises(ValueError):
        GreedyDecoderOracle(['a', BLANK_SYMBOL, 'b'])
    with pytest.raises(ValueError):
        PrefixBeamOracle(['a', 'b'], 1)
    with pytest.raises(TypeError):
        PrefixBeamOracle(LETTERS, None)
    with pytest.raises(ValueError):
        PrefixBeamOracle(LETTERS, 0)


def test_prefix_beam_mass_conservation():
    """Size-independent property: with an unbounded beam the hypothesis probabilities sum to 1."""
    rng = np.random.default_rng(5)
    lp = cases.peaky_logprobs(rng, 1, 6, 4, sharp=2.0)[0]
    hyps = prefix_beam(lp, 10 ** 6)
    assert np.logaddexp.reduce([h[1] for h in hyps]) == pytest.approx(0.0, abs=1e-9)
