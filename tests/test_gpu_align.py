"""GPU parity of the CTC forced alignment (b200ocr_force_align behind pero_ocr_b200.force_alignment) against the
reference's known-answer tests, outputs of the unmodified reference (tests/golden/align.npz) and the oracle on
seeded batches.  Bar: identical integer paths (the costs are float64 sums in the same order, ties by the same rule)."""
import numpy as np
import pytest

from oracle import cases
from oracle.align_oracle import align_cases
from oracle.align_oracle import align_text as oracle_align_text
from oracle.align_oracle import force_align as oracle_force_align
from tests.align_kats import KATS
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('kat', KATS, ids=lambda k: k[0])
def test_reference_known_answers(kat, dtype):
    from pero_ocr_b200.force_alignment import force_align
    name, neg, text, blank, want = kat
    neg = np.asarray(neg, dtype=dtype)
    if want == 'ValueError':
        with pytest.raises(ValueError):
            force_align(neg, text, blank)
    else:
        assert force_align(neg, text, blank) == want


def test_matches_reference_golden(golden_dir):
    from pero_ocr_b200.force_alignment import align_text, force_align
    gold = load_golden(golden_dir, 'align.npz')
    for name, neg, labels, blank in align_cases():
        assert force_align(neg, labels, blank) == list(gold[f'sym_{name}']), name
        assert force_align(neg, labels, blank, return_seq_positions=True) == list(gold[f'pos_{name}']), name
        got = align_text(neg, np.array(labels), blank)
        assert got.dtype == np.int32 and np.array_equal(got, gold[f'chr_{name}']), name


def test_batch_of_ragged_lines_matches_oracle():
    """One launch: 24 lines of the recogniser's frame count (T = 336, C = 120), ragged transcriptions and frame
    counts, one line that cannot be aligned (text longer than its frames) and one with an invalid symbol."""
    from pero_ocr_b200.force_alignment import force_align_batch
    rng = np.random.default_rng(33)
    n, t, c = 24, 336, 120
    lp = cases.peaky_logprobs(rng, n, t, c, sharp=10.0).astype(np.float32)
    neg = -lp
    frames = rng.integers(40, t + 1, n)
    frames[0] = t
    labels = []
    for i in range(n):
        best = lp[i, :frames[i]].argmax(axis=1)
        text = [int(v) for k, v in enumerate(best) if v != c - 1 and (k == 0 or best[k - 1] != v)] or [5]
        if i % 5 == 0:
            text = [int(v) for v in rng.integers(0, c - 1, max(1, len(text) // 2))]      # a wrong transcription
        labels.append(text)
    frames[3], labels[3] = 4, [1, 2, 3, 4, 5, 6]                                          # cannot fit: status 1
    labels[7] = [4, c - 1, 9]                                                             # blank inside: status 2
    res = force_align_batch(neg, labels, c - 1, n_frames=frames, want_char_positions=True)
    for i in range(n):
        if i == 3:
            assert res['status'][i] == 1
            continue
        if i == 7:
            assert res['status'][i] == 2
            continue
        assert res['status'][i] == 0
        x = neg[i, :frames[i]]
        assert list(res['symbols'][i, :frames[i]]) == oracle_force_align(x, labels[i], c - 1), i
        assert list(res['positions'][i, :frames[i]]) == oracle_force_align(x, labels[i], c - 1, True), i
        assert (res['symbols'][i, frames[i]:] == -1).all()
        assert np.array_equal(res['char_positions'][i, :len(labels[i])], oracle_align_text(x, labels[i], c - 1)), i


def test_char_confidences_match_reference_golden(golden_dir):
    """b200ocr_char_confidence against get_line_confidence of the unmodified reference.  float32 arithmetic with CUDA's
    expf instead of NumPy's: 2e-6 absolute on probabilities in [0, 1]."""
    from pero_ocr_b200.confidence_estimation import get_line_confidence, line_confidences_batch
    gold = load_golden(golden_dir, 'align.npz')
    import types
    seen = 0
    for name, neg, labels, blank in align_cases():
        if f'conf_{name}' not in gold.files:
            continue
        lp = (-neg).astype(np.float32)
        line = types.SimpleNamespace(logits=np.zeros(lp.shape, dtype=np.float32))
        got = get_line_confidence(line, np.array(labels), log_probs=lp)
        assert got.shape == gold[f'conf_{name}'].shape
        assert np.abs(got - gold[f'conf_{name}']).max() <= 2e-6, name
        # with the alignment handed in (the reference's aligned_letters argument)
        got2 = get_line_confidence(line, np.array(labels), aligned_letters=gold[f'chr_{name}'], log_probs=lp)
        assert np.abs(got2 - gold[f'conf_{name}']).max() <= 2e-6, name
        seen += 1
    assert seen >= 6


def test_char_confidences_batch_matches_oracle():
    from oracle.confidence_oracle import line_confidence
    from pero_ocr_b200.confidence_estimation import line_confidences_batch
    rng = np.random.default_rng(41)
    n, t, c = 12, 200, 60
    lp = cases.peaky_logprobs(rng, n, t, c, sharp=10.0).astype(np.float32)
    frames = rng.integers(60, t + 1, n)
    labels = []
    for i in range(n):
        best = lp[i, :frames[i]].argmax(axis=1)
        labels.append([int(v) for k, v in enumerate(best) if v != c - 1 and (k == 0 or best[k - 1] != v)] or [3])
    conf, status = line_confidences_batch(lp, labels, n_frames=frames)
    assert (status == 0).all()
    for i in range(n):
        want = line_confidence(lp[i, :frames[i]], labels[i])
        assert np.abs(conf[i] - want).max() <= 2e-6, i
