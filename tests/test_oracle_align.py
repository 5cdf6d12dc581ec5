"""CPU: oracle/align_oracle.py (restatement of pero_ocr/core/force_alignment.py) against the reference's own
known-answer tests and against outputs of the unmodified reference on seeded cases (tests/golden/align.npz)."""
import numpy as np
import pytest

from oracle.align_oracle import align_cases, align_text, force_align
from tests.align_kats import KATS
from tests.util import load_golden


@pytest.mark.parametrize('kat', KATS, ids=lambda k: k[0])
def test_reference_known_answers(kat):
    name, neg, text, blank, want = kat
    neg = np.asarray(neg, dtype=np.float64)
    if want == 'ValueError':
        with pytest.raises(ValueError):
            force_align(neg, text, blank)
    else:
        assert force_align(neg, text, blank) == want


def test_oracle_matches_reference_golden(golden_dir):
    gold = load_golden(golden_dir, 'align.npz')
    for name, neg, labels, blank in align_cases():
        assert force_align(neg, labels, blank) == list(gold[f'sym_{name}']), name
        assert force_align(neg, labels, blank, return_seq_positions=True) == list(gold[f'pos_{name}']), name
        assert np.array_equal(align_text(neg, np.array(labels), blank), gold[f'chr_{name}']), name


def test_char_confidence_oracle_matches_reference_golden(golden_dir):
    """oracle/confidence_oracle.py restates get_line_confidence (confidence_estimation.py:73-104)."""
    from oracle.confidence_oracle import line_confidence
    gold = load_golden(golden_dir, 'align.npz')
    seen = 0
    for name, neg, labels, blank in align_cases():
        if f'conf_{name}' not in gold.files:
            continue
        got = line_confidence((-neg).astype(np.float32), labels)
        assert np.array_equal(got, gold[f'conf_{name}']), name
        seen += 1
    assert seen >= 6
