"""TEST INFRASTRUCTURE ONLY -- CPU restatement of pero-ocr's per-character confidence.

Follows get_line_confidence (pero_ocr/core/confidence_estimation.py:73-104).  Pinned against outputs of the unmodified
reference function on the seeded cases of oracle/align_oracle.py (tests/golden/align.npz, keys conf_*; written by
oracle/make_golden.py: golden_align)."""
import numpy as np

from .align_oracle import align_text


def line_confidence(log_probs, labels, aligned_letters=None):
    """log_probs [T, C] (blank last); labels: int sequence.  -> float64 [len(labels)]."""
    log_probs = np.asarray(log_probs)
    if aligned_letters is None:
        aligned_letters = align_text(-log_probs, np.asarray(labels), log_probs.shape[1] - 1)
    frames = np.concatenate([aligned_letters, [1000]])
    probs = np.exp(log_probs)
    out = np.zeros(len(labels))
    border = 0
    for i, label in enumerate(labels):
        own = probs[frames[i], label]
        nxt = (frames[i] + 1 + frames[i + 1]) // 2
        window = np.copy(probs[border:nxt])
        window[:, label] = 0
        if i > 0:
            window[:, labels[i - 1]] = 0
        if i + 1 < len(labels):
            window[:, labels[i + 1]] = 0
        out[i] = max(0, own - window[:, :-1].max())
        border = nxt
    return out
