"""CPU: the line cropper's host geometry and the oracle's restatement of cv2.remap, against outputs of the
unmodified reference EngineLineCropper stored in tests/golden/cropper.npz (oracle/make_golden.py: golden_cropper)."""
import hashlib

import numpy as np
import pytest

from oracle.crop_oracle import CROP_CASES, page_image, remap_bilinear_u8
from tests.util import load_golden


def _geometry(kw, baseline, heights):
    from pero_ocr_b200.cropper import B200LineCropper
    c = B200LineCropper(**kw)
    return c.get_crop_inputs(baseline, heights, c.line_height)


@pytest.mark.parametrize('case', [c for c in CROP_CASES if c[0] != 'degenerate_single_point'], ids=lambda c: c[0])
def test_crop_geometry_is_bit_identical_to_reference(golden_dir, case):
    """B200LineCropper.get_crop_inputs restates crop_engine.py:54-99: same float32 map, bit for bit."""
    name, kw, baseline, heights = case
    gold = load_golden(golden_dir, 'cropper.npz')
    m = _geometry(kw, baseline, heights)
    assert m.dtype == np.float32 and list(m.shape) == list(gold[f'mapshape_{name}'])
    assert np.array_equal(m[:, ::4], gold[f'map_{name}'])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(m).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, gold[f'mapsha_{name}'])


@pytest.mark.parametrize('case', [c for c in CROP_CASES if c[0] != 'degenerate_single_point'], ids=lambda c: c[0])
def test_oracle_remap_matches_reference_crops(golden_dir, case):
    """oracle.crop_oracle.remap_bilinear_u8 (OpenCV's 8-bit fixed-point bilinear remap, restated) reproduces the
    reference's crops byte for byte -- including the lines that leave the page (constant border)."""
    name, kw, baseline, heights = case
    gold = load_golden(golden_dir, 'cropper.npz')
    crop = remap_bilinear_u8(page_image(), _geometry(kw, baseline, heights))
    assert crop.dtype == np.uint8 and np.array_equal(crop, gold[f'crop_{name}'])


def test_degenerate_baseline_gives_the_reference_fallback(golden_dir):
    """A baseline without extent: the reference's crop() swallows the failure and returns zeros [H, 32, 3]
    (crop_engine.py:16-22); the geometry itself yields an empty map."""
    name, kw, baseline, heights = [c for c in CROP_CASES if c[0] == 'degenerate_single_point'][0]
    gold = load_golden(golden_dir, 'cropper.npz')
    assert gold[f'crop_{name}'].shape == (40, 32, 3) and not gold[f'crop_{name}'].any()
    try:
        m = _geometry(kw, baseline, heights)
        assert m.shape[1] == 0
    except Exception:
        pass                                             # raising is the other way the reference reaches its fallback


def test_oracle_remap_against_cv2_on_random_maps():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    coords = np.stack([rng.random((40, 300)) * 150 - 10, rng.random((40, 300)) * 110 - 10], axis=2).astype(np.float32)
    coords[0, :6, 0] = [0.0, 129.0, 128.999, -1.0, 64.5, 64.015625]        # exact pixels, last column, half steps
    coords[0, :6, 1] = [0.0, 89.0, 88.5, 3.0, 0.5, 7.984375]
    want = cv2.remap(img, coords[..., 0], coords[..., 1], interpolation=cv2.INTER_LINEAR,
                     borderMode=cv2.BORDER_CONSTANT)
    assert np.array_equal(remap_bilinear_u8(img, coords), want)


@pytest.mark.parametrize('case', [c for c in CROP_CASES if c[1]['poly'] and c[0] != 'degenerate_single_point'],
                         ids=lambda c: c[0])
def test_device_geometry_formulas_reproduce_reference_maps(golden_dir, case):
    """The per-pixel float64 arithmetic of b200ocr_remap_poly_lines (restated in oracle.crop_oracle.poly_map_columns)
    on the parameters B200LineCropper.poly_params hands to the device reproduces the reference's float32 map bit for
    bit -- checked on every 4th column (the columns stored in the fixture) and on the last one."""
    from oracle.crop_oracle import poly_map_columns
    from pero_ocr_b200.cropper import B200LineCropper
    name, kw, baseline, heights = case
    gold = load_golden(golden_dir, 'cropper.npz')
    c = B200LineCropper(**kw)
    line, offsets = c.poly_params(baseline, heights)
    shape = list(gold[f'mapshape_{name}'])
    assert [len(offsets), line.n_out, 2] == shape
    cols = list(range(0, line.n_out, 4))
    got = poly_map_columns(line, offsets, cols[::3] + [cols[-1]])
    want = gold[f'map_{name}'][:, list(range(len(cols)))[::3] + [len(cols) - 1]]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    full = _geometry(kw, baseline, heights)
    last = poly_map_columns(line, offsets, [line.n_out - 1])
    assert np.array_equal(last[:, 0].view(np.uint32), full[:, -1].view(np.uint32))


def test_poly_params_of_a_degenerate_baseline_is_the_zero_crop():
    from pero_ocr_b200.cropper import B200LineCropper
    name, kw, baseline, heights = [c for c in CROP_CASES if c[0] == 'degenerate_single_point'][0]
    line, offsets = B200LineCropper(**kw).poly_params(baseline, heights)
    assert line.ncoef == 0 and line.n_out == 32 and not offsets.any()
