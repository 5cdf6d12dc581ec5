"""GPU parity of the line cropper's device resampling (b200ocr_remap_lines) against crops produced by the unmodified
reference EngineLineCropper (tests/golden/cropper.npz) and against the oracle's restatement of cv2.remap on seeded
maps.  Bar: byte-exact."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.crop_oracle import CROP_CASES, page_image, remap_bilinear_u8
from tests.util import load_golden, make_case_net, write_engine_json

pytestmark = pytest.mark.gpu


def test_crops_match_reference_golden(golden_dir):
    from pero_ocr_b200.cropper import B200LineCropper
    gold = load_golden(golden_dir, 'cropper.npz')
    img = page_image()
    for name, kw, baseline, heights in CROP_CASES:
        crop = B200LineCropper(**kw).crop(img, baseline, heights)
        want = gold[f'crop_{name}']
        assert crop.dtype == np.uint8 and crop.shape == want.shape, name
        assert np.array_equal(crop, want), name


def test_whole_page_launch_matches_oracle_on_random_maps():
    """One launch, ragged widths, maps that leave the page on every side, exact-pixel / half-step / NaN / huge
    coordinates; written into a padded batch like the recogniser's (pad 32, zero fill)."""
    from pero_ocr_b200.cropper import DevicePage, remap_into
    rng = np.random.default_rng(17)
    img = rng.integers(0, 256, (211, 333, 3), dtype=np.uint8)
    widths = [1, 37, 256, 300, 0, 129]
    maps = []
    for w in widths:
        m = np.stack([rng.random((40, w)) * 360 - 14, rng.random((40, w)) * 240 - 14], axis=2).astype(np.float32)
        maps.append(m)
    maps[2][0, :8, 0] = [0.0, 332.0, 331.999, -1.0, 64.5, 64.015625, np.nan, 3e9]
    maps[2][0, :8, 1] = [0.0, 210.0, 209.5, 3.0, 0.5, 7.984375, 5.0, 5.0]
    page = DevicePage(img)
    out = torch.full((len(maps), 40, 320, 3), 0xAB, dtype=torch.uint8, device='cuda')   # stale bytes must vanish
    remap_into(page, maps, out, 32)
    got = out.cpu().numpy()
    for i, (w, m) in enumerate(zip(widths, maps)):
        keep = min(w, 320 - 32)
        want = np.zeros((40, 320, 3), dtype=np.uint8)
        if keep:
            want[:, 32:32 + keep] = remap_bilinear_u8(img, m)[:, :keep]
        assert np.array_equal(got[i], want), i


def test_process_line_maps_equals_process_lines_on_reference_crops(tmp_path, golden_dir):
    """Cropping on the device straight into the recogniser batch gives exactly what the reference pipeline gives:
    LineCropper.process_page crops (golden) -> process_lines."""
    from pero_ocr_b200.cropper import B200LineCropper, DevicePage
    from pero_ocr_b200.engine import B200EngineLineOCR
    gold = load_golden(golden_dir, 'cropper.npz')
    js = write_engine_json(tmp_path, 'lstm')
    eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=4, module=make_case_net('lstm'))
    use = [c for c in CROP_CASES if c[1]['line_height'] == 40 and c[0] != 'degenerate_single_point']
    cropper = B200LineCropper(line_height=40)
    maps = []
    for name, kw, baseline, heights in use:
        c = B200LineCropper(**kw)
        maps.append(c.get_crop_inputs(baseline, heights, 40))
    page = DevicePage(page_image())
    tr_a, lg_a, co_a = eng.process_line_maps(page, maps, sparse_logits=False)
    tr_b, lg_b, co_b = eng.process_lines([gold[f'crop_{n}'] for n, *_ in use], sparse_logits=False)
    assert tr_a == tr_b and co_a == co_b
    for a, b in zip(lg_a, lg_b):
        assert np.array_equal(a, b)                    # identical input bytes -> identical logits


def test_device_geometry_crops_match_reference_golden(golden_dir):
    """b200ocr_remap_poly_lines: the sampling maps are evaluated on the device from the fitted polynomial (float64, in
    NumPy's order of operations) -- the crops must still be the reference's, byte for byte."""
    from pero_ocr_b200.cropper import B200LineCropper
    gold = load_golden(golden_dir, 'cropper.npz')
    img = page_image()
    seen = 0
    for name, kw, baseline, heights in CROP_CASES:
        if not kw['poly']:
            continue
        crop = B200LineCropper(**kw).crop_page_poly(img, [(baseline, heights)])[0]
        want = gold[f'crop_{name}']
        assert crop.shape == want.shape and np.array_equal(crop, want), name
        seen += 1
    assert seen >= 5


def test_process_baselines_equals_process_lines_on_reference_crops(tmp_path, golden_dir):
    from pero_ocr_b200.cropper import B200LineCropper, DevicePage
    from pero_ocr_b200.engine import B200EngineLineOCR
    gold = load_golden(golden_dir, 'cropper.npz')
    js = write_engine_json(tmp_path, 'lstm')
    eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=4, module=make_case_net('lstm'))
    use = [c for c in CROP_CASES if c[1] == dict(line_height=40, poly=2, scale=1)]
    assert len(use) >= 3                      # includes the degenerate baseline -> the reference's zero crop
    cropper = B200LineCropper(line_height=40, poly=2, scale=1)
    page = DevicePage(page_image())
    tr_a, lg_a, co_a = eng.process_baselines(page, [(b, h) for _, _, b, h in use], cropper, sparse_logits=False)
    tr_b, lg_b, co_b = eng.process_lines([gold[f'crop_{n}'] for n, *_ in use], sparse_logits=False)
    assert tr_a == tr_b and co_a == co_b
    for a, b in zip(lg_a, lg_b):
        assert np.array_equal(a, b)


def test_pad_lines_builds_the_reference_batch():
    """b200ocr_pad_lines against the reference's own padding / stacking (line_ocr_engine.py:121-127): ragged widths,
    a 1-px line, a line wider than the batch (cut on the right), through the engine's packed staging."""
    import ctypes as C
    from pero_ocr_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(21)
    widths = [200, 1, 97, 264, 33, 300]
    out_w, pad, height = 264, 32, 40
    lines = [rng.integers(0, 256, (height, w, 3), dtype=np.uint8) for w in widths]
    want = np.zeros((len(lines), height, out_w, 3), dtype=np.uint8)
    for i, l in enumerate(lines):
        w = min(l.shape[1], out_w - pad)
        want[i, :, pad:pad + w] = l[:, :w]
    offs = np.zeros(len(lines) + 1, dtype=np.int64)
    np.cumsum([l.nbytes for l in lines], out=offs[1:])
    packed = torch.from_numpy(np.concatenate([l.reshape(-1) for l in lines])).cuda()
    d_off = torch.from_numpy(offs[:-1].copy()).cuda()
    d_w = torch.tensor(widths, dtype=torch.int32, device='cuda')
    out = torch.full((len(lines), height, out_w, 3), 0xAB, dtype=torch.uint8, device='cuda')
    _lib.check(lib.b200ocr_pad_lines(packed.data_ptr(), d_off.data_ptr(), d_w.data_ptr(), len(lines), height,
                                     out.data_ptr(), out_w, pad, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)


def test_process_pages_pipeline_equals_page_by_page(tmp_path):
    """process_pages (upload / baseline fits / ParseNet of the following pages overlapped with the recognition of the
    current one) returns, page for page, exactly what DevicePage + process_baselines + get_maps return one at a time."""
    from oracle.nets import make_net
    from pero_ocr_b200.cropper import B200LineCropper, DevicePage
    from pero_ocr_b200.engine import B200EngineLineOCR
    from pero_ocr_b200.parsenet import B200ParseNet
    js = write_engine_json(tmp_path, 'lstm')
    eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=8, module=make_case_net('lstm'))
    pn = B200ParseNet(None, torch.device('cuda', 0), downsample=2, adaptive_downsample=False,
                      module=make_net('parsenet', seed=cases.PARSENET_CASE['seed']))
    cropper = B200LineCropper(line_height=40, poly=2, scale=1)
    rng = np.random.default_rng(31)
    pages = []
    for p in range(5):
        img = rng.integers(0, 256, (300 + 40 * p, 520, 3), dtype=np.uint8)
        lines = [([[20, 40 + 50 * i], [250, 44 + 50 * i + p], [500 - 30 * i, 40 + 50 * i]], [22, 10]) for i in range(3 + p % 2)]
        pages.append((img, lines))
    got = list(eng.process_pages(iter(pages), cropper, parsenet=pn, parsenet_downsample=2, prefetch=2, sparse_logits=False))
    assert len(got) == len(pages)
    for (img, lines), (tr, lg, co, maps) in zip(pages, got):
        tr_b, lg_b, co_b = eng.process_baselines(DevicePage(img), lines, cropper, sparse_logits=False)
        assert tr == tr_b and co == co_b
        for a, b in zip(lg, lg_b):
            assert np.array_equal(a, b)
        assert np.array_equal(maps, pn.get_maps(img, 2))
