"""CPU, build container only: the host mirror of the engine dropped into the UNMODIFIED reference's PageParser at the
B3 seam of SURVEY.md 8(b) (`page_parser.ocr.ocr_engine`, page_parser.py:418-430), next to the reference's own
PytorchEngineLineOCR hosting the same weights.  The device step is played by the torch-CPU oracle (there is no GPU
here), so this pins the INTERFACE: what PageOCR.process_page, TextLine.get_dense_logits / get_full_logprobs and
PageParser.update_confidences read from the engine's results.  Skipped where /root/reference does not exist (the GPU
box); the GPU parity tests cover the device step against the committed golden vectors instead."""
import configparser
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import cases
from tests.test_host_logic import _host_only_engine
from tests.util import make_case_net

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'pero_ocr')),
                                reason='the reference tree exists only in the build container')


def _page_layout(PageLayout, RegionLayout, TextLine, shape):
    layout = PageLayout(id='p', page_size=shape[:2])
    region = RegionLayout('r', np.array([[0, 0], [shape[1], 0], [shape[1], shape[0]], [0, shape[0]]], dtype=float))
    for i, (x0, x1, y, dy) in enumerate([(40, 1350, 80, 4), (60, 700, 160, -3), (30, 400, 240, 0), (100, 1200, 330, 7)]):
        base = np.array([[x0, y], [(x0 + x1) / 2, y + dy], [x1, y]], dtype=float)
        region.lines.append(TextLine(id=f'l{i}', baseline=base, heights=[24.0, 10.0]))
    layout.regions.append(region)
    return layout


def test_engine_object_drops_into_the_unmodified_page_parser(tmp_path):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from oracle.make_golden import _export_engine, _install_stubs
    _install_stubs()
    from pero_ocr.core.layout import PageLayout, RegionLayout, TextLine
    from pero_ocr.document_ocr.page_parser import PageParser
    spec = cases.ENGINE_CASES['lstm']
    _export_engine(str(tmp_path), make_case_net('lstm'), 'ocr', spec['classes'] - 2)
    cfg = configparser.ConfigParser()
    cfg.read_dict({'PAGE_PARSER': {'RUN_LAYOUT_PARSER': 'no', 'RUN_LINE_CROPPER': 'yes', 'RUN_OCR': 'yes',
                                   'RUN_DECODER': 'no'},
                   'LINE_CROPPER': {'INTERP': '2', 'LINE_SCALE': '1', 'LINE_HEIGHT': '40'},
                   'OCR': {'OCR_JSON': 'ocr.json'}})
    with contextlib.redirect_stdout(io.StringIO()):
        parser = PageParser(cfg, device=torch.device('cpu'), config_path=str(tmp_path))
    img = np.random.default_rng(12).integers(0, 256, (420, 1400, 3), dtype=np.uint8)
    with contextlib.redirect_stdout(io.StringIO()):
        want = parser.process_page(img, _page_layout(PageLayout, RegionLayout, TextLine, img.shape))
        # --- the seam: only the engine object changes; budget as the reference engine object's
        ours = _host_only_engine('lstm')
        ours.batch_size = parser.ocr.ocr_engine.batch_size
        ours.max_input_horizontal_pixels = parser.ocr.ocr_engine.max_input_horizontal_pixels
        assert list(ours.characters) == list(parser.ocr.ocr_engine.characters)
        parser.ocr.ocr_engine = ours
        got = parser.process_page(img, _page_layout(PageLayout, RegionLayout, TextLine, img.shape))
    lines_a, lines_b = list(want.lines_iterator()), list(got.lines_iterator())
    assert len(lines_a) == len(lines_b) == 4
    for a, b in zip(lines_a, lines_b):
        assert a.transcription == b.transcription
        assert list(a.logit_coords) == list(b.logit_coords)
        assert list(a.characters) == list(b.characters)
        assert a.logits.shape == b.logits.shape and type(a.logits) is type(b.logits)
        assert np.array_equal(a.logits.indptr, b.logits.indptr) and np.array_equal(a.logits.indices, b.logits.indices)
        np.testing.assert_allclose(a.logits.data, b.logits.data, atol=2e-5)
        np.testing.assert_allclose(a.get_full_logprobs(), b.get_full_logprobs(), atol=2e-5)
        assert abs(a.transcription_confidence - b.transcription_confidence) <= 1e-5


def test_parsenet_host_logic_matches_the_unmodified_torch_parsenet(tmp_path):
    """B5 seam: B200ParseNet's host side -- INTER_AREA downscale, x64 canvas, crop back, the adaptive
    second pass and its `last_downsample` state (torch_parsenet.py:37-93) -- against the unmodified TorchParseNet, both
    around the same deterministic stand-in for the conv forward."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from pero_ocr.layout_engines.torch_parsenet import TorchParseNet
    from pero_ocr_b200.parsenet import B200ParseNet
    ck = os.path.join(str(tmp_path), 'pn.pt')
    torch.jit.script(make_case_net('lstm').agg_act).save(ck + '.cpu')          # any loadable blob: replaced below

    def fake(x):                       # f32 [1,3,H,W] in [0,1] -> ([1,5,H,W], aux); line height 30 px, lines where bright
        g = x.mean(dim=1, keepdim=True)
        return torch.cat([g * 0 + 30.0, g * 0 + 5.0, (g > 0.45).float(), g, 1 - g], dim=1), None

    ref = TorchParseNet(ck, torch.device('cpu'), downsample=2, max_mp=5, detection_threshold=0.2)
    ref.net = fake
    ours = B200ParseNet.__new__(B200ParseNet)                                  # host logic only: no device here
    for name in ('max_megapixels', 'detection_threshold', 'adaptive_downsample', 'init_downsample', 'last_downsample',
                 'downsample_line_pixel_adapt_threshold', 'min_line_processing_height', 'max_line_processing_height',
                 'optimal_line_processing_height', 'min_downsample', 'max_downsample'):
        setattr(ours, name, getattr(ref, name))
    ours.net = lambda canvas: fake(torch.from_numpy(canvas).float().permute(0, 3, 1, 2) * (1 / 255.))[0]
    rng = np.random.default_rng(3)
    for shape in ((333, 517, 3), (640, 256, 3), (100, 90, 3)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        with contextlib.redirect_stdout(io.StringIO()):
            want_map, want_ds = ref.get_maps_with_optimal_resolution(img)
            plain = ref.get_maps(img, 3)
        got_map, got_ds = ours.get_maps_with_optimal_resolution(img)
        assert got_ds == want_ds and ours.last_downsample == ref.last_downsample
        assert got_map.shape == want_map.shape and np.array_equal(got_map, want_map)
        assert np.array_equal(ours.get_maps(img, 3), plain)
        assert ours.get_med_height(got_map) == ref.get_med_height(want_map)
    assert ref.last_downsample != 2                        # the adaptive second pass did run and moved the state
