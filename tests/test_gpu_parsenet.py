"""GPU parity: ParseNet conv forward behind the TorchParseNet interface vs the reference's get_maps output."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.nets import make_net
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def test_get_maps_matches_reference_golden(golden_dir):
    from pero_ocr_b200.parsenet import B200ParseNet
    spec = cases.PARSENET_CASE
    net = make_net('parsenet', seed=spec['seed'])
    pn = B200ParseNet(None, torch.device('cuda', 0), downsample=spec['downsample'], adaptive_downsample=False,
                      module=net)
    img = cases.parsenet_image()
    maps = pn.get_maps(img, spec['downsample'])
    gold = load_golden(golden_dir, 'parsenet.npz')['maps']
    assert maps.shape == gold.shape and maps.dtype == np.float32
    assert np.abs(maps - gold).max() <= 1e-3
    maps2, ds = pn.get_maps_with_optimal_resolution(img)
    assert ds == spec['downsample'] and np.array_equal(maps, maps2)


def test_page_size_maps_and_adaptive_second_pass(golden_dir):
    """BASELINE config 4's size: a 3000 x 4000 page at DOWNSAMPLE 4 (768 x 1024 canvas) and the adaptive second pass
    (torch_parsenet.py:60-93) against the unmodified TorchParseNet.get_maps / get_maps_with_optimal_resolution."""
    from pero_ocr_b200.parsenet import B200ParseNet
    spec = cases.PARSENET_PAGE_CASE
    gold = load_golden(golden_dir, 'parsenet_page.npz')
    pn = B200ParseNet(None, torch.device('cuda', 0), downsample=spec['downsample'], adaptive_downsample=True,
                      module=cases.parsenet_page_net())
    img = cases.parsenet_image(spec)
    st = spec['stride']
    first = pn.get_maps(img, spec['downsample'])
    assert list(first.shape) == list(gold['first_shape'])
    assert np.abs(first[::st, ::st] - gold['first']).max() <= 1e-3 * max(1.0, float(np.abs(gold['first']).max()) / 20)
    assert pn.get_med_height(first) == pytest.approx(float(gold['med_height']), rel=1e-4)
    maps, used = pn.get_maps_with_optimal_resolution(img)
    assert used == pytest.approx(float(gold['used_downsample']), rel=1e-4)
    assert pn.last_downsample == pytest.approx(float(gold['last_downsample']), rel=1e-4)
    assert list(maps.shape) == list(gold['maps_shape'])
    # the second pass resizes by 1 / used: a last-digit difference of the median moves a few INTER_AREA weights
    assert np.abs(maps[::st, ::st] - gold['maps']).max() <= 5e-3
