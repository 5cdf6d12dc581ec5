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
