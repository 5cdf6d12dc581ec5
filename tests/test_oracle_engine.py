"""CPU: the forward/engine oracle (oracle/forward_oracle.py) against outputs of the unmodified reference classes
stored in tests/golden/ (PytorchEngineLineOCR.process_lines, TorchParseNet.get_maps, PageParser confidence)."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.forward_oracle import (OracleEngine, dense_logits, full_logprobs, line_confidence,
                                   line_confident_enough, parsenet_forward, sparsify_logits)
from oracle.nets import make_net


def _engine(kind):
    spec = cases.ENGINE_CASES[kind]
    net = make_net(kind, spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'], **spec['net_kw'])
    return OracleEngine(dict(net.state_dict()), cases.json_characters(spec['classes'] - 2), kind=kind,
                        batch_size=spec['engine_batch_size'])


@pytest.mark.parametrize('kind', ['lstm', 'transformer'])
def test_process_lines_matches_reference(golden_dir, kind):
    gold = np.load(os.path.join(golden_dir, f'engine_{kind}.npz'))
    eng = _engine(kind)
    lines = cases.engine_lines(kind)
    tr, lg, co = eng.process_lines(lines, sparse_logits=False)
    assert tr == list(gold['transcriptions'])
    for i in range(len(lines)):
        np.testing.assert_allclose(lg[i], gold[f'logits_{i}'], atol=2e-5)
        assert list(co[i]) == list(gold[f'coords_{i}'])
    tr2, lg2, co2 = eng.process_lines(lines, sparse_logits=True)
    for i in range(len(lines)):
        assert np.array_equal(lg2[i].indptr, gold[f'csc_indptr_{i}'])
        assert np.array_equal(lg2[i].indices, gold[f'csc_indices_{i}'])
        np.testing.assert_allclose(lg2[i].data, gold[f'csc_data_{i}'], atol=2e-5)
    tr3, lg3, co3 = eng.process_lines(lines, sparse_logits=False, tight_crop_logits=True)
    for i in range(len(lines)):
        np.testing.assert_allclose(lg3[i], gold[f'tight_{i}'], atol=2e-5)
        assert co3[i] == [None, None]
    tr4, lg4, co4 = eng.process_lines(lines, no_logits=True)
    assert tr4 == tr and all(x is None for x in lg4)


def test_parsenet_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, 'parsenet.npz'))['maps']
    spec = cases.PARSENET_CASE
    net = make_net('parsenet', seed=spec['seed'])
    img = cases.parsenet_image()
    small = cv2.resize(img, (0, 0), fx=1 / spec['downsample'], fy=1 / spec['downsample'], interpolation=cv2.INTER_AREA)
    h64, w64 = -(-small.shape[0] // 64) * 64, -(-small.shape[1] // 64) * 64
    canvas = np.zeros((1, h64, w64, 3), dtype=np.uint8)
    canvas[0, :small.shape[0], :small.shape[1]] = small
    x = torch.from_numpy(canvas).float().permute(0, 3, 1, 2) * (1 / 255.)
    with torch.no_grad():
        y = parsenet_forward(dict(net.state_dict()), x)
    maps = y.permute(0, 2, 3, 1).numpy()[0, :small.shape[0], :small.shape[1]]
    np.testing.assert_allclose(maps, gold, atol=1e-5)


def test_confidence_chain_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, 'confidence.npz'))
    for i, m in enumerate(cases.confidence_logits()):
        sp = sparsify_logits(m)
        d = dense_logits(sp)
        assert np.array_equal(d, gold[f'dense_{i}'])
        np.testing.assert_allclose(full_logprobs(sp), gold[f'logprobs_{i}'], atol=1e-5)
        assert line_confidence(d) == pytest.approx(float(gold['confidence'][i]), rel=1e-6)
        assert line_confident_enough(full_logprobs(sp), 0.5) == bool(gold['confident_enough_0.5'][i])
