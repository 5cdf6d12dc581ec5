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
    net_kind = spec.get('net', kind)
    net = make_net(net_kind, spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'], **spec['net_kw'])
    return OracleEngine(dict(net.state_dict()), cases.json_characters(spec.get('json_chars', spec['classes'] - 2)),
                        kind=net_kind, batch_size=spec['engine_batch_size'])


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


def test_full_width_lines_match_reference(golden_dir):
    """Config-2 width: 8 lines of up to 1280 px through the unmodified PytorchEngineLineOCR (several of its batches)."""
    gold = np.load(os.path.join(golden_dir, 'engine_lstm_wide.npz'))
    eng = _engine('lstm_wide')
    lines = cases.engine_lines('lstm_wide')
    tr, lg, co = eng.process_lines(lines, sparse_logits=False)
    assert tr == list(gold['transcriptions'])
    for i in range(len(lines)):
        assert list(co[i]) == list(gold[f'coords_{i}'])
        if f'logits_{i}' in gold:
            np.testing.assert_allclose(lg[i], gold[f'logits_{i}'], atol=2e-5)
    best = np.concatenate([l.argmax(axis=1) for l in lg])
    srt = np.concatenate([np.sort(l, axis=1)[:, -2:] for l in lg])
    decided = (srt[:, 1] - srt[:, 0]) > 1e-4
    assert np.array_equal(best[decided], gold['best_path'][decided])


def test_checkpoint_class_convention_matches_reference(golden_dir):
    """A net that emits len(JSON characters) + 1 classes (the convention of real pero checkpoints: U+200B shares the
    blank's slot) through the engine oracle and the decoder oracles, letters as decoder_factory builds them
    (decoding_itf.py:49-50), against the unmodified reference's engine + decoders."""
    from oracle.decoders_oracle import greedy, prefix_beam
    gold = np.load(os.path.join(golden_dir, 'engine_lstm_c119.npz'))
    spec = cases.ENGINE_CASES['lstm_c119']
    eng = _engine('lstm_c119')
    assert len(eng.characters) == spec['classes']                       # 119 JSON characters + U+200B
    lines = cases.engine_lines('lstm_c119')
    tr, lg, co = eng.process_lines(lines, sparse_logits=True)
    assert tr == list(gold['transcriptions'])
    letters = cases.json_characters(spec['json_chars']) + [cases.BLANK]
    for i in range(len(lines)):
        lp = full_logprobs(lg[i])[co[i][0]:co[i][1]]
        assert greedy(lp, letters)[0] == str(gold['decoder_greedy'][i])
        hyps = prefix_beam(lp.astype(np.float64), 4)
        best = max(hyps, key=lambda h: h[1])
        assert ''.join(letters[c] for c in best[0]) == str(gold['decoder_beam4'][i])


def test_second_recogniser_family_walk_and_golden(golden_dir):
    """netdesc.describe_line_net is driven by module types and shapes: the second family (other names, nesting,
    LeakyReLU slopes, Conv1d head) walks to the expected layer list, whose restated forward (plain torch on the walked
    specs) reproduces the unmodified reference engine's golden logits; a tree registered out of order is refused."""
    from pero_ocr_b200 import _lib, netdesc, synthetic
    import torch.nn.functional as F
    gold = np.load(os.path.join(golden_dir, 'engine_lstm_alt.npz'))
    spec = cases.ENGINE_CASES['lstm_alt']
    net = make_net('lstm_alt', spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'])
    layers, n_classes = netdesc.describe_line_net(net)
    assert n_classes == 120
    assert [l['kind'] for l in layers] == [_lib.CONV_FIRST] + [_lib.CONV] * 5 + [_lib.BILSTM, _lib.CTC_HEAD]
    assert [l.get('act_slope') for l in layers[:6]] == [0.1, 0.1, None, 0.2, 0.2, 0.3]
    assert [(l['pool_h'], l['pool_w']) for l in layers[:6]] == [(1, 1), (2, 2), (1, 1), (2, 2), (2, 1), (1, 1)]

    def forward(x):                              # the layer list as the engine executes it
        for l in layers:
            if l['kind'] in (_lib.CONV_FIRST, _lib.CONV):
                x = F.conv2d(x, torch.from_numpy(l['weight']), torch.from_numpy(l['bias']), padding=(l['pad_h'], l['pad_w']))
                x = F.relu(x) if l['act'] == _lib.ACT_RELU else F.leaky_relu(x, l.get('act_slope', 0.01))
                if (l['pool_h'], l['pool_w']) != (1, 1):
                    x = F.max_pool2d(x, (l['pool_h'], l['pool_w']))
                if 'post_scale' in l:
                    x = x * torch.from_numpy(l['post_scale'])[None, :, None, None] + torch.from_numpy(l['post_shift'])[None, :, None, None]
            elif l['kind'] == _lib.BILSTM:
                seq = x.squeeze(2).permute(2, 0, 1) if x.dim() == 4 else x
                lstm = torch.nn.LSTM(l['cin'], l['hidden'], bidirectional=True)
                lstm.load_state_dict({f'{k}_l0{sfx}': torch.from_numpy(l[short][d]) for k, short in
                                      (('weight_ih', 'w_ih'), ('weight_hh', 'w_hh'), ('bias_ih', 'b_ih'), ('bias_hh', 'b_hh'))
                                      for d, sfx in ((0, ''), (1, '_reverse'))})
                x, _ = lstm(seq)
            else:
                x = (x @ torch.from_numpy(l['weight']).T + torch.from_numpy(l['bias'])).permute(1, 2, 0)
        return x

    line = cases.engine_lines('lstm_alt')[4]      # the widest line: the width of its reference batch is its own
    batch = np.zeros((1, 40, int(np.ceil(line.shape[1] / 32) * 32) + 64, 3), dtype=np.uint8)
    batch[0, :, 32:32 + line.shape[1]] = line
    with torch.no_grad():
        got = forward(torch.from_numpy(batch).float().div(255.0).permute(0, 3, 1, 2))[0].numpy().T
    np.testing.assert_allclose(got, gold['logits_4'], atol=3e-5)
    # the same modules registered in another order: channel counts do not chain
    wrong = synthetic.LineNetLSTMAlt(120)
    feats = list(wrong.features.named_children())
    wrong.features = torch.nn.Sequential(*[m for _, m in (feats[2], feats[1], feats[0], feats[3])])
    with pytest.raises(ValueError):
        netdesc.describe_line_net(wrong)


def test_embedding_fold_equals_the_conditioned_forward(golden_dir):
    """netdesc folds the gathered embedding of `model(x, ids)` (one id per batch, pytorch_ocr_engine.py:64-66) into the
    aggregation layer's post-activation shift: the shift is table[id] for an int id and the last row for "mean", and
    the module's own forward with that id reproduces the unmodified reference engine's golden logits."""
    from pero_ocr_b200 import netdesc
    gold = np.load(os.path.join(golden_dir, 'engine_lstm_embed.npz'))
    spec = cases.ENGINE_CASES['lstm_embed']
    net = make_net('lstm_embed', spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'], **spec['net_kw'])
    table = net.embeddings_layer.weight.detach().numpy()
    for embed_id, row in ((2, 2), ('mean', 5), (0, 0)):
        layers, _ = netdesc.describe_line_net(net, embed_id=embed_id)
        agg = [l for l in layers if 'embedding_table' in l]
        assert len(agg) == 1 and np.array_equal(agg[0]['post_shift'], table[row]) and (agg[0]['post_scale'] == 1).all()
        assert netdesc.resolve_embed_id(embed_id, table.shape[0]) == row
    with pytest.raises(ValueError):
        netdesc.describe_line_net(net)
    with pytest.raises(IndexError):
        netdesc.describe_line_net(net, embed_id=6)
    line = cases.engine_lines('lstm_embed')[0]          # the widest line of the case
    batch = np.zeros((1, 40, int(np.ceil(line.shape[1] / 32) * 32) + 64, 3), dtype=np.uint8)
    batch[0, :, 32:32 + line.shape[1]] = line
    x = torch.from_numpy(batch).float().div(255.0).permute(0, 3, 1, 2)
    with torch.no_grad():
        for embed_id, key in ((2, 'logits_0'), (5, 'mean_logits_0'), (0, 'id0_logits_0')):
            got = net(x, torch.tensor([embed_id]))[0].numpy().T
            np.testing.assert_allclose(got, gold[key], atol=3e-5)


def test_netdesc_refuses_what_the_kernels_do_not_run():
    """Error behaviour of the type-driven walk: every unsupported shape is a ValueError at construction time, never a
    silent fallback (there is no generic executor behind it)."""
    from torch import nn
    from pero_ocr_b200 import netdesc

    def line_net(front, agg=None, seq=None, head=None):
        m = nn.Module()
        m.front = nn.Sequential(*front)
        m.agg = agg if agg is not None else nn.Sequential(nn.Conv2d(64, 128, (5, 1)), nn.ReLU())
        m.seq = seq if seq is not None else nn.LSTM(128, 256, bidirectional=True)
        m.head = head if head is not None else nn.Linear(512, 10)
        return m

    ok_front = [nn.Conv2d(3, 64, 3, padding=1), nn.ReLU()]
    layers, classes = netdesc.describe_line_net(line_net(ok_front))
    assert classes == 10 and len(layers) == 4
    bad = {
        '5x5 frontend conv': line_net([nn.Conv2d(3, 64, 5, padding=2), nn.ReLU()]),
        'unpadded frontend conv': line_net([nn.Conv2d(3, 64, 3, padding=0), nn.ReLU()]),
        'strided frontend conv': line_net([nn.Conv2d(3, 64, 3, padding=1, stride=2), nn.ReLU()]),
        '3x3 max-pool': line_net(ok_front + [nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 3)]),
        'overlapping max-pool': line_net(ok_front + [nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 1)]),
        'pool straight after the first conv': line_net(ok_front + [nn.MaxPool2d(2, 2)]),
        'two activations in a row': line_net(ok_front + [nn.ReLU()]),
        'negative LeakyReLU slope': line_net([nn.Conv2d(3, 64, 3, padding=1), nn.LeakyReLU(-0.1)]),
        'other activation': line_net([nn.Conv2d(3, 64, 3, padding=1), nn.Tanh()]),
        'unidirectional LSTM': line_net(ok_front, seq=nn.LSTM(128, 256, bidirectional=False), head=nn.Linear(256, 10)),
        'GRU': line_net(ok_front, seq=nn.GRU(128, 256, bidirectional=True)),
        'channel mismatch': line_net(ok_front, agg=nn.Sequential(nn.Conv2d(32, 128, (5, 1)), nn.ReLU())),
        'head width mismatch': line_net(ok_front, head=nn.Linear(256, 10)),
        'wide head kernel': line_net(ok_front, head=nn.Conv1d(512, 10, 3)),
        'no head': line_net(ok_front, head=nn.Identity()),
    }
    for name, net in bad.items():
        with pytest.raises(ValueError):
            netdesc.describe_line_net(net)
            pytest.fail(f'{name} was accepted')
