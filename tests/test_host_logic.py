"""CPU: host-side logic of the engine mirror (batching, padding, cropping, logit_coords, sparsification, ordering)
with the device step replaced by the torch-CPU oracle -- compared with the unmodified reference's outputs."""
import contextlib
import os
import re

import numpy as np
import pytest
import torch

from oracle import cases
from oracle.forward_oracle import greedy_ctc_indices, sparsify_logits
from tests.util import load_golden, make_case_net


def _host_only_engine(kind):
    from pero_ocr_b200.engine import B200EngineLineOCR
    spec = cases.ENGINE_CASES[kind]
    net = make_case_net(kind)
    eng = object.__new__(B200EngineLineOCR)
    eng.line_px_height = 40
    eng.line_padding_px = 32
    eng.net_subsampling = 4
    eng.batch_size = spec['engine_batch_size']
    eng.max_input_horizontal_pixels = 480 * eng.batch_size
    eng.characters = cases.json_characters(spec['classes'] - 2) + [u'​']
    eng.want_confidence = False
    eng.host_threads = 3
    eng.host_ms = {'stage': 0.0, 'wait': 0.0, 'finish': 0.0}
    eng._executor = None
    eng._models = [None]                   # one native engine: two batches in flight
    eng.h2d_bytes = eng.d2h_bytes = 0
    eng._device_ctx = contextlib.nullcontext
    store = {}

    def submit(k, shape, fill, no_logits, sparse_ranges=None, device_fill=None, packed=None):
        batch = np.empty(shape, dtype=np.uint8)
        batch[...] = 0xAB                      # stale garbage: the staging must overwrite every byte
        if packed is not None:                 # b200ocr_pad_lines' role (remap.cu; GPU-tested in test_gpu_cropper.py)
            batch[...] = 0
            for i, line in enumerate(packed):
                w = min(line.shape[1], shape[2] - eng.line_padding_px)
                batch[i, :, eng.line_padding_px:eng.line_padding_px + w] = line[:, :w]
        else:
            fill(batch)
        with torch.no_grad():
            logits = net(torch.from_numpy(batch).float().div(255.0).permute(0, 3, 1, 2)).numpy()
        ids = greedy_ctc_indices(logits)
        T = logits.shape[2]
        labels = np.full((shape[0], T), -1, dtype=np.int32)
        for i, v in enumerate(ids):
            labels[i, :len(v)] = v
        store[k] = dict(labels=labels, lengths=np.array([len(v) for v in ids], dtype=np.int32),
                        logits=np.ascontiguousarray(logits.transpose(0, 2, 1)))
        if sparse_ranges is not None:           # the device sparsifier's role, played by the oracle's NumPy pass
            lo, hi = sparse_ranges
            store[k]['sparse'] = [sparsify_logits(store[k]['logits'][i, lo[i]:hi[i]]) for i in range(shape[0])]
        return k, None

    eng._submit = submit
    eng._collect = lambda ticket: store[ticket[0]]
    return eng


@pytest.mark.parametrize('kind', ['lstm', 'transformer'])
def test_process_lines_host_logic_matches_reference(golden_dir, kind):
    gold = load_golden(golden_dir, f'engine_{kind}.npz')
    eng = _host_only_engine(kind)
    lines = cases.engine_lines(kind)
    tr, lg, co = eng.process_lines(lines, sparse_logits=False)
    assert tr == list(gold['transcriptions'])
    for i in range(len(lines)):
        np.testing.assert_allclose(lg[i], gold[f'logits_{i}'], atol=2e-5)
        assert list(co[i]) == list(gold[f'coords_{i}'])
    _, lg2, _ = eng.process_lines(lines, sparse_logits=True)
    for i in range(len(lines)):
        assert np.array_equal(lg2[i].indptr, gold[f'csc_indptr_{i}'])
        assert np.array_equal(lg2[i].indices, gold[f'csc_indices_{i}'])
    _, lg3, co3 = eng.process_lines(lines, sparse_logits=False, tight_crop_logits=True)
    for i in range(len(lines)):
        np.testing.assert_allclose(lg3[i], gold[f'tight_{i}'], atol=2e-5)
        assert co3[i] == [None, None]
    ids, _, _ = eng.process_lines(lines, no_logits=True, return_ids=True)
    assert [''.join(eng.characters[c] for c in v) for v in ids] == tr


def test_library_exports_every_declared_symbol():
    """include/b200_lineocr.h and libb200_lineocr.so agree (no compute call: there is no GPU here)."""
    from pero_ocr_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, 'include', 'b200_lineocr.h')).read()
    declared = set(re.findall(r'\b(b200ocr_[a-z_0-9]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pero_ocr_b200 import B200Error, netdesc
    from pero_ocr_b200.engine import LineRecognizer
    from pero_ocr_b200.decoders import GreedyDecoder
    layers, _ = netdesc.describe_line_net(make_case_net('lstm'))
    with pytest.raises(B200Error):
        LineRecognizer(layers)
    with pytest.raises(B200Error):
        GreedyDecoder(['a', '<BLANK>'])(np.log(np.array([[0.5, 0.5]])))


def test_product_package_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'pero_ocr_b200')
    for name in os.listdir(pkg):
        if name.endswith('.py'):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), name


def test_threaded_batch_padding_is_equivalent():
    """process_lines pads large batches with several host threads: same batch bytes, hence same results."""
    rng = np.random.default_rng(8)
    lines = [cases.line_crop(rng, int(w)) for w in rng.integers(20, 97, 70)]
    outs = []
    for threads in (1, 3):
        eng = _host_only_engine('lstm')
        eng.host_threads = threads
        eng.max_input_horizontal_pixels = 96 * 80          # one batch holds all 70 lines
        outs.append(eng.process_lines(lines, sparse_logits=False))
    assert outs[0][0] == outs[1][0] and outs[0][2] == outs[1][2]
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)


def test_process_lines_edge_cases_follow_the_reference():
    """Empty input -> three empty lists; a zero-width crop divides by zero in the batch-size computation exactly as
    in the reference (line_ocr_engine.py:84: `max_input_horizontal_pixels // max_width`); a 1-px line is one batch."""
    eng = _host_only_engine('lstm')
    assert eng.process_lines([]) == ([], [], [])
    with pytest.raises(ZeroDivisionError):
        eng.process_lines([np.zeros((40, 0, 3), dtype=np.uint8)])
    tr, lg, co = eng.process_lines([np.zeros((40, 1, 3), dtype=np.uint8)], sparse_logits=False)
    assert len(tr) == 1 and lg[0].shape[0] == (32 + 64) // 4 and co[0] == [8, 8]
    with pytest.raises(ValueError):
        eng.process_lines([np.zeros((39, 50, 3), dtype=np.uint8)])


def test_ar_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pero_ocr_b200 import B200Error, netdesc
    from pero_ocr_b200.transformer_engine import ARLineRecognizer
    net, dec, sd = cases.ar_state_dict()
    layers, decoder = netdesc.describe_transformer_ocr(sd, cases.AR_NET_CONFIG, 40)
    with pytest.raises(B200Error):
        ARLineRecognizer(layers, decoder)


def test_model_seam_recovers_the_byte_batch_exactly():
    """B1 seam (LineRecognizer.__call__): run_ocr's float input uint8 / 255 maps back to the same bytes for all 256
    values, in the NHWC layout the engine consumes."""
    from pero_ocr_b200.engine import LineRecognizer
    rng = np.random.default_rng(0)
    batch = rng.integers(0, 256, (2, 5, 64, 3), dtype=np.uint8)
    batch[0, 0, :, 0] = np.arange(64) * 4
    batch[0, 1, :, 1] = 255 - np.arange(64) * 4
    batch[1, 0, :64, 2] = np.arange(192, 256)
    x = torch.from_numpy(batch).float()
    x /= 255.0                                                   # pytorch_ocr_engine.py:61-62
    back = LineRecognizer.bytes_from_unit_floats(x.permute(0, 3, 1, 2))
    assert back.dtype == torch.uint8 and back.is_contiguous() and np.array_equal(back.numpy(), batch)
    every = torch.arange(256, dtype=torch.float32).div(255.0).view(1, 1, 1, 256).expand(1, 3, 1, 256)
    assert np.array_equal(LineRecognizer.bytes_from_unit_floats(every)[0, 0, :, 0].numpy(), np.arange(256))


def test_pinned_pool_lends_and_recycles_blocks():
    """sparse_logits.PinnedPool without a GPU (registration stubbed): arrays made over a block keep it on loan; it
    returns when the last one dies; the cap refuses instead of growing."""
    import gc
    from pero_ocr_b200.sparse_logits import PinnedPool, _buffer_of
    registered = []
    pool = PinnedPool(cap_bytes=32 << 20, register=lambda mem, size: registered.append(size), unregister=lambda mem: None)
    assert PinnedPool.block_size(1) == 2 << 20 and PinnedPool.block_size(44_000_000) <= 1.125 * 44_000_000
    blk = pool.take(5 << 20)
    whole = np.frombuffer(blk, dtype=np.int32, count=1000)
    assert whole.flags.writeable
    whole[:] = np.arange(1000)
    part = np.frombuffer(_buffer_of(whole), dtype=np.int32, count=10, offset=40)
    del blk, whole
    gc.collect()
    assert not pool.free and list(part) == list(range(10, 20))                 # still on loan through `part`
    del part
    gc.collect()
    assert sum(len(v) for v in pool.free.values()) == 1
    again = pool.take(5 << 20)
    assert pool.stats == {'hits': 1, 'new': 1, 'refused': 0} and registered == [again.nbytes]
    assert pool.take(64 << 20) is None and pool.stats['refused'] == 1
    del again
    pool.trim()
    assert pool.registered == 0 and not pool.free
    # a closed pool lends nothing, and a block that outlives it is unregistered when its last array dies
    released = []
    pool2 = PinnedPool(cap_bytes=32 << 20, register=lambda mem, size: None, unregister=lambda mem: released.append(1))
    late = np.frombuffer(pool2.take(3 << 20), dtype=np.uint8, count=16)
    pool2.close()
    assert pool2.take(3 << 20) is None and pool2.registered > 0 and not released
    del late
    gc.collect()
    assert pool2.registered == 0 and released == [1]


def test_batch_pipeline_runs_across_jobs_in_order():
    """_run_jobs (the driver behind process_lines / process_baselines / process_pages): several jobs through ONE batch
    pipeline -- every job's results are those of running it alone, jobs complete in order, the next job is pulled and
    its first batch submitted BEFORE the previous job's last batch is collected, and empty jobs pass through."""
    eng = _host_only_engine('lstm')
    lines = cases.engine_lines('lstm')
    alone = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    alone3 = eng.process_lines([l.copy() for l in lines[:3]], sparse_logits=False)
    events = []
    submit, collect = eng._submit, eng._collect

    def traced_submit(k, *a, **kw):
        events.append(('submit', k))
        return submit(k, *a, **kw)

    def traced_collect(ticket):
        events.append(('collect', ticket[0]))
        return collect(ticket)

    eng._submit, eng._collect = traced_submit, traced_collect
    pulled = []

    def jobs():
        for j, part in enumerate((lines, [], lines[:3], lines)):
            pulled.append((j, len(events)))
            part = list(part)
            yield [l.shape[1] for l in part], (lambda chunk, width, part=part: {'packed': [part[i] for i in chunk]})

    results = list(eng._run_jobs(jobs(), False, False, False, False))
    assert len(results) == 4 and results[1] == ([], [], [])
    for got, want in ((results[0], alone), (results[2], alone3), (results[3], alone)):
        assert got[0] == want[0] and len(got[1]) == len(want[1])
        for a, b in zip(got[1], want[1]):
            assert np.array_equal(a, b)
        assert [list(c) for c in got[2]] == [list(c) for c in want[2]]
    n_batches = sum(1 for e in events if e[0] == 'submit')
    assert n_batches == sum(1 for e in events if e[0] == 'collect')
    # job 3 (the last) was pulled while batches of earlier jobs were still uncollected
    at = pulled[3][1]
    assert sum(1 for e in events[:at] if e[0] == 'submit') > sum(1 for e in events[:at] if e[0] == 'collect')
    # slots rotate over 2 x replicas and a slot is never resubmitted before it was collected
    busy = set()
    for kind, k in events:
        if kind == 'submit':
            assert k not in busy
            busy.add(k)
        else:
            busy.discard(k)
