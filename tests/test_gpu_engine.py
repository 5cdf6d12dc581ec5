"""GPU parity: the CUDA recogniser behind B200EngineLineOCR against (1) outputs of the unmodified reference
PytorchEngineLineOCR stored in tests/golden/engine_*.npz and (2) the torch-CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star: "bit-exact CTC argmax indices; logits within 1e-3 fp32"):
  * precision 'fp16x3' (default): |logit - reference| <= 1e-3 absolute; per-frame argmax identical on every frame
    whose reference top-2 margin exceeds 2e-3 (closer calls are numerically undecidable between any two fp32
    implementations); transcriptions identical.
  * precision 'fp16f8' (fp16 pass + e5m2 first-order correction pass, 2 pass-equivalents): the same bars as 'fp16x3'.
  * precision 'fp16' (single-pass, same 10-bit mantissa as the reference's own cuDNN-TF32 GPU path): 3e-3 * max|logit|.
"""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.forward_oracle import OracleEngine, dense_logits, greedy_ctc_indices, line_confidence, sparsify_logits
from tests.util import load_golden, make_case_net, write_engine_json

pytestmark = pytest.mark.gpu

TOL = 1e-3
MARGIN = 2e-3


def _engine(tmp_path, kind, precision='fp16x3', batch_size=None):
    from pero_ocr_b200.engine import B200EngineLineOCR
    spec = cases.ENGINE_CASES[kind]
    js = write_engine_json(tmp_path, kind)
    return B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=batch_size or spec['engine_batch_size'],
                             precision=precision, module=make_case_net(kind))


@pytest.mark.parametrize('precision', ['fp16x3', 'fp16f8'])
@pytest.mark.parametrize('kind', ['lstm', 'transformer'])
def test_process_lines_matches_reference_golden(tmp_path, golden_dir, kind, precision):
    gold = load_golden(golden_dir, f'engine_{kind}.npz')
    eng = _engine(tmp_path, kind, precision=precision)
    lines = cases.engine_lines(kind)
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    assert eng.characters == list(gold['chars'])
    worst = 0.0
    for i in range(len(lines)):
        ref = gold[f'logits_{i}']
        assert lg[i].shape == ref.shape and lg[i].dtype == np.float32
        worst = max(worst, float(np.abs(lg[i] - ref).max()))
        assert list(co[i]) == list(gold[f'coords_{i}'])
        # bit-exact argmax wherever the reference's own decision is numerically meaningful
        srt = np.sort(ref, axis=1)
        decided = (srt[:, -1] - srt[:, -2]) > MARGIN
        assert np.array_equal(lg[i].argmax(axis=1)[decided], ref.argmax(axis=1)[decided])
    assert worst <= TOL, worst
    assert tr == list(gold['transcriptions'])
    # tight crop / no-logits variants (line_ocr_engine.py:145-150, 143-144)
    tr2, lg2, co2 = eng.process_lines([l.copy() for l in lines], sparse_logits=False, tight_crop_logits=True)
    for i in range(len(lines)):
        assert co2[i] == [None, None]
        assert np.abs(lg2[i] - gold[f'tight_{i}']).max() <= TOL
    tr3, lg3, co3 = eng.process_lines([l.copy() for l in lines], no_logits=True)
    assert tr3 == tr and all(x is None for x in lg3) and all(x is None for x in co3)


@pytest.mark.parametrize('kind', ['lstm', 'transformer'])
def test_sparse_logits_match_reference_golden(tmp_path, golden_dir, kind):
    from scipy import sparse
    gold = load_golden(golden_dir, f'engine_{kind}.npz')
    eng = _engine(tmp_path, kind)
    lines = cases.engine_lines(kind)
    _, lg, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    for i in range(len(lines)):
        assert sparse.issparse(lg[i]) and lg[i].format == 'csc'
        ref = sparse.csc_matrix((gold[f'csc_data_{i}'], gold[f'csc_indices_{i}'], gold[f'csc_indptr_{i}']),
                                shape=lg[i].shape).toarray()
        got = lg[i].toarray()
        both = (ref != 0) & (got != 0)
        assert np.abs(got[both] - ref[both]).max() <= TOL
        # the keep/drop pattern may differ only for probabilities within rounding of the 1e-4 threshold
        assert ((ref != 0) != (got != 0)).mean() < 1e-3


def test_checkpoint_file_route(tmp_path, golden_dir):
    """Same TorchScript checkpoint file the reference engine loads (pytorch_ocr_engine.py:52-57)."""
    from pero_ocr_b200.engine import B200EngineLineOCR
    net = make_case_net('lstm')
    torch.jit.script(net).save(str(tmp_path / 'ck.pt'))
    js = write_engine_json(tmp_path, 'lstm', checkpoint='ck.pt')
    eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=1)
    gold = load_golden(golden_dir, 'engine_lstm.npz')
    tr, _, _ = eng.process_lines(cases.engine_lines('lstm'), no_logits=True)
    assert tr == list(gold['transcriptions'])


def test_single_pass_fp16_mode(tmp_path, golden_dir):
    gold = load_golden(golden_dir, 'engine_lstm.npz')
    eng = _engine(tmp_path, 'lstm', precision='fp16')
    lines = cases.engine_lines('lstm')
    tr, lg, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    scale = max(float(np.abs(gold[f'logits_{i}']).max()) for i in range(len(lines)))
    agree = total = 0
    for i in range(len(lines)):
        ref = gold[f'logits_{i}']
        assert np.abs(lg[i] - ref).max() <= 3e-3 * scale
        agree += int((lg[i].argmax(axis=1) == ref.argmax(axis=1)).sum())
        total += ref.shape[0]
    assert agree / total > 0.97


@pytest.mark.parametrize('precision', ['fp16x3', 'fp16f8'])
@pytest.mark.parametrize('kind', ['lstm', 'transformer'])
def test_tensor_core_kernels_match_cuda_core_cross_check(kind, precision):
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    net = make_case_net(kind)
    layers, _ = netdesc.describe_line_net(net)
    eng = LineRecognizer(layers, precision=precision)
    rng = np.random.default_rng(3)
    crops = torch.from_numpy(rng.integers(0, 256, (5, 40, 328, 3), dtype=np.uint8)).cuda()
    a = eng.forward(crops, want_logits=True, want_best_path=True)
    a = {k: v.clone() for k, v in a.items()}
    eng.use_reference_kernels(True)
    b = eng.forward(crops, want_logits=True, want_best_path=True, out={})
    torch.cuda.synchronize()
    assert (a['logits'] - b['logits']).abs().max().item() <= TOL


def test_fused_confidence_matches_reference_chain(tmp_path):
    """confidence output == PageParser.compute_line_confidence on the engine's own sparsified logits."""
    eng = _engine(tmp_path, 'lstm', batch_size=8)
    eng.want_confidence = True
    lines = cases.engine_lines('lstm')
    _, lg, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    for i in range(len(lines)):
        want = line_confidence(dense_logits(lg[i]))
        assert eng.last_line_confidences[i] == pytest.approx(want, rel=2e-4)


def test_error_behaviour(tmp_path):
    from pero_ocr_b200 import B200Error
    eng = _engine(tmp_path, 'lstm')
    with pytest.raises(B200Error):
        eng.run_ocr(np.zeros((1, 32, 64, 3), dtype=np.uint8))        # wrong line height
    with pytest.raises(ValueError):
        eng.run_ocr(np.zeros((1, 40, 64, 1), dtype=np.uint8))        # not 3 channels
    tr, lg, co = eng.process_lines([])
    assert tr == [] and lg == [] and co == []


def test_full_size_properties():
    """BASELINE.json config 2 size (256 x 40 x 1344): determinism and batch independence -- a line's result must
    not depend on which other lines share its batch (the reference has the same property by construction)."""
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    net = make_case_net('lstm')
    layers, _ = netdesc.describe_line_net(net)
    eng = LineRecognizer(layers, precision='fp16x3')
    crops = np.zeros((256, 40, 1344, 3), dtype=np.uint8)
    crops[:, :, 32:-32] = cases.bench_crops(256, 1280, seed=0)
    d = torch.from_numpy(crops).cuda()
    a = {k: v.clone() for k, v in eng.forward(d, want_logits=True).items()}
    b = {k: v.clone() for k, v in eng.forward(d, want_logits=True, out={}).items()}
    assert torch.equal(a['labels'], b['labels']) and torch.equal(a['logits'], b['logits'])
    sub = eng.forward(d[100:107].contiguous(), want_logits=True, out={})
    torch.cuda.synchronize()
    assert torch.equal(sub['labels'], a['labels'][100:107])
    assert (sub['logits'] - a['logits'][100:107]).abs().max().item() <= 1e-5
    # oracle on a bounded sample of the same batch
    with torch.no_grad():
        x = torch.from_numpy(crops[:2]).float().div(255.0).permute(0, 3, 1, 2)
        ref = net(x).numpy()
    got = a['logits'][:2].cpu().numpy()
    assert np.abs(got - ref.transpose(0, 2, 1)).max() <= TOL
    want = greedy_ctc_indices(ref)
    lab, ln = a['labels'][:2].cpu().numpy(), a['lengths'][:2].cpu().numpy()
    srt = np.sort(ref, axis=1)
    if ((srt[:, -1] - srt[:, -2]) > MARGIN).all():
        assert [list(lab[i, :ln[i]]) for i in range(2)] == [list(w) for w in want]


@pytest.mark.parametrize('precision', ['fp16x3', 'fp16f8'])
@pytest.mark.parametrize('width', [136, 264, 520])
def test_halo_kernel_matches_per_tap_kernel(width, precision):
    """The halo-reuse 3x3 kernel (shifted swizzle-128B views) and the per-tap TMA kernel are two implementations of
    the same contraction: identical operands, fp32 accumulation in a different order."""
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    net = make_case_net('lstm')
    layers, _ = netdesc.describe_line_net(net)
    eng = LineRecognizer(layers, precision=precision)
    rng = np.random.default_rng(width)
    crops = torch.from_numpy(rng.integers(0, 256, (3, 40, width, 3), dtype=np.uint8)).cuda()
    a = {k: v.clone() for k, v in eng.forward(crops, want_logits=True).items()}
    eng.set_flag(1, 0)
    b = eng.forward(crops, want_logits=True, out={})
    torch.cuda.synchronize()
    assert (a['logits'] - b['logits']).abs().max().item() <= 2e-4
    assert torch.equal(a['labels'], b['labels'])


@pytest.mark.parametrize('width', [40, 136, 264, 128, 272, 1344])
def test_first_conv_tensor_core_kernel_is_fp32_grade(width):
    """conv_first.cu (uint8 pixels as exact fp16 operands, weights split hi + lo per channel scale, 1/255 in the epilogue)
    against torch fp32 conv2d(x / 255) and against the CUDA-core fp32 cross-check kernel, read back after the first
    layer -- the tcgen05 kernel (widths that are multiples of 16: TMA-staged patch) and the mma.sync kernel under its
    staging variants (flag 4; other widths fall back to plain loads).  The fp16x3 record (hi + lo) resolves ~2^-22 of
    the value, so the bar is a few fp32 ulps of the largest activation."""
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    net = make_case_net('lstm')
    layers, _ = netdesc.describe_line_net(net)
    eng = LineRecognizer(layers, precision='fp16x3')
    rng = np.random.default_rng(width)
    crops = rng.integers(0, 256, (3, 40, width, 3), dtype=np.uint8)
    crops[0, :, :7] = 255                                   # saturated block at the left edge
    crops[1] = 0                                            # all-zero line: output = act(bias)
    d = torch.from_numpy(crops).cuda()
    conv = net.conv[0]
    with torch.no_grad():
        x = torch.from_numpy(crops).float().div(255.0).permute(0, 3, 1, 2)
        want = torch.relu(conv(x)).permute(0, 2, 3, 1).numpy()
    scale = float(np.abs(want).max())
    for staging in (3, 2, 1, 0):                            # 3 = tcgen05 (the default), 2 / 1 / 0 = mma.sync variants
        eng.set_flag(4, staging)
        got = eng.debug_forward_prefix(d, 1)                # [n, h, w, 64] fp32
        assert np.abs(got - want).max() <= 4e-6 * max(scale, 1.0), staging
    eng.set_flag(4, 3)
    eng.use_reference_kernels(True)
    ref_kernel = eng.debug_forward_prefix(d, 1)
    assert np.abs(ref_kernel - want).max() <= 4e-6 * max(scale, 1.0)


@pytest.mark.parametrize('fmt_precision', ['fp16f8', 'fp16'])
def test_first_conv_with_32_channels_and_leaky_slope(fmt_precision):
    """The first conv alone with 32 output channels and LeakyReLU(0.2), both kernels, in the one- and two-plane record
    formats (read back through the record: the bar is the record's resolution)."""
    from pero_ocr_b200 import _lib
    from pero_ocr_b200.engine import LineRecognizer
    rng = np.random.default_rng(9)
    wgt = (rng.standard_normal((32, 3, 3, 3)) * 0.3).astype(np.float32)
    bias = (rng.standard_normal(32) * 0.1).astype(np.float32)
    layers = [dict(kind=_lib.CONV_FIRST, cin=3, cout=32, kh=3, kw=3, pad_h=1, pad_w=1, act=_lib.ACT_LEAKY_RELU, act_slope=0.2,
                   pool_h=1, pool_w=1, weight=wgt, bias=bias)]
    eng = LineRecognizer(layers, precision=fmt_precision)
    crops = rng.integers(0, 256, (2, 40, 160, 3), dtype=np.uint8)
    d = torch.from_numpy(crops).cuda()
    with torch.no_grad():
        x = torch.from_numpy(crops).float().div(255.0).permute(0, 3, 1, 2)
        want = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, torch.from_numpy(wgt), torch.from_numpy(bias),
                                                                        padding=1), 0.2).permute(0, 2, 3, 1).numpy()
    tol = (2.0 ** -11 if fmt_precision == 'fp16' else 2.0 ** -13) * float(np.abs(want).max())
    for staging in (3, 2):
        eng.set_flag(4, staging)
        got = eng.debug_forward_prefix(d, 1)
        assert got.shape == want.shape and np.abs(got - want).max() <= tol, (staging, float(np.abs(got - want).max()), tol)


def test_transformer_variant_on_a_very_long_line():
    """A 2016 px line (T = 520 frames) exceeds what the attention kernel can keep of K / V in shared memory; the
    in-place variant must give the oracle's logits too (same bars as the other engine tests)."""
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    net = make_case_net('transformer')
    layers, _ = netdesc.describe_line_net(net)
    eng = LineRecognizer(layers)
    rng = np.random.default_rng(77)
    crops = np.zeros((2, 40, 2080, 3), dtype=np.uint8)
    crops[:, :, 32:-32] = np.repeat(rng.integers(0, 256, (2, 40, 2016, 1), dtype=np.uint8), 3, axis=3)
    with torch.no_grad():
        ref = net(torch.from_numpy(crops).float().div(255.0).permute(0, 3, 1, 2)).numpy()       # [N, C, T]
    out = eng.forward(torch.from_numpy(crops).cuda(), want_logits=True, want_best_path=True)
    torch.cuda.synchronize()
    got = out['logits'].cpu().numpy()
    assert got.shape == (2, 520, 120)
    assert np.abs(got - ref.transpose(0, 2, 1)).max() <= TOL
    srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > MARGIN                                                # [N, T]
    assert np.array_equal(out['best_path'].cpu().numpy()[decided], ref.argmax(axis=1)[decided])


# ---- per-layer correction modes of the fp16f8 precision (b200ocr_set_layer_correction) ---------------------------

def _case_recognizer(kind='lstm', precision='fp16f8'):
    from pero_ocr_b200 import netdesc
    from pero_ocr_b200.engine import LineRecognizer
    layers, _ = netdesc.describe_line_net(make_case_net(kind))
    return LineRecognizer(layers, precision=precision), layers


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_layer_correction_modes_match_cuda_core_cross_check(mode):
    """Both (0), weight-side-only (1) and no (2) e5m2 correction, layer by layer: the halo kernel's chunk / K-step
    skips (cin = 64 and 128) and the per-tap kernel's shortened e5m2 pass against the CUDA-core kernel walking the
    same operand bytes.  Each convolution is isolated on IDENTICAL inputs (debug flag 5: only that layer runs on the
    cross-check kernel) -- end to end, the two paths' independent fp16 roundings of every activation would otherwise
    show up as ~5e-4 of logit noise once the activation-side correction is off."""
    from pero_ocr_b200 import _lib
    eng, layers = _case_recognizer()
    convs = [i for i, l in enumerate(layers) if l['kind'] == _lib.CONV]
    for i in convs:
        eng.set_layer_correction(i, mode)
    rng = np.random.default_rng(11)
    crops = torch.from_numpy(rng.integers(0, 256, (3, 40, 328, 3), dtype=np.uint8)).cuda()
    worst = 0.0
    for i in convs:
        eng.set_flag(5, -1)
        a = eng.debug_forward_prefix(crops, i + 1)
        eng.set_flag(5, i)
        b = eng.debug_forward_prefix(crops, i + 1)
        # the read-back is the record's value hi + lo' (lo' has 3 significant bits: steps of 2^-14 of the element), so
        # two fp32 results a rounding apart may read back one such step apart; a wrong chunk / K-step would be >= 1e-3
        # (plus, on small elements of deep layers, the fp32 summation-order noise of K = 4608 terms)
        scale = max(1.0, float(np.sqrt((b.astype(np.float64) ** 2).mean())))
        err = (np.abs(a - b) - 5e-5 * scale) / (np.abs(b) + 1e-6)
        worst = max(worst, float(err.max()))
        assert err.max() <= 2.0 ** -12, (i, float(err.max()))
    eng.set_flag(5, -1)
    print(f'correction mode {mode}: tcgen05 vs CUDA-core cross-check per layer, worst relative difference {worst:.2e}')
    total, per = eng.executed_passes(3, 328)
    want = {0: 2.0, 1: 1.5, 2: 1.0}[mode]
    assert all(per[i] == pytest.approx(want) for i in convs)
    assert all(per[i] == pytest.approx(2.0) for i, l in enumerate(layers) if l['kind'] in (_lib.BILSTM, _lib.CTC_HEAD))


def test_no_correction_equals_single_pass_fp16():
    """CORR_NONE leaves the fp16 hi * hi pass: through the convolution stack that is the arithmetic of precision
    'fp16' exactly (the accumulator scale 2^11 is a power of two), read back after the last frontend layer."""
    from pero_ocr_b200 import _lib
    eng, layers = _case_recognizer()
    convs = [i for i, l in enumerate(layers) if l['kind'] == _lib.CONV]
    for i in convs:
        eng.set_layer_correction(i, _lib.CORR_NONE)
    plain, _ = _case_recognizer(precision='fp16')
    rng = np.random.default_rng(12)
    crops = torch.from_numpy(rng.integers(0, 256, (3, 40, 264, 3), dtype=np.uint8)).cuda()
    last = convs[-1] + 1
    a = eng.debug_forward_prefix(crops, last)          # hi + lo' of the records: the fp32 value before rounding
    b = plain.debug_forward_prefix(crops, last)        # fp16 records
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= 2e-3 * max(1.0, float(np.abs(b).max()))
    assert np.abs(a.astype(np.float16).astype(np.float32) - b).max() <= 2e-3 * max(1.0, float(np.abs(b).max()))


def test_weight_only_preset_holds_the_parity_bar(tmp_path, golden_dir):
    """precision 'fp16f8w' (weight-side correction only in the deep 3x3 layers) against the unmodified reference's
    outputs: the same 1e-3 logit bar, per-frame argmax identical on every decided frame (transcriptions identical
    whenever no undecidable frame flipped); 1.7 instead of 2 pass-equivalents."""
    gold = load_golden(golden_dir, 'engine_lstm.npz')
    eng = _engine(tmp_path, 'lstm', precision='fp16f8w')
    lines = cases.engine_lines('lstm')
    tr, lg, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    worst, flipped = 0.0, 0
    for i in range(len(lines)):
        ref = gold[f'logits_{i}']
        worst = max(worst, float(np.abs(lg[i] - ref).max()))
        srt = np.sort(ref, axis=1)
        margin = srt[:, -1] - srt[:, -2]
        flips = lg[i].argmax(axis=1) != ref.argmax(axis=1)
        assert not (flips & (margin > MARGIN)).any()
        flipped += int(flips.sum())
    print(f'fp16f8w: worst |logit - reference| = {worst:.2e}, {flipped} undecidable frames flipped')
    assert worst <= TOL, worst
    if not flipped:
        assert tr == list(gold['transcriptions'])
    total, _ = eng.model.executed_passes(8, 512)
    assert 1.6 < total < 1.8


def test_autotune_precision_respects_its_budget():
    eng, layers = _case_recognizer()
    rep = eng.autotune_precision(budget=3e-4)
    assert rep['max_abs_dev_vs_full_correction'] <= 3e-4
    assert rep['executed_passes'] <= 2.0
    rng = np.random.default_rng(13)
    crops = torch.from_numpy(rng.integers(0, 256, (4, 40, 264, 3), dtype=np.uint8)).cuda()
    net = make_case_net('lstm')
    with torch.no_grad():
        ref = net(torch.from_numpy(crops.cpu().numpy()).float().div(255.0).permute(0, 3, 1, 2)).numpy()
    got = eng.forward(crops, want_logits=True)['logits'].cpu().numpy()
    assert np.abs(got - ref.transpose(0, 2, 1)).max() <= TOL
    strict = eng.autotune_precision(budget=0.0)
    assert strict['weight_only_layers'] == [] and strict['executed_passes'] == pytest.approx(2.0)


# ---- config-2 width under the default precision, and the class-count convention of real checkpoints ---------------

@pytest.mark.parametrize('autotune', [None, 5e-4])
def test_full_width_lines_match_reference_golden(tmp_path, golden_dir, autotune):
    """8 lines of up to 1280 px (T = 336: BASELINE config 2's shape) against the unmodified PytorchEngineLineOCR, in the
    default precision fp16f8 -- and with the per-layer corrections chosen by autotune_precision(5e-4), the arithmetic
    bench.py times.  Logits within 1e-3; per-frame argmax == the reference's stored best_path on every decided frame;
    the exempt (undecidable) frames are counted and printed."""
    gold = load_golden(golden_dir, 'engine_lstm_wide.npz')
    eng = _engine(tmp_path, 'lstm_wide', precision='fp16f8', batch_size=8)     # the reference's budget: same batches,
    if autotune:                                                                # hence the same logit frames per line
        rep = eng.model.autotune_precision(budget=autotune)
        print('autotune:', rep['weight_only_layers'], f"{rep['executed_passes']:.3f} passes")
    lines = cases.engine_lines('lstm_wide')
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    worst = 0.0
    for i in range(len(lines)):
        assert list(co[i]) == list(gold[f'coords_{i}'])
        if f'logits_{i}' in gold:
            worst = max(worst, float(np.abs(lg[i] - gold[f'logits_{i}']).max()))
    assert worst <= TOL, worst
    ours = np.concatenate([l.argmax(axis=1) for l in lg])
    ref_best = gold['best_path']
    assert ours.shape == ref_best.shape
    srt = np.concatenate([np.sort(l, axis=1)[:, -2:] for l in lg])
    margin = srt[:, 1] - srt[:, 0]
    flips = ours != ref_best
    hist = {f'<{b:g}': int((margin < b).sum()) for b in (1e-4, 3e-4, 1e-3, 2e-3)}
    print(f'{ours.size} frames; margin histogram {hist}; argmax differs from the reference on {int(flips.sum())} frames, '
          f'largest margin among them {float(margin[flips].max()) if flips.any() else 0.0:.1e}; worst |logit - ref| {worst:.2e}')
    assert not (flips & (margin > MARGIN)).any()
    same = [a == b for a, b in zip(tr, gold['transcriptions'])]
    if not flips.any():
        assert all(same)


def test_checkpoint_class_convention(tmp_path, golden_dir):
    """Net with len(JSON characters) + 1 classes (real pero checkpoints; U+200B shares the blank's slot): the engine
    accepts it, transcribes like the unmodified reference engine, and the decoders built the way decoder_factory
    builds them (JSON characters + '<BLANK>', decoding_itf.py:49-50) reproduce the reference decoders' results on
    the engine's own logits, host chain and fused device chain alike."""
    from pero_ocr_b200.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder, GreedyDecoder
    from oracle.forward_oracle import full_logprobs
    gold = load_golden(golden_dir, 'engine_lstm_c119.npz')
    spec = cases.ENGINE_CASES['lstm_c119']
    eng = _engine(tmp_path, 'lstm_c119')
    assert eng.num_classes == spec['classes'] == len(eng.characters)
    lines = cases.engine_lines('lstm_c119')
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    assert tr == list(gold['transcriptions'])
    letters = cases.json_characters(spec['json_chars']) + [BLANK_SYMBOL]
    gd, bd = GreedyDecoder(letters), CTCPrefixLogRawNumpyDecoder(letters, 4)
    for i in range(len(lines)):
        lp = full_logprobs(lg[i])[co[i][0]:co[i][1]]
        assert gd(lp).best_hyp() == str(gold['decoder_greedy'][i])
        assert bd(lp.astype(np.float64)).best_hyp() == str(gold['decoder_beam4'][i])
    bags = eng.decode_lines([l.copy() for l in lines], bd)
    assert [b.best_hyp() for b in bags] == [str(x) for x in gold['decoder_beam4']]
    with pytest.raises(ValueError):
        eng.decode_lines(lines, CTCPrefixLogRawNumpyDecoder(letters[:-2] + [BLANK_SYMBOL], 4))


@pytest.mark.parametrize('precision', ['fp16f8', 'fp16x3'])
@pytest.mark.parametrize('width', [40, 264, 1344, 1536])
def test_tensor_core_attention_matches_cuda_core_attention(width, precision):
    """attention_tc.cu (tcgen05: S = Q K^T and O = P V with hi / lo operand splits, P kept in tensor memory) against the
    fp32 CUDA-core attention kernel on the same QKV rows (debug flag 8), for T = 10, 66 (one partial query tile),
    336 (BASELINE config 3: three query tiles, 96-key last chunk) and 384 (the kernel's limit), and against the
    torch-CPU fp32 module."""
    eng, _ = _case_recognizer('transformer', precision)
    rng = np.random.default_rng(width)
    n = 3 if width < 1000 else 2
    crops = np.repeat(rng.integers(0, 256, (n, 40, width, 1), dtype=np.uint8), 3, axis=3)
    crops[-1, :, width // 2:] = 0                          # a line that ends half way: trailing padding frames
    d = torch.from_numpy(crops).cuda()
    a = eng.forward(d, want_logits=True, out={})['logits'].clone()
    eng.set_flag(8, 0)
    b = eng.forward(d, want_logits=True, out={})['logits'].clone()
    torch.cuda.synchronize()
    diff = (a - b).abs().max().item()
    print(f'T = {width // 4}: tcgen05 vs CUDA-core attention, max |d logit| = {diff:.2e}')
    # fp16x3 records resolve 2^-22 of a value; an fp16f8 record's e5m2 residual 2^-14 -- two fp32 results a rounding
    # apart may land one such step apart, which the layers downstream turn into a few 1e-4 at |logit| <= 8.8
    assert diff <= (1e-4 if precision == 'fp16x3' else 6e-4)
    net = make_case_net('transformer')
    with torch.no_grad():
        ref = net(torch.from_numpy(crops).float().div(255.0).permute(0, 3, 1, 2)).permute(0, 2, 1).numpy()
    assert np.abs(a.cpu().numpy() - ref).max() <= TOL


def test_recycled_pinned_result_blocks(tmp_path):
    """Sparse logits of a batch whose CSC parts exceed the pool's threshold (a flat net: every class kept) come back
    identical through recycled page-locked blocks and through fresh pageable arrays; dropping a call's matrices returns
    the blocks; a pool that is too small falls back without an error."""
    import gc
    from pero_ocr_b200.engine import B200EngineLineOCR
    from oracle.nets import make_net
    spec = cases.ENGINE_CASES['lstm']
    js = write_engine_json(tmp_path, 'lstm')
    net = make_net('lstm', spec['classes'], seed=3, out_gain=0.05, **spec['net_kw'])       # near-uniform posteriors
    rng = np.random.default_rng(5)
    lines = [rng.integers(0, 256, size=(40, 600 + 8 * (i % 5), 3), dtype=np.uint8) for i in range(48)]
    engines = {name: B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=64, precision='fp16f8', module=net,
                                       pinned_logit_bytes=cap)
               for name, cap in (('pool', 1 << 30), ('fresh', 0), ('tiny', 3 << 20))}
    out = {name: e.process_lines([l.copy() for l in lines]) for name, e in engines.items()}
    pool = engines['pool'].pinned_pool
    assert pool.stats['new'] >= 2 and pool.stats['refused'] == 0 and engines['fresh'].pinned_pool is None
    assert engines['tiny'].pinned_pool.stats['refused'] >= 1
    kept = np.mean([m.nnz / m.shape[0] for m in out['pool'][1]])
    assert kept > 60, kept                                              # the case is the dense worst case
    for name in ('fresh', 'tiny'):
        assert out[name][0] == out['pool'][0]
        for a, b in zip(out[name][1], out['pool'][1]):
            assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data)
    first = [m.copy() for m in out['pool'][1]]
    held = pool.registered
    del out, a, b
    gc.collect()
    assert sum(len(v) for v in pool.free.values()) == pool.stats['new']          # everything came back
    again = engines['pool'].process_lines([l.copy() for l in lines])
    assert pool.stats['hits'] >= 2 and pool.registered == held                    # served from the returned blocks
    for a, b in zip(again[1], first):
        assert np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)


def test_two_linked_replicas_give_the_single_engine_results(tmp_path, golden_dir):
    """replicas=2: batches alternate between two native engines on two streams, linked by b200ocr_run_after; results
    (strings, sparse logits, coords) must be those of one engine on one stream, in the caller's order, and unlinking or
    destroying one engine of the ring must leave the other usable."""
    from pero_ocr_b200.engine import B200EngineLineOCR
    spec = cases.ENGINE_CASES['lstm']
    js = write_engine_json(tmp_path, 'lstm')
    net = make_case_net('lstm')
    rng = np.random.default_rng(11)
    lines = [rng.integers(0, 256, size=(40, int(w), 3), dtype=np.uint8) for w in rng.integers(40, 900, size=70)]
    one = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=8, precision='fp16f8', module=net)
    two = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=8, precision='fp16f8', module=net, replicas=2)
    assert len(two._models) == 2
    a = one.process_lines([l.copy() for l in lines])
    for _ in range(2):                                   # second pass: every slot and both engines reused
        b = two.process_lines([l.copy() for l in lines])
        assert a[0] == b[0] and [list(c) for c in a[2]] == [list(c) for c in b[2]]
        for x, y in zip(a[1], b[1]):
            assert x.shape == y.shape and np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices)
            assert np.array_equal(x.data, y.data)
    two._models[0].run_after(None)                       # open the ring: still correct
    assert two.process_lines([l.copy() for l in lines], no_logits=True)[0] == a[0]
    two._models[0].run_after(two._models[1])
    first = two._models.pop(0)                           # destroy one engine of the ring while linked
    two.model = two._models[0]
    two._run_streams = None
    two._slots = None
    first.close()
    assert two.process_lines([l.copy() for l in lines], no_logits=True)[0] == a[0]


@pytest.mark.parametrize('precision', ['fp16x3', 'fp16f8'])
def test_second_recogniser_family_matches_reference_golden(tmp_path, golden_dir, precision):
    """A recogniser with another module tree (nested blocks under other names, LeakyReLU slopes 0.1 / 0.2 / 0.3, 384-wide
    aggregation, ONE BiLSTM layer, Conv1d head; synthetic.LineNetLSTMAlt) hosted by the unmodified
    PytorchEngineLineOCR (tests/golden/engine_lstm_alt.npz) against the CUDA engine built by the type-driven walk
    of netdesc.describe_line_net."""
    gold = load_golden(golden_dir, 'engine_lstm_alt.npz')
    eng = _engine(tmp_path, 'lstm_alt', precision=precision)
    lines = cases.engine_lines('lstm_alt')
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    worst = 0.0
    for i in range(len(lines)):
        ref = gold[f'logits_{i}']
        assert lg[i].shape == ref.shape
        worst = max(worst, float(np.abs(lg[i] - ref).max()))
        assert list(co[i]) == list(gold[f'coords_{i}'])
        srt = np.sort(ref, axis=1)
        decided = (srt[:, -1] - srt[:, -2]) > MARGIN
        assert np.array_equal(lg[i].argmax(axis=1)[decided], ref.argmax(axis=1)[decided])
    print(f'second family, {precision}: max |d logit| = {worst:.2e}')
    assert worst <= TOL, worst
    assert tr == list(gold['transcriptions'])
    # the scripted checkpoint file walks to the same engine
    scripted = torch.jit.script(make_case_net('lstm_alt'))
    from pero_ocr_b200 import netdesc
    a, _ = netdesc.describe_line_net(make_case_net('lstm_alt'))
    b, _ = netdesc.describe_line_net(scripted)
    assert [(x['kind'], x.get('act'), x.get('act_slope'), x.get('pool_h'), x.get('pool_w')) for x in a] == \
           [(x['kind'], x.get('act'), x.get('act_slope'), x.get('pool_h'), x.get('pool_w')) for x in b]


def test_embedding_conditioned_recogniser_matches_reference_golden(tmp_path, golden_dir):
    """`model(batch, ids)` with the JSON's embed_id (pytorch_ocr_engine.py:46-50, 64-66): the unmodified reference engine
    on synthetic.LineNetLSTMEmbed with embed_id 2, with "mean", and with the attribute reassigned to 0 on the live
    engine (user_scripts/select_embed_id.py:79-80) -- against the CUDA engine, two replicas."""
    from pero_ocr_b200.engine import B200EngineLineOCR
    gold = load_golden(golden_dir, 'engine_lstm_embed.npz')
    spec = cases.ENGINE_CASES['lstm_embed']
    net = make_case_net('lstm_embed')
    lines = cases.engine_lines('lstm_embed')

    def build(embed_id, **kw):
        js = write_engine_json(tmp_path, 'lstm_embed', embed_id=embed_id)
        return B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=spec['engine_batch_size'], precision='fp16f8',
                                 module=net, **kw)

    def check(eng, tr_key, logit_key):
        tr, lg, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
        assert np.abs(lg[0] - gold[logit_key]).max() <= TOL
        assert tr == list(gold[tr_key])

    eng = build(None, replicas=2)                       # the case's own id (2) from the JSON
    assert eng.embed_id == 2 and eng.get_mean_embed_id() == int(gold['mean_embed_id'])
    check(eng, 'transcriptions', 'logits_0')
    eng.embed_id = 0                                    # reassigned on the live engine: both replicas follow
    check(eng, 'id0_transcriptions', 'id0_logits_0')
    eng.embed_id = 'mean'
    check(eng, 'mean_transcriptions', 'mean_logits_0')
    check(build('mean'), 'mean_transcriptions', 'mean_logits_0')
    with pytest.raises(IndexError):
        eng.embed_id = 6
    with pytest.raises(ValueError):                     # a table but no id / an id but no table
        B200EngineLineOCR(write_engine_json(tmp_path, 'lstm'), torch.device('cuda', 0), module=net)
    with pytest.raises(ValueError):
        B200EngineLineOCR(write_engine_json(tmp_path, 'lstm', embed_id=1), torch.device('cuda', 0), module=make_case_net('lstm'))


def test_other_lstm_hidden_size_matches_reference_golden(tmp_path, golden_dir):
    """BiLSTM hidden size 128: the tcgen05 cluster kernel is built for 256, other sizes run on the generic fp32
    recurrence kernel (kernels.cu: lstm_ref_kernel) behind the same tensor-core input projections -- against the
    unmodified reference engine's golden."""
    gold = load_golden(golden_dir, 'engine_lstm_h128.npz')
    eng = _engine(tmp_path, 'lstm_h128', precision='fp16f8')
    lines = cases.engine_lines('lstm_h128')
    tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
    worst = max(float(np.abs(lg[i] - gold[f'logits_{i}']).max()) for i in range(len(lines)))
    print(f'hidden 128: max |d logit| = {worst:.2e}')
    assert worst <= TOL, worst
    assert tr == list(gold['transcriptions'])
    assert [list(c) for c in co] == [list(gold[f'coords_{i}']) for i in range(len(lines))]
