"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* by running the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

What is executed from the reference (nothing is copied; the tree is imported read-only):
  * pero_ocr.ocr_engine.pytorch_ocr_engine.PytorchEngineLineOCR / greedy_decode_ctc, hosting TorchScript
    exports of oracle/nets.py (the reference ships no recogniser definition, pytorch_ocr_engine.py:52-57)
  * pero_ocr.ocr_engine.transformer.build_net -> ConvolutionalEncoder + LineSelfAttentionEncoder, loaded with
    the same seeded weights, to pin oracle/nets.py's restatement of that architecture
  * pero_ocr.decoding.decoders.GreedyDecoder / CTCPrefixLogRawNumpyDecoder
  * pero_ocr.layout_engines.torch_parsenet.TorchParseNet.get_maps
  * pero_ocr.ocr_engine.transformer_ocr_engine.TransformerEngineLineOCR.transcribe_batch (autoregressive decoder)
  * pero_ocr.ocr_engine.line_ocr_engine.BaseEngineLineOCR.process_lines (model_type "transformer": split / merge of
    long lines) and find_best_overlap, around a deterministic run_ocr stand-in
  * pero_ocr.core.force_alignment.force_align / align_text
  * pero_ocr.core.crop_engine.EngineLineCropper.crop / get_crop_inputs (cv2.remap underneath)
  * pero_ocr.document_ocr.page_parser.PageParser.compute_line_confidence / line_confident_enough and
    pero_ocr.core.layout.TextLine.get_dense_logits / get_full_logprobs (imported behind stub modules for the
    absent lxml / shapely / skimage / arabic_reshaper packages -- SURVEY.md appendix A.4)

Inputs and weights are regenerated from seeds by the tests; only reference OUTPUTS are stored.
"""
import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

from . import cases                                   # noqa: E402
from .nets import make_net, seeded_state_dict         # noqa: E402


def _install_stubs():
    """Stand-ins for packages absent from this image; only import-time names are provided."""
    import xml.etree.ElementTree as ET

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    lx = mod('lxml')
    lx.etree = mod('lxml.etree', **{k: getattr(ET, k) for k in dir(ET) if not k.startswith('_')})
    geo = mod('shapely.geometry', LineString=_Dummy, Polygon=_Dummy, MultiPoint=_Dummy, MultiLineString=_Dummy,
              MultiPolygon=_Dummy, Point=_Dummy, GeometryCollection=_Dummy)
    geo.polygon = mod('shapely.geometry.polygon', Polygon=_Dummy)
    sh = mod('shapely', geometry=geo)
    sh.ops = mod('shapely.ops', unary_union=None, polygonize=None, nearest_points=None)
    sh.affinity = mod('shapely.affinity')
    sh.errors = mod('shapely.errors', TopologicalError=Exception)
    sk = mod('skimage')
    sk.draw = mod('skimage.draw', polygon2mask=None, polygon=None, line=None)
    sk.measure = mod('skimage.measure')
    mod('arabic_reshaper', ArabicReshaper=_Dummy)


def _export_engine(tmp, net, name, n_chars, extra=None):
    """TorchScript export + engine JSON in the form PytorchEngineLineOCR reads (line_ocr_engine.py:19-46)."""
    ck = os.path.join(tmp, name + '.pt')
    scripted = torch.jit.script(net)
    scripted.save(ck)
    scripted.save(ck + '.cpu')                       # CPU suffix rule, pytorch_ocr_engine.py:53-54
    js = os.path.join(tmp, name + '.json')
    with open(js, 'w', encoding='utf8') as f:
        json.dump({'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': name + '.pt',
                   'characters': cases.json_characters(n_chars), 'net_name': 'B200_GOLDEN', **(extra or {})}, f)
    return js


def golden_engine(kind, tmp):
    from pero_ocr.ocr_engine.pytorch_ocr_engine import PytorchEngineLineOCR
    spec = cases.ENGINE_CASES[kind]
    net = make_net(spec.get('net', kind), spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'], **spec['net_kw'])
    extra = {'embed_id': str(spec['embed_id'])} if 'embed_id' in spec else None
    js = _export_engine(tmp, net, kind, spec.get('json_chars', spec['classes'] - 2), extra)
    eng = PytorchEngineLineOCR(js, torch.device('cpu'), batch_size=spec['engine_batch_size'])
    lines = cases.engine_lines(kind)
    with contextlib.redirect_stdout(io.StringIO()):
        tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
        tr_s, lg_s, co_s = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
        tr_t, lg_t, co_t = eng.process_lines([l.copy() for l in lines], sparse_logits=False, tight_crop_logits=True)
    assert tr == tr_s == tr_t
    out = {'n': np.int64(len(lines)), 'chars': np.array(eng.characters)}
    keep = spec.get('store_logits', len(lines))            # fixture size: full logits of the first `keep` lines only
    for i in range(len(lines)):
        out[f'coords_{i}'] = np.array(co[i], dtype=np.int64)
        if i >= keep:
            continue
        out[f'logits_{i}'] = lg[i].astype(np.float32)
        if keep == len(lines):
            out[f'tight_{i}'] = lg_t[i].astype(np.float32)
            sp = lg_s[i]
            out[f'csc_data_{i}'] = sp.data
            out[f'csc_indices_{i}'] = sp.indices
            out[f'csc_indptr_{i}'] = sp.indptr
    out['transcriptions'] = np.array(tr)
    if 'embed_id' in spec:
        # the same lines under "embed_id": "mean" (pytorch_ocr_engine.py:46-50) and under another id assigned to the live
        # engine, the way user_scripts/select_embed_id.py:79-80 does
        js_mean = _export_engine(tmp, net, kind + '_mean', spec.get('json_chars', spec['classes'] - 2), {'embed_id': 'mean'})
        eng_mean = PytorchEngineLineOCR(js_mean, torch.device('cpu'), batch_size=spec['engine_batch_size'])
        out['mean_embed_id'] = np.int64(eng_mean.embed_id)
        with contextlib.redirect_stdout(io.StringIO()):
            tr_m, lg_m, _ = eng_mean.process_lines([l.copy() for l in lines], sparse_logits=False)
            eng.embed_id = 0
            tr_0, lg_0, _ = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
        out['mean_transcriptions'], out['id0_transcriptions'] = np.array(tr_m), np.array(tr_0)
        out['mean_logits_0'], out['id0_logits_0'] = lg_m[0].astype(np.float32), lg_0[0].astype(np.float32)
    if 'json_chars' in spec:
        # the reference's own decoder chain on the reference's own logits, letters as decoder_factory builds them
        # (decoding_itf.py:49-50): JSON characters + '<BLANK>'
        _install_stubs()
        from pero_ocr.core.layout import TextLine
        from pero_ocr.decoding.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder, GreedyDecoder
        letters = cases.json_characters(spec['json_chars']) + [BLANK_SYMBOL]
        gd, bd = GreedyDecoder(letters), CTCPrefixLogRawNumpyDecoder(letters, k=4)
        g_out, b_out = [], []
        for i in range(len(lines)):
            line = TextLine(id=str(i), logits=lg_s[i], logit_coords=co_s[i])
            lp = line.get_full_logprobs()[co_s[i][0]:co_s[i][1]]
            g_out.append(gd(lp).best_hyp())
            b_out.append(bd(lp.astype(np.float64)).best_hyp())
        out['decoder_greedy'] = np.array(g_out)
        out['decoder_beam4'] = np.array(b_out)
    # raw per-frame argmax of the full batch logits ("bit-exact CTC argmax indices")
    out['best_path'] = np.concatenate([l.argmax(axis=1).astype(np.int32) for l in lg])
    np.savez_compressed(os.path.join(GOLDEN, f'engine_{kind}.npz'), **out)
    margins = np.concatenate([np.sort(l, axis=1)[:, -1] - np.sort(l, axis=1)[:, -2] for l in lg])
    return {'lines': len(lines), 'transcription_lengths': [len(t) for t in tr],
            'top2_margin_median': float(np.median(margins)), 'top2_margin_min': float(margins.min()),
            'logit_absmax': float(max(np.abs(l).max() for l in lg))}


def check_reference_architecture():
    """oracle/nets.py vs the reference's own ConvolutionalEncoder + LineSelfAttentionEncoder (same weights)."""
    import torchvision
    from pero_ocr.ocr_engine import transformer as ref_tr
    orig = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **k: orig(weights=None)   # no network here; random init
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ref_tr.build_net({'dim_model': 512, 'dim_ff': 2048, 'heads': 8, 'encoder_layers': 2,
                                    'decoder_layers': 1, 'conv_subsampling': [8, 4]}, input_height=40,
                                   input_channels=3, nb_output_symbols=118).eval()
    finally:
        torchvision.models.vgg16 = orig
    ours = make_net('transformer', 120, seed=3, layers=2)
    sd = ours.state_dict()
    # map our parameter names onto the reference module tree
    ref_front = ref.encoder_frontend
    ref_convs = [m for m in ref_front.blocks_2d.modules() if isinstance(m, torch.nn.Conv2d)]
    our_convs = [m for m in ours.conv if isinstance(m, torch.nn.Conv2d)]
    assert len(ref_convs) == len(our_convs) == 9
    for r, o in zip(ref_convs, our_convs):
        assert r.weight.shape == o.weight.shape
        r.load_state_dict(o.state_dict())
    ref_bn = [m for m in ref_front.blocks_2d.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert len(ref_bn) == 1
    ref_bn[0].load_state_dict([m for m in ours.conv if isinstance(m, torch.nn.BatchNorm2d)][0].state_dict())
    ref_front.aggregation_conv[0].load_state_dict(ours.agg.state_dict())
    ref.encoder.input_norm.load_state_dict(ours.input_norm.state_dict())
    ref.encoder.trans_encoder.load_state_dict(ours.trans_encoder.state_dict())
    # structural check of the layer list (activations / pool strides), SURVEY.md section 8(a) a6
    ref_layers = [type(m).__name__ + (str(tuple(m.kernel_size)) if isinstance(m, torch.nn.MaxPool2d) else '')
                  for m in ref_front.blocks_2d.modules()
                  if isinstance(m, (torch.nn.ReLU, torch.nn.LeakyReLU, torch.nn.MaxPool2d))]
    ref_layers = [l for l in ref_layers if l != 'MaxPool2d(1, 1)']        # identity pool after block 4
    our_layers = [type(m).__name__ + (str(tuple(m.kernel_size)) if isinstance(m, torch.nn.MaxPool2d) else '')
                  for m in ours.conv if isinstance(m, (torch.nn.ReLU, torch.nn.LeakyReLU, torch.nn.MaxPool2d))]
    assert ref_layers == our_layers, (ref_layers, our_layers)
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.random((2, 3, 40, 192), dtype=np.float32))
    with torch.no_grad():
        enc_ref = ref.encode(x)                                       # [T,N,512]
        y = ours.agg_act(ours.agg(ours.conv(x))).squeeze(2).permute(2, 0, 1)
        y = ours.input_norm(y)
        enc_ours = ours.trans_encoder(y + ours.pe[:y.size(0)])
    diff = float((enc_ref - enc_ours).abs().max())
    assert diff < 1e-5, diff
    return {'encode_max_abs_diff_vs_reference_TransformerOCR.encode': diff, 'layer_list': our_layers}


def golden_decoders():
    from pero_ocr.decoding.decoders import GreedyDecoder, CTCPrefixLogRawNumpyDecoder, BLANK_SYMBOL
    from pero_ocr.ocr_engine.pytorch_ocr_engine import greedy_decode_ctc
    out = {}
    # BASELINE.json config 1: 128 x (256 x 120)
    raw, lp, letters = cases.config1_logits()
    gd = GreedyDecoder(letters)
    g_str, g_sc = [], []
    for m in lp:
        boh = gd(m)
        h = list(boh)[0]
        g_str.append(h.transcript)
        g_sc.append(h.vis_sc)
    t_str = greedy_decode_ctc(torch.from_numpy(raw.transpose(0, 2, 1).copy()), letters[:-1] + [''])
    assert t_str == g_str
    out['config1_greedy'] = np.array(g_str)
    out['config1_greedy_score'] = np.array(g_sc, dtype=np.float64)
    bd = CTCPrefixLogRawNumpyDecoder(letters, k=16)
    nb = cases.CONFIG1_BEAM_LINES
    for i in range(nb):
        boh = bd(lp[i].astype(np.float64))
        out[f'config1_beam_best_{i}'] = np.array(boh.best_hyp())
        out[f'config1_beam_hyps_{i}'] = np.array([h.transcript for h in boh])
        out[f'config1_beam_scores_{i}'] = np.array([h.vis_sc for h in boh], dtype=np.float64)
    # peaky logits (what a trained net emits): few relevant characters per frame, k = 1, 4, 16
    for name, (lp2, letters2) in cases.peaky_cases().items():
        for k in (1, 4, 16):
            bd2 = CTCPrefixLogRawNumpyDecoder(letters2, k=k)
            for i, m in enumerate(lp2):
                boh = bd2(m)
                out[f'{name}_k{k}_best_{i}'] = np.array(boh.best_hyp())
                out[f'{name}_k{k}_hyps_{i}'] = np.array([h.transcript for h in boh])
                out[f'{name}_k{k}_scores_{i}'] = np.array([h.vis_sc for h in boh], dtype=np.float64)
        g2 = GreedyDecoder(letters2)
        out[f'{name}_greedy'] = np.array([list(g2(m))[0].transcript for m in lp2])
        out[f'{name}_greedy_score'] = np.array([list(g2(m))[0].vis_sc for m in lp2], dtype=np.float64)
    # greedy_decode_ctc tie / NaN / first-frame semantics
    for name, (arr, chars) in cases.greedy_edge_cases().items():
        out[f'edge_{name}'] = np.array(greedy_decode_ctc(torch.from_numpy(arr.copy()), chars))
    np.savez_compressed(os.path.join(GOLDEN, 'decoders.npz'), **out)
    return {'config1_lines': len(g_str), 'beam_lines': nb}


def golden_confidence():
    _install_stubs()
    from pero_ocr.document_ocr.page_parser import PageParser, line_confident_enough
    from pero_ocr.core.layout import TextLine
    from pero_ocr.ocr_engine.softmax import softmax
    from scipy import sparse
    out = {}
    lg = cases.confidence_logits()
    conf, dense_all, lp_all, enough = [], [], [], []
    for i, m in enumerate(lg):
        m = m.copy()
        probs = softmax(m, axis=1)
        m[probs < 0.0001] = 0                                        # line_ocr_engine.py:168-171
        line = TextLine(id=str(i), logits=sparse.csc_matrix(m))
        conf.append(PageParser.compute_line_confidence(line))
        dense_all.append(line.get_dense_logits())
        lp_all.append(line.get_full_logprobs())
        enough.append(bool(line_confident_enough(lp_all[-1], 0.5)))
    out['confidence'] = np.array(conf, dtype=np.float64)
    out['confident_enough_0.5'] = np.array(enough)
    for i in range(len(lg)):
        out[f'dense_{i}'] = dense_all[i]
        out[f'logprobs_{i}'] = lp_all[i]
    np.savez_compressed(os.path.join(GOLDEN, 'confidence.npz'), **out)
    return {'lines': len(lg)}


def golden_parsenet(tmp):
    from pero_ocr.layout_engines.torch_parsenet import TorchParseNet
    spec = cases.PARSENET_CASE
    net = make_net('parsenet', seed=spec['seed'])
    path = os.path.join(tmp, 'parsenet.pt')
    s = torch.jit.script(net)
    s.save(path + '.cpu')
    pn = TorchParseNet(path, torch.device('cpu'), downsample=spec['downsample'], adaptive_downsample=False)
    img = cases.parsenet_image()
    with contextlib.redirect_stdout(io.StringIO()):
        maps = pn.get_maps(img, spec['downsample'])
        maps2, ds2 = pn.get_maps_with_optimal_resolution(img)
    assert np.array_equal(maps, maps2) and ds2 == spec['downsample']
    np.savez_compressed(os.path.join(GOLDEN, 'parsenet.npz'), maps=maps.astype(np.float32))
    return {'maps_shape': list(maps.shape), 'absmax': float(np.abs(maps).max())}


def golden_parsenet_page(tmp):
    """TorchParseNet.get_maps_with_optimal_resolution of the unmodified reference at BASELINE config 4's size (3000 x
    4000 page, DOWNSAMPLE 4 -> 768 x 1024 canvas) with the adaptive second pass taken (torch_parsenet.py:60-93).  The
    maps are stored subsampled (every `stride`-th pixel of both passes) -- 15 MB of float32 otherwise."""
    from pero_ocr.layout_engines.torch_parsenet import TorchParseNet
    spec = cases.PARSENET_PAGE_CASE
    net = cases.parsenet_page_net()
    path = os.path.join(tmp, 'parsenet_page.pt')
    torch.jit.script(net).save(path + '.cpu')
    pn = TorchParseNet(path, torch.device('cpu'), downsample=spec['downsample'], adaptive_downsample=True)
    img = cases.parsenet_image(spec)
    st = spec['stride']
    with contextlib.redirect_stdout(io.StringIO()):
        first = pn.get_maps(img, spec['downsample'])
        med = pn.get_med_height(first)
        maps, used = pn.get_maps_with_optimal_resolution(img)
    assert used != spec['downsample'], 'the case must take the second pass'
    np.savez_compressed(os.path.join(GOLDEN, 'parsenet_page.npz'), first=first[::st, ::st].astype(np.float32),
                        first_shape=np.array(first.shape), maps=maps[::st, ::st].astype(np.float32),
                        maps_shape=np.array(maps.shape), used_downsample=np.float64(used), med_height=np.float64(med),
                        last_downsample=np.float64(pn.last_downsample))
    return {'first_shape': list(first.shape), 'second_shape': list(maps.shape), 'used_downsample': float(used),
            'med_height_first_pass': float(med)}


def golden_cropper():
    """EngineLineCropper.crop on a seeded noise page: coordinate maps (get_crop_inputs) and crops (cv2.remap)."""
    from pero_ocr.core.crop_engine import EngineLineCropper
    from oracle.crop_oracle import CROP_CASES, page_image
    img = page_image()
    out, info = {}, {}
    for name, kw, baseline, heights in CROP_CASES:
        cropper = EngineLineCropper(**kw)
        crop = cropper.crop(img, baseline, heights)
        out[f'crop_{name}'] = crop
        try:
            full = cropper.get_crop_inputs(baseline, heights, cropper.line_height)
            out[f'map_{name}'] = full[:, ::4]          # every 4th column + a digest of the whole map (fixture size)
            out[f'mapsha_{name}'] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest(),
                                                   dtype=np.uint8)
            out[f'mapshape_{name}'] = np.array(full.shape)
        except Exception as exc:                       # the reference's crop() swallows this and returns zeros
            info[name] = f'geometry fails: {type(exc).__name__}'
        info.setdefault(name, list(crop.shape))
    np.savez_compressed(os.path.join(GOLDEN, 'cropper.npz'), **out)
    return info


def golden_align():
    """force_align / align_text of the unmodified reference on the seeded cases of oracle/align_oracle.py."""
    from pero_ocr.core.force_alignment import align_text, force_align
    from oracle.align_oracle import align_cases
    out, info = {}, {}
    for name, neg, labels, blank in align_cases():
        out[f'sym_{name}'] = np.asarray(force_align(neg, labels, blank), dtype=np.int32)
        out[f'pos_{name}'] = np.asarray(force_align(neg, labels, blank, return_seq_positions=True), dtype=np.int32)
        out[f'chr_{name}'] = np.asarray(align_text(neg, np.array(labels), blank), dtype=np.int32)
        if blank == neg.shape[1] - 1 and name != 'all_ties':
            # get_line_confidence (confidence_estimation.py:73-104) on the same line: blank must be the last class
            from pero_ocr.core.confidence_estimation import get_line_confidence
            line = types.SimpleNamespace(logits=np.zeros(neg.shape, dtype=np.float32))
            lp = (-neg).astype(np.float32)
            out[f'conf_{name}'] = np.asarray(get_line_confidence(line, np.array(labels), log_probs=lp), dtype=np.float64)
        info[name] = [int(neg.shape[0]), len(labels), str(neg.dtype)]
    np.savez_compressed(os.path.join(GOLDEN, 'align.npz'), **out)
    return info


def golden_ar_decoder():
    """TransformerEngineLineOCR.transcribe_batch (autoregressive greedy decoding with the cached decoder) of the
    unmodified reference, hosting the seeded encoder of oracle/nets.py and the seeded decoder of oracle/ar_oracle.py;
    one golden file per case of cases.AR_CASES."""
    import torchvision
    from pero_ocr.ocr_engine import transformer as ref_tr
    from pero_ocr.ocr_engine.transformer_ocr_engine import TransformerEngineLineOCR
    from oracle.ar_oracle import ar_decoder_state
    report = {}
    for name, spec in cases.AR_CASES.items():
        orig = torchvision.models.vgg16
        torchvision.models.vgg16 = lambda pretrained=False, **k: orig(weights=None)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ref = ref_tr.build_net(cases.ar_net_config(spec), input_height=40, input_channels=3,
                                       nb_output_symbols=spec['classes'] - 2).eval()
        finally:
            torchvision.models.vgg16 = orig
        ours = make_net('transformer', 120, seed=spec['encoder_seed'], layers=2)
        ref_front = ref.encoder_frontend
        for r, o in zip([m for m in ref_front.blocks_2d.modules() if isinstance(m, torch.nn.Conv2d)],
                        [m for m in ours.conv if isinstance(m, torch.nn.Conv2d)]):
            r.load_state_dict(o.state_dict())
        [m for m in ref_front.blocks_2d.modules() if isinstance(m, torch.nn.BatchNorm2d)][0].load_state_dict(
            [m for m in ours.conv if isinstance(m, torch.nn.BatchNorm2d)][0].state_dict())
        ref_front.aggregation_conv[0].load_state_dict(ours.agg.state_dict())
        ref.encoder.input_norm.load_state_dict(ours.input_norm.state_dict())
        ref.encoder.trans_encoder.load_state_dict(ours.trans_encoder.state_dict())
        dec = ar_decoder_state(seed=spec['decoder_seed'], layers=spec['decoder_layers'], classes=spec['classes'])
        missing, unexpected = ref.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()}, strict=False)
        assert not unexpected and not [k for k in missing if k.startswith(('trans_decoder', 'dec_'))], (missing, unexpected)
        fake = types.SimpleNamespace(net=ref, device=torch.device('cpu'), sentence_boundary_ind=spec['classes'] - 2,
                                     ignore_ind=spec['classes'] - 1)
        fake.postprocess_decoded = lambda *a, fake=fake: TransformerEngineLineOCR.postprocess_decoded(fake, *a)
        inputs = cases.ar_inputs(spec)                               # uint8 [N, 3, 40, W]
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            outs, logits = TransformerEngineLineOCR.transcribe_batch(fake, inputs, is_cached=True)
        n = len(outs)
        width = max([len(o) for o in outs] + [1])
        toks = np.full((n, width), -1, dtype=np.int64)
        for i, o in enumerate(outs):
            toks[i, :len(o)] = o.numpy()
        np.savez_compressed(os.path.join(GOLDEN, spec['golden']), tokens=toks,
                            lengths=np.array([len(o) for o in outs]), logits=logits.numpy().astype(np.float32))
        srt = np.sort(logits.numpy(), axis=2)
        report[name] = {'lines': n, 'steps': int(logits.shape[1]), 'lengths': [len(o) for o in outs],
                        'top2_margin_min': float((srt[..., -1] - srt[..., -2]).min())}
    return report


def golden_ar_host(tmp):
    """BaseEngineLineOCR.process_lines with model_type "transformer" (line_ocr_engine.py:57-211: split of lines wider
    than max_line_width, merge on the best overlap, logit_coords [0, len]) of the unmodified reference around the
    deterministic run_ocr stand-in of oracle/cases.py; plus the key / shape list of the state dict that the unmodified
    transformer.build_net produces (the checkpoint format of TransformerEngineLineOCR)."""
    import torchvision
    from pero_ocr.ocr_engine import transformer as ref_tr
    from pero_ocr.ocr_engine.line_ocr_engine import BaseEngineLineOCR, find_best_overlap
    spec = cases.AR_HOST_CASE
    chars = cases.json_characters(spec['classes'] - 2)
    js = os.path.join(tmp, 'ar_host.json')
    with open(js, 'w', encoding='utf8') as f:
        json.dump({'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': 'none.pt', 'characters': chars,
                   'net_name': 'x', 'max_line_width': spec['max_line_width']}, f)
    eng = BaseEngineLineOCR(js, torch.device('cpu'), batch_size=spec['batch_size'], model_type='transformer')
    eng.run_ocr = cases.ar_host_fake_run_ocr(chars)
    eng.net_subsampling = 4      # TransformerEngineLineOCR never sets it: tight_crop_logits raises AttributeError there
    lines = cases.ar_host_lines()
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        tr, lg, co = eng.process_lines([l.copy() for l in lines], sparse_logits=False)
        tr_t, lg_t, co_t = eng.process_lines([l.copy() for l in lines], sparse_logits=False, tight_crop_logits=True)
        tr_s, lg_s, co_s = eng.process_lines([l.copy() for l in lines], sparse_logits=True)
    out['transcriptions'] = np.array(tr)
    for i in range(len(lines)):
        out[f'logits_{i}'] = lg[i]
        out[f'coords_{i}'] = np.array(co[i])
        out[f'tight_{i}'] = lg_t[i]
        out[f'sparse_{i}'] = lg_s[i].toarray()
    pairs = [('hello world', 'world peace'), ('abcabc', 'bcabcd'), ('xyz', 'abc'), ('aaaa', 'aa'), ('a', 'b'),
             ('', 'abc'), ('same', 'same')]
    out['overlap_pairs'] = np.array([list(p) for p in pairs])
    out['overlaps'] = np.array([find_best_overlap(a, b) for a, b in pairs])
    np.savez_compressed(os.path.join(GOLDEN, 'ar_host.npz'), **out)
    orig = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **k: orig(weights=None)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ref_tr.build_net(cases.AR_NET_CONFIG, input_height=40, input_channels=3,
                                   nb_output_symbols=cases.AR_CASE['classes'] - 2)
    finally:
        torchvision.models.vgg16 = orig
    keys = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLDEN, 'transformer_ocr_keys.json'), 'w') as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    return {'lines': len(lines), 'transcriptions': tr, 'state_dict_keys': len(keys)}


def main():
    sys.path.insert(0, REF)
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    report = {'reference': 'DCGM/pero-ocr v0.7.0 @ /root/reference', 'torch': torch.__version__,
              'numpy': np.__version__}
    only = set(sys.argv[1:])                    # e.g. `python -m oracle.make_golden cropper` refreshes one part
    if only:
        with open(os.path.join(GOLDEN, 'REPORT.json')) as f:
            report = json.load(f)
    with tempfile.TemporaryDirectory() as tmp:
        parts = [('architecture_check', check_reference_architecture), ('decoders', golden_decoders),
                 ('engine_lstm', lambda: golden_engine('lstm', tmp)),
                 ('engine_transformer', lambda: golden_engine('transformer', tmp)),
                 ('engine_lstm_wide', lambda: golden_engine('lstm_wide', tmp)),
                 ('engine_lstm_c119', lambda: golden_engine('lstm_c119', tmp)),
                 ('engine_lstm_alt', lambda: golden_engine('lstm_alt', tmp)),
                 ('engine_lstm_embed', lambda: golden_engine('lstm_embed', tmp)),
                 ('engine_lstm_h128', lambda: golden_engine('lstm_h128', tmp)),
                 ('parsenet', lambda: golden_parsenet(tmp)), ('parsenet_page', lambda: golden_parsenet_page(tmp)),
                 ('confidence', golden_confidence),
                 ('cropper', golden_cropper), ('align', golden_align), ('ar_decoder', golden_ar_decoder),
                 ('ar_host', lambda: golden_ar_host(tmp))]
        for name, fn in parts:
            if not only or name in only:
                report[name] = fn()
    with open(os.path.join(GOLDEN, 'REPORT.json'), 'w') as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
