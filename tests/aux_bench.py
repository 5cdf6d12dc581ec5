"""Measurement of the SURVEY 8(f) rows built next to the hot path: line cropper, logit sparsification, forced alignment.

    python -m tests.aux_bench > profiles/<round>_aux_bench.json      (lives under tests/: it runs the oracle as the CPU side)

For each: device time of the kernel(s) by CUDA events on the launching stream (inputs resident, warm-up first),
algorithmic bytes moved against the measured HBM copy bandwidth (all three are byte/integer work bound by memory or
by a sequential chain, not by the tensor pipe), and the reference algorithm on one host core beside it (cv2.remap
itself when importable -- it is the reference's own callee -- else the NumPy oracle; the oracle ports otherwise).
One JSON object per line on stdout.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases                                                     # noqa: E402
from oracle.align_oracle import force_align as oracle_force_align            # noqa: E402
from oracle.crop_oracle import remap_bilinear_u8                             # noqa: E402
from oracle.forward_oracle import sparsify_logits                            # noqa: E402
from pero_ocr_b200.cropper import B200LineCropper, DevicePage, remap_into     # noqa: E402
from pero_ocr_b200.force_alignment import force_align_batch                  # noqa: E402
from pero_ocr_b200.sparse_logits import csc_lines, sparsify_device           # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('hbm_gbs', 6450.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_cropper(peak):
    """BASELINE config-4 shaped: one 3000 x 4000 page, 60 baselines of ~1280 output pixels, line height 40."""
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (3000, 4000, 3), dtype=np.uint8)
    cropper = B200LineCropper(line_height=40, poly=2, scale=1)
    lines = []
    for i in range(60):
        y = 60 + i * 48
        lines.append(([[100, y], [1400, y + rng.integers(-6, 7)], [2700, y + rng.integers(-6, 7)]], [26, 14]))
    t0 = time.perf_counter()
    maps = [cropper.get_crop_inputs(b, h, 40) for b, h in lines]
    geom_ms = 1e3 * (time.perf_counter() - t0)
    page = DevicePage(img)
    width = int(np.ceil(max(m.shape[1] for m in maps) / 32.0) * 32) + 64
    out = torch.empty((len(maps), 40, width, 3), dtype=torch.uint8, device='cuda')
    coords, offs, widths, _ = page.stage_maps(maps)
    from pero_ocr_b200 import _lib
    import ctypes as C
    lib = _lib.load_library()
    stream = torch.cuda.current_stream().cuda_stream

    def kernel():
        _lib.check(lib.b200ocr_remap_lines(page.image.data_ptr(), 3000, 4000, coords.data_ptr(), offs.data_ptr(),
                                           widths.data_ptr(), len(maps), 40, out.data_ptr(), width, 32,
                                           C.c_void_p(stream)))
    ms = timed(kernel, reps=20)
    px = sum(40 * m.shape[1] for m in maps)
    algo_bytes = px * (8 + 3 + 3) + (out.numel() - px * 3)      # map read + source pixel (each used ~once) + store; padding
    e2e_ms = timed(lambda: remap_into(page, maps, out, 32), reps=5)
    # CPU: the reference's callee on one core
    try:
        import cv2
        cv2.setNumThreads(1)
        t0 = time.perf_counter()
        for m in maps[:20]:
            cv2.remap(img, m[..., 0], m[..., 1], interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        cpu_ms, cpu_kind = 1e3 * (time.perf_counter() - t0) / 20, 'cv2.remap 4.x, 1 thread (the reference\'s callee)'
    except ImportError:
        t0 = time.perf_counter()
        for m in maps[:4]:
            remap_bilinear_u8(img, m)
        cpu_ms, cpu_kind = 1e3 * (time.perf_counter() - t0) / 4, 'NumPy oracle, 1 core'
    got = out.cpu().numpy()
    ok = bool(np.array_equal(got[7, :, 32:32 + maps[7].shape[1]], remap_bilinear_u8(img, maps[7])))
    # the same page with the maps evaluated on the device (b200ocr_remap_poly_lines): host fits the polynomial only
    from pero_ocr_b200.cropper import remap_poly_into
    t0 = time.perf_counter()
    prepared = [cropper.poly_params(b, h) for b, h in lines]
    params_ms = 1e3 * (time.perf_counter() - t0)
    par, offs_ = [p for p, _ in prepared], np.stack([o for _, o in prepared])
    out2 = torch.empty_like(out)
    poly_ms = timed(lambda: remap_poly_into(page, par, offs_, out2, 32), reps=10)
    same = bool(torch.equal(out, out2))
    return {'op': 'line cropper (b200ocr_remap_lines)', 'workload': '60 lines x 40 x ~1300 px from one 3000x4000 page',
            'gpu_ms_per_page': ms, 'gpu_us_per_line': 1e3 * ms / len(maps),
            'gpu_ms_per_page_with_map_upload': e2e_ms, 'host_geometry_ms_per_page': geom_ms,
            'device_geometry': {'gpu_ms_per_page_incl_param_upload': poly_ms, 'host_polyfit_ms_per_page': params_ms,
                                'bytes_uploaded_per_line': 112 + 40 * 8, 'identical_to_map_path': same},
            'cpu_ms_per_line': cpu_ms, 'cpu_kind': cpu_kind,
            'roofline': {'bound': 'hbm', 'achieved': algo_bytes / (ms * 1e-3) / 1e9, 'peak': peak[0], 'unit': 'GB/s',
                         'frac': algo_bytes / (ms * 1e-3) / 1e9 / peak[0], 'peak_source': peak[1],
                         'algorithmic_bytes_per_pixel': 14},
            'parity_spot_check_vs_oracle': ok}


def bench_sparsify(peak):
    """256 lines x 336 frames x 120 classes of trained-net-like (peaky) logits."""
    rng = np.random.default_rng(2)
    lp = cases.peaky_logprobs(rng, 16, 336, 120, sharp=11.0).astype(np.float32)
    raw = np.tile(lp, (16, 1, 1)) + 3.0                       # raw logits: log-probs shifted (softmax invariant)
    x = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    sp = sparsify_device(x)
    ms = timed(lambda: sparsify_device(x, out=sp), reps=10)
    total = int(sp.base.cpu()[-1])
    algo_bytes = 2 * x.numel() * 4 + total * 8 + sp.indptr.numel() * 4
    t0 = time.perf_counter()
    host = [sparsify_logits(raw[i]) for i in range(16)]
    cpu_ms = 1e3 * (time.perf_counter() - t0) / 16
    got = csc_lines(sp)
    same = all((got[i] != host[i]).nnz <= 2 for i in range(16))
    return {'op': 'logit sparsification (b200ocr_sparsify_logits)', 'workload': '256 x 336 x 120 peaky logits',
            'gpu_ms_per_batch': ms, 'gpu_us_per_line': 1e3 * ms / 256, 'entries_kept_per_frame': total / (256 * 336),
            'd2h_bytes_per_line_sparse': (total * 8 + sp.indptr.numel() * 4) / 256, 'd2h_bytes_per_line_dense': 336 * 120 * 4,
            'cpu_ms_per_line': cpu_ms, 'cpu_kind': 'NumPy + scipy.sparse oracle (the reference\'s pass), 1 core',
            'roofline': {'bound': 'hbm', 'achieved': algo_bytes / (ms * 1e-3) / 1e9, 'peak': peak[0], 'unit': 'GB/s',
                         'frac': algo_bytes / (ms * 1e-3) / 1e9 / peak[0], 'peak_source': peak[1],
                         'algorithmic_bytes_per_line': algo_bytes / 256},
            'parity_spot_check_vs_oracle': bool(same)}


def bench_align(peak):
    """256 lines x 336 frames x 120 classes, transcriptions = the greedy text of each line (~40-60 symbols)."""
    rng = np.random.default_rng(3)
    lp = cases.peaky_logprobs(rng, 32, 336, 120, sharp=10.0).astype(np.float32)
    neg = np.ascontiguousarray(np.tile(-lp, (8, 1, 1)))
    labels = []
    for i in range(256):
        best = lp[i % 32].argmax(axis=1)
        labels.append([int(v) for k, v in enumerate(best) if v != 119 and (k == 0 or best[k - 1] != v)] or [1])
    x = torch.from_numpy(neg).cuda()
    res = force_align_batch(x, labels, 119, want_char_positions=True)
    t0 = time.perf_counter()
    for _ in range(3):
        force_align_batch(x, labels, 119, want_char_positions=True)
    wall_ms = 1e3 * (time.perf_counter() - t0) / 3               # includes label upload + result download
    t0 = time.perf_counter()
    want = [oracle_force_align(neg[i], labels[i], 119) for i in range(2)]
    cpu_ms = 1e3 * (time.perf_counter() - t0) / 2
    ok = all(list(res['symbols'][i]) == want[i] for i in range(2)) and int(res['status'].max()) == 0
    mean_len = float(np.mean([len(l) for l in labels]))
    return {'op': 'CTC forced alignment (b200ocr_force_align)', 'workload': f'256 lines x 336 frames x 120 classes, mean text length {mean_len:.0f}',
            'gpu_ms_per_batch_incl_transfers': wall_ms, 'gpu_us_per_line': 1e3 * wall_ms / 256,
            'cpu_ms_per_line': cpu_ms, 'cpu_kind': 'pure-Python oracle restatement of force_alignment.py, 1 core',
            'bound': 'latency of the T sequential frames (one CTA per line; 256 lines = 2 waves)',
            'parity_spot_check_vs_oracle': bool(ok)}


def bench_config3(peak):
    """BASELINE config 3: Transformer-encoder variant + CTC prefix beam (k = 16), batch = 256 synthetic 40x1280 crops,
    through engine.decode_lines (host crops in, BagOfHypotheses out; logits never leave the GPU)."""
    import tempfile
    from pero_ocr_b200 import synthetic
    from pero_ocr_b200.decoders import BLANK_SYMBOL, CTCPrefixLogRawNumpyDecoder
    from pero_ocr_b200.engine import B200EngineLineOCR
    net = synthetic.make_net('transformer', 120, seed=0, out_gain=2.5, layers=2)
    tmp = tempfile.mkdtemp()
    js = os.path.join(tmp, 'ocr.json')
    with open(js, 'w', encoding='utf8') as f:
        json.dump({'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': 'unused.pt',
                   'characters': synthetic.json_characters(118), 'net_name': 'B200_AUX'}, f)
    eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=8, module=net)
    eng.max_input_horizontal_pixels = 256 * 1280
    dec = CTCPrefixLogRawNumpyDecoder(eng.characters + [BLANK_SYMBOL], 16)
    lines = list(synthetic.bench_crops(512, 1280, seed=0))
    eng.decode_lines(lines[:256], dec)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bags = eng.decode_lines(lines, dec)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.process_lines(lines, no_logits=True)
    dt_fwd = time.perf_counter() - t0
    return {'op': 'config 3: Transformer-encoder variant (2 layers) + CTC prefix beam k=16, decode_lines',
            'workload': '512 lines of 40x1280 in batches of 256', 'lines_per_s': len(lines) / dt,
            'ms_per_256_lines': 1e3 * dt / 2, 'forward_only_lines_per_s (process_lines, greedy)': len(lines) / dt_fwd,
            'hypotheses_per_line': float(np.mean([len(b) for b in bags]))}


def bench_config4(peak, pages=16):
    """BASELINE config 4: synthetic 4000x3000 pages through the ParseNet forward (conv-only stand-in, DOWNSAMPLE = 4 ->
    net input [1,3,768,1024]) + line OCR of an injected fixed layout of 60 baselines per page (random-init
    maps give arbitrary line counts, so the layout is fixed), the lines cropped on the device from the uploaded page
    (process_baselines: polynomial fit on the host, maps + resampling + recogniser on the GPU).  The CPU geometry
    between the two (cnn_layout_engine.py, shapely) is outside the path and not timed."""
    import tempfile
    from pero_ocr_b200 import synthetic
    from pero_ocr_b200.engine import B200EngineLineOCR
    from pero_ocr_b200.parsenet import B200ParseNet
    dev = torch.device('cuda', 0)
    pn = B200ParseNet(None, dev, downsample=4, adaptive_downsample=False, module=synthetic.make_net('parsenet', seed=1))
    net = synthetic.make_net('lstm', 120, seed=0, out_gain=2.5)
    tmp = tempfile.mkdtemp()
    js = os.path.join(tmp, 'ocr.json')
    with open(js, 'w', encoding='utf8') as f:
        json.dump({'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': 'unused.pt',
                   'characters': synthetic.json_characters(118), 'net_name': 'B200_AUX'}, f)
    eng = B200EngineLineOCR(js, dev, batch_size=8, module=net)
    eng.max_input_horizontal_pixels = 64 * 1408
    cropper = B200LineCropper(line_height=40, poly=2, scale=1)
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, (3000, 4000, 3), dtype=np.uint8) for _ in range(2)]     # alternated: 36 MB each
    lines = []
    for i in range(60):
        y = 60 + i * 48
        lines.append(([[100, y], [1400, y + rng.integers(-6, 7)], [2700, y + rng.integers(-6, 7)]], [26, 14]))
    t = {'parsenet': 0.0, 'upload': 0.0, 'ocr': 0.0}

    def one_page(img):
        t0 = time.perf_counter()
        maps = pn.get_maps(img, 4)                      # INTER_AREA resize on the host + conv forward + D2H of 5 maps
        t1 = time.perf_counter()
        page = DevicePage(img)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tr, _, _ = eng.process_baselines(page, lines, cropper, no_logits=True)
        t3 = time.perf_counter()
        t['parsenet'] += t1 - t0; t['upload'] += t2 - t1; t['ocr'] += t3 - t2
        return maps.shape, len(tr)

    one_page(imgs[0])
    one_page(imgs[1])
    for k in t:
        t[k] = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(pages):
        shape, n_lines = one_page(imgs[i & 1])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # device time of the two forwards alone
    canvas = np.zeros((1, 768, 1024, 3), dtype=np.uint8)
    canvas[0, :750, :1000] = imgs[0][::4, ::4][:750, :1000]
    pn_ms = timed(lambda: pn.net(canvas), reps=5)
    return {'op': 'config 4: page pipeline (ParseNet forward + device cropper + line OCR)',
            'workload': f'{pages} synthetic 3000x4000 pages, ParseNet stand-in at downsample 4 (maps {list(shape)}), '
                        f'{n_lines} injected baselines per page -> 40 x ~1300 px crops, CNN+BiLSTM recogniser',
            'pages_per_s': pages / dt, 'lines_per_s': pages * n_lines / dt, 'ms_per_page': 1e3 * dt / pages,
            'ms_per_page_parsenet_get_maps (host resize + upload + conv forward + D2H)': 1e3 * t['parsenet'] / pages,
            'ms_per_page_parsenet_forward_incl_canvas_upload': pn_ms,
            'ms_per_page_upload_page_image': 1e3 * t['upload'] / pages,
            'ms_per_page_crop_and_ocr (process_baselines)': 1e3 * t['ocr'] / pages}


def bench_incumbent(peak, steps=10):
    """The reference's own GPU path as the incumbent: what PytorchEngineLineOCR(json, cuda).run_ocr
    executes on a 256 x 40 x 1344 batch -- H2D of the uint8 batch, `/255`, NHWC -> NCHW, the module in PyTorch eager
    (cuDNN convolutions with TF32 allowed: torch's default, cuDNN LSTM, cuBLAS), greedy_decode_ctc's tensor part and its
    `.cpu()` (pytorch_ocr_engine.py:13-34, 59-74) -- restated here around the same seeded nn.Module, because the
    reference package itself cannot travel to the GPU box.  Library code on purpose: this is the baseline, not the
    product.  `with_logits` adds run_ocr's second result (the [N,C,T] -> [N,T,C] logits download, :72)."""
    from pero_ocr_b200 import synthetic
    net = synthetic.make_net('lstm', 120, seed=0, out_gain=2.5).cuda()
    crops = synthetic.bench_crops(256, 1280, seed=0)
    batch = np.zeros((256, 40, 1344, 3), dtype=np.uint8)
    batch[:, :, 32:32 + 1280] = crops
    host = torch.from_numpy(batch).pin_memory()
    chars = synthetic.json_characters(118) + ['\u200b']

    def run_ocr(with_logits):
        with torch.no_grad():
            x = host.to('cuda', non_blocking=True).float()
            x /= 255.0
            logits = net(x.permute(0, 3, 1, 2))                                   # [N,C,T]
            sp = torch.cat((logits[:, :, 0:1], logits), dim=2)                    # greedy_decode_ctc, :19-27
            sp[:, :, 0] = -1000
            sp[:, -1, 0] = 1000
            best = torch.argmax(sp, 1) + 1
            mask = best[:, :-1] == best[:, 1:]
            best = best[:, 1:]
            best[mask] = 0
            best[best == sp.shape[1]] = 0
            best = best.cpu().numpy() - 1
            out = [''.join(chars[c] for c in line[np.nonzero(line >= 0)]) for line in best]
            lg = logits.permute(0, 2, 1).cpu().numpy() if with_logits else None
        return out, lg

    res = {}
    for name, with_logits in (('no_logits', False), ('with_logits', True)):
        for _ in range(3):
            run_ocr(with_logits)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            run_ocr(with_logits)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = {'lines_per_s': 256 * steps / dt, 'ms_per_256_lines': 1e3 * dt / steps}
    return {'op': 'incumbent: the reference\'s GPU path (PyTorch eager run_ocr on the same net, cuDNN/cuBLAS)',
            'workload': 'config 2: 256 x 40 x 1344 uint8 batch from pinned host memory per step',
            'torch': torch.__version__, 'cudnn_allow_tf32': bool(torch.backends.cudnn.allow_tf32),
            'matmul_allow_tf32': bool(torch.backends.cuda.matmul.allow_tf32), **res}


def main():
    assert torch.cuda.is_available()
    peak = hbm_peak()
    benches = {'cropper': bench_cropper, 'sparsify': bench_sparsify, 'align': bench_align, 'config3': bench_config3,
               'config4': bench_config4, 'incumbent': bench_incumbent}
    for name in (sys.argv[1:] or list(benches)):             # e.g. `python -m tests.aux_bench config4`
        print(json.dumps(benches[name](peak)), flush=True)


if __name__ == '__main__':
    main()
