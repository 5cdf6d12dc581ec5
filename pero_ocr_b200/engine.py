"""Host-side mirror of the reference's line-OCR engine surface, driving libb200_lineocr.so.

Replaces (same names, argument meaning and return types):
  * ``BaseEngineLineOCR.__init__`` / ``process_lines``   pero_ocr/ocr_engine/line_ocr_engine.py:17-55, 57-177
  * ``PytorchEngineLineOCR.__init__`` / ``run_ocr``      pero_ocr/ocr_engine/pytorch_ocr_engine.py:37-74
  * ``greedy_decode_ctc``'s id -> string join            pero_ocr/ocr_engine/pytorch_ocr_engine.py:28-34
  * ``PageParser.compute_line_confidence``               pero_ocr/document_ocr/page_parser.py:486-496 (fused, optional)

PyTorch appears only as the owner of device/pinned buffers and of the CUDA stream.
"""
import ctypes as C
import json
import math
import time
from os.path import dirname, isabs, join, realpath

import numpy as np

from . import _lib, netdesc
from ._lib import DEFAULT_PRECISION
from .sparse_logits import csc_lines, sparsify_device


class LineRecognizer:
    """Thin handle over one native engine (one per device, not thread-safe -- like the reference engines)."""

    def __init__(self, layers, precision=DEFAULT_PRECISION, line_height=40, device=0):
        import torch
        self._lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise _lib.B200Error('no CUDA device: the B200 line recogniser has no CPU fallback')
        self.torch = torch
        self.device = torch.device('cuda', device)
        self.precision = precision
        self.line_height = line_height
        desc, keep = netdesc.to_ctypes(layers, precision, line_height, device)
        handle = C.c_void_p()
        torch.cuda.set_device(self.device)
        _lib.check(self._lib.b200ocr_create(C.byref(desc), C.byref(handle)))
        del keep
        self._h = handle
        self.num_classes = [l for l in layers if l['kind'] == _lib.CTC_HEAD][-1]['cout'] if any(
            l['kind'] == _lib.CTC_HEAD for l in layers) else None
        self._reserved = (0, 0)
        self._layers = layers
        self._kinds = [l['kind'] for l in layers]
        self.corrections = {}
        if precision == 'fp16f8w':       # the library's preset (engine.cu: b200ocr_create)
            self.corrections = {i: _lib.CORR_WEIGHT for i, l in enumerate(layers)
                                if l['kind'] == _lib.CONV and (l.get('kh') or 1) * (l.get('kw') or 1) == 9 and l['cin'] >= 256}

    def close(self):
        if getattr(self, '_h', None):
            self._lib.b200ocr_destroy(self._h)
            self._h = None

    __del__ = close

    def reserve(self, max_lines, max_width):
        if max_lines > self._reserved[0] or max_width > self._reserved[1]:
            n, w = max(max_lines, self._reserved[0]), max(max_width, self._reserved[1])
            _lib.check(self._lib.b200ocr_reserve(self._h, n, w), self._h)
            self._reserved = (n, w)

    def use_reference_kernels(self, on):
        _lib.check(self._lib.b200ocr_debug_use_reference_kernels(self._h, 1 if on else 0), self._h)

    def set_embedding(self, embed_id):
        """Embedding-conditioned recogniser (netdesc.describe_line_net): switch to the table row `embed_id` (int or
        "mean") -- b200ocr_set_layer_post_shift on the aggregation layer."""
        for i, l in enumerate(self._layers):
            if 'embedding_table' in l:
                table = l['embedding_table']
                shift = np.ascontiguousarray(l['embedding_base_shift'] + table[netdesc.resolve_embed_id(embed_id, table.shape[0])],
                                             dtype=np.float32)
                _lib.check(self._lib.b200ocr_set_layer_post_shift(self._h, i, shift.ctypes.data_as(C.c_void_p)), self._h)
                return
        raise ValueError('the recogniser has no embedding table')

    def run_after(self, other):
        """b200ocr_run_after: this engine's forwards start once `other`'s latest forward has finished its convolutional
        front end (None unlinks)."""
        _lib.check(self._lib.b200ocr_run_after(self._h, other._h if other is not None else None), self._h)

    def set_flag(self, flag, value):
        _lib.check(self._lib.b200ocr_debug_set_flag(self._h, int(flag), int(value)), self._h)

    @property
    def launch_count(self):
        return int(self._lib.b200ocr_launch_count(self._h))

    def flops(self, n, w):
        g = C.c_double()
        total = self._lib.b200ocr_forward_flops(self._h, n, w, C.byref(g))
        return float(total), float(g.value)

    def set_layer_correction(self, layer, mode):
        """fp16f8 engines: correction terms of one layer's e5m2 pass (_lib.CORR_BOTH / CORR_WEIGHT / CORR_NONE)."""
        _lib.check(self._lib.b200ocr_set_layer_correction(self._h, int(layer), int(mode)), self._h)
        self.corrections[int(layer)] = int(mode)

    def executed_passes(self, n, w):
        """(FLOP-weighted tensor-core pass-equivalents of the GEMM layers, per-layer list)."""
        per = np.zeros(len(self._kinds), dtype=np.float32)
        total = self._lib.b200ocr_executed_passes(self._h, n, w, len(per), per.ctypes.data_as(C.c_void_p))
        return float(total), per.tolist()

    def autotune_precision(self, budget=3e-4, sample=None, candidates=None):
        """Chooses, per layer, how much of the fp16f8 correction pass to execute, by MEASURED logit error: the
        calibration batch is run with both correction terms everywhere (the parity-grade arithmetic, ~1e-4 of the
        fp32 oracle), then layers are switched to the weight-side-only correction one at a time -- most tensor-core
        work first -- and a switch is kept while max |logits - full-correction logits| over the calibration batch stays
        within `budget` (absolute, in logit units).  `sample`: CUDA uint8 [n,H,w,3] crops representative of the
        workload (default: 8 seeded noise lines of 512 px, the bench's input statistics).  Returns a dict with the
        chosen modes, the measured deviation and the executed pass-equivalents.  Plain device work through the C ABI:
        nothing here consults the oracle."""
        torch = self.torch
        if self.precision not in ('fp16f8', 'fp16f8w'):
            raise ValueError('per-layer corrections exist in the fp16f8 precision only')
        if sample is None:
            g = torch.Generator(device='cpu').manual_seed(1234)
            gray = torch.randint(0, 256, (8, self.line_height, 512, 1), generator=g, dtype=torch.uint8)
            sample = gray.expand(-1, -1, -1, 3).contiguous().to(self.device)
        n, _, w, _ = sample.shape
        if candidates is None:
            _, per = self.executed_passes(n, w)
            flops = self.layer_gemm_flops(n, w)
            candidates = [i for i in np.argsort(flops)[::-1].tolist() if flops[i] > 0 and self._kinds[i] == _lib.CONV]
        for i in self.corrections:
            self.set_layer_correction(i, _lib.CORR_BOTH)
        base = self.forward(sample, want_logits=True)['logits'].clone()
        chosen, worst = [], 0.0
        for i in candidates:
            self.set_layer_correction(i, _lib.CORR_WEIGHT)
            dev = float((self.forward(sample, want_logits=True)['logits'] - base).abs().max().item())
            if dev <= budget:
                chosen.append(i)
                worst = dev
            else:
                self.set_layer_correction(i, _lib.CORR_BOTH)
        total, per = self.executed_passes(n, w)
        return {'weight_only_layers': sorted(chosen), 'max_abs_dev_vs_full_correction': worst, 'budget': budget,
                'executed_passes': total, 'per_layer_passes': per, 'calibration': [int(n), int(w)]}

    def layer_gemm_flops(self, n, w):
        """Algorithmic GEMM FLOPs of each layer at (n, w) (0 for layers without a contraction)."""
        out, h, cw = [], self.line_height, w
        for l in self._layers:
            k = l['kind']
            if k == _lib.CONV:
                kh, kw_ = l.get('kh', 1) or 1, l.get('kw', 1) or 1
                ho, wo = h + 2 * l.get('pad_h', 0) - kh + 1, cw + 2 * l.get('pad_w', 0) - kw_ + 1
                out.append(2.0 * n * ho * wo * l['cin'] * l['cout'] * kh * kw_)
                h, cw = ho // max(1, l.get('pool_h', 1) or 1), wo // max(1, l.get('pool_w', 1) or 1)
            elif k == _lib.BILSTM:
                out.append(2.0 * n * cw * l['cin'] * 8 * l['hidden'])
            elif k == _lib.CTC_HEAD:
                out.append(2.0 * n * cw * l['cin'] * l['cout'])
            else:
                if k == _lib.CONV_FIRST:
                    pass
                out.append(0.0)
        return out

    def forward(self, crops, want_logits=True, want_confidence=False, want_best_path=False, out=None):
        """crops: CUDA uint8 tensor [N,H,W,3] (contiguous).  Returns dict of CUDA tensors; stream-ordered on torch's
        current stream, no synchronisation."""
        torch = self.torch
        assert crops.is_cuda and crops.dtype == torch.uint8 and crops.is_contiguous() and crops.dim() == 4
        n, h, w, ch = crops.shape
        if ch != 3:
            raise ValueError('line crops need three colour channels')
        self.reserve(n, w)
        T = w // 4
        o = out if out is not None else {}
        dev = crops.device
        if 'labels' not in o or o['labels'].shape != (n, T):
            o['labels'] = torch.empty((n, T), dtype=torch.int32, device=dev)
            o['lengths'] = torch.empty((n,), dtype=torch.int32, device=dev)
        if want_logits and ('logits' not in o or o['logits'].shape != (n, T, self.num_classes)):
            o['logits'] = torch.empty((n, T, self.num_classes), dtype=torch.float32, device=dev)
        if want_confidence and ('confidence' not in o or o['confidence'].shape != (n,)):
            o['confidence'] = torch.empty((n,), dtype=torch.float32, device=dev)
        if want_best_path and ('best_path' not in o or o['best_path'].shape != (n, T)):
            o['best_path'] = torch.empty((n, T), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self._lib.b200ocr_forward(
            self._h, crops.data_ptr(), n, h, w,
            o['logits'].data_ptr() if want_logits else None,
            o['labels'].data_ptr(), o['lengths'].data_ptr(),
            o['confidence'].data_ptr() if want_confidence else None,
            o['best_path'].data_ptr() if want_best_path else None,
            C.c_void_p(stream)), self._h)
        return o

    @staticmethod
    def bytes_from_unit_floats(x):
        """f32 [N,3,H,W] in [0,1] (what run_ocr feeds the blob: uint8 / 255, NCHW, pytorch_ocr_engine.py:61-62) ->
        the uint8 [N,H,W,3] batch it came from.  Exact for every k / 255: the engine's input IS the byte batch, the
        `/255` happens inside the first convolution."""
        import torch
        return (x * 255.0).round_().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()

    def __call__(self, x):
        """The B1 seam of SURVEY.md 8(b): a stand-in for the TorchScript blob, `model(x: f32[N,3,H,W] in [0,1], CUDA)
        -> f32[N,C,T]` raw logits with the blank last (pytorch_ocr_engine.py:64-69).  Stream-ordered, no sync."""
        out = self.forward(self.bytes_from_unit_floats(x), want_logits=True)
        return out['logits'].permute(0, 2, 1)

    def profile(self, on):
        _lib.check(self._lib.b200ocr_profile(self._h, 1 if on else 0), self._h)

    def profile_read(self, capacity=4096):
        tags = np.zeros(capacity, dtype=np.int32)
        layers = np.zeros(capacity, dtype=np.int32)
        ms = np.zeros(capacity, dtype=np.float32)
        count = C.c_int32()
        _lib.check(self._lib.b200ocr_profile_read(self._h, capacity, tags.ctypes.data_as(C.c_void_p),
                                                  layers.ctypes.data_as(C.c_void_p), ms.ctypes.data_as(C.c_void_p),
                                                  C.byref(count)), self._h)
        k = min(count.value, capacity)
        return tags[:k], layers[:k], ms[:k]

    def profile_read_since(self, reference, capacity=4096):
        """Timeline variant: `reference` a recorded torch.cuda.Event(enable_timing=True); -> tags, layers, start_ms,
        end_ms relative to it."""
        tags = np.zeros(capacity, dtype=np.int32)
        layers = np.zeros(capacity, dtype=np.int32)
        t0 = np.zeros(capacity, dtype=np.float32)
        t1 = np.zeros(capacity, dtype=np.float32)
        count = C.c_int32()
        _lib.check(self._lib.b200ocr_profile_read_since(self._h, C.c_void_p(reference.cuda_event), capacity,
                                                        tags.ctypes.data_as(C.c_void_p), layers.ctypes.data_as(C.c_void_p),
                                                        t0.ctypes.data_as(C.c_void_p), t1.ctypes.data_as(C.c_void_p),
                                                        C.byref(count)), self._h)
        k = min(count.value, capacity)
        return tags[:k], layers[:k], t0[:k], t1[:k]

    def debug_forward_prefix(self, crops, n_layers):
        n, h, w, _ = crops.shape
        self.reserve(n, w)
        cap = n * h * w * 64 * 2
        buf = np.empty(cap, dtype=np.float32)
        written = C.c_int64()
        shape = (C.c_int32 * 4)()
        stream = self.torch.cuda.current_stream(crops.device).cuda_stream
        _lib.check(self._lib.b200ocr_debug_forward_prefix(
            self._h, crops.data_ptr(), n, h, w, n_layers, buf.ctypes.data_as(C.c_void_p), cap, C.byref(written),
            shape, C.c_void_p(stream)), self._h)
        return buf[:written.value].reshape(tuple(shape)).copy()


class B200EngineLineOCR:
    """Drop-in for ``PytorchEngineLineOCR(json_def, device, batch_size)``.

    The engine JSON is the reference's (keys ``line_px_height, line_vertical_scale, checkpoint, characters,
    net_name``; optional ``max_line_width``); ``checkpoint`` is the same TorchScript file the reference loads
    (the ``.cpu`` suffix rule does not apply -- the weights are read once on the host and packed for the GPU).
    Embedding-conditioned nets -- ``model(batch, ids)`` with the JSON's ``embed_id`` (an int or "mean",
    line_ocr_engine.py:36-42) -- are supported in the form netdesc.describe_line_net documents; ``embed_id`` may be
    reassigned between calls like the reference's attribute (user_scripts/select_embed_id.py:79-80).
    """

    def __init__(self, json_def, device=None, batch_size=8, precision=DEFAULT_PRECISION, module=None, replicas=1,
                 pinned_logit_bytes=4 << 30):
        import torch
        with open(json_def, 'r', encoding='utf8') as f:
            self.config = json.load(f)
        self.line_px_height = self.config['line_px_height']
        self.line_vertical_scale = self.config['line_vertical_scale']
        ck = self.config['checkpoint']
        self.checkpoint = ck if isabs(ck) else realpath(join(dirname(json_def), ck))
        self.characters = list(self.config['characters']) + [u'​']   # pytorch_ocr_engine.py:42
        self.net_name = self.config['net_name']
        self.embed_num = int(self.config['embed_num']) if 'embed_num' in self.config else None
        self._models = []
        self._embed_id = None
        if 'embed_id' in self.config:                                       # line_ocr_engine.py:36-42
            self._embed_id = 'mean' if self.config['embed_id'] == 'mean' else int(self.config['embed_id'])
        self.max_line_width = int(self.config.get('max_line_width', 1e10))
        self.model_type = 'ctc'
        self.device = device if device is not None else torch.device('cuda', 0)
        if self.device.type != 'cuda':
            raise _lib.B200Error('B200EngineLineOCR runs on a CUDA device only (no CPU fallback)')
        self.batch_size = batch_size
        self.line_padding_px = 32                                           # line_ocr_engine.py:54
        self.max_input_horizontal_pixels = 480 * batch_size                 # line_ocr_engine.py:55
        self.net_subsampling = 4                                            # pytorch_ocr_engine.py:41
        if module is None:
            module = torch.jit.load(self.checkpoint, map_location='cpu')
        layers, n_classes = netdesc.describe_line_net(module, embed_id=self._embed_id)
        # The blank is the LAST class (greedy_decode_ctc, pytorch_ocr_engine.py:27) and `characters` must name every
        # other one.  pero checkpoints emit len(JSON characters) + 1 classes -- decoder_factory's letters are the JSON
        # characters + '<BLANK>' (decoding_itf.py:49-50), so the U+200B appended at :42 sits in the blank's slot and is
        # never emitted; a net with one class more (U+200B as a real symbol, blank after it) decodes through the same
        # table.  The reference itself checks nothing (a too-small table is an IndexError in the join at :33).
        if not (len(self.characters) <= n_classes <= len(self.characters) + 1):
            raise ValueError(f'net emits {n_classes} classes; the engine JSON names {len(self.characters) - 1} characters '
                             f'(+ U+200B), which fits {len(self.characters)} or {len(self.characters) + 1} classes with '
                             f'the blank last')
        self.num_classes = n_classes
        self.model = LineRecognizer(layers, precision=precision, line_height=self.line_px_height,
                                    device=self.device.index or 0)
        # replicas = 2: a second native engine (own weights copy and workspace) on a second stream.  process_lines then
        # alternates its batches between the two, so the latency-bound tail of batch i (BiLSTM recurrence: 64 of the
        # 148 SMs) overlaps the convolutions of batch i+1; the persistent conv kernels draw tiles dynamically
        # (csrc/tilesched.cuh), which keeps the sharing work-conserving.
        self._models = [self.model] + [LineRecognizer(layers, precision=precision, line_height=self.line_px_height,
                                                      device=self.device.index or 0) for _ in range(max(1, replicas) - 1)]
        # the replicas pass the GPU to one another for their convolutional front ends (b200ocr_run_after), so that the
        # recurrence of batch i runs beside the convolutions of batch i+1 instead of beside the recurrence of batch i+1
        import os as _os
        if len(self._models) > 1 and _os.environ.get('B200OCR_LINK_REPLICAS', '1') != '0':
            for i, m in enumerate(self._models):
                m.run_after(self._models[i - 1])
        self._slots = None
        self._copy_stream = None
        # host threads that pad a batch into pinned memory (process_lines): a few when the process has the cores for
        # it, none when several ranks share the host (torchrun exports LOCAL_WORLD_SIZE)
        import os
        ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
        total = os.cpu_count() or 2
        try:
            mine = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            mine = total
        if mine < total:                 # the launcher pinned this process: those cores are ours
            cores = mine
        elif ranks > 1:                  # several ranks share the host unpinned: an equal share
            cores = total // ranks
        else:                            # the only process: leave most cores to the caller
            cores = total // 4
        self.host_threads = max(1, min(4, cores))
        self.host_ms = {'stage': 0.0, 'wait': 0.0, 'finish': 0.0}      # where process_lines spends host time
        self._executor = None
        self.want_confidence = False
        self.last_confidences = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        # page-locked, recycled host memory for the sparse logits handed to the caller (sparse_logits.PinnedPool):
        # at most `pinned_logit_bytes` registered at any time (0: always fresh pageable arrays)
        from .sparse_logits import PinnedPool
        self.pinned_pool = PinnedPool(pinned_logit_bytes) if pinned_logit_bytes > 0 else None

    def _device_ctx(self):
        return self.model.torch.cuda.device(self.device)

    @property
    def embed_id(self):
        """The reference's attribute (line_ocr_engine.py:36-42; pytorch_ocr_engine.py:46-47 resolves "mean" to the last
        row at construction): the table row of an embedding-conditioned recogniser, None otherwise.  Assignable."""
        return self._embed_id

    @embed_id.setter
    def embed_id(self, value):
        if value is None:
            if self._embed_id is not None:
                raise ValueError('an embedding-conditioned recogniser needs an embed_id')
            return
        value = 'mean' if value == 'mean' else int(value)
        with self._device_ctx():
            for m in self._models:
                m.set_embedding(value)
        self._embed_id = value

    def get_mean_embed_id(self):
        """pytorch_ocr_engine.py:49-50."""
        for l in self.model._layers:
            if 'embedding_table' in l:
                return l['embedding_table'].shape[0] - 1
        raise AttributeError('the recogniser has no embeddings_layer')

    def autotune_precision(self, budget=3e-4, sample=None):
        """LineRecognizer.autotune_precision on the first native engine; the chosen per-layer modes are applied to
        every replica."""
        with self._device_ctx():
            rep = self.model.autotune_precision(budget=budget, sample=sample)
            for other in self._models[1:]:
                for i in set(other.corrections) | set(self.model.corrections):
                    other.set_layer_correction(i, self.model.corrections.get(i, _lib.CORR_BOTH))
        return rep

    def _pool(self):
        if getattr(self, '_executor', None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._executor = ThreadPoolExecutor(max_workers=max(1, self.host_threads))
        return self._executor

    # ---- device step --------------------------------------------------------------------------------------
    def _slot(self, k):
        """Per-slot pinned staging, device input and output buffers (two slots: batch i+1 is staged on the host
        and copied on a side stream while batch i computes)."""
        if self._slots is None:
            torch = self.model.torch
            self._slots = [dict(pin=None, dev=None, outs={}, host={}, done=torch.cuda.Event(), h2d=torch.cuda.Event())
                           for _ in range(2 * len(self._models))]
            self._copy_stream = torch.cuda.Stream(self.device)
            self._run_streams = [torch.cuda.Stream(self.device) for _ in self._models] if len(self._models) > 1 else None
        return self._slots[k]

    def _stage_packed(self, sl, crops, shape, dev, main):
        """Host crops -> device batch without a padded host copy: every crop is copied ONCE, contiguously, into a
        pinned buffer ([offsets i64 n | widths i32 n | crops back to back]), one H2D copy on the side stream brings it
        over, and b200ocr_pad_lines spreads it into the zero-padded batch (line_ocr_engine.py:121-127)."""
        torch = self.model.torch
        n, height, width, _ = shape
        pad = self.line_padding_px
        t0 = time.perf_counter()
        widths = np.empty(n, dtype=np.int32)
        for i, line in enumerate(crops):
            if line.ndim != 3 or line.shape[0] != height or line.shape[2] != 3:
                raise ValueError(f'line crops must be [{height}, w, 3] uint8, got {line.shape}')
            widths[i] = line.shape[1]
        head = (n * 12 + 255) // 256 * 256
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(widths.astype(np.int64) * (3 * height), out=offs[1:])
        total = head + int(offs[n])
        if sl.get('pk_pin') is None or sl['pk_pin'].numel() < total:
            cap = int(total * 1.25) + 4096
            sl['pk_pin'] = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            sl['pk_dev'] = torch.empty(cap, dtype=torch.uint8, device=self.device)
        pin = sl['pk_pin'].numpy()
        pin[:n * 8].view(np.int64)[:] = offs[:n]
        pin[n * 8:n * 12].view(np.int32)[:] = widths
        body = pin[head:total]

        def copy_range(lo, hi):
            for i in range(lo, hi):
                line = crops[i]
                dst = body[offs[i]:offs[i + 1]].reshape(line.shape)
                if line.dtype != np.uint8:
                    raise ValueError('line crops must be uint8')
                np.copyto(dst, line)

        workers = min(self.host_threads, max(1, n // 16))
        if workers <= 1:
            copy_range(0, n)
        else:
            step = (n + workers - 1) // workers
            for j in [self._pool().submit(copy_range, lo, min(n, lo + step)) for lo in range(0, n, step)]:
                j.result()
        self.host_ms['stage'] += 1e3 * (time.perf_counter() - t0)
        pk = sl['pk_dev']
        with torch.cuda.stream(self._copy_stream):
            pk[:total].copy_(sl['pk_pin'][:total], non_blocking=True)
            sl['h2d'].record(self._copy_stream)
        main.wait_event(sl['h2d'])
        base = pk.data_ptr()
        _lib.check(self.model._lib.b200ocr_pad_lines(base + head, base, base + n * 8, n, height, dev.data_ptr(), width,
                                                     pad, C.c_void_p(main.cuda_stream)))
        self.h2d_bytes += total

    def _submit(self, k, shape, fill, no_logits, sparse_ranges=None, device_fill=None, packed=None, beam=None):
        """Stage one padded uint8 batch of `shape` -- filled in place by `fill(view)` in pinned host memory and copied
        to the device on the side stream, or spread on the device from `packed` host crops (_stage_packed), or
        produced on the device by `device_fill(dev_batch)` -- run the forward on the current stream and start the
        device->host copies of the results."""
        torch = self.model.torch
        sl = self._slot(k)
        n_bytes = int(np.prod(shape))
        if sl['dev'] is None or sl['dev'].numel() < n_bytes:
            sl['dev'] = torch.empty(n_bytes, dtype=torch.uint8, device=self.device)
            sl['pin'] = None
        dev = sl['dev'][:n_bytes].view(shape)
        model = self._models[k % len(self._models)]
        if self._run_streams is not None:
            # this slot's own stream: ordered after whatever the caller queued on the current stream so far
            main = self._run_streams[k % len(self._run_streams)]
            main.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(main):
                return self._submit_on(sl, k, model, main, dev, shape, fill, no_logits, sparse_ranges, device_fill, packed,
                                       beam)
        main = torch.cuda.current_stream(self.device)
        return self._submit_on(sl, k, model, main, dev, shape, fill, no_logits, sparse_ranges, device_fill, packed, beam)

    def _submit_on(self, sl, k, model, main, dev, shape, fill, no_logits, sparse_ranges, device_fill, packed, beam=None):
        torch = self.model.torch
        n_bytes = int(np.prod(shape))
        if device_fill is not None:
            self.h2d_bytes += int(device_fill(dev))
        elif packed is not None:
            self._stage_packed(sl, packed, shape, dev, main)
        else:
            if sl['pin'] is None or sl['pin'].numel() < n_bytes:
                sl['pin'] = torch.empty(sl['dev'].numel(), dtype=torch.uint8, pin_memory=True)
            pin = sl['pin'][:n_bytes].view(shape)
            t0 = time.perf_counter()
            fill(pin.numpy())
            self.host_ms['stage'] += 1e3 * (time.perf_counter() - t0)
            with torch.cuda.stream(self._copy_stream):
                dev.copy_(pin, non_blocking=True)
                sl['h2d'].record(self._copy_stream)
            main.wait_event(sl['h2d'])
            self.h2d_bytes += n_bytes
        t_launch = time.perf_counter()
        o = model.forward(dev, want_logits=(not no_logits) or beam is not None, want_confidence=self.want_confidence,
                          out=sl['outs'])
        sl['outs'] = o
        sl['sparse'] = None
        dense = not no_logits and sparse_ranges is None and beam is None
        if beam is not None:
            # PageDecoder's chain on the device (decode_lines): dense logits -> what get_full_logprobs() returns after the
            # sparsification -> prefix beam search on the frames [logit_coords[0], logit_coords[1]) of every line
            from .decoders import full_logprobs_device, prefix_beam_device_ranges
            beam_k, lo, hi = beam
            lp = full_logprobs_device(o['logits'])
            lo_d = torch.from_numpy(np.asarray(lo, dtype=np.int32)).to(self.device, non_blocking=True)
            hi_d = torch.from_numpy(np.asarray(hi, dtype=np.int32)).to(self.device, non_blocking=True)
            b_labels, b_lengths, b_scores, b_status = prefix_beam_device_ranges(lp, beam_k, lo_d, hi_d)
            o = dict(o, beam_labels=b_labels, beam_lengths=b_lengths, beam_scores=b_scores, beam_status=b_status)
            sl['beam_keep'] = (lp, lo_d, hi_d)          # alive until this slot is used again
        if not no_logits and sparse_ranges is not None:
            # softmax threshold + CSC on the device (line_ocr_engine.py:152-156, 168-172): only the surviving entries
            # are copied back, in _collect
            lo, hi = sparse_ranges
            sl['sparse'] = sparsify_device(o['logits'], lo, hi, out=sl.get('sparse_buf'))
            sl['sparse_buf'] = sl['sparse']
            sl['sparse'].prefetch_meta(sl.setdefault('sparse_pin', {}))
        names = ['labels', 'lengths'] + (['logits'] if dense else []) + (['confidence'] if self.want_confidence else [])
        if beam is not None:
            names += ['beam_labels', 'beam_lengths', 'beam_scores', 'beam_status']
        for name in names:
            t = o[name]
            h = sl['host'].get(name)
            if h is None or h.shape != t.shape:
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                sl['host'][name] = h
            h.copy_(t, non_blocking=True)
            self.d2h_bytes += h.numel() * h.element_size()
        sl['done'].record(main)
        self.host_ms['launch'] = self.host_ms.get('launch', 0.0) + 1e3 * (time.perf_counter() - t_launch)
        return k, names

    def _collect(self, ticket):
        k, names = ticket
        sl = self._slots[k]
        t0 = time.perf_counter()
        sl['done'].synchronize()
        self.host_ms['wait'] += 1e3 * (time.perf_counter() - t0)
        res = {name: sl['host'][name].numpy() for name in names}
        if sl.get('sparse') is not None:
            t1 = time.perf_counter()
            fetched = sl['sparse'].fetch(self._copy_stream, pinned=sl.get('sparse_pin'), pool=self._pool(),
                                         blocks=getattr(self, 'pinned_pool', None))
            self.host_ms['fetch'] = self.host_ms.get('fetch', 0.0) + 1e3 * (time.perf_counter() - t1)
            self.d2h_bytes += sum(int(getattr(a, 'array', a).nbytes) for a in fetched)
            t1 = time.perf_counter()
            res['sparse'] = csc_lines(sl['sparse'], fetched)
            self.host_ms['csc'] = self.host_ms.get('csc', 0.0) + 1e3 * (time.perf_counter() - t1)
        return res

    def _decode_ids(self, labels, lengths):
        """label ids -> strings (the join of greedy_decode_ctc, pytorch_ocr_engine.py:28-34).  A 256-line batch is
        ~50k symbols: a per-symbol Python loop costs ~10 ms per batch on the thread that also feeds the GPU, so
        single-code-point alphabets go through one NumPy gather + UTF-32 decode per line."""
        chars = self.characters
        table = getattr(self, '_codepoints', None)
        if table is None or len(table) != len(chars):
            table = (np.array([ord(c) for c in chars], dtype=np.uint32)
                     if all(isinstance(c, str) and len(c) == 1 for c in chars) else False)
            self._codepoints = table
        lens = np.asarray(lengths).tolist()
        if table is not False:
            codes = table[np.asarray(labels).clip(0)]                      # unused tail entries are -1
            try:
                return [codes[i, :ln].tobytes().decode('utf-32-le') for i, ln in enumerate(lens)]
            except UnicodeDecodeError:                                     # lone surrogates in the alphabet
                pass
        get = chars.__getitem__
        return [''.join(map(get, row[:ln])) for row, ln in zip(np.asarray(labels).tolist(), lens)]

    def run_ocr(self, batch_data, no_logits=False):
        """np.uint8 [N,H,W,3] -> (list[str], np.float32 [N,T,C])   (pytorch_ocr_engine.py:59-74)."""
        batch_data = np.asarray(batch_data)
        if batch_data.ndim != 4 or batch_data.shape[3] != 3:
            raise ValueError('line crops need three colour channels')
        with self._device_ctx():
            def fill(view):
                view[...] = batch_data
            res = self._collect(self._submit(0, batch_data.shape, fill, no_logits))
        if self.want_confidence:
            self.last_confidences = res['confidence'].copy()
        logits = None if no_logits else res['logits'].copy()
        return self._decode_ids(res['labels'], res['lengths']), logits

    # ---- batching ------------------------------------------------------------------------------------------
    def _batches(self, widths):
        """Widest-first batches under the pixel budget (line_ocr_engine.py:79-90)."""
        pending = sorted(range(len(widths)), key=lambda i: -widths[i])     # stable: ties keep input order
        while pending:
            widest = int(math.ceil(widths[pending[0]] / 32.0) * 32)
            take = max(1, self.max_input_horizontal_pixels // widest)
            chunk, pending = pending[:take], pending[take:]
            yield chunk, widest

    def _run_batches(self, widths, stager, sparse_logits, tight_crop_logits, no_logits, return_ids):
        """Common driver of process_lines / process_line_maps / process_baselines: one job through _run_jobs."""
        gen = self._run_jobs([(widths, stager)], sparse_logits, tight_crop_logits, no_logits, return_ids)
        try:
            return next(gen)
        finally:
            gen.close()         # unwinds the pipeline's frames now: they reference the results (and through them the
                                # page-locked blocks) and sit in a reference cycle that only the cyclic GC would free

    def _run_jobs(self, jobs, sparse_logits, tight_crop_logits, no_logits, return_ids):
        """Generator over `jobs` = iterable of (widths, stager): the lines of one call (or one page) each.
        `stager(chunk, width)` returns the keyword arguments (`fill=` host stager, `packed=` host crops or
        `device_fill=` device stager) that produce the padded batch of `chunk`.  Yields (transcriptions, logits,
        logit_coords) per job, in order.  The batch pipeline runs ACROSS jobs: the next job is pulled from the
        iterable and its first batches are submitted before the last batches of the current one are collected, so
        consecutive pages overlap on the GPU like consecutive batches of one call do."""
        pad, sub, height = self.line_padding_px, self.net_subsampling, self.line_px_height
        budget = self.max_input_horizontal_pixels

        def finish(job, chunk, res):
            widths = job['widths']
            labels, lengths = res['labels'], res['lengths']
            if return_ids:
                texts = [labels[s, :lengths[s]].copy() for s in range(len(chunk))]
            else:
                texts = self._decode_ids(labels, lengths)
            for slot, idx in enumerate(chunk):
                job['transcriptions'][idx] = texts[slot]
                if self.want_confidence:
                    job['confidences'][idx] = float(res['confidence'][slot])
                if no_logits:
                    continue
                lo, hi = int(pad // sub), int((pad + widths[idx]) // sub)
                if sparse_logits:
                    job['coords'][idx] = [None, None] if tight_crop_logits else [lo, hi]
                    job['logits'][idx] = res['sparse'][slot]
                    continue
                line_logits = res['logits'][slot]
                if tight_crop_logits:
                    line_logits = line_logits[lo:hi]
                    job['coords'][idx] = [None, None]
                else:
                    job['coords'][idx] = [lo, hi]
                job['logits'][idx] = line_logits.copy()
            job['open'] -= 1

        def collect_oldest(in_flight):
            job, chunk, ticket = in_flight.pop(0)
            with self._device_ctx():
                res = self._collect(ticket)
            t0 = time.perf_counter()
            finish(job, chunk, res)
            self.host_ms['finish'] += 1e3 * (time.perf_counter() - t0)

        def completed(order):
            """Jobs at the head of `order` whose batches have all been submitted and collected."""
            while order and order[0]['submitted'] and order[0]['open'] == 0:
                job = order.pop(0)
                self.last_line_confidences = job['confidences'] if self.want_confidence else None
                yield job['transcriptions'], job['logits'], job['coords']

        # batches in flight before the oldest is collected: one per native engine (replica), so that with two replicas
        # the tail of batch i, the head of batch i+1 and the staging of batch i+2 are all under way at once
        depth = len(self._models)
        nslots = 2 * depth
        in_flight, order = [], []
        bi = 0

        def ready():
            """Collects, without waiting, the batches at the head of the pipeline whose results have arrived."""
            slots = getattr(self, '_slots', None)
            while in_flight and slots is not None and slots[in_flight[0][2][0]]['done'].query():
                collect_oldest(in_flight)

        jobs = iter(jobs)
        while True:
            # pulling the next job may block (process_pages: the page's preparation): hand out what is finished first
            ready()
            yield from completed(order)
            nxt = next(jobs, None)
            if nxt is None:
                break
            widths, stager = nxt
            count = len(widths)
            job = {'widths': widths, 'transcriptions': [None] * count, 'logits': [None] * count,
                   'coords': [None] * count, 'confidences': [None] * count, 'open': 0, 'submitted': False}
            order.append(job)
            for chunk, widest in self._batches(widths):
                full_w = widest + 2 * pad
                width = full_w
                if full_w > budget:
                    print(f'WARNING: Line too long for OCR engine. Cropping from {full_w} px down to {budget}.')
                    width = budget
                ranges = None
                if sparse_logits and not no_logits:
                    t_all = width // sub
                    if tight_crop_logits:
                        ranges = ([pad // sub] * len(chunk), [(pad + widths[i]) // sub for i in chunk])
                    else:
                        ranges = ([0] * len(chunk), [t_all] * len(chunk))
                kw = stager(chunk, width)
                with self._device_ctx():
                    ticket = self._submit(bi % nslots, (len(chunk), height, width, 3), kw.get('fill'), no_logits, ranges,
                                          device_fill=kw.get('device_fill'), packed=kw.get('packed'))
                bi += 1
                job['open'] += 1
                in_flight.append((job, chunk, ticket))
                if len(in_flight) > depth:
                    collect_oldest(in_flight)
                    yield from completed(order)
            job['submitted'] = True
            yield from completed(order)
        while in_flight:
            collect_oldest(in_flight)
            yield from completed(order)

    def process_lines(self, lines, sparse_logits=True, tight_crop_logits=False, no_logits=False, return_ids=False):
        """list of [H,w,3] uint8 crops -> (transcriptions, logits, logit_coords); semantics of
        line_ocr_engine.py:57-177 for model_type 'ctc': widest-first batches under a pixel budget, 32 px zero
        padding on both sides, over-budget batches cropped, logits sparsified at softmax p < 1e-4.
        Batches are double-buffered: batch i+1 is staged and uploaded while batch i runs on the GPU; the crops cross
        PCIe packed and are zero-padded on the device (b200ocr_pad_lines)."""

        def stager(chunk, width):
            return {'packed': [lines[i] for i in chunk]}

        return self._run_batches([l.shape[1] for l in lines], stager, sparse_logits, tight_crop_logits, no_logits,
                                 return_ids)

    def process_baselines(self, page, lines, cropper, sparse_logits=True, tight_crop_logits=False, no_logits=False,
                          return_ids=False, prepared=None):
        """The page path end to end on the device: `lines` = [(baseline, heights), ...] of one page (`page` a
        cropper.DevicePage), `cropper` a B200LineCropper with a polynomial baseline fit (poly > 0).  Only ~200 bytes
        of line parameters per line are uploaded; the sampling maps, the bilinear resampling, the padding of the
        batch and the recogniser all run on the GPU.  Results are those of LineCropper.process_page followed by
        process_lines (page_parser.py:384-393, 418-430) on the reference's crops."""
        if cropper.line_height != self.line_px_height:
            raise ValueError('cropper and recogniser disagree on the line height')
        if prepared is None:
            prepared = [cropper.poly_params(b, h) for b, h in lines]
        widths, stager = self._baseline_job(page, prepared)
        return self._run_batches(widths, stager, sparse_logits, tight_crop_logits, no_logits, return_ids)

    def _baseline_job(self, page, prepared):
        """(widths, stager) of one page for _run_jobs: the crops are resampled on the device straight into the batch."""
        from .cropper import remap_poly_into

        def stager(chunk, width):
            def device_fill(dev_batch):
                return remap_poly_into(page, [prepared[i][0] for i in chunk], np.stack([prepared[i][1] for i in chunk]),
                                       dev_batch, self.line_padding_px)
            return {'device_fill': device_fill}

        return [p[0].n_out for p in prepared], stager

    def process_pages(self, pages, cropper, parsenet=None, parsenet_downsample=None, prefetch=2, **kw):
        """Pages through the page path with the stages of consecutive pages overlapped -- what
        PageParser.process_page (page_parser.py:515-531) does page after page, synchronously: `pages` is an iterable of
        (image uint8 [H, W, 3], [(baseline, heights), ...]); yields, in order, (transcriptions, logits, logit_coords,
        maps) per page (`maps` = parsenet.get_maps(image, parsenet_downsample) or None).  While page i is recognised
        on the main stream, up to `prefetch` following pages are prepared on host threads: image upload on a side
        stream, polynomial fits of the baselines (crop_engine.py:54-73), and the ParseNet forward on its own engine.
        Results are exactly those of process_baselines / get_maps called page by page."""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        from .cropper import DevicePage
        torch = self.model.torch
        lock = threading.Lock()
        fit_lock = threading.Lock()
        prefetch = max(1, int(prefetch))
        # side streams and worker threads live as long as the engine: torch's caching allocator keeps one pool per
        # stream, so fresh streams per call would mean fresh cudaMallocs for every page image
        if len(getattr(self, '_page_streams', ())) < prefetch:
            with self._device_ctx():
                self._page_streams = [torch.cuda.Stream(self.device) for _ in range(prefetch)]
        if getattr(self, '_page_workers', None) is None or self._page_workers._max_workers < prefetch:
            self._page_workers = ThreadPoolExecutor(max_workers=max(prefetch, 3), thread_name_prefix='b200ocr-page')
        streams, pool = self._page_streams[:prefetch], self._page_workers
        page_ms = self.page_ms = {'upload': 0.0, 'parsenet': 0.0, 'fit': 0.0, 'starved': 0.0}   # host time per stage

        def prepare(k, image, lines):
            t0 = time.perf_counter()
            with torch.cuda.device(self.device), torch.cuda.stream(streams[k % len(streams)]):
                page = DevicePage(image, self.device)
                ready = torch.cuda.Event()
                ready.record(streams[k % len(streams)])
                t1 = time.perf_counter()
                maps = None
                if parsenet is not None:                         # serialises its own forwards (parsenet.py)
                    maps = parsenet.get_maps(image, parsenet_downsample or parsenet.init_downsample)
            t2 = time.perf_counter()
            with fit_lock:          # interpreter-bound NumPy: three fits at once take longer than three in a row
                t2b = time.perf_counter()
                fitted = [cropper.poly_params(b, h) for b, h in lines]
            t3 = time.perf_counter()
            with lock:
                page_ms['upload'] += 1e3 * (t1 - t0); page_ms['parsenet'] += 1e3 * (t2 - t1); page_ms['fit'] += 1e3 * (t3 - t2b)
            return page, fitted, ready, maps

        if cropper.line_height != self.line_px_height:
            raise ValueError('cropper and recogniser disagree on the line height')
        it = iter(pages)
        maps_of = []                                   # per page, in order: the ParseNet maps (or None)

        def page_jobs():
            pending, k = [], 0
            first = next(it, None)
            if first is not None:
                # the first page is prepared alone: three preparations started together share the interpreter and
                # each takes three times as long, which is pure latency while the GPU still has nothing to do
                pending.append(pool.submit(prepare, k, first[0], first[1]))
                k += 1
            while pending:
                t0 = time.perf_counter()
                page, fitted, ready, maps = pending.pop(0).result()
                page_ms['starved'] += 1e3 * (time.perf_counter() - t0)
                while len(pending) < prefetch:
                    nxt = next(it, None)
                    if nxt is None:
                        break
                    pending.append(pool.submit(prepare, k, nxt[0], nxt[1]))
                    k += 1
                with self._device_ctx():
                    torch.cuda.current_stream(self.device).wait_event(ready)
                maps_of.append(maps)
                yield self._baseline_job(page, fitted)

        # one batch pipeline across the pages: the first batches of page i+1 are on the GPU before the last ones of
        # page i are collected
        flags = {'sparse_logits': True, 'tight_crop_logits': False, 'no_logits': False, 'return_ids': False}
        flags.update(kw)
        for tr, lg, co in self._run_jobs(page_jobs(), **flags):
            yield tr, lg, co, maps_of.pop(0)

    def decode_lines(self, lines, decoder):
        """Recognise and beam-decode in one pass on the device: the work of PageOCR.process_page followed by
        PageDecoder.process_page with a CTC prefix decoder without LM (page_parser.py:418-430 and 108-142) --
        raw logits -> sparsification semantics -> -80 fill -> log-softmax (core/layout.py:65-72) -> the slice
        [logit_coords[0]:logit_coords[1]] -> prefix beam search -- without the logits ever leaving the GPU.
        `decoder`: pero_ocr_b200.decoders.CTCPrefixLogRawNumpyDecoder built on the net's classes (characters +
        '<BLANK>', the letters decoder_factory builds, decoding_itf.py:49-50).  Batches are staged, run and collected
        through the same double-buffered pipeline as process_lines.
        -> list of BagOfHypotheses, one per line (what `decoder(logprobs)` returns in the reference chain)."""
        from .decoders import CTCPrefixLogRawNumpyDecoder
        if not isinstance(decoder, CTCPrefixLogRawNumpyDecoder):
            raise TypeError('decode_lines fuses the GPU prefix beam decoder (CTCPrefixLogRawNumpyDecoder, lm=None)')
        if len(decoder._letters) != self.num_classes:
            raise ValueError(f'decoder has {len(decoder._letters)} letters (blank included), the net emits '
                             f'{self.num_classes} classes')
        pad, sub, height = self.line_padding_px, self.net_subsampling, self.line_px_height
        budget = self.max_input_horizontal_pixels
        widths = [l.shape[1] for l in lines]
        bags = [None] * len(lines)

        def finish(chunk, res):
            for slot, bag in enumerate(decoder.bags_from_host(res['beam_labels'], res['beam_lengths'], res['beam_scores'],
                                                              res['beam_status'])):
                bags[chunk[slot]] = bag

        depth = len(self._models)
        nslots = 2 * depth
        in_flight = []
        with self._device_ctx():
            for bi, (chunk, widest) in enumerate(self._batches(widths)):
                width = min(widest + 2 * pad, budget)
                lo = [pad // sub] * len(chunk)
                hi = [(pad + widths[i]) // sub for i in chunk]
                ticket = self._submit(bi % nslots, (len(chunk), height, width, 3), None, True,
                                      packed=[lines[i] for i in chunk], beam=(decoder._k, lo, hi))
                in_flight.append((chunk, ticket))
                if len(in_flight) > depth:
                    done_chunk, done_ticket = in_flight.pop(0)
                    finish(done_chunk, self._collect(done_ticket))
            for done_chunk, done_ticket in in_flight:
                finish(done_chunk, self._collect(done_ticket))
        return bags

    def process_line_maps(self, page, maps, sparse_logits=True, tight_crop_logits=False, no_logits=False,
                          return_ids=False):
        """Lines given as sampling maps instead of pixels: `page` is a cropper.DevicePage (the page image, uploaded
        once), `maps` a list of float32 [H, w, 2] source-coordinate maps (B200LineCropper.get_crop_inputs, i.e.
        crop_engine.py:54-99).  Each batch is resampled on the device straight into the padded recogniser batch
        (b200ocr_remap_lines) -- LineCropper.process_page (page_parser.py:384-393) + process_lines without the host
        crops, their padding copy and their upload.  Results are those of process_lines on the reference's crops."""
        from .cropper import remap_into
        height = self.line_px_height
        for m in maps:
            if m.ndim != 3 or m.shape[0] != height or m.shape[2] != 2:
                raise ValueError(f'line maps must be [{height}, w, 2] float32, got {m.shape}')

        def stager(chunk, width):
            def device_fill(dev_batch):
                return remap_into(page, [maps[i] for i in chunk], dev_batch, self.line_padding_px)
            return {'device_fill': device_fill}

        return self._run_batches([m.shape[1] for m in maps], stager, sparse_logits, tight_crop_logits, no_logits,
                                 return_ids)
