"""Host-side mirror of the reference's autoregressive Transformer line-OCR engine, driving libb200_lineocr.so.

Replaces (same names, argument meaning and return types):
  * ``TransformerEngineLineOCR.__init__`` / ``run_ocr`` / ``transcribe_batch`` / ``postprocess_decoded`` / ``decode``
                                                  pero_ocr/ocr_engine/transformer_ocr_engine.py:12-110
  * ``BaseEngineLineOCR.process_lines`` for ``model_type == "transformer"`` (splitting of lines wider than
    ``max_line_width`` into overlapping parts and the merge of their transcriptions)
                                                  pero_ocr/ocr_engine/line_ocr_engine.py:57-177
  * ``merge_transcriptions_and_logits`` / ``find_best_overlap``        line_ocr_engine.py:180-211
  * ``levenshtein_distance`` (unit costs)                              pero_ocr/sequence_alignment.py:4-13

The encoder (``TransformerOCR.encode``) is the engine's layer walk; the greedy token loop with cached attention runs
inside one C-ABI call (``b200ocr_ar_transcribe``).  PyTorch appears only as the owner of device buffers and of the
CUDA stream and as the reader of the checkpoint file (``torch.load`` of a state dict, like the reference).
"""
import ctypes as C
import json

import numpy as np
from scipy import sparse

from . import _lib, netdesc
from ._lib import DEFAULT_PRECISION
from .engine import LineRecognizer

MIN_WIDTH = 1088          # transformer_ocr_engine.py:36-40
MAX_SEQ_LEN = 2000        # transformer.py:12 (build_net's default, which the reference engine never overrides)


class ARLineRecognizer(LineRecognizer):
    """Native engine holding the encoder half (layer walk) + the attached autoregressive decoder."""

    def __init__(self, layers, decoder, precision=DEFAULT_PRECISION, line_height=40, device=0):
        super().__init__(layers, precision=precision, line_height=line_height, device=device)
        desc, keep = netdesc.ar_to_ctypes(decoder)
        _lib.check(self._lib.b200ocr_ar_attach(self._h, C.byref(desc)), self._h)
        del keep
        self.num_classes = int(decoder['classes'])
        self._ar_reserved = (0, 0, 0)
        import os
        if os.environ.get('B200OCR_AR_LINEAR'):        # bring-up override: A/B of the step-projection kernels
            self.set_flag(2, int(os.environ['B200OCR_AR_LINEAR']))

    def reserve_ar(self, max_lines, max_width, max_steps):
        r = self._ar_reserved
        if max_lines > r[0] or max_width > r[1] or max_steps > r[2]:
            r = (max(max_lines, r[0]), max(max_width, r[1]), max(max_steps, r[2]))
            _lib.check(self._lib.b200ocr_ar_reserve(self._h, *r), self._h)
            self._ar_reserved = r
            self._reserved = (max(self._reserved[0], r[0]), max(self._reserved[1], r[1]))

    def transcribe(self, crops, start_token, max_steps=None, want_logits=True, check_every=4):
        """crops: CUDA uint8 [N,H,W,3] -> (tokens int32 CUDA [steps, N], logits float32 CUDA [N, steps, C] or None,
        steps).  Row s of `tokens` is the greedy choice at step s; `steps` is the number of iterations the
        reference's loop executes on this batch (transformer_ocr_engine.py:64-84).  Synchronises the stream."""
        torch = self.torch
        assert crops.is_cuda and crops.dtype == torch.uint8 and crops.is_contiguous() and crops.dim() == 4
        n, h, w, ch = crops.shape
        if ch != 3:
            raise ValueError('line crops need three colour channels')
        if max_steps is None:
            max_steps = w // 4 + 1                     # :79: `len(partial_transcripts) > inputs.shape[-1] // 4`
        self.reserve_ar(n, w, max_steps)
        dev = crops.device
        # persistent output buffers per shape: the library replays the launches of a position as a CUDA graph that was
        # captured with these addresses; the caller receives copies of the used prefix
        bufs = self.__dict__.setdefault('_ar_out', {})
        key = (n, max_steps, bool(want_logits), dev.index)
        if key not in bufs:
            if len(bufs) >= 4:
                bufs.clear()
            bufs[key] = (torch.empty((max_steps, n), dtype=torch.int32, device=dev),
                         torch.empty((n, max_steps, self.num_classes), dtype=torch.float32, device=dev) if want_logits else None)
        tokens, logits = bufs[key]
        steps = C.c_int32(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self._lib.b200ocr_ar_transcribe(
            self._h, crops.data_ptr(), n, h, w, int(start_token), int(max_steps), int(check_every), tokens.data_ptr(),
            logits.data_ptr() if want_logits else None, C.byref(steps), C.c_void_p(stream)), self._h)
        k = int(steps.value)
        return tokens[:k].clone(), (logits[:, :k].clone() if want_logits else None), k


def softmax(x, axis):
    """pero_ocr/ocr_engine/softmax.py:4-46 for a 2-D float array (theta = 1): same operations in the same dtype."""
    y = x - np.expand_dims(np.max(x, axis=axis), axis)
    y = np.exp(y)
    return y / np.expand_dims(np.sum(y, axis=axis), axis)


def levenshtein_distance(source, target):
    """Unit-cost edit distance (pero_ocr/sequence_alignment.py:4-13 with its default costs)."""
    dist = list(range(len(target) + 1))
    for s in source:
        prev_diag, dist[0] = dist[0], dist[0] + 1
        for j, t in enumerate(target, start=1):
            cur = min(dist[j] + 1, dist[j - 1] + 1, prev_diag + (t != s))
            prev_diag, dist[j] = dist[j], cur
    return dist[-1]


def find_best_overlap(text1, text2):
    """line_ocr_engine.py:196-211: the overlap length (>= 1) whose suffix / prefix pair has the lowest CER below 1."""
    best_cer, best_overlap = 1, 0
    for i in range(1, min(len(text1), len(text2)) + 1):
        cer = levenshtein_distance(list(text1[-i:]), list(text2[:i])) / i
        if cer < best_cer:
            best_cer, best_overlap = cer, i
    return best_overlap


def merge_transcriptions_and_logits(transcription_parts, logits_parts):
    """line_ocr_engine.py:180-193, slice arithmetic included (`-overlap // 2` floors: an overlap of 0 keeps nothing
    of the text so far -- the reference's behaviour)."""
    shrinked = [lg[:len(tr)] for tr, lg in zip(transcription_parts, logits_parts)]
    result_transcription, result_logits = transcription_parts[0], shrinked[0]
    for transcription, logits in zip(transcription_parts[1:], shrinked[1:]):
        overlap = find_best_overlap(result_transcription, transcription)
        result_transcription = result_transcription[:-overlap // 2] + transcription[overlap // 2:]
        result_logits = np.concatenate([result_logits[:-overlap // 2], logits[overlap // 2:]], axis=0)
    return result_transcription, result_logits


def postprocess_decoded(transcripts, ignore_ind, sentence_boundary_ind):
    """transformer_ocr_engine.py:91-104 on a host array [N, steps]: symbols up to the first sentence boundary, the
    ignore symbol skipped."""
    outputs = []
    for line in np.asarray(transcripts):
        stop = np.flatnonzero(line == sentence_boundary_ind)
        line = line[:stop[0]] if stop.size else line
        outputs.append(line[line != ignore_ind].astype(np.int64))
    return outputs


class B200TransformerEngineLineOCR:
    """Drop-in for ``TransformerEngineLineOCR(json_def, device, batch_size)``.

    The engine JSON is the reference's (``net_name`` = the ``build_net`` config: dim_model, dim_ff, heads,
    encoder_layers, decoder_layers, conv_subsampling; optional ``max_line_width``); ``checkpoint`` is the state dict
    of ``TransformerOCR`` that the reference ``torch.load``s (transformer_ocr_engine.py:29).  ``state_dict`` may be
    given directly instead of the file."""

    def __init__(self, json_def, device=None, batch_size=4, precision=DEFAULT_PRECISION, state_dict=None,
                 check_every=4):
        import torch
        from os.path import dirname, isabs, join, realpath
        with open(json_def, 'r', encoding='utf8') as f:
            self.config = json.load(f)
        self.line_px_height = self.config['line_px_height']
        self.line_vertical_scale = self.config['line_vertical_scale']
        ck = self.config['checkpoint']
        self.checkpoint = ck if isabs(ck) else realpath(join(dirname(json_def), ck))
        self.characters = list(self.config['characters']) + [u'\u200B', '']     # transformer_ocr_engine.py:16
        self.sentence_boundary_ind = len(self.characters) - 2                   # :18
        self.ignore_ind = len(self.characters) - 1                              # :19
        self.net_name = self.config['net_name']
        self.embed_num = int(self.config['embed_num']) if 'embed_num' in self.config else None
        self.embed_id = None
        self.max_line_width = 1e10                                              # line_ocr_engine.py:44-46
        if 'max_line_width' in self.config:
            self.max_line_width = int(self.config['max_line_width'])
            # batches of split lines are max_line_width + 128 px wide (process_lines below); the recogniser takes
            # widths that are multiples of 8 (16-byte pixel rows).  The reference accepts any value: fail here, with
            # the reason, rather than at the first over-wide line
            if self.max_line_width % 8:
                raise ValueError(f'max_line_width must be a multiple of 8 for the B200 engine (got {self.max_line_width})')
        self.model_type = 'transformer'
        self.device = device if device is not None else torch.device('cuda', 0)
        if self.device.type != 'cuda':
            raise _lib.B200Error('B200TransformerEngineLineOCR runs on a CUDA device only (no CPU fallback)')
        self.batch_size = batch_size
        self.line_padding_px = 32
        self.max_input_horizontal_pixels = 480 * batch_size
        self.net_subsampling = 4
        self.check_every = check_every
        if state_dict is None:
            state_dict = torch.load(self.checkpoint, map_location='cpu')
        layers, decoder = netdesc.describe_transformer_ocr(state_dict, self.net_name, self.line_px_height)
        if decoder['classes'] != len(self.characters):
            raise ValueError(f'net emits {decoder["classes"]} classes, engine JSON implies {len(self.characters)}')
        self.net = ARLineRecognizer(layers, decoder, precision=precision, line_height=self.line_px_height,
                                    device=self.device.index or 0)
        self.model = self.net
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ---- device step --------------------------------------------------------------------------------------
    def transcribe_batch(self, inputs, is_cached=True, no_logits=False):
        """np.uint8 [N,3,H,W] (the reference's NCHW batch) -> (list of int64 label arrays, np.float32 [N,steps,C]);
        transformer_ocr_engine.py:49-89.  Only the cached decoding path exists here (`is_cached` is accepted for
        signature compatibility; the uncached path computes the same function)."""
        torch = self.net.torch
        nhwc = np.ascontiguousarray(np.transpose(np.asarray(inputs), (0, 2, 3, 1)))
        # the reference's cached attention holds max_seq_len = 2000 positions (transformer.py:12, 241): longer memories
        # (lines of 8000 px and more) fail its assertion
        assert nhwc.shape[2] // 4 < MAX_SEQ_LEN, f'MHA: Sequence longer than {MAX_SEQ_LEN} logits'
        with torch.cuda.device(self.device):
            dev = torch.from_numpy(nhwc).to(self.device)
            self.h2d_bytes += nhwc.nbytes
            tokens, logits, steps = self.net.transcribe(dev, self.sentence_boundary_ind, max_steps=nhwc.shape[2] // 4 + 1,
                                                        want_logits=not no_logits, check_every=self.check_every)
            toks = tokens[:steps - 1].cpu().numpy()                   # partial_transcripts[1:]  (:86)
            lg = None if no_logits else logits.cpu().numpy()
        self.d2h_bytes += toks.nbytes + (0 if lg is None else lg.nbytes)
        outs = postprocess_decoded(toks.T, self.ignore_ind, self.sentence_boundary_ind)
        return outs, lg

    def run_ocr(self, batch_data, no_logits=False):
        """np.uint8 [N,H,W,3] -> (list[str], np.float32 [N,steps,C])   (transformer_ocr_engine.py:33-47): batches
        narrower than 1088 px are centred in a 1088 px canvas first."""
        batch_data = np.asarray(batch_data)
        if batch_data.ndim != 4 or batch_data.shape[3] != 3:
            raise ValueError('line crops need three colour channels')
        batch_data = np.transpose(batch_data, (0, 3, 1, 2))
        if batch_data.shape[3] < MIN_WIDTH:
            wide = np.zeros(batch_data.shape[:3] + (MIN_WIDTH,), dtype=batch_data.dtype)
            s = (MIN_WIDTH - batch_data.shape[3]) // 2
            wide[:, :, :, s:s + batch_data.shape[3]] = batch_data
            batch_data = wide
        labels, logits = self.transcribe_batch(batch_data, is_cached=True, no_logits=no_logits)
        return self.decode(labels), logits

    def decode(self, labels):
        return [''.join(self.characters[c] for c in line_labels) for line_labels in labels]

    # ---- batching ------------------------------------------------------------------------------------------
    def process_lines(self, lines, sparse_logits=True, tight_crop_logits=False, no_logits=False):
        """list of [H,w,3] uint8 crops -> (transcriptions, logits, logit_coords): line_ocr_engine.py:57-177 for
        model_type "transformer" -- widest-first batches under the pixel budget, lines wider than `max_line_width`
        split into parts overlapping by a quarter and merged on the best-matching overlap, logit_coords
        [0, len(transcription)]."""
        count = len(lines)
        all_transcriptions, all_logits, all_logit_coords = [None] * count, [None] * count, [None] * count
        pad, sub = self.line_padding_px, self.net_subsampling
        line_ids = [i for i, _ in sorted(enumerate(lines), key=lambda x: -x[1].shape[1])]
        while line_ids:
            max_width = int(np.ceil(lines[line_ids[0]].shape[1] / 32.0) * 32)
            max_width = min(max_width, self.max_line_width + 2 * pad)
            batch_size = int(max(1, self.max_input_horizontal_pixels // max_width))
            batch_line_ids, line_ids = line_ids[:batch_size], line_ids[batch_size:]
            # lines wider than max_line_width become windows of that width advancing by three quarters of it; the last
            # window is the first one that reaches the end of the line (line_ocr_engine.py:95-117)
            window = self.max_line_width
            stride = window - window // 4
            batch_images, spans = [], []
            for i in batch_line_ids:
                image = lines[i]
                if image.shape[0] != self.line_px_height or image.ndim != 3 or image.shape[2] != 3:
                    raise ValueError(f'line crops must be [{self.line_px_height}, w, 3] uint8, got {image.shape}')
                starts = [0]
                if image.shape[1] > window:
                    while starts[-1] + window < image.shape[1]:
                        starts.append(starts[-1] + stride)
                    batch_images += [image[:, s:s + window, :] for s in starts]
                else:
                    batch_images.append(image)
                spans.append(len(starts))
            batch_data = np.zeros([len(batch_images), self.line_px_height, int(max_width) + 2 * pad, 3], dtype=np.uint8)
            for data, image in zip(batch_data, batch_images):
                data[:, pad:pad + image.shape[1], :] = image
            if batch_data.shape[2] > self.max_input_horizontal_pixels:
                print(f'WARNING: Line too long for OCR engine. Cropping from {batch_data.shape[2]} px down to '
                      f'{self.max_input_horizontal_pixels}.')
                batch_data = batch_data[:, :, :self.max_input_horizontal_pixels]
            if no_logits:            # the logits are only sliced along the text by the merge: skip computing / copying them
                out_transcriptions, _ = self.run_ocr(batch_data, no_logits=True)
                out_logits = [np.zeros((len(t), 0), dtype=np.float32) for t in out_transcriptions]
            else:
                out_transcriptions, out_logits = self.run_ocr(batch_data)
            merged_transcriptions, merged_logits = [], []
            start = 0
            for span in spans:
                tr, lg = merge_transcriptions_and_logits(out_transcriptions[start:start + span],
                                                         out_logits[start:start + span])
                merged_transcriptions.append(tr)
                merged_logits.append(lg)
                start += span
            for ids, transcription, line_logits in zip(batch_line_ids, merged_transcriptions, merged_logits):
                all_transcriptions[ids] = transcription
                if no_logits:
                    continue
                if tight_crop_logits:
                    line_logits = line_logits[int(pad // sub):int((pad + lines[ids].shape[1]) // sub)]
                    all_logit_coords[ids] = [None, None]
                else:
                    all_logit_coords[ids] = [0, len(transcription)]
                if sparse_logits:
                    line_logits = np.array(line_logits, dtype=np.float32)
                    line_logits[softmax(line_logits, axis=1) < 0.0001] = 0
                    line_logits = sparse.csc_matrix(line_logits)
                all_logits[ids] = line_logits
        return all_transcriptions, all_logits, all_logit_coords
