"""GPU CTC decoders behind the reference's decoder interface.

Replaces (same constructor arguments, same call signature, same result object protocol):
  * ``GreedyDecoder``                      pero_ocr/decoding/decoders.py:35-62
  * ``CTCPrefixLogRawNumpyDecoder`` (lm=None) pero_ocr/decoding/decoders.py:170-299
  * ``BagOfHypotheses``                    pero_ocr/decoding/bag_of_hypotheses.py:11-65
  * ``greedy_decode_ctc``                  pero_ocr/ocr_engine/pytorch_ocr_engine.py:13-34

``decoder(logprobs[T,C]) -> BagOfHypotheses`` decodes one line like the reference; ``decode_batch`` takes many lines
at once (one CTA per line on the GPU), which is how PageDecoder-style callers should use it.
"""
import ctypes as C
import math
from collections import namedtuple

import numpy as np

from . import _lib

BLANK_SYMBOL = '<BLANK>'
Hypothese = namedtuple('Hypothese', 'transcript vis_sc lm_sc')


class BagOfHypotheses:
    def __init__(self, lm_weight=1.0):
        self._hyps = []
        self.lm_weight = lm_weight

    def add(self, transcript, visual_sc, lm_sc=None):
        self._hyps.append(Hypothese(transcript, visual_sc, lm_sc))

    def sort(self):
        self._hyps.sort(key=lambda h: h.vis_sc, reverse=True)

    def __iter__(self):
        return iter(self._hyps)

    def __len__(self):
        return len(self._hyps)

    def total_scores(self):
        return [h.vis_sc + (self.lm_weight * h.lm_sc if h.lm_sc is not None else 0.0) for h in self._hyps]

    def posteriors(self):
        scores = np.asarray(self.total_scores(), dtype=np.float64)
        return list(scores - np.logaddexp.reduce(scores))

    def confidence(self):
        return math.exp(max(self.posteriors()))

    def transcript_confidence(self, transcript):
        for h, p in zip(self._hyps, self.posteriors()):
            if h.transcript == transcript:
                return math.exp(p)
        return 0.0

    def best_hyp(self):
        return max(self._hyps, key=lambda h: h.vis_sc + (h.lm_sc if h.lm_sc is not None else 0)).transcript


def _check_letters(letters):
    seen, dup = set(), []
    for x in letters:
        if x in seen:
            dup.append(x)
        seen.add(x)
    if dup:
        raise ValueError(f'Letters contain these duplicit elements: {dup}')
    at = letters.index(BLANK_SYMBOL)
    if at != len(letters) - 1:
        raise ValueError(f"Expected {BLANK_SYMBOL} as the last of letters, it's instead at position {at}")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.B200Error('no CUDA device: the B200 CTC decoders have no CPU fallback')
    return torch


def _stream(torch, device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def greedy_ids_device(x, layout='ntc', want_confidence=False):
    """Device-resident form: x is a contiguous CUDA float32 tensor; returns CUDA tensors, no synchronisation."""
    torch = _torch()
    lib = _lib.load_library()
    x = x.contiguous()
    dev = x.device
    if layout == 'ntc':
        n, t, c = x.shape
        lay = 0
    else:
        n, c, t = x.shape
        lay = 1
    out = dict(labels=torch.empty((n, t), dtype=torch.int32, device=dev),
               lengths=torch.empty((n,), dtype=torch.int32, device=dev),
               best_path=torch.empty((n, t), dtype=torch.int32, device=dev),
               frame_max=torch.empty((n, t), dtype=torch.float32, device=dev),
               frame_lse=torch.empty((n, t), dtype=torch.float32, device=dev))
    if want_confidence:
        out['confidence'] = torch.empty((n,), dtype=torch.float32, device=dev)
    _lib.check(lib.b200ocr_ctc_greedy(x.data_ptr(), n, t, c, lay, out['labels'].data_ptr(), out['lengths'].data_ptr(),
                                      out['confidence'].data_ptr() if want_confidence else None,
                                      out['best_path'].data_ptr(), out['frame_max'].data_ptr(),
                                      out['frame_lse'].data_ptr(), _stream(torch, dev)))
    return out


def prefix_beam_device(x, k):
    """x: contiguous CUDA float64 [N,T,C] normalised log-probs -> CUDA (labels [N,k,T], lengths [N,k], scores [N,k],
    status [N]); no synchronisation."""
    torch = _torch()
    lib = _lib.load_library()
    x = x.contiguous()
    n, t, c = x.shape
    dev = x.device
    labels = torch.empty((n, k, t), dtype=torch.int32, device=dev)
    lengths = torch.empty((n, k), dtype=torch.int32, device=dev)
    scores = torch.empty((n, k), dtype=torch.float64, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.b200ocr_ctc_prefix_beam(x.data_ptr(), n, t, c, k, labels.data_ptr(), lengths.data_ptr(),
                                           scores.data_ptr(), status.data_ptr(), _stream(torch, dev)))
    return labels, lengths, scores, status


def full_logprobs_device(logits):
    """logits: contiguous CUDA float32 [N,T,C] raw recogniser outputs -> CUDA float64 [N,T,C]: what
    ``TextLine.get_full_logprobs()`` returns after the engine's sparsification (line_ocr_engine.py:168-172,
    core/layout.py:65-72), computed on the device; no synchronisation."""
    torch = _torch()
    lib = _lib.load_library()
    logits = logits.contiguous()
    n, t, c = logits.shape
    out = torch.empty((n, t, c), dtype=torch.float64, device=logits.device)
    _lib.check(lib.b200ocr_full_logprobs(logits.data_ptr(), n, t, c, out.data_ptr(), _stream(torch, logits.device)))
    return out


def prefix_beam_device_ranges(x, k, t_lo, t_hi):
    """prefix_beam_device on the frame range [t_lo[i], t_hi[i]) of every line (CUDA int32 [N] tensors)."""
    torch = _torch()
    lib = _lib.load_library()
    x = x.contiguous()
    n, t, c = x.shape
    dev = x.device
    labels = torch.empty((n, k, t), dtype=torch.int32, device=dev)
    lengths = torch.empty((n, k), dtype=torch.int32, device=dev)
    scores = torch.empty((n, k), dtype=torch.float64, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.b200ocr_ctc_prefix_beam_ranges(x.data_ptr(), n, t, c, k, t_lo.data_ptr(), t_hi.data_ptr(),
                                                  labels.data_ptr(), lengths.data_ptr(), scores.data_ptr(),
                                                  status.data_ptr(), _stream(torch, dev)))
    return labels, lengths, scores, status


def greedy_ids(scores, layout='ntc', want_confidence=False, device=None):
    """scores: np.ndarray or CUDA tensor, [N,T,C] ('ntc') or [N,C,T] ('nct'), blank = last class.
    -> dict(labels [N,T] i32 left-packed -1 padded, lengths [N], best_path [N,T], frame_max, frame_lse[, confidence])
    as numpy arrays.  Tie / NaN rules = torch.argmax (first maximal index, NaN maximal)."""
    torch = _torch()
    lib = _lib.load_library()
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    x = scores if hasattr(scores, 'is_cuda') else torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
    x = x.to(dev, dtype=torch.float32).contiguous()
    if layout == 'ntc':
        n, t, c = x.shape
        lay = 0
    else:
        n, c, t = x.shape
        lay = 1
    labels = torch.empty((n, t), dtype=torch.int32, device=dev)
    lengths = torch.empty((n,), dtype=torch.int32, device=dev)
    best = torch.empty((n, t), dtype=torch.int32, device=dev)
    fmax = torch.empty((n, t), dtype=torch.float32, device=dev)
    flse = torch.empty((n, t), dtype=torch.float32, device=dev)
    conf = torch.empty((n,), dtype=torch.float32, device=dev) if want_confidence else None
    _lib.check(lib.b200ocr_ctc_greedy(x.data_ptr(), n, t, c, lay, labels.data_ptr(), lengths.data_ptr(),
                                      conf.data_ptr() if conf is not None else None, best.data_ptr(),
                                      fmax.data_ptr(), flse.data_ptr(), _stream(torch, dev)))
    out = dict(labels=labels.cpu().numpy(), lengths=lengths.cpu().numpy(), best_path=best.cpu().numpy(),
               frame_max=fmax.cpu().numpy(), frame_lse=flse.cpu().numpy())
    if conf is not None:
        out['confidence'] = conf.cpu().numpy()
    return out


def greedy_decode_ctc(scores_probs, chars):
    """[N,C,T] (or [C,T]) scores -> list of strings; blank is the last class (pytorch_ocr_engine.py:13-34)."""
    x = scores_probs
    if len(x.shape) == 2:
        x = x[None]
    r = greedy_ids(x, layout='nct')
    return [''.join(chars[c] for c in row[:ln]) for row, ln in zip(r['labels'], r['lengths'])]


class GreedyDecoder:
    def __init__(self, letters, symbol_separator=''):
        _check_letters(letters)
        self._letters = letters
        self._blank_ind = letters.index(BLANK_SYMBOL)
        self.symbol_separator = symbol_separator

    def decode_batch(self, batch, max_unnormalization=1e-5):
        """batch: list of [T_i,C] log-prob arrays (ragged) or one [N,T,C] array."""
        mats = [np.asarray(m) for m in batch]
        if not mats:
            return []
        t_max = max(m.shape[0] for m in mats)
        c = mats[0].shape[1]
        packed = np.full((len(mats), t_max, c), 0.0, dtype=np.float32)
        packed[:, :, c - 1] = 1.0                       # padding frames decode to blank
        for i, m in enumerate(mats):
            packed[i, :m.shape[0]] = m
        r = greedy_ids(packed, layout='ntc')
        bags = []
        for i, m in enumerate(mats):
            t = m.shape[0]
            # normalisation gate of the reference (decoders.py:49-51): exp(lse) must be 1 within 1e-5
            if t and np.max(np.abs(np.exp(r['frame_lse'][i, :t].astype(np.float64)) - 1)) > max_unnormalization:
                raise ValueError('Expected properly normalized logits')
            ids = r['labels'][i, :r['lengths'][i]]
            text = self.symbol_separator.join(self._letters[k] for k in ids)
            maxes = r['frame_max'][i, :t].astype(np.float64)
            bag = BagOfHypotheses()
            bag.add(text, float(np.logaddexp.reduce(maxes)) if t else -np.inf)   # sic: decoders.py:60
            bags.append(bag)
        return bags

    def __call__(self, logits, max_unnormalization=1e-5):
        return self.decode_batch([logits], max_unnormalization)[0]


class CTCPrefixLogRawNumpyDecoder:
    """Prefix beam search without a language model (``lm=None``); passing an LM raises."""

    def __init__(self, letters, k, lm=None, lm_scale=1.0, insertion_bonus=0.0, symbol_separator=''):
        _check_letters(letters)
        if not isinstance(k, int):
            raise TypeError("Beam size 'k' has to be int, got {} instead (value: {}).".format(type(k), k))
        if k < 1:
            raise ValueError("Beam size 'k' has to be positive, got {} instead.".format(k))
        if lm is not None:
            raise NotImplementedError('LM fusion depends on the un-vendored brnolm package and is out of scope')
        self._letters = letters
        self._k = k
        self._blank_ind = letters.index(BLANK_SYMBOL)
        self.symbol_separator = symbol_separator

    def decode_batch(self, batch):
        torch = _torch()
        lib = _lib.load_library()
        mats = [np.asarray(m, dtype=np.float64) for m in batch]
        if not mats:
            return []
        dev = torch.device('cuda', torch.cuda.current_device())
        bags = [None] * len(mats)
        # one launch per distinct length: padding frames would change the search
        by_len = {}
        for i, m in enumerate(mats):
            by_len.setdefault(m.shape[0], []).append(i)
        for t, idxs in by_len.items():
            c = mats[idxs[0]].shape[1]
            n, k = len(idxs), self._k
            if t == 0:
                for i in idxs:
                    bag = BagOfHypotheses()
                    bag.add('', 0.0, 0)
                    bags[i] = bag
                continue
            # C order explicitly: get_full_logprobs() of a CSC matrix is Fortran-ordered (scipy's toarray), and np.stack
            # of Fortran-ordered inputs keeps that order -- the kernel reads [n][t][c]
            x = torch.from_numpy(np.ascontiguousarray(np.stack([mats[i] for i in idxs]))).to(dev)
            labels = torch.empty((n, k, t), dtype=torch.int32, device=dev)
            lengths = torch.empty((n, k), dtype=torch.int32, device=dev)
            scores = torch.empty((n, k), dtype=torch.float64, device=dev)
            status = torch.empty((n,), dtype=torch.int32, device=dev)
            _lib.check(lib.b200ocr_ctc_prefix_beam(x.data_ptr(), n, t, c, k, labels.data_ptr(), lengths.data_ptr(),
                                                   scores.data_ptr(), status.data_ptr(), _stream(torch, dev)))
            for row, bag in enumerate(self.bags_from_device(labels, lengths, scores, status)):
                bags[idxs[row]] = bag
        return bags

    def bags_from_device(self, labels, lengths, scores, status):
        """Device results of b200ocr_ctc_prefix_beam[_ranges] -> list of BagOfHypotheses (one per line)."""
        return self.bags_from_host(labels.cpu().numpy(), lengths.cpu().numpy(), scores.cpu().numpy(),
                                   status.cpu().numpy())

    def bags_from_host(self, labels, lengths, scores, status):
        """The same on host arrays (labels [n, k, T] int32, lengths [n, k], scores [n, k], status [n]).  Alphabets of
        single code points are joined with one NumPy gather + UTF-32 decode per hypothesis (a 256-line batch at k = 16
        is 4096 hypotheses: a per-symbol Python join costs ~20 ms of the feeding thread per batch)."""
        table = getattr(self, '_codepoints', None)
        if table is None:
            ok = self.symbol_separator == '' and all(isinstance(c, str) and len(c) == 1 for c in self._letters[:-1])
            table = np.array([ord(c) for c in self._letters[:-1]] + [0], dtype=np.uint32) if ok else False
            self._codepoints = table
        if table is not False:
            bags = []
            lens = np.asarray(lengths)
            for row in range(labels.shape[0]):
                if status[row] != 0:
                    raise ValueError('Expected properly normalized logits')
                bag = BagOfHypotheses()
                codes = table[np.asarray(labels[row]).clip(0, len(table) - 1)]
                try:
                    for b in range(labels.shape[1]):
                        ln = int(lens[row, b])
                        if ln >= 0:
                            bag.add(codes[b, :ln].tobytes().decode('utf-32-le'), float(scores[row, b]), 0)
                except UnicodeDecodeError:
                    table = self._codepoints = False
                    break
                bag.sort()
                bags.append(bag)
            else:
                return bags
        bags = []
        for row in range(labels.shape[0]):
            if status[row] != 0:
                raise ValueError('Expected properly normalized logits')
            bag = BagOfHypotheses()
            for b in range(labels.shape[1]):
                ln = lengths[row, b]
                if ln < 0:
                    continue
                text = self.symbol_separator.join(self._letters[j] for j in labels[row, b, :ln])
                bag.add(text, float(scores[row, b]), 0)
            bag.sort()
            bags.append(bag)
        return bags

    def __call__(self, logits, model_eos=False, max_unnormalization=1e-5, return_h=False, init_h=None):
        if model_eos or return_h or init_h is not None:
            raise NotImplementedError('LM-related options need the un-vendored brnolm package')
        return self.decode_batch([logits])[0]
