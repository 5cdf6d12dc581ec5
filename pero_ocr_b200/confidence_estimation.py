"""Per-character confidences on the GPU: host mirror of ``pero_ocr.core.confidence_estimation.get_line_confidence``
(confidence_estimation.py:73-104), the consumer of ``align_text`` in the ALTO export (core/layout.py:489-519).

``line_confidences_batch`` runs forced alignment (``b200ocr_force_align``) and the confidence kernel
(``b200ocr_char_confidence``) for all lines of a page in two launches.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib


def line_confidences_batch(log_probs, labels, n_frames=None, aligned_letters=None):
    """log_probs: [N, T, C] float32 log-probabilities (blank = last class; NumPy or CUDA tensor); labels: list of N int
    sequences; aligned_letters: optional list of N frame-index arrays (else computed by forced alignment).
    -> (list of N float64 arrays of per-character confidences, status int32 [N] of the alignment)."""
    import torch
    lib = _lib.load_library()
    if not torch.cuda.is_available():
        raise _lib.B200Error('no CUDA device: the B200 confidence estimation has no CPU fallback')
    x = log_probs if torch.is_tensor(log_probs) else torch.from_numpy(np.ascontiguousarray(log_probs, dtype=np.float32))
    x = x.to(torch.float32).cuda().contiguous()
    n, t, c = x.shape
    dev = x.device
    l_max = max([len(l) for l in labels] + [1])
    lab = np.full((n, l_max), -1, dtype=np.int32)
    for i, l in enumerate(labels):
        lab[i, :len(l)] = np.asarray(l, dtype=np.int64)
    lens = np.array([len(l) for l in labels], dtype=np.int32)
    d_lab, d_len = torch.from_numpy(lab).to(dev), torch.from_numpy(lens).to(dev)
    d_nf = torch.from_numpy(np.asarray(n_frames, dtype=np.int32)).to(dev) if n_frames is not None else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    status = torch.zeros((n,), dtype=torch.int32, device=dev)
    if aligned_letters is None:
        neg = (-x).contiguous()
        pos = torch.empty((n, t), dtype=torch.int32, device=dev)
        chp = torch.empty((n, l_max), dtype=torch.int32, device=dev)
        _lib.check(lib.b200ocr_force_align(
            neg.data_ptr(), 0, n, t, c, d_nf.data_ptr() if d_nf is not None else None, d_lab.data_ptr(), l_max,
            d_len.data_ptr(), c - 1, None, pos.data_ptr(), chp.data_ptr(), status.data_ptr(), C.c_void_p(stream)))
    else:
        al = np.full((n, l_max), -1, dtype=np.int32)
        for i, a in enumerate(aligned_letters):
            al[i, :len(a)] = np.asarray(a, dtype=np.int64)
        chp = torch.from_numpy(al).to(dev)
    conf = torch.empty((n, l_max), dtype=torch.float32, device=dev)
    _lib.check(lib.b200ocr_char_confidence(
        x.data_ptr(), n, t, c, d_nf.data_ptr() if d_nf is not None else None, d_lab.data_ptr(), l_max, d_len.data_ptr(),
        chp.data_ptr(), conf.data_ptr(), C.c_void_p(stream)))
    host = conf.cpu().numpy()
    return [host[i, :lens[i]].astype(np.float64) for i in range(n)], status.cpu().numpy()


def get_line_confidence(line, labels, aligned_letters=None, log_probs=None):
    """confidence_estimation.py:73-104 for one line (`line` needs `.logits` and, when log_probs is None,
    `.get_full_logprobs()` -- the TextLine protocol).  The one-output-per-label shortcut of the autoregressive
    transformer engine (:76-77) is outside the CTC path."""
    if line.logits.shape[0] == len(labels):
        raise NotImplementedError('per-label logits of the autoregressive transformer engine are outside the CTC path')
    if log_probs is None:
        log_probs = line.get_full_logprobs()
    log_probs = np.asarray(log_probs)
    aligned = None if aligned_letters is None else [aligned_letters]
    conf, status = line_confidences_batch(log_probs[None], [list(labels)], aligned_letters=aligned)
    if status[0] == 1:
        raise ValueError('It was not possible to align the states with the logits, best path has cost of np.inf')
    if status[0] == 2:
        raise ValueError('invalid transcription for forced alignment (empty, or contains the blank symbol)')
    return conf[0]
