"""CTC forced alignment on the GPU: host mirror of ``pero_ocr.core.force_alignment``.

Same function names, argument meaning and error behaviour as the reference module (force_alignment.py:13-35,
152-165); the Viterbi runs in ``b200ocr_force_align`` for a whole batch of lines at once (``force_align_batch``)
instead of a Python loop per frame and line.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib


def force_align_batch(neg_logprobs, labels, blank_symbol, n_frames=None, want_char_positions=False):
    """neg_logprobs: [N, T, C] float32 / float64 (NumPy or CUDA tensor); labels: list of N int sequences.
    -> dict(symbols [N, T] int32, positions [N, T] int32 (-1 = blank), status [N] int32, char_positions [N, Lmax]
    int32 if asked) as NumPy arrays; entries past a line's frames / characters are -1.  status: 0 ok, 1 no alignment
    with finite cost, 2 invalid transcription."""
    import torch
    lib = _lib.load_library()
    if not torch.cuda.is_available():
        raise _lib.B200Error('no CUDA device: the B200 forced alignment has no CPU fallback')
    x = neg_logprobs if torch.is_tensor(neg_logprobs) else torch.from_numpy(np.ascontiguousarray(neg_logprobs))
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64)
    x = x.cuda().contiguous()
    n, t, c = x.shape
    dev = x.device
    l_max = max([len(l) for l in labels] + [1])
    lab = np.full((n, l_max), -1, dtype=np.int32)
    for i, l in enumerate(labels):
        lab[i, :len(l)] = np.asarray(l, dtype=np.int64)
    lens = np.array([len(l) for l in labels], dtype=np.int32)
    d_lab, d_len = torch.from_numpy(lab).to(dev), torch.from_numpy(lens).to(dev)
    d_nf = torch.from_numpy(np.asarray(n_frames, dtype=np.int32)).to(dev) if n_frames is not None else None
    sym = torch.empty((n, t), dtype=torch.int32, device=dev)
    pos = torch.empty((n, t), dtype=torch.int32, device=dev)
    chp = torch.empty((n, l_max), dtype=torch.int32, device=dev) if want_char_positions else None
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.b200ocr_force_align(
        x.data_ptr(), 1 if x.dtype == torch.float64 else 0, n, t, c,
        d_nf.data_ptr() if d_nf is not None else None, d_lab.data_ptr(), l_max, d_len.data_ptr(), int(blank_symbol),
        sym.data_ptr(), pos.data_ptr(), chp.data_ptr() if chp is not None else None, status.data_ptr(),
        C.c_void_p(stream)))
    out = dict(symbols=sym.cpu().numpy(), positions=pos.cpu().numpy(), status=status.cpu().numpy())
    if chp is not None:
        out['char_positions'] = chp.cpu().numpy()
    return out


def _raise_for(status, symbols_seq, blank_symbol):
    if status == 2:
        if len(symbols_seq) < 1:
            raise ValueError("Cannot construct a CTC 'HMM' from an empty string")
        raise ValueError(f'The blank symbol {blank_symbol} is present in the non blank seq {list(symbols_seq)}')
    if status == 1:
        raise ValueError('It was not possible to align the states with the logits, best path has cost of np.inf')


def force_align(neg_logprobs, symbols_seq, blank_symbol, return_seq_positions=False):
    """force_alignment.py:13-35 for one line: list of per-frame symbols (CTC blanks included) of the most probable
    path, or per-frame character indices (-1 on blanks).  Raises ValueError like the reference."""
    neg = np.asarray(neg_logprobs)
    symbols_seq = [int(s) for s in symbols_seq]
    if len(symbols_seq) < 1:
        _raise_for(2, symbols_seq, blank_symbol)
    res = force_align_batch(neg[None], [symbols_seq], blank_symbol)
    _raise_for(int(res['status'][0]), symbols_seq, blank_symbol)
    row = res['positions'][0] if return_seq_positions else res['symbols'][0]
    return [int(v) for v in row[:neg.shape[0]]]


def align_text(neg_logprobs, transcription, blank_symbol):
    """force_alignment.py:152-165: int32 [len(transcription)] frame index of every character."""
    neg = np.asarray(neg_logprobs)
    transcription = [int(s) for s in np.asarray(transcription).reshape(-1)]
    if len(transcription) < 1:
        _raise_for(2, transcription, blank_symbol)
    res = force_align_batch(neg[None], [transcription], blank_symbol, want_char_positions=True)
    _raise_for(int(res['status'][0]), transcription, blank_symbol)
    return res['char_positions'][0, :len(transcription)].astype(np.int32)
