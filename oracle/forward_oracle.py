"""TEST INFRASTRUCTURE ONLY -- functional torch-CPU restatement of the recogniser hot path.

Restates, op by op:
  * ``PytorchEngineLineOCR.run_ocr``        pero_ocr/ocr_engine/pytorch_ocr_engine.py:59-74
  * ``greedy_decode_ctc``                   pero_ocr/ocr_engine/pytorch_ocr_engine.py:13-34
  * ``BaseEngineLineOCR.process_lines``     pero_ocr/ocr_engine/line_ocr_engine.py:57-177 (CTC branch)
  * ``softmax``                             pero_ocr/ocr_engine/softmax.py:4-46
  * ``TextLine.get_dense_logits/get_full_logprobs``   pero_ocr/core/layout.py:65-72
  * ``PageParser.compute_line_confidence/get_prob``   pero_ocr/document_ocr/page_parser.py:486-496, 437-450
and the synthetic nets of ``oracle/nets.py`` as explicit functional ops on a plain state dict (so the
restatement does not depend on nn.Module internals).  ``storage`` = 'fp16' additionally models the CUDA
path's numerics (fp16 operands, fp32 accumulate, fp16 activations between layers) so tests can separate
"kernel bug" from "expected fp16 rounding".

Parity is pinned by tests/golden/* produced from the unmodified reference classes (oracle/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .nets import VGG_FRONTEND

LINE_PADDING_PX = 32      # line_ocr_engine.py:54
NET_SUBSAMPLING = 4       # pytorch_ocr_engine.py:41


def _q(x, storage):
    if storage == 'fp16':
        return x.half().float()
    return x


def frontend_forward(sd, x, storage='fp32'):
    """x: f32[N,3,H,W] in [0,1] -> f32[N,512,T]."""
    y = x
    ci = 0
    idx = 0
    for cin, cout, act, pool in VGG_FRONTEND:
        w, b = sd[f'conv.{idx}.weight'], sd[f'conv.{idx}.bias']
        if ci == 0:
            y = F.conv2d(y, w, b, padding=1)           # first layer is computed in fp32 on CUDA cores
        else:
            y = F.conv2d(_q(y, storage), _q(w, storage), b, padding=1)
        y = F.relu(y) if act == 'relu' else F.leaky_relu(y, 0.01)
        idx += 2
        if pool is not None:
            y = F.max_pool2d(y, pool, pool)
            idx += 1
        ci += 1
    g, bt = sd[f'conv.{idx}.weight'], sd[f'conv.{idx}.bias']
    mu, var = sd[f'conv.{idx}.running_mean'], sd[f'conv.{idx}.running_var']
    y = F.batch_norm(y, mu, var, g, bt, training=False, eps=1e-5)
    y = F.conv2d(_q(y, storage), _q(sd['agg.weight'], storage), sd['agg.bias'])
    y = F.leaky_relu(y, 0.01)
    return y.squeeze(2)


def lstm_layer_dir(x, w_ih, w_hh, b_ih, b_hh, reverse, storage):
    """x: [T,N,D] -> [T,N,H]; PyTorch gate order i,f,g,o (nn.LSTM docs)."""
    T, N, _ = x.shape
    H = w_hh.shape[1]
    pre = torch.matmul(_q(x, storage), _q(w_ih, storage).t()) + (b_ih + b_hh)
    h = torch.zeros(N, H)
    c = torch.zeros(N, H)
    out = torch.empty(T, N, H)
    w_hh_q = _q(w_hh, storage).t().contiguous()
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = pre[t] + torch.matmul(_q(h, storage), w_hh_q)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out


def lstm_net_forward(sd, x, storage='fp32'):
    """LineNetLSTM as functional ops.  x: f32[N,3,H,W] -> logits f32[N,C,T]."""
    y = frontend_forward(sd, x, storage).permute(2, 0, 1)        # [T,N,512]
    layer = 0
    while f'lstm.weight_ih_l{layer}' in sd:
        outs = []
        for sfx, rev in (('', False), ('_reverse', True)):
            outs.append(lstm_layer_dir(y, sd[f'lstm.weight_ih_l{layer}{sfx}'], sd[f'lstm.weight_hh_l{layer}{sfx}'],
                                       sd[f'lstm.bias_ih_l{layer}{sfx}'], sd[f'lstm.bias_hh_l{layer}{sfx}'], rev, storage))
        y = torch.cat(outs, dim=2)
        layer += 1
    y = torch.matmul(_q(y, storage), _q(sd['out.weight'], storage).t()) + sd['out.bias']
    return y.permute(1, 2, 0)


def transformer_net_forward(sd, x, heads=8, storage='fp32'):
    """LineNetTransformer as functional ops (post-LN encoder layers, ReLU FFN; nn.TransformerEncoderLayer)."""
    y = frontend_forward(sd, x, storage).permute(2, 0, 1)        # [T,N,D]
    T, N, D = y.shape
    y = F.layer_norm(y, (D,), sd['input_norm.weight'], sd['input_norm.bias'], 1e-5)
    pe = torch.zeros(T, D)
    position = torch.arange(0, T, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, D, 2).float() * (-math.log(10000.0) / D))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    y = y + pe.unsqueeze(1)
    layer = 0
    dh = D // heads
    while f'trans_encoder.layers.{layer}.self_attn.in_proj_weight' in sd:
        p = f'trans_encoder.layers.{layer}.'
        qkv = torch.matmul(_q(y, storage), _q(sd[p + 'self_attn.in_proj_weight'], storage).t()) + sd[p + 'self_attn.in_proj_bias']
        q, k, v = qkv.chunk(3, dim=2)
        q = q.reshape(T, N * heads, dh).transpose(0, 1)
        k = k.reshape(T, N * heads, dh).transpose(0, 1)
        v = v.reshape(T, N * heads, dh).transpose(0, 1)
        att = torch.softmax(torch.bmm(_q(q, storage), _q(k, storage).transpose(1, 2)) / math.sqrt(dh), dim=2)
        o = torch.bmm(_q(att, storage), _q(v, storage)).transpose(0, 1).reshape(T, N, D)
        o = torch.matmul(_q(o, storage), _q(sd[p + 'self_attn.out_proj.weight'], storage).t()) + sd[p + 'self_attn.out_proj.bias']
        y = F.layer_norm(y + o, (D,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-5)
        f = F.relu(torch.matmul(_q(y, storage), _q(sd[p + 'linear1.weight'], storage).t()) + sd[p + 'linear1.bias'])
        f = torch.matmul(_q(f, storage), _q(sd[p + 'linear2.weight'], storage).t()) + sd[p + 'linear2.bias']
        y = F.layer_norm(y + f, (D,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-5)
        layer += 1
    y = torch.matmul(_q(y, storage), _q(sd['out.weight'], storage).t()) + sd['out.bias']
    return y.permute(1, 2, 0)


def parsenet_forward(sd, x, storage='fp32'):
    def c(name, t, first=False):
        if first:
            return F.conv2d(t, sd[name + '.weight'], sd[name + '.bias'], padding=1)
        return F.conv2d(_q(t, storage), _q(sd[name + '.weight'], storage), sd[name + '.bias'], padding=1)
    y = F.relu(c('e1', x, True))
    y = F.max_pool2d(F.relu(c('e2', y)), 2, 2)
    y = F.relu(c('e3', y))
    y = F.max_pool2d(F.relu(c('e4', y)), 2, 2)
    y = F.relu(c('d1', y))
    y = F.relu(c('d2', y))
    y = c('head', y)
    return F.interpolate(y, scale_factor=4.0, mode='nearest')


# --------------------------------------------------------------------------------------------
# decode / engine restatements
# --------------------------------------------------------------------------------------------

def greedy_ctc_indices(logits_nct: np.ndarray):
    """Label ids after CTC collapse, per line.  logits [N,C,T], blank = C-1.

    pytorch_ocr_engine.py:13-34: a virtual all-blank frame is prepended, per-frame argmax (first maximal
    index wins; NaN counts as maximal -- torch.argmax semantics), a label equal to its predecessor is
    dropped, blanks are dropped.
    """
    N, C, T = logits_nct.shape
    out = []
    for n in range(N):
        a = logits_nct[n]
        nan_any = np.isnan(a).any(axis=0)
        best = np.argmax(np.where(np.isnan(a), np.inf, a), axis=0)
        if nan_any.any():
            first_nan = np.argmax(np.isnan(a), axis=0)
            best = np.where(nan_any, first_nan, best)
        prev = np.concatenate([[C - 1], best[:-1]])
        keep = (best != prev) & (best != C - 1)
        out.append(best[keep].astype(np.int32))
    return out


def greedy_ctc_strings(logits_nct, chars):
    return [''.join(chars[c] for c in ids) for ids in greedy_ctc_indices(logits_nct)]


def softmax_np(x, axis):
    """ocr_engine/softmax.py:4-46 with theta=1."""
    y = x - np.expand_dims(np.max(x, axis=axis), axis)
    y = np.exp(y)
    return y / np.expand_dims(np.sum(y, axis=axis), axis)


def sparsify_logits(line_logits: np.ndarray):
    """line_ocr_engine.py:168-172 -- zero raw logits whose softmax prob < 1e-4 (mutates a copy), CSC."""
    from scipy import sparse
    line_logits = line_logits.copy()
    probs = softmax_np(line_logits, axis=1)
    line_logits[probs < 0.0001] = 0
    return sparse.csc_matrix(line_logits)


def dense_logits(sparse_logits, zero_logit_value=-80):
    """core/layout.py:65-68."""
    d = sparse_logits.toarray()
    d[d == 0] = zero_logit_value
    return d


def full_logprobs(sparse_logits, zero_logit_value=-80):
    """core/layout.py:70-72."""
    d = dense_logits(sparse_logits, zero_logit_value)
    from scipy.special import logsumexp
    return d - logsumexp(d, axis=1)[:, np.newaxis]


def line_confidence(dense):
    """PageParser.compute_line_confidence (page_parser.py:486-496) + get_prob (:437-450): the minimum over
    runs of equal per-frame best id (blank runs included) of the run's maximal best-class probability,
    starting from 1.  `dense` = TextLine.get_dense_logits() output [T,C]."""
    log_probs = dense - np.logaddexp.reduce(dense, axis=1)[:, np.newaxis]
    best_ids = np.argmax(log_probs, axis=-1)
    best_probs = np.exp(np.max(log_probs, axis=-1))
    worst, run_id, run_p = 1, -1, 1
    for i, p in zip(best_ids, best_probs):
        if i != run_id:
            worst = min(worst, run_p)
            run_id, run_p = i, p
        else:
            run_p = max(run_p, p)
    return min(worst, run_p)


def line_confident_enough(logprobs, threshold):
    """page_parser.py:81-86."""
    lp = logprobs - np.logaddexp.reduce(logprobs, axis=1)[:, np.newaxis]
    return bool(np.exp(np.min(np.max(lp, axis=-1))) > threshold)


class OracleEngine:
    """CPU restatement of BaseEngineLineOCR.process_lines + PytorchEngineLineOCR.run_ocr (CTC model type)."""

    def __init__(self, sd, characters, kind='lstm', line_px_height=40, batch_size=8, storage='fp32'):
        self.sd = sd
        self.kind = kind
        self.storage = storage
        self.characters = list(characters) + [u'​']           # pytorch_ocr_engine.py:42
        self.line_px_height = line_px_height
        self.line_padding_px = LINE_PADDING_PX
        self.net_subsampling = NET_SUBSAMPLING
        self.batch_size = batch_size
        self.max_input_horizontal_pixels = 480 * batch_size         # line_ocr_engine.py:55

    def model(self, x):
        with torch.no_grad():
            if self.kind == 'lstm':
                return lstm_net_forward(self.sd, x, self.storage)
            return transformer_net_forward(self.sd, x, storage=self.storage)

    def run_ocr(self, batch_data):
        x = torch.from_numpy(batch_data).float() / 255.0
        x = x.permute(0, 3, 1, 2)
        logits = self.model(x).numpy()
        decoded = greedy_ctc_strings(logits, self.characters)
        return decoded, np.ascontiguousarray(logits.transpose(0, 2, 1))

    def process_lines(self, lines, sparse_logits=True, tight_crop_logits=False, no_logits=False):
        n = len(lines)
        trans, logits_out, coords = [None] * n, [None] * n, [None] * n
        order = sorted(range(n), key=lambda i: -lines[i].shape[1])
        while order:
            max_width = int(np.ceil(lines[order[0]].shape[1] / 32.0) * 32)
            bs = max(1, self.max_input_horizontal_pixels // max_width)
            ids, order = order[:bs], order[bs:]
            batch = np.zeros([len(ids), self.line_px_height, max_width + 2 * self.line_padding_px, 3], dtype=np.uint8)
            for row, i in zip(batch, ids):
                row[:, self.line_padding_px:self.line_padding_px + lines[i].shape[1], :] = lines[i]
            if batch.shape[2] > self.max_input_horizontal_pixels:
                batch = batch[:, :, :self.max_input_horizontal_pixels]
            out_t, out_l = self.run_ocr(batch)
            for i, t, ll in zip(ids, out_t, out_l):
                trans[i] = t
                if no_logits:
                    continue
                a = self.line_padding_px // self.net_subsampling
                b = (self.line_padding_px + lines[i].shape[1]) // self.net_subsampling
                if tight_crop_logits:
                    ll = ll[a:b]
                    coords[i] = [None, None]
                else:
                    coords[i] = [a, b]
                logits_out[i] = sparsify_logits(ll) if sparse_logits else ll
        return trans, logits_out, coords
