"""TEST INFRASTRUCTURE ONLY -- CPU restatement of pero-ocr's autoregressive Transformer decoding.

Follows ``TransformerEngineLineOCR.transcribe_batch`` / ``postprocess_decoded``
(pero_ocr/ocr_engine/transformer_ocr_engine.py:49-104) with the cached inference path of the decoder:
``Decoder.infer`` (pero_ocr/ocr_engine/transformer.py:467-486), ``DecoderLayer.infer`` (:418-462) and
``CustomMultiheadAttention.cached_forward`` (:183-305), ``PositionalEncoding`` (:316-332), embedding + output
projection of ``TransformerOCR`` (:489-546).  One token position is processed per step; keys / values of the
self-attention accumulate in a cache, keys / values of the encoder-decoder attention are projected once.

Pinned: tests/golden/ar_decoder.npz and ar_decoder_wide.npz hold token sequences and logits produced by the UNMODIFIED reference classes
(``transformer.build_net`` + ``TransformerEngineLineOCR.transcribe_batch``) hosting the seeded weights of
``ar_decoder_state`` / ``pero_ocr_b200.synthetic`` (oracle/make_golden.py: golden_ar_decoder); tests/test_oracle_ar.py
checks this restatement against them; tests/test_zz_gpu_ar_decoder.py checks the device path
(b200ocr_ar_transcribe behind B200TransformerEngineLineOCR) against the same golden and against this oracle.
"""
import math
from collections import OrderedDict

import numpy as np

D_MODEL = 512


def ar_decoder_state(seed=5, layers=2, dim_ff=2048, classes=32, out_gain=3.0):
    """Seeded decoder parameters under the reference's state-dict names (TransformerOCR: ``trans_decoder.layers.i.*``,
    ``dec_embeder.weight``, ``dec_out_proj.*``).  float32 NumPy arrays."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()

    def lin(name, out_f, in_f, gain=1.0):
        b = gain / math.sqrt(in_f)
        sd[name + 'weight'] = rng.uniform(-b, b, (out_f, in_f)).astype(np.float32)
        sd[name + 'bias'] = rng.uniform(-0.05, 0.05, (out_f,)).astype(np.float32)

    for i in range(layers):
        p = f'trans_decoder.layers.{i}.'
        for att in ('self_attn.', 'multihead_attn.'):
            b = 1.0 / math.sqrt(D_MODEL)
            sd[p + att + 'in_proj_weight'] = rng.uniform(-b, b, (3 * D_MODEL, D_MODEL)).astype(np.float32)
            sd[p + att + 'in_proj_bias'] = rng.uniform(-0.05, 0.05, (3 * D_MODEL,)).astype(np.float32)
            lin(p + att + 'out_proj.', D_MODEL, D_MODEL)
        lin(p + 'linear1.', dim_ff, D_MODEL, gain=math.sqrt(2.0))
        lin(p + 'linear2.', D_MODEL, dim_ff)
        for k in ('norm1.', 'norm2.', 'norm3.'):
            sd[p + k + 'weight'] = rng.uniform(0.8, 1.2, (D_MODEL,)).astype(np.float32)
            sd[p + k + 'bias'] = rng.uniform(-0.1, 0.1, (D_MODEL,)).astype(np.float32)
    sd['dec_embeder.weight'] = rng.standard_normal((classes, D_MODEL)).astype(np.float32)
    lin('dec_out_proj.', classes, D_MODEL, gain=out_gain)
    return sd


def positional_encoding(max_len, d_model=D_MODEL):
    """transformer.py:321-328, float32 like torch."""
    pe = np.zeros((max_len, d_model), dtype=np.float32)
    position = np.arange(0, max_len, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, d_model, 2, dtype=np.float32) * np.float32(-math.log(10000.0) / d_model))
    pe[:, 0::2] = np.sin(position * div)
    pe[:, 1::2] = np.cos(position * div)
    return pe


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _attend(q, k, v, heads):
    """q [N, D] (one position), k / v [S, N, D] -> [N, D]; q is scaled by head_dim ** -0.5 (:268-271)."""
    n, d = q.shape
    hd = d // heads
    qh = (q * np.float32(float(hd) ** -0.5)).reshape(n, heads, hd)
    kh = k.reshape(k.shape[0], n, heads, hd)
    vh = v.reshape(v.shape[0], n, heads, hd)
    scores = np.einsum('nhd,snhd->nhs', qh, kh)
    scores = scores - scores.max(axis=-1, keepdims=True)
    w = np.exp(scores)
    w = w / w.sum(axis=-1, keepdims=True)
    return np.einsum('nhs,snhd->nhd', w, vh).reshape(n, d)


def greedy_transcribe(memory, sd, layers, heads, sentence_boundary_ind, max_len_px):
    """memory: encoder output [T, N, D] float32 (TransformerOCR.encode); sd: decoder state (ar_decoder_state names).
    -> (tokens int64 [steps_kept, N] = the reference's partial_transcripts[1:], logits float32 [N, steps, C]).
    Loop control of transformer_ocr_engine.py:61-89: start from the sentence-boundary token, stop when every line has
    emitted it at least once, or when the transcript exceeds max_len_px // 4 tokens."""
    memory = np.asarray(memory, dtype=np.float32)
    t_mem, n, d = memory.shape
    pe = positional_encoding(max_len_px // 4 + 8, d)
    W = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
    # encoder-decoder attention: K / V of the memory, projected once per layer (:239-249)
    mem_kv = []
    for i in range(layers):
        p = f'trans_decoder.layers.{i}.multihead_attn.'
        kv = memory @ W[p + 'in_proj_weight'][d:].T + W[p + 'in_proj_bias'][d:]
        mem_kv.append((kv[..., :d], kv[..., d:]))
    self_k = [[] for _ in range(layers)]
    self_v = [[] for _ in range(layers)]
    partial = [np.full((n,), sentence_boundary_ind, dtype=np.int64)]
    alive = np.ones((n,), dtype=np.int64)
    all_logits = []
    while True:
        s = len(partial) - 1                                           # position being decoded
        x = W['dec_embeder.weight'][partial[-1]] + pe[s]               # [N, D]
        for i in range(layers):
            p = f'trans_decoder.layers.{i}.'
            qkv = x @ W[p + 'self_attn.in_proj_weight'].T + W[p + 'self_attn.in_proj_bias']
            self_k[i].append(qkv[:, d:2 * d])
            self_v[i].append(qkv[:, 2 * d:])
            a = _attend(qkv[:, :d], np.stack(self_k[i]), np.stack(self_v[i]), heads)
            a = a @ W[p + 'self_attn.out_proj.weight'].T + W[p + 'self_attn.out_proj.bias']
            x = _layer_norm(x + a, W[p + 'norm1.weight'], W[p + 'norm1.bias'])
            q = x @ W[p + 'multihead_attn.in_proj_weight'][:d].T + W[p + 'multihead_attn.in_proj_bias'][:d]
            a = _attend(q, mem_kv[i][0], mem_kv[i][1], heads)
            a = a @ W[p + 'multihead_attn.out_proj.weight'].T + W[p + 'multihead_attn.out_proj.bias']
            x = _layer_norm(x + a, W[p + 'norm2.weight'], W[p + 'norm2.bias'])
            f = np.maximum(x @ W[p + 'linear1.weight'].T + W[p + 'linear1.bias'], 0)
            f = f @ W[p + 'linear2.weight'].T + W[p + 'linear2.bias']
            x = _layer_norm(x + f, W[p + 'norm3.weight'], W[p + 'norm3.bias']).astype(np.float32)
        logits = x @ W['dec_out_proj.weight'].T + W['dec_out_proj.bias']
        all_logits.append(logits.astype(np.float32))
        samples = logits.argmax(axis=-1)
        alive = alive * (samples != sentence_boundary_ind)
        if alive.sum() == 0:
            break
        if len(partial) > max_len_px // 4:
            break
        partial.append(samples.astype(np.int64))
    tokens = np.stack(partial[1:]) if len(partial) > 1 else np.zeros((0, n), dtype=np.int64)
    return tokens, np.stack(all_logits).transpose(1, 0, 2)


def postprocess_decoded(tokens, ignore_ind, sentence_boundary_ind):
    """transformer_ocr_engine.py:91-104: per line, symbols up to the first sentence boundary, ignore symbol skipped."""
    out = []
    for line in np.asarray(tokens).T:
        keep = []
        for s in line:
            if s == sentence_boundary_ind:
                break
            if s == ignore_ind:
                continue
            keep.append(int(s))
        out.append(keep)
    return out
