"""Synthetic random-init networks and inputs for benchmarks, smoke tests and parity tests.

The reference ships no CNN+BiLSTM definition or checkpoint: the recogniser is an opaque TorchScript blob loaded
at ``pero_ocr/ocr_engine/pytorch_ocr_engine.py:52-57`` and called as ``model(x: f32[N,3,40,W]) -> f32[N,C,W/4]``
(``:64-69``); BASELINE.json asks for "random-init" nets.  The nn.Modules below are seeded stand-ins with that
contract -- plain PyTorch *definitions* (they are what gets scripted into a checkpoint file, hosted by the
unmodified reference engine for golden vectors, and walked by ``netdesc`` to build the CUDA engine).  Nothing here
executes on the product path.

The convolutional frontend restates the layer list the reference builds in
``pero_ocr/ocr_engine/transformer.py:75-148`` (VGG16 ``features[:17]`` with the pool strides rewritten for
subsampling (8, 4), then a 256->512->512 LeakyReLU block, BatchNorm2d(512)) and ``:335-363`` (5x1 aggregation conv
+ LeakyReLU).  ``oracle/make_golden.py`` checks this restatement against the reference's own
``ConvolutionalEncoder`` loaded with the same weights.

Weights are drawn from ``numpy.random.default_rng(seed)`` (PCG64 -- stable across numpy/torch versions and
machines) in a fixed order, so the GPU box rebuilds the exact same parameters without shipping 100 MB of floats.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

# (cin, cout, act, pool_after)   act: 'relu' | 'lrelu'
VGG_FRONTEND = [
    (3, 64, 'relu', None),
    (64, 64, 'relu', (2, 2)),
    (64, 128, 'relu', None),
    (128, 128, 'relu', (2, 2)),
    (128, 256, 'relu', None),
    (256, 256, 'relu', None),
    (256, 256, 'relu', (2, 1)),
    (256, 512, 'lrelu', None),
    (512, 512, 'lrelu', None),
]
LINE_HEIGHT = 40
AGG_HEIGHT = LINE_HEIGHT // 8
D_MODEL = 512


def build_frontend_modules():
    layers = []
    for cin, cout, act, pool in VGG_FRONTEND:
        layers.append(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1))
        layers.append(nn.ReLU() if act == 'relu' else nn.LeakyReLU())
        if pool is not None:
            layers.append(nn.MaxPool2d(kernel_size=pool, stride=pool))
    layers.append(nn.BatchNorm2d(512))
    return nn.Sequential(*layers)


class LineNetLSTM(nn.Module):
    """VGG frontend -> 2-layer BiLSTM -> linear CTC head.  f32[N,3,40,W] -> f32[N,C,W/4]."""

    def __init__(self, num_classes: int = 120, hidden: int = 256, lstm_layers: int = 2):
        super().__init__()
        self.conv = build_frontend_modules()
        self.agg = nn.Conv2d(512, D_MODEL, kernel_size=(AGG_HEIGHT, 1))
        self.agg_act = nn.LeakyReLU()
        self.lstm = nn.LSTM(D_MODEL, hidden, num_layers=lstm_layers, bidirectional=True)
        self.out = nn.Linear(2 * hidden, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.agg_act(self.agg(self.conv(x)))      # [N,512,1,T]
        y = y.squeeze(2).permute(2, 0, 1)             # [T,N,512]
        y, _ = self.lstm(y)                           # [T,N,2H]
        y = self.out(y)                               # [T,N,C]
        return y.permute(1, 2, 0)                     # [N,C,T]


class LineNetLSTMEmbed(LineNetLSTM):
    """Embedding-conditioned recogniser, the form the reference engine hosts as ``model(x, ids)`` with a learned table
    ``model.embeddings_layer`` (pytorch_ocr_engine.py:46-50, 64-66; the architecture itself lives in the opaque
    checkpoint): one vector per style id, gathered per line and added to the aggregated features of every frame
    before the recurrence."""

    def __init__(self, num_classes: int = 120, num_embeddings: int = 6, **kw):
        super().__init__(num_classes, **kw)
        self.embeddings_layer = nn.Embedding(num_embeddings, D_MODEL)

    def forward(self, x: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        y = self.agg_act(self.agg(self.conv(x))).squeeze(2)          # [N,512,T]
        y = y + self.embeddings_layer(ids)[:, :, None]
        y, _ = self.lstm(y.permute(2, 0, 1))
        return self.out(y).permute(1, 2, 0)


class LineNetLSTMAlt(nn.Module):
    """A second recogniser family with the same contract (f32[N,3,40,W] -> f32[N,C,W/4]) and a different module tree:
    nested blocks under other names, LeakyReLU slopes 0.1 / 0.2 / 0.3, BatchNorm after pooled blocks, a Dropout, a
    384-wide aggregation, ONE BiLSTM layer and a 1x1 Conv1d as the CTC head.  Exists to pin that the engine builds from
    module types and shapes, not from attribute names (netdesc.describe_line_net)."""

    def __init__(self, num_classes: int = 120, hidden: int = 256):
        super().__init__()
        self.features = nn.Sequential(OrderedDict([
            ('stem', nn.Sequential(nn.Conv2d(3, 64, 3, padding=1), nn.LeakyReLU(0.1))),
            ('stage1', nn.Sequential(nn.Conv2d(64, 64, 3, padding=1), nn.LeakyReLU(0.1), nn.MaxPool2d(2, 2),
                                     nn.BatchNorm2d(64))),
            ('stage2', nn.Sequential(nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(), nn.Conv2d(128, 128, 3, padding=1),
                                     nn.LeakyReLU(0.2), nn.MaxPool2d(2, 2))),
            ('stage3', nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.LeakyReLU(0.2), nn.MaxPool2d((2, 1), (2, 1)),
                                     nn.BatchNorm2d(256), nn.Dropout2d(0.1))),
        ]))
        self.collapse = nn.Sequential(nn.Conv2d(256, 384, kernel_size=(AGG_HEIGHT, 1)), nn.LeakyReLU(0.3))
        self.rnn = nn.LSTM(384, hidden, num_layers=1, bidirectional=True)
        self.classifier = nn.Conv1d(2 * hidden, num_classes, kernel_size=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.collapse(self.features(x)).squeeze(2)     # [N,384,T]
        y, _ = self.rnn(y.permute(2, 0, 1))                # [T,N,2H]
        return self.classifier(y.permute(1, 2, 0))         # [N,C,T]


class LineNetTransformer(nn.Module):
    """VGG frontend -> LayerNorm + sinusoid PE + post-LN TransformerEncoder -> linear CTC head.

    Encoder half of the reference's ``TransformerOCR`` (``transformer.py:366-385, 548-555``)
    with a CTC head in place of the autoregressive decoder.
    """

    def __init__(self, num_classes: int = 120, layers: int = 2, heads: int = 8, dim_ff: int = 2048,
                 max_len: int = 2000):
        super().__init__()
        self.conv = build_frontend_modules()
        self.agg = nn.Conv2d(512, D_MODEL, kernel_size=(AGG_HEIGHT, 1))
        self.agg_act = nn.LeakyReLU()
        self.input_norm = nn.LayerNorm(D_MODEL, eps=1e-5)
        enc_layer = nn.TransformerEncoderLayer(D_MODEL, heads, dim_feedforward=dim_ff, dropout=0.0)
        self.trans_encoder = nn.TransformerEncoder(enc_layer, num_layers=layers, enable_nested_tensor=False)
        pe = torch.zeros(max_len, D_MODEL)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, D_MODEL, 2).float() * (-math.log(10000.0) / D_MODEL))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(1), persistent=False)
        self.out = nn.Linear(D_MODEL, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.agg_act(self.agg(self.conv(x)))
        y = y.squeeze(2).permute(2, 0, 1)             # [T,N,512]
        y = self.input_norm(y)
        y = y + self.pe[:y.size(0)]
        y = self.trans_encoder(y)
        y = self.out(y)
        return y.permute(1, 2, 0)


class ParseNetStandIn(nn.Module):
    """Conv-only stand-in for the opaque ParseNet blob (``layout_engines/torch_parsenet.py:11-15, 51-53``):
    f32[1,3,H,W] (H, W multiples of 64) -> (f32[1,5,H,W], aux).  Encoder: 3 conv+pool levels, decoder:
    nearest upsampling + conv, 5 output maps (``cnn_layout_engine.py:129-131``)."""

    def __init__(self, base: int = 64):
        super().__init__()
        b = base
        self.e1 = nn.Conv2d(3, b, 3, padding=1)
        self.e2 = nn.Conv2d(b, b, 3, padding=1)
        self.e3 = nn.Conv2d(b, 2 * b, 3, padding=1)
        self.e4 = nn.Conv2d(2 * b, 2 * b, 3, padding=1)
        self.d1 = nn.Conv2d(2 * b, b, 3, padding=1)
        self.d2 = nn.Conv2d(b, b, 3, padding=1)
        self.head = nn.Conv2d(b, 5, 3, padding=1)
        self.pool = nn.MaxPool2d(2, 2)
        self.act = nn.ReLU()

    def forward(self, x: torch.Tensor):
        y = self.act(self.e1(x))
        y = self.pool(self.act(self.e2(y)))           # /2
        y = self.act(self.e3(y))
        y = self.pool(self.act(self.e4(y)))           # /4
        y = self.act(self.d1(y))
        y = self.act(self.d2(y))
        y = self.head(y)                               # [1,5,H/4,W/4]
        y = torch.nn.functional.interpolate(y, scale_factor=4.0, mode='nearest')
        return y, y[:, :1]


# --------------------------------------------------------------------------------------------
# Seeded parameters (numpy PCG64) -- same dict feeds the oracle modules and the CUDA engine.
# --------------------------------------------------------------------------------------------

def _fan_in(shape):
    f = 1
    for s in shape[1:]:
        f *= s
    return f


def seeded_state_dict(module: nn.Module, seed: int = 0, out_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Draw every parameter/buffer of ``module`` from default_rng(seed) in state_dict order.

    conv / linear weights: He-uniform (gain sqrt(2)) so activations stay O(1) through the ReLU stack;
    biases U(-0.1, 0.1); BatchNorm: gamma U(0.5,1.5), beta U(-0.2,0.2), mean U(-0.2,0.2), var U(0.5,1.5);
    LayerNorm gamma U(0.8,1.2), beta U(-0.1,0.1); LSTM / attention in_proj: U(-1/sqrt(fan_in), ..).
    ``out_gain`` scales the CTC head ('out.weight') so the top-2 logit margin is well above numeric error.
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        leaf = name.split('.')[-1]
        if leaf == 'num_batches_tracked':
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        is_bn = leaf in ('running_mean', 'running_var') or (
            len(shape) == 1 and name.rsplit('.', 1)[0] + '.running_mean' in module.state_dict())
        is_ln = ('norm' in name) and len(shape) == 1 and not is_bn
        if leaf == 'running_var':
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf == 'running_mean':
            a = rng.uniform(-0.2, 0.2, shape)
        elif is_bn and leaf == 'weight':
            a = rng.uniform(0.5, 1.5, shape)
        elif is_bn and leaf == 'bias':
            a = rng.uniform(-0.2, 0.2, shape)
        elif is_ln and leaf == 'weight':
            a = rng.uniform(0.8, 1.2, shape)
        elif is_ln and leaf == 'bias':
            a = rng.uniform(-0.1, 0.1, shape)
        elif name.startswith('lstm.') or 'in_proj' in name or leaf.startswith(('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')):
            if len(shape) >= 2:
                b = 1.0 / math.sqrt(shape[1])
            else:
                b = 0.05
            a = rng.uniform(-b, b, shape)
        elif len(shape) >= 2:
            gain = math.sqrt(2.0)
            b = gain * math.sqrt(3.0 / _fan_in(shape))
            if name in ('out.weight', 'classifier.weight'):
                b = out_gain * math.sqrt(3.0 / _fan_in(shape))
            a = rng.uniform(-b, b, shape)
        else:
            a = rng.uniform(-0.1, 0.1, shape)
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


def make_net(kind: str = 'lstm', num_classes: int = 120, seed: int = 0, out_gain: float = 1.0, **kw) -> nn.Module:
    if kind == 'lstm':
        net = LineNetLSTM(num_classes, **kw)
    elif kind == 'lstm_alt':
        net = LineNetLSTMAlt(num_classes, **kw)
    elif kind == 'lstm_embed':
        net = LineNetLSTMEmbed(num_classes, **kw)
    elif kind == 'transformer':
        net = LineNetTransformer(num_classes, **kw)
    elif kind == 'parsenet':
        net = ParseNetStandIn(**kw)
    else:
        raise ValueError(kind)
    net.load_state_dict(seeded_state_dict(net, seed, out_gain))
    return net.eval()


def json_characters(n):
    """`n` distinct printable characters for an engine JSON (the engine appends U+200B and blank is last)."""
    return [chr(0x100 + i) for i in range(n)]


def bench_crops(n, width=1280, seed=0, height=40):
    """BASELINE.json configs 2/3/5: n "gray" 40 x width crops [n,H,W,3] u8 (equal channels), default_rng(seed)."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 256, (n, height, width), dtype=np.uint8)
    return np.repeat(g[:, :, :, None], 3, axis=3)


def transformer_ocr_state(encoder_net: "LineNetTransformer", decoder_state) -> "OrderedDict[str, torch.Tensor]":
    """Re-keys the encoder half of a ``LineNetTransformer`` + a decoder state (reference names already:
    ``trans_decoder.layers.i.*``, ``dec_embeder.weight``, ``dec_out_proj.*``) into the state dict of the reference's
    ``TransformerOCR`` as ``transformer.build_net`` lays it out (``pero_ocr/ocr_engine/transformer.py:75-148,
    335-385``): VGG convolutions directly in ``encoder_frontend.blocks_2d.blocks_2d`` (each followed by its ReLU, a
    pool by a Dropout), the 256->512->512 block as a nested Sequential, then the BatchNorm.  This is the checkpoint
    format ``TransformerEngineLineOCR`` loads (``transformer_ocr_engine.py:29``); tests/golden/transformer_ocr_keys.json
    holds the key / shape list of the unmodified reference for it."""
    src = encoder_net.state_dict()
    convs = [k[:-len('.weight')] for k, v in src.items() if k.startswith('conv.') and k.endswith('.weight') and v.dim() == 4]
    bn = [k[:-len('.running_mean')] for k in src if k.startswith('conv.') and k.endswith('.running_mean')]
    assert len(convs) == len(VGG_FRONTEND) and len(bn) == 1
    out = OrderedDict()
    prefix = 'encoder_frontend.blocks_2d.blocks_2d.'
    idx = 0
    block = None
    for name, (cin, cout, act, pool) in zip(convs, VGG_FRONTEND):
        if act == 'relu':
            dst = f'{prefix}{idx}'
            idx += 4 if pool is not None else 2          # conv, ReLU (, MaxPool2d, Dropout)
        else:
            if block is None:
                block = [idx, 0]
            dst = f'{prefix}{block[0]}.{block[1]}'
            block[1] += 2                                # conv, LeakyReLU inside create_vgg_block_2d
        out[dst + '.weight'] = src[name + '.weight']
        out[dst + '.bias'] = src[name + '.bias']
    bn_dst = f'{prefix}{block[0] + 1}'
    for leaf in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked'):
        out[f'{bn_dst}.{leaf}'] = src[f'{bn[0]}.{leaf}']
    out['encoder_frontend.aggregation_conv.0.weight'] = src['agg.weight']
    out['encoder_frontend.aggregation_conv.0.bias'] = src['agg.bias']
    for k, v in src.items():
        if k.startswith('trans_encoder.') or k.startswith('input_norm.'):
            out['encoder.' + k] = v
    for k, v in decoder_state.items():
        out[k] = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))
    return out


# --------------------------------------------------------------------------------------------
# Synthetic decoder inputs (BASELINE.json config 1 and trained-net-like log-probabilities)
# --------------------------------------------------------------------------------------------
BLANK = '<BLANK>'


def config1_logits():
    """BASELINE.json config 1: 128 lines x T=256 x C=120, blank last.  -> (raw f32, log-probs, letters)"""
    from scipy.special import log_softmax
    rng = np.random.default_rng(1234)
    raw = (rng.standard_normal((128, 256, 120)) * 4).astype(np.float32)
    lp = log_softmax(raw, axis=2)
    letters = [chr(0x100 + i) for i in range(119)] + [BLANK]
    return raw, lp, letters


def peaky_logprobs(rng, n, t, c, sharp=9.0, p_blank=0.55, p_repeat=0.3):
    """Log-probabilities shaped like a trained CTC net's: one dominant class per frame, blank-heavy,
    with repeats, plus low-level noise so that a few classes pass the decoder's > -10 relevance gate."""
    from scipy.special import log_softmax
    out = np.empty((n, t, c), dtype=np.float64)
    for i in range(n):
        raw = rng.standard_normal((t, c)) * 1.5
        prev = c - 1
        for f in range(t):
            u = rng.random()
            if u < p_blank:
                k = c - 1
            elif u < p_blank + p_repeat and prev != c - 1:
                k = prev
            else:
                k = int(rng.integers(0, c - 1))
            raw[f, k] += sharp * rng.uniform(0.4, 1.0)
            prev = k
        out[i] = log_softmax(raw, axis=1)
    return out
