"""Checkpoint -> layer list for b200ocr_create.

The reference hosts the recogniser as an opaque TorchScript blob (pero_ocr/ocr_engine/pytorch_ocr_engine.py:52-57)
and ParseNet likewise (pero_ocr/layout_engines/torch_parsenet.py:11-15).  A TorchScript module still exposes its
module tree (``original_name``, ``named_children``) and ``state_dict()``, so the same checkpoint file the reference
loads is walked here and turned into the flat layer list of ``include/b200_lineocr.h``.  Supported vocabulary:

  line recogniser:  any module tree (names and nesting are free) whose leaf modules, in registration order, are
                    Conv2d(3x3, pad 1) / ReLU / LeakyReLU(any positive slope) / MaxPool2d / BatchNorm2d / Dropout,
                    a Conv2d(k = (H/8, 1)) + activation, then either a bidirectional nn.LSTM or LayerNorm +
                    post-LN nn.TransformerEncoder (ReLU), then Linear / Conv1d(k = 1) as the CTC head.
  page detector:    ``e1 e2 | pool | e3 e4 | pool | d1 d2 | head`` 3x3 convs + nearest x4 upsampling.

  TransformerOCR:   the state dict ``TransformerEngineLineOCR`` loads (transformer_ocr_engine.py:21-29) plus the
                    ``net_name`` config of ``transformer.build_net`` (transformer.py:12-48) -> encoder layer list
                    (no CTC head) + decoder description for ``b200ocr_ar_attach`` (``describe_transformer_ocr``).

Anything else raises -- there is no generic fallback executor.
"""
import json
import re

import numpy as np

from . import _lib


def _name(m):
    return getattr(m, 'original_name', type(m).__name__)


def _f32(t):
    if not hasattr(t, 'detach'):
        return np.ascontiguousarray(np.asarray(t, dtype=np.float32))
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def _act_code(m):
    """-> (activation code, LeakyReLU negative slope or 0)."""
    n = _name(m)
    if n == 'ReLU':
        return _lib.ACT_RELU, 0.0
    if n == 'LeakyReLU':
        slope = float(getattr(m, 'negative_slope', 0.01))
        if slope == 0.0:
            return _lib.ACT_RELU, 0.0
        if slope < 0.0:
            raise ValueError(f'LeakyReLU slope {slope} < 0: the max-pool is fused BEFORE the activation, which needs a '
                             f'monotone activation')
        return _lib.ACT_LEAKY_RELU, slope
    raise ValueError(f'unsupported activation {n}')


def _set_act(spec, m):
    spec['act'], slope = _act_code(m)
    if slope and slope != 0.01:              # 0.01 is the kernels' default (nn.LeakyReLU's own)
        spec['act_slope'] = slope


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


def _conv_spec(sd, prefix, first):
    w = _f32(sd[prefix + '.weight'])
    b = _f32(sd[prefix + '.bias']) if prefix + '.bias' in sd else None
    cout, cin, kh, kw = w.shape
    return dict(kind=_lib.CONV_FIRST if first else _lib.CONV, cin=cin, cout=cout, kh=kh, kw=kw,
                pad_h=(kh - 1) // 2, pad_w=(kw - 1) // 2, act=_lib.ACT_NONE, pool_h=1, pool_w=1, weight=w, bias=b)


_LEAF_TYPES = {'Embedding', 'Conv2d', 'Conv1d', 'ReLU', 'LeakyReLU', 'MaxPool2d', 'BatchNorm2d', 'Dropout', 'Dropout2d', 'Identity',
               'LSTM', 'Linear', 'LayerNorm', 'TransformerEncoder'}


def _leaves(module, prefix=''):
    """(state-dict prefix, module) of the leaf modules in registration order; containers are expanded, whatever
    their names and nesting."""
    for name, m in module.named_children():
        key = prefix + name
        if _name(m) in _LEAF_TYPES or not any(True for _ in m.named_children()):
            yield key, m
        else:
            yield from _leaves(m, key + '.')


def describe_line_net(module, embed_id=None):
    """nn.Module or TorchScript module -> (layer specs, num_classes).

    The walk is driven by module TYPES and tensor shapes, not by attribute names: the leaf modules are taken in
    registration order (the order `nn.Sequential` runs them and the order every recogniser family seen so far
    registers them; the channel counts of consecutive layers are checked against each other, which catches a tree
    registered out of order) and must spell
        (Conv2d 3x3 pad 1 [ReLU | LeakyReLU] [MaxPool2d] [BatchNorm2d] | Dropout)+
        Conv2d (k, 1) pad 0 [ReLU | LeakyReLU]                       -- height aggregation
        LSTM (bidirectional, any number of layers)  |  LayerNorm + TransformerEncoder (post-LN, ReLU)
        Linear | Conv1d k = 1 | Conv2d 1x1                           -- CTC head
    An ``nn.Embedding`` anywhere in the tree makes it an embedding-conditioned recogniser, ``model(x, ids)``
    (pytorch_ocr_engine.py:64-66), of the form "one gathered vector per line added to the aggregated features": the
    reference passes ONE id for the whole batch, so the vector is a per-channel shift after the aggregation's
    activation.  It is folded into that layer's post-affine for `embed_id` (an int, or "mean" = the table's last row,
    pytorch_ocr_engine.py:49-50); the layer spec keeps the table (``embedding_table``) and the shift without it
    (``embedding_base_shift``) so that the id can be changed later (LineRecognizer.set_embedding).
    """
    sd = {k: v for k, v in module.state_dict().items()}
    layers = []
    last_conv = None
    stage = 'frontend'                       # -> 'agg' -> 'sequence' -> 'head' -> 'done'
    feat = None
    pending_norm = None
    embedding = None
    for key, m in _leaves(module):
        n = _name(m)
        if n in ('Dropout', 'Dropout2d', 'Identity'):
            continue
        if n == 'Embedding':
            if embedding is not None:
                raise ValueError('more than one embedding table')
            embedding = _f32(sd[key + '.weight'])
            continue
        if stage == 'done':
            raise ValueError(f'module {key} ({n}) after the CTC head')
        if n == 'Conv2d' and stage in ('frontend', 'agg'):
            spec = _conv_spec(sd, key, first=not layers)
            if layers and spec['cin'] != layers[-1]['cout']:
                raise ValueError(f'{key}: {spec["cin"]} input channels after a layer with {layers[-1]["cout"]} outputs '
                                 f'(module tree not registered in forward order?)')
            if (spec['kh'], spec['kw']) == (3, 3) and stage == 'frontend':
                if _pair(getattr(m, 'padding', (1, 1))) != (1, 1) or _pair(getattr(m, 'stride', (1, 1))) != (1, 1):
                    raise ValueError(f'{key}: frontend convolutions must have padding 1 and stride 1')
                layers.append(spec)
                last_conv = spec
            elif spec['kw'] == 1 and stage == 'frontend' and layers:
                if _pair(getattr(m, 'padding', (0, 0))) != (0, 0) or _pair(getattr(m, 'stride', (1, 1))) != (1, 1):
                    raise ValueError(f'{key}: the height-aggregation convolution must have no padding and stride 1')
                spec['pad_h'] = spec['pad_w'] = 0
                layers.append(spec)
                last_conv = spec
                stage = 'agg'
                feat = spec['cout']
            else:
                raise ValueError(f'{key}: unsupported convolution {spec["kh"]}x{spec["kw"]} at this position')
        elif n in ('ReLU', 'LeakyReLU') and stage in ('frontend', 'agg'):
            if last_conv is None or last_conv['act'] != _lib.ACT_NONE or 'post_scale' in last_conv or \
                    (last_conv['pool_h'], last_conv['pool_w']) != (1, 1):
                raise ValueError(f'{key}: an activation must directly follow a convolution')
            _set_act(last_conv, m)
        elif n == 'MaxPool2d' and stage == 'frontend':
            ph, pw = _pair(m.kernel_size)
            if (ph, pw) == (1, 1):
                continue
            st = getattr(m, 'stride', None)
            if st is not None and _pair(st) != (ph, pw):
                raise ValueError(f'{key}: max-pool stride must equal its kernel')
            if last_conv is None or (last_conv['pool_h'], last_conv['pool_w']) != (1, 1) or 'post_scale' in last_conv \
                    or ph not in (1, 2) or pw not in (1, 2) or last_conv['kind'] == _lib.CONV_FIRST:
                raise ValueError('max-pool must be 2x2 / 2x1 / 1x2 and follow a tensor-core convolution')
            last_conv['pool_h'], last_conv['pool_w'] = ph, pw
        elif n == 'BatchNorm2d' and stage == 'frontend':
            # eval-mode BN folded to y * scale + shift, applied after activation (+pool) of the preceding conv
            g, b = _f32(sd[key + '.weight']), _f32(sd[key + '.bias'])
            mu, var = _f32(sd[key + '.running_mean']), _f32(sd[key + '.running_var'])
            eps = float(getattr(m, 'eps', 1e-5))
            scale = (g.astype(np.float64) / np.sqrt(var.astype(np.float64) + eps))
            shift = b.astype(np.float64) - mu.astype(np.float64) * scale
            if last_conv is None or last_conv['kind'] == _lib.CONV_FIRST or 'post_scale' in last_conv:
                raise ValueError('BatchNorm must follow a tensor-core convolution')
            last_conv['post_scale'] = scale.astype(np.float32)
            last_conv['post_shift'] = shift.astype(np.float32)
        elif n == 'LSTM' and stage == 'agg':
            if not bool(getattr(m, 'bidirectional', f'{key}.weight_ih_l0_reverse' in sd)):
                raise ValueError(f'{key}: the recurrence kernel is bidirectional')
            layer, cin = 0, feat
            while f'{key}.weight_ih_l{layer}' in sd:
                spec = dict(kind=_lib.BILSTM, cin=cin)
                for name, sfx in (('w_ih', 'weight_ih'), ('w_hh', 'weight_hh'), ('b_ih', 'bias_ih'), ('b_hh', 'bias_hh')):
                    spec[name] = [_f32(sd[f'{key}.{sfx}_l{layer}']), _f32(sd[f'{key}.{sfx}_l{layer}_reverse'])]
                spec['hidden'] = spec['w_hh'][0].shape[1]
                if spec['w_ih'][0].shape[1] != cin:
                    raise ValueError(f'{key} layer {layer}: input width {spec["w_ih"][0].shape[1]} != {cin}')
                layers.append(spec)
                cin = 2 * spec['hidden']
                layer += 1
            feat = cin
            stage = 'head'
        elif n == 'LayerNorm' and stage == 'agg':
            pending_norm = key
        elif n == 'TransformerEncoder' and stage == 'agg' and pending_norm is not None:
            layers.append(dict(kind=_lib.LN_PE, cin=feat, norm1_w=_f32(sd[pending_norm + '.weight']),
                               norm1_b=_f32(sd[pending_norm + '.bias'])))
            layer = 0
            while f'{key}.layers.{layer}.linear1.weight' in sd:
                p = f'{key}.layers.{layer}.'
                heads = int(getattr(module, 'num_heads', 8))
                try:
                    heads = int(dict(m.named_children())['layers'][layer].self_attn.num_heads)
                except Exception:
                    pass
                layers.append(dict(
                    kind=_lib.TRANSFORMER_LAYER, cin=feat, heads=heads, dim_ff=sd[p + 'linear1.weight'].shape[0],
                    in_proj_w=_f32(sd[p + 'self_attn.in_proj_weight']), in_proj_b=_f32(sd[p + 'self_attn.in_proj_bias']),
                    out_proj_w=_f32(sd[p + 'self_attn.out_proj.weight']), out_proj_b=_f32(sd[p + 'self_attn.out_proj.bias']),
                    lin1_w=_f32(sd[p + 'linear1.weight']), lin1_b=_f32(sd[p + 'linear1.bias']),
                    lin2_w=_f32(sd[p + 'linear2.weight']), lin2_b=_f32(sd[p + 'linear2.bias']),
                    norm1_w=_f32(sd[p + 'norm1.weight']), norm1_b=_f32(sd[p + 'norm1.bias']),
                    norm2_w=_f32(sd[p + 'norm2.weight']), norm2_b=_f32(sd[p + 'norm2.bias'])))
                layer += 1
            stage = 'head'
        elif n in ('Linear', 'Conv1d', 'Conv2d') and stage == 'head':
            ow = _f32(sd[key + '.weight'])
            if ow.ndim > 2:
                if any(d != 1 for d in ow.shape[2:]):
                    raise ValueError(f'{key}: a convolutional CTC head must have kernel size 1')
                ow = np.ascontiguousarray(ow.reshape(ow.shape[0], ow.shape[1]))
            if ow.shape[1] != feat:
                raise ValueError(f'{key}: head input width {ow.shape[1]} != {feat}')
            ob = _f32(sd[key + '.bias']) if key + '.bias' in sd else np.zeros(ow.shape[0], dtype=np.float32)
            layers.append(dict(kind=_lib.CTC_HEAD, cin=feat, cout=ow.shape[0], kh=1, kw=1, weight=ow, bias=ob))
            stage = 'done'
        else:
            raise ValueError(f'unsupported module {key} ({n}) in the {stage} part of the recogniser')
    if stage != 'done':
        raise ValueError('the recogniser must end in a sequence encoder (bidirectional LSTM, or LayerNorm + '
                         'TransformerEncoder) and a linear CTC head')
    if embedding is not None:
        if embed_id is None:
            raise ValueError('the recogniser takes an embedding id: set "embed_id" in the engine JSON '
                             '(line_ocr_engine.py:36-42)')
        agg = [l for l in layers if l['kind'] == _lib.CONV][-1]
        if embedding.shape[1] != agg['cout']:
            raise ValueError(f'embedding width {embedding.shape[1]} != aggregated feature width {agg["cout"]}')
        agg['embedding_table'] = embedding
        agg['embedding_base_shift'] = agg.get('post_shift', np.zeros(agg['cout'], dtype=np.float32)).copy()
        agg.setdefault('post_scale', np.ones(agg['cout'], dtype=np.float32))
        agg['post_shift'] = agg['embedding_base_shift'] + embedding[resolve_embed_id(embed_id, embedding.shape[0])]
    elif embed_id is not None:
        raise ValueError('"embed_id" is set but the recogniser has no embedding table')
    return layers, layers[-1]['cout']


def resolve_embed_id(embed_id, rows):
    """int or "mean" (= the last row, pytorch_ocr_engine.py:49-50) -> row index, range-checked."""
    idx = rows - 1 if embed_id == 'mean' else int(embed_id)
    if not 0 <= idx < rows:
        raise IndexError(f'embed_id {embed_id} outside the table of {rows} embeddings')
    return idx


def describe_transformer_ocr(state_dict, net_config, line_height=40):
    """State dict of the reference's ``TransformerOCR`` (transformer.py:489-546) + the ``net_name`` config that
    ``build_net`` reads (:12-20) -> (encoder layer specs for b200ocr_create, decoder spec for b200ocr_ar_attach).

    The convolutional frontend is rebuilt from the key structure the way ``VGG_conv_module.__init__`` builds it
    (:75-148): convolutions directly in ``blocks_2d`` come from VGG16 (ReLU), a gap of four module indices after one
    means a (max-pool, dropout) pair follows it, nested ``Sequential`` blocks are ``create_vgg_block_2d(norm='none')``
    (:51-72: conv + LeakyReLU twice, then a max-pool) followed by a BatchNorm2d; pool strides replay the subsampling
    bookkeeping (:104-139) from ``conv_subsampling``."""
    cfg = json.loads(net_config) if isinstance(net_config, str) else dict(net_config)
    sd = {k: _f32(v) for k, v in state_dict.items() if not k.endswith('num_batches_tracked')}
    d_model, dim_ff, heads = int(cfg['dim_model']), int(cfg['dim_ff']), int(cfg['heads'])
    sub_v, sub_h = cfg['conv_subsampling']
    prefix = 'encoder_frontend.blocks_2d.blocks_2d.'
    mods = {}                                   # (idx,) or (idx, sub) -> module key prefix
    for k in sd:
        if k.startswith(prefix):
            m = re.match(r'(\d+)(?:\.(\d+))?\.(weight|bias|running_mean|running_var)$', k[len(prefix):])
            if not m:
                raise ValueError(f'unexpected frontend parameter {k}')
            key = (int(m.group(1)),) if m.group(2) is None else (int(m.group(1)), int(m.group(2)))
            mods[key] = prefix + '.'.join(str(i) for i in key)
    top = sorted({k[0] for k in mods})
    layers = []
    cur_v = cur_h = 1

    def next_stride():
        nonlocal cur_v, cur_h
        sv = 2 if (sub_v is None or cur_v < sub_v) else 1           # transformer.py:106-114 / 128-136
        sh = 2 if cur_h < sub_h else 1
        cur_v, cur_h = cur_v * sv, cur_h * sh
        return sv, sh

    def set_pool(spec, stride):
        if stride == (1, 1):
            return
        if spec['kind'] == _lib.CONV_FIRST or stride[0] not in (1, 2) or stride[1] not in (1, 2):
            raise ValueError('max-pool must be 2x2 / 2x1 / 1x2 and follow a tensor-core convolution')
        spec['pool_h'], spec['pool_w'] = stride

    for pos, idx in enumerate(top):
        nested = sorted(k for k in mods if k[0] == idx and len(k) == 2)
        nxt = top[pos + 1] if pos + 1 < len(top) else None
        if nested:                                                   # create_vgg_block_2d(norm='none')
            for key in nested:
                spec = _conv_spec(sd, mods[key], first=False)
                spec['act'] = _lib.ACT_LEAKY_RELU
                layers.append(spec)
            set_pool(layers[-1], next_stride())
        elif mods[(idx,)] + '.running_mean' in sd:                   # BatchNorm2d after a block (:141-144)
            key = mods[(idx,)]
            g, b, mu, var = (sd[key + s].astype(np.float64) for s in ('.weight', '.bias', '.running_mean', '.running_var'))
            scale = g / np.sqrt(var + 1e-5)
            if not layers or layers[-1]['kind'] == _lib.CONV_FIRST or 'post_scale' in layers[-1]:
                raise ValueError('BatchNorm must follow a tensor-core convolution')
            layers[-1]['post_scale'] = scale.astype(np.float32)
            layers[-1]['post_shift'] = (b - mu * scale).astype(np.float32)
        else:                                                        # VGG16 conv + ReLU
            spec = _conv_spec(sd, mods[(idx,)], first=not layers)
            spec['act'] = _lib.ACT_RELU
            layers.append(spec)
            if nxt is not None and nxt - idx == 4:                   # conv, ReLU, MaxPool2d, Dropout
                set_pool(spec, next_stride())
            elif nxt is not None and nxt - idx != 2:
                raise ValueError(f'unexpected module spacing in the VGG frontend at index {idx}')
    for spec in layers:
        if (spec['kh'], spec['kw']) != (3, 3):
            raise ValueError('frontend convolutions must be 3x3')
    if (sub_v is not None and cur_v != sub_v) or cur_h != sub_h:
        raise ValueError(f'frontend subsampling ({cur_v}, {cur_h}) does not reach conv_subsampling {cfg["conv_subsampling"]}')
    agg = _conv_spec(sd, 'encoder_frontend.aggregation_conv.0', first=False)
    if agg['kh'] != line_height // cur_v or agg['kw'] != 1 or agg['cout'] != d_model:
        raise ValueError('aggregation convolution does not match line height / subsampling / dim_model')
    agg['pad_h'] = agg['pad_w'] = 0
    agg['act'] = _lib.ACT_LEAKY_RELU
    layers.append(agg)
    layers.append(dict(kind=_lib.LN_PE, cin=d_model, norm1_w=sd['encoder.input_norm.weight'],
                       norm1_b=sd['encoder.input_norm.bias']))
    for i in range(int(cfg['encoder_layers'])):
        p = f'encoder.trans_encoder.layers.{i}.'
        layers.append(dict(
            kind=_lib.TRANSFORMER_LAYER, cin=d_model, heads=heads, dim_ff=sd[p + 'linear1.weight'].shape[0],
            in_proj_w=sd[p + 'self_attn.in_proj_weight'], in_proj_b=sd[p + 'self_attn.in_proj_bias'],
            out_proj_w=sd[p + 'self_attn.out_proj.weight'], out_proj_b=sd[p + 'self_attn.out_proj.bias'],
            lin1_w=sd[p + 'linear1.weight'], lin1_b=sd[p + 'linear1.bias'],
            lin2_w=sd[p + 'linear2.weight'], lin2_b=sd[p + 'linear2.bias'],
            norm1_w=sd[p + 'norm1.weight'], norm1_b=sd[p + 'norm1.bias'],
            norm2_w=sd[p + 'norm2.weight'], norm2_b=sd[p + 'norm2.bias']))
    if f'encoder.trans_encoder.layers.{int(cfg["encoder_layers"])}.linear1.weight' in sd:
        raise ValueError('state dict has more encoder layers than the net config')
    dec_layers = []
    for i in range(int(cfg['decoder_layers'])):
        p = f'trans_decoder.layers.{i}.'
        dec_layers.append(dict(
            self_in_w=sd[p + 'self_attn.in_proj_weight'], self_in_b=sd[p + 'self_attn.in_proj_bias'],
            self_out_w=sd[p + 'self_attn.out_proj.weight'], self_out_b=sd[p + 'self_attn.out_proj.bias'],
            cross_in_w=sd[p + 'multihead_attn.in_proj_weight'], cross_in_b=sd[p + 'multihead_attn.in_proj_bias'],
            cross_out_w=sd[p + 'multihead_attn.out_proj.weight'], cross_out_b=sd[p + 'multihead_attn.out_proj.bias'],
            lin1_w=sd[p + 'linear1.weight'], lin1_b=sd[p + 'linear1.bias'],
            lin2_w=sd[p + 'linear2.weight'], lin2_b=sd[p + 'linear2.bias'],
            norm1_w=sd[p + 'norm1.weight'], norm1_b=sd[p + 'norm1.bias'],
            norm2_w=sd[p + 'norm2.weight'], norm2_b=sd[p + 'norm2.bias'],
            norm3_w=sd[p + 'norm3.weight'], norm3_b=sd[p + 'norm3.bias']))
    if f'trans_decoder.layers.{int(cfg["decoder_layers"])}.linear1.weight' in sd:
        raise ValueError('state dict has more decoder layers than the net config')
    for i, ly in enumerate(dec_layers):
        if ly['lin1_w'].shape != (dim_ff, d_model) or ly['self_in_w'].shape != (3 * d_model, d_model):
            raise ValueError(f'decoder layer {i}: parameter shapes do not match the net config')
    classes = sd['dec_embeder.weight'].shape[0]
    if sd['dec_out_proj.weight'].shape != (classes, d_model):
        raise ValueError('dec_out_proj does not match dec_embeder')
    decoder = dict(layers=dec_layers, heads=heads, dim_ff=dim_ff, classes=classes, d_model=d_model,
                   embed=sd['dec_embeder.weight'], out_w=sd['dec_out_proj.weight'], out_b=sd['dec_out_proj.bias'])
    return layers, decoder


def ar_to_ctypes(decoder):
    """Decoder spec of describe_transformer_ocr -> (ArDesc, keepalive list) for b200ocr_ar_attach."""
    import ctypes as C
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_float))

    arr = (_lib.ArLayer * len(decoder['layers']))()
    for i, spec in enumerate(decoder['layers']):
        for key, _ in _lib.ArLayer._fields_:
            setattr(arr[i], key, ptr(spec[key]))
    desc = _lib.ArDesc(len(decoder['layers']), int(decoder['heads']), int(decoder['dim_ff']), int(decoder['classes']),
                       arr, ptr(decoder['embed']), ptr(decoder['out_w']), ptr(decoder['out_b']))
    keep.append(arr)
    return desc, keep


def describe_parsenet(module):
    sd = {k: v for k, v in module.state_dict().items()}
    layers = []
    for name, pool in (('e1', 1), ('e2', 2), ('e3', 1), ('e4', 2), ('d1', 1), ('d2', 1)):
        spec = _conv_spec(sd, name, first=(name == 'e1'))
        spec['act'] = _lib.ACT_RELU
        spec['pool_h'] = spec['pool_w'] = pool
        layers.append(spec)
    layers.append(_conv_spec(sd, 'head', first=False))
    layers.append(dict(kind=_lib.UPSAMPLE, pool_h=4, pool_w=4, cin=layers[-1]['cout'], cout=layers[-1]['cout']))
    return layers


def to_ctypes(layers, precision, line_height, device):
    """-> (NetDesc, keepalive list).  Arrays must outlive b200ocr_create (it copies everything to the device)."""
    import ctypes as C
    keep = []
    arr = (_lib.Layer * len(layers))()

    def ptr(a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_float))

    for i, spec in enumerate(layers):
        ly = arr[i]
        for key in ('kind', 'cin', 'cout', 'kh', 'kw', 'pad_h', 'pad_w', 'act', 'pool_h', 'pool_w', 'hidden', 'heads',
                    'dim_ff'):
            if key in spec:
                setattr(ly, key, int(spec[key]))
        ly.act_slope = float(spec.get('act_slope', 0.0))
        for key in ('weight', 'bias', 'post_scale', 'post_shift', 'in_proj_w', 'in_proj_b', 'out_proj_w', 'out_proj_b',
                    'lin1_w', 'lin1_b', 'lin2_w', 'lin2_b', 'norm1_w', 'norm1_b', 'norm2_w', 'norm2_b'):
            if spec.get(key) is not None:
                setattr(ly, key, ptr(spec[key]))
        for key in ('w_ih', 'w_hh', 'b_ih', 'b_hh'):
            if key in spec:
                pair = getattr(ly, key)
                pair[0], pair[1] = ptr(spec[key][0]), ptr(spec[key][1])
    desc = _lib.NetDesc(len(layers), arr, _lib.PRECISIONS[precision], int(line_height), int(device))
    keep.append(arr)
    return desc, keep
