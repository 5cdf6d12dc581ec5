"""Checkpoint -> layer list for b200ocr_create.

The reference hosts the recogniser as an opaque TorchScript blob (pero_ocr/ocr_engine/pytorch_ocr_engine.py:52-57)
and ParseNet likewise (pero_ocr/layout_engines/torch_parsenet.py:11-15).  A TorchScript module still exposes its
module tree (``original_name``, ``named_children``) and ``state_dict()``, so the same checkpoint file the reference
loads is walked here and turned into the flat layer list of ``include/b200_lineocr.h``.  Supported vocabulary:

  line recogniser:  ``conv`` = Sequential of Conv2d(3x3, pad 1) / ReLU / LeakyReLU / MaxPool2d / BatchNorm2d,
                    ``agg`` = Conv2d(k = (H/8, 1)), ``agg_act``, then either ``lstm`` (bidirectional nn.LSTM) or
                    ``input_norm`` + ``trans_encoder`` (post-LN nn.TransformerEncoder, ReLU), then ``out`` (Linear).
  page detector:    ``e1 e2 | pool | e3 e4 | pool | d1 d2 | head`` 3x3 convs + nearest x4 upsampling.

Anything else raises -- there is no generic fallback executor.
"""
import numpy as np

from . import _lib


def _name(m):
    return getattr(m, 'original_name', type(m).__name__)


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def _act_code(m):
    n = _name(m)
    if n == 'ReLU':
        return _lib.ACT_RELU
    if n == 'LeakyReLU':
        slope = float(getattr(m, 'negative_slope', 0.01))
        if abs(slope - 0.01) > 1e-12:
            raise ValueError(f'LeakyReLU slope {slope} not supported (kernels use 0.01)')
        return _lib.ACT_LEAKY_RELU
    raise ValueError(f'unsupported activation {n}')


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


def _conv_spec(sd, prefix, first):
    w = _f32(sd[prefix + '.weight'])
    b = _f32(sd[prefix + '.bias']) if prefix + '.bias' in sd else None
    cout, cin, kh, kw = w.shape
    return dict(kind=_lib.CONV_FIRST if first else _lib.CONV, cin=cin, cout=cout, kh=kh, kw=kw,
                pad_h=(kh - 1) // 2, pad_w=(kw - 1) // 2, act=_lib.ACT_NONE, pool_h=1, pool_w=1, weight=w, bias=b)


def describe_line_net(module):
    """nn.Module or TorchScript module following the vocabulary above -> (layer specs, num_classes)."""
    sd = {k: v for k, v in module.state_dict().items()}
    children = dict(module.named_children())
    layers = []
    last_conv = None
    for idx, (cname, m) in enumerate(children['conv'].named_children()):
        n = _name(m)
        key = f'conv.{cname}'
        if n == 'Conv2d':
            spec = _conv_spec(sd, key, first=(last_conv is None and not layers))
            if (spec['kh'], spec['kw']) != (3, 3):
                raise ValueError('frontend convolutions must be 3x3')
            layers.append(spec)
            last_conv = spec
        elif n in ('ReLU', 'LeakyReLU'):
            if last_conv is None or last_conv['act'] != _lib.ACT_NONE or 'post_scale' in last_conv:
                raise ValueError('activation must directly follow a convolution')
            last_conv['act'] = _act_code(m)
        elif n == 'MaxPool2d':
            ph, pw = _pair(m.kernel_size)
            if (ph, pw) == (1, 1):
                continue
            if last_conv is None or (last_conv['pool_h'], last_conv['pool_w']) != (1, 1) or 'post_scale' in last_conv \
                    or ph not in (1, 2) or pw not in (1, 2) or last_conv['kind'] == _lib.CONV_FIRST:
                raise ValueError('max-pool must be 2x2 / 2x1 / 1x2 and follow a tensor-core convolution')
            last_conv['pool_h'], last_conv['pool_w'] = ph, pw
        elif n == 'BatchNorm2d':
            # eval-mode BN folded to y * scale + shift, applied after activation (+pool) of the preceding conv
            g, b = _f32(sd[key + '.weight']), _f32(sd[key + '.bias'])
            mu, var = _f32(sd[key + '.running_mean']), _f32(sd[key + '.running_var'])
            eps = float(getattr(m, 'eps', 1e-5))
            scale = (g.astype(np.float64) / np.sqrt(var.astype(np.float64) + eps))
            shift = b.astype(np.float64) - mu.astype(np.float64) * scale
            if last_conv is None or last_conv['kind'] == _lib.CONV_FIRST:
                raise ValueError('BatchNorm must follow a tensor-core convolution')
            last_conv['post_scale'] = scale.astype(np.float32)
            last_conv['post_shift'] = shift.astype(np.float32)
        elif n == 'Dropout':
            continue
        else:
            raise ValueError(f'unsupported frontend module {n}')
    agg = _conv_spec(sd, 'agg', first=False)
    agg['pad_h'] = agg['pad_w'] = 0
    agg['act'] = _act_code(children['agg_act'])
    layers.append(agg)
    d_model = agg['cout']
    if 'lstm' in children:
        layer = 0
        cin = d_model
        while f'lstm.weight_ih_l{layer}' in sd:
            spec = dict(kind=_lib.BILSTM, cin=cin)
            for key, sfx in (('w_ih', 'weight_ih'), ('w_hh', 'weight_hh'), ('b_ih', 'bias_ih'), ('b_hh', 'bias_hh')):
                spec[key] = [_f32(sd[f'lstm.{sfx}_l{layer}']), _f32(sd[f'lstm.{sfx}_l{layer}_reverse'])]
            spec['hidden'] = spec['w_hh'][0].shape[1]
            layers.append(spec)
            cin = 2 * spec['hidden']
            layer += 1
        feat = cin
    elif 'trans_encoder' in children:
        layers.append(dict(kind=_lib.LN_PE, cin=d_model, norm1_w=_f32(sd['input_norm.weight']),
                           norm1_b=_f32(sd['input_norm.bias'])))
        layer = 0
        while f'trans_encoder.layers.{layer}.linear1.weight' in sd:
            p = f'trans_encoder.layers.{layer}.'
            heads = int(getattr(module, 'num_heads', 8))
            try:
                heads = int(dict(children['trans_encoder'].named_children())['layers'][layer].self_attn.num_heads)
            except Exception:
                pass
            layers.append(dict(
                kind=_lib.TRANSFORMER_LAYER, cin=d_model, heads=heads, dim_ff=sd[p + 'linear1.weight'].shape[0],
                in_proj_w=_f32(sd[p + 'self_attn.in_proj_weight']), in_proj_b=_f32(sd[p + 'self_attn.in_proj_bias']),
                out_proj_w=_f32(sd[p + 'self_attn.out_proj.weight']), out_proj_b=_f32(sd[p + 'self_attn.out_proj.bias']),
                lin1_w=_f32(sd[p + 'linear1.weight']), lin1_b=_f32(sd[p + 'linear1.bias']),
                lin2_w=_f32(sd[p + 'linear2.weight']), lin2_b=_f32(sd[p + 'linear2.bias']),
                norm1_w=_f32(sd[p + 'norm1.weight']), norm1_b=_f32(sd[p + 'norm1.bias']),
                norm2_w=_f32(sd[p + 'norm2.weight']), norm2_b=_f32(sd[p + 'norm2.bias'])))
            layer += 1
        feat = d_model
    else:
        raise ValueError('sequence encoder must be `lstm` or `input_norm` + `trans_encoder`')
    ow, ob = _f32(sd['out.weight']), _f32(sd['out.bias'])
    layers.append(dict(kind=_lib.CTC_HEAD, cin=feat, cout=ow.shape[0], kh=1, kw=1, weight=ow, bias=ob))
    return layers, ow.shape[0]


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
