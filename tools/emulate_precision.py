"""CPU emulation of the operand-splitting schemes of the tensor-core path (design aid, not product code).

    python tools/emulate_precision.py [width] [lines]

Every dense contraction F(a, w) of the recogniser is evaluated as
  fp16    : F(h(a), h(w))
  fp16x3  : F(ah, wh) + F(ah, wl) + F(al, wh)                       (ah = fp16(a), al = fp16(a - ah))
  fp16f8  : F(ah, wh) + 2^-11 [ F(q(al 2^11), q(wh)) + F(q(ah), q(wl 2^11)) ]      (q = e5m2 rounding)
and the logits are compared with the fp32 module.  fp32 accumulation everywhere (as in TMEM).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pero_ocr_b200 import synthetic  # noqa: E402

S = 2048.0


def h(x):
    return x.half().float()


def q(x, fmt):
    return x.to(fmt).float()


def contract(F, a, w, mode, fmt=torch.float8_e5m2):
    if mode == 'fp32':
        return F(a, w)
    ah, wh = h(a), h(w)
    if mode == 'fp16':
        return F(ah, wh)
    al, wl = h(a - ah), h(w - wh)
    if mode == 'fp16x3':
        return F(ah, wh) + F(ah, wl) + F(al, wh)
    if mode == 'fp16x2a':   # activations split, weights single
        return F(ah, wh) + F(al, wh)
    if mode == 'fp16f8':
        return F(ah, wh) + (F(q(al * S, fmt), q(wh, fmt)) + F(q(ah, fmt), q(wl * S, fmt))) / S
    raise ValueError(mode)


def forward(net, x, mode, fmt=torch.float8_e5m2):
    mods = list(net.conv)
    y = x
    i = 0
    first = True
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Conv2d):
            if first:       # first conv runs in fp32 on CUDA cores / exact integer operands
                y = Fn.conv2d(y, m.weight, m.bias, padding=1)
                first = False
            else:
                y = contract(lambda a, w: Fn.conv2d(a, w, None, padding=1), y, m.weight, mode, fmt) + m.bias.view(1, -1, 1, 1)
        else:
            y = m(y)
        i += 1
    y = contract(lambda a, w: Fn.conv2d(a, w, None), y, net.agg.weight, mode, fmt) + net.agg.bias.view(1, -1, 1, 1)
    y = net.agg_act(y)
    y = y.squeeze(2).permute(2, 0, 1)          # [T,N,512]
    T, N, _ = y.shape
    H = net.lstm.hidden_size
    for layer in range(net.lstm.num_layers):
        outs = []
        for d, suf in enumerate(['', '_reverse']):
            w_ih = getattr(net.lstm, f'weight_ih_l{layer}{suf}')
            w_hh = getattr(net.lstm, f'weight_hh_l{layer}{suf}')
            b = getattr(net.lstm, f'bias_ih_l{layer}{suf}') + getattr(net.lstm, f'bias_hh_l{layer}{suf}')
            pre = contract(lambda a, w: a @ w.t(), y.reshape(T * N, -1), w_ih, mode, fmt).view(T, N, 4 * H) + b
            hs = torch.zeros(N, H)
            c = torch.zeros(N, H)
            seq = [None] * T
            order = range(T) if d == 0 else range(T - 1, -1, -1)
            for t in order:
                g = pre[t] + contract(lambda a, w: a @ w.t(), hs, w_hh, mode, fmt)
                i_, f_, g_, o_ = g.chunk(4, dim=1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                hs = torch.sigmoid(o_) * torch.tanh(c)
                seq[t] = hs
            outs.append(torch.stack(seq))
        y = torch.cat(outs, dim=2)
    y = contract(lambda a, w: a @ w.t(), y.reshape(T * N, -1), net.out.weight, mode, fmt).view(T, N, -1) + net.out.bias
    return y.permute(1, 2, 0)


def main():
    width = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lines = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
    crops = synthetic.bench_crops(lines, width, seed=3)
    x = torch.from_numpy(crops).float().div(255.0).permute(0, 3, 1, 2)
    with torch.no_grad():
        ref = net(x)
        chk = forward(net, x, 'fp32')
        print('restatement vs module: %.2e   max|logit| %.2f' % (float((ref - chk).abs().max()), float(ref.abs().max())))
        for mode, fmt in [('fp16', None), ('fp16x2a', None), ('fp16x3', None), ('fp16f8', torch.float8_e5m2),
                          ('fp16f8', torch.float8_e4m3fn)]:
            out = forward(net, x, mode, fmt)
            d = (out - ref).abs()
            agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
            print('%-8s %-22s max|d| %.2e  rms %.2e  argmax agreement %.4f' % (mode, str(fmt), float(d.max()),
                                                                             float(d.pow(2).mean().sqrt()), agree))


if __name__ == '__main__':
    main()
