"""CPU study (design aid, not product code): per-layer sensitivity of the recogniser's logits to the precision of each
dense contraction, on top of tools/emulate_precision.py.

    python tools/emulate_mixed_precision.py [width] [lines]

Modes per contraction: fp16 (1 tensor-core pass), fp16f8 (fp16 pass + both e5m2 first-order corrections: 2
pass-equivalents), fp16f8a (fp16 pass + the ACTIVATION correction al*wh only: 1.5), fp16f8w (+ the WEIGHT correction
ah*wl only: 1.5).  Question: which layers need which corrections for the logits to stay within the parity bar.
"""
import os
import sys

import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pero_ocr_b200 import synthetic  # noqa: E402
from emulate_precision import S, h, q  # noqa: E402

E5 = torch.float8_e5m2
COST = {'fp32': 0.0, 'fp16': 1.0, 'fp16f8': 2.0, 'fp16f8a': 1.5, 'fp16f8w': 1.5, 'fp16x3': 3.0}


def contract(F, a, w, mode):
    if mode == 'fp32':
        return F(a, w)
    ah, wh = h(a), h(w)
    al, wl = h(a - ah), h(w - wh)
    out = F(ah, wh)
    if mode == 'fp16':
        return out
    if mode == 'fp16x3':
        return out + F(ah, wl) + F(al, wh)
    if mode in ('fp16f8', 'fp16f8a'):
        out = out + F(q(al * S, E5), q(wh, E5)) / S
    if mode in ('fp16f8', 'fp16f8w'):
        out = out + F(q(ah, E5), q(wl * S, E5)) / S
    return out


def layer_names(net):
    names = [f'conv{i}' for i, m in enumerate([m for m in net.conv if isinstance(m, torch.nn.Conv2d)]) if i > 0]
    names += ['agg']
    for layer in range(net.lstm.num_layers):
        names += [f'lstm{layer}_ih', f'lstm{layer}_hh']
    return names + ['out']


def layer_flops(net, width):
    """GFLOP per line of each contraction (2 * MACs), same order as layer_names."""
    fl, hh, ww = [], 40, width
    convs = [m for m in net.conv]
    first = True
    for i, m in enumerate(convs):
        if isinstance(m, torch.nn.Conv2d):
            if not first:
                fl.append(2.0 * hh * ww * m.in_channels * m.out_channels * 9 / 1e9)
            first = False
        elif isinstance(m, torch.nn.MaxPool2d):
            k = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size, m.kernel_size)
            hh, ww = hh // k[0], ww // k[1]
    fl.append(2.0 * ww * 512 * 512 * 5 / 1e9)
    H = net.lstm.hidden_size
    d_in = 512
    for layer in range(net.lstm.num_layers):
        fl.append(2.0 * ww * d_in * 8 * H / 1e9)
        fl.append(2.0 * ww * H * 8 * H / 1e9)
        d_in = 2 * H
    fl.append(2.0 * ww * 2 * H * net.out.out_features / 1e9)
    return fl


def forward(net, x, modes):
    it = iter(modes)
    y = x
    first = True
    for m in net.conv:
        if isinstance(m, torch.nn.Conv2d):
            if first:
                y = Fn.conv2d(y, m.weight, m.bias, padding=1)
                first = False
            else:
                y = contract(lambda a, w: Fn.conv2d(a, w, None, padding=1), y, m.weight, next(it)) + m.bias.view(1, -1, 1, 1)
        else:
            y = m(y)
    y = contract(lambda a, w: Fn.conv2d(a, w, None), y, net.agg.weight, next(it)) + net.agg.bias.view(1, -1, 1, 1)
    y = net.agg_act(y).squeeze(2).permute(2, 0, 1)
    T, N, _ = y.shape
    H = net.lstm.hidden_size
    for layer in range(net.lstm.num_layers):
        m_ih, m_hh = next(it), next(it)
        outs = []
        for d, suf in enumerate(['', '_reverse']):
            w_ih = getattr(net.lstm, f'weight_ih_l{layer}{suf}')
            w_hh = getattr(net.lstm, f'weight_hh_l{layer}{suf}')
            b = getattr(net.lstm, f'bias_ih_l{layer}{suf}') + getattr(net.lstm, f'bias_hh_l{layer}{suf}')
            pre = contract(lambda a, w: a @ w.t(), y.reshape(T * N, -1), w_ih, m_ih).view(T, N, 4 * H) + b
            hs, c = torch.zeros(N, H), torch.zeros(N, H)
            seq = [None] * T
            for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
                g = pre[t] + contract(lambda a, w: a @ w.t(), hs, w_hh, m_hh)
                i_, f_, g_, o_ = g.chunk(4, dim=1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                hs = torch.sigmoid(o_) * torch.tanh(c)
                seq[t] = hs
            outs.append(torch.stack(seq))
        y = torch.cat(outs, dim=2)
    y = contract(lambda a, w: a @ w.t(), y.reshape(T * N, -1), net.out.weight, next(it)).view(T, N, -1) + net.out.bias
    return y.permute(1, 2, 0)


def report(tag, out, ref, cost):
    d = (out - ref).abs()
    print('%-44s max|d| %.2e  rms %.2e  argmax agree %.4f  tensor passes (FLOP-weighted) %.2f' % (
        tag, float(d.max()), float(d.pow(2).mean().sqrt()), float((out.argmax(1) == ref.argmax(1)).float().mean()), cost))


def main():
    width = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lines = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
    names = layer_names(net)
    fl = layer_flops(net, width)
    x = torch.from_numpy(synthetic.bench_crops(lines, width, seed=3)).float().div(255.0).permute(0, 3, 1, 2)
    hh_idx = [i for i, n in enumerate(names) if n.endswith('_hh')]

    def cost(modes):      # the recurrence runs its own three-pass kernel: not part of the GEMM budget
        idx = [i for i in range(len(names)) if i not in hh_idx]
        return sum(fl[i] * COST[modes[i]] for i in idx) / sum(fl[i] for i in idx)

    def base(mode):
        return [('fp16x3' if i in hh_idx else mode) for i in range(len(names))]

    with torch.no_grad():
        ref = net(x)
        print('max|logit| %.2f;  GFLOP/line per contraction: %s' % (float(ref.abs().max()),
              ', '.join('%s %.2f' % (n, f) for n, f in zip(names, fl))))
        for mode in ('fp16', 'fp16f8a', 'fp16f8w', 'fp16f8'):
            m = base(mode)
            report('all ' + mode, forward(net, x, m), ref, cost(m))
        print('--- one contraction degraded to fp16 / fp16f8a / fp16f8w, the rest fp16f8')
        for i, n in enumerate(names):
            if i in hh_idx:
                continue
            for mode in ('fp16', 'fp16f8a', 'fp16f8w'):
                m = base('fp16f8')
                m[i] = mode
                report(f'{n} -> {mode}', forward(net, x, m), ref, cost(m))


if __name__ == '__main__':
    main()
