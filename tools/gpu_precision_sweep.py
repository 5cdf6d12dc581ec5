"""GPU session helper: what each arithmetic choice of the fp16f8 engine costs and buys on the config-2 net.

    python tools/gpu_precision_sweep.py > gpurun_out/<tag>_precision_sweep.json

For autotune budgets 0 .. 1e-3: the layers switched to the weight-side-only correction, executed pass-equivalents,
device-resident lines/s (256 x 40 x 1344, CUDA events), and the error that matters -- max |logit - torch-CPU fp32
module| and per-frame argmax agreement on 6 full-width lines (the CPU side is the seeded nn.Module of
pero_ocr_b200/synthetic.py itself, i.e. what the reference engine would host).  Also the BiLSTM recurrence with and
without the fp16 residual of h_t (debug flag 3).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pero_ocr_b200 import _lib, netdesc, synthetic          # noqa: E402
from pero_ocr_b200.engine import LineRecognizer             # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    net = synthetic.make_net('lstm', 120, seed=0, out_gain=6.0)
    layers, _ = netdesc.describe_line_net(net)
    rec = LineRecognizer(layers, precision='fp16f8')
    crops = np.zeros((256, 40, 1344, 3), dtype=np.uint8)
    crops[:, :, 32:-32] = synthetic.bench_crops(256, 1280, seed=0)
    dev = torch.from_numpy(crops).cuda()
    n_ref = 6
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = net(torch.from_numpy(crops[:n_ref]).float().div(255.0).permute(0, 3, 1, 2)).permute(0, 2, 1).numpy()
    srt = np.sort(ref, axis=2)
    margin = srt[..., -1] - srt[..., -2]
    outs = {}
    report = {'reference_lines': n_ref, 'logit_absmax': float(np.abs(ref).max()),
              'top2_margin_quantiles': {q: float(np.quantile(margin, q)) for q in (0.001, 0.01, 0.1, 0.5)}, 'budgets': []}

    def measure(tag, extra):
        got = rec.forward(dev[:n_ref].contiguous(), want_logits=True, out={})['logits'].cpu().numpy()
        ms = timed(lambda: rec.forward(dev, want_logits=False, out=outs))
        flips = got.argmax(2) != ref.argmax(2)
        total, per = rec.executed_passes(256, 1344)
        row = {'tag': tag, 'max_abs_err_vs_fp32': float(np.abs(got - ref).max()), 'rms_err': float(np.sqrt(((got - ref) ** 2).mean())),
               'argmax_flips': int(flips.sum()), 'frames': int(flips.size),
               'largest_margin_of_a_flipped_frame': float(margin[flips].max()) if flips.any() else 0.0,
               'executed_passes': total, 'ms_per_step': ms, 'lines_per_s': 256 / (ms / 1e3)}
        row.update(extra)
        return row

    for budget in (0.0, 1e-4, 2e-4, 3e-4, 5e-4, 1e-3):
        tuned = rec.autotune_precision(budget=budget)
        report['budgets'].append(measure(f'autotune budget {budget:g}', {'autotune': tuned}))
    # every conv + projection weight-only / no correction at all
    gemm = [i for i, l in enumerate(layers) if l['kind'] in (_lib.CONV, _lib.BILSTM, _lib.CTC_HEAD)]
    for name, mode in (('all layers weight-only', _lib.CORR_WEIGHT), ('all layers plain fp16', _lib.CORR_NONE)):
        for i in gemm:
            rec.set_layer_correction(i, mode)
        report['budgets'].append(measure(name, {}))
    for i in gemm:
        rec.set_layer_correction(i, _lib.CORR_BOTH)
    # BiLSTM recurrence: h_t residual exchanged (flag 3 = 1, three passes) or not (default)
    lstm = {}
    for flag in (0, 1):
        rec.set_flag(3, flag)
        row = measure(f'lstm h residual {"on" if flag else "off"}', {})
        rec.profile(True)
        rec.forward(dev, want_logits=False, out=outs)
        tags, lidx, ms = rec.profile_read()
        rec.profile(False)
        row['lstm_ms_per_step'] = float(ms[tags == 2].sum())
        lstm['h_residual_on' if flag else 'h_residual_off'] = row
    rec.set_flag(3, 0)
    report['lstm'] = lstm
    print(json.dumps(report, indent=1))


if __name__ == '__main__':
    main()
