"""GPU bring-up diagnostics (run on the B200 box): per-layer tcgen05 kernels vs CUDA-core cross-check kernels vs the
torch-CPU oracle.  Not a test -- prints a report."""
import sys
import os
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases
from oracle.nets import make_net
from oracle.forward_oracle import greedy_ctc_indices
from pero_ocr_b200 import netdesc
from pero_ocr_b200.engine import LineRecognizer


def oracle_prefix(net, x, n_layers, layers):
    """Activation after the first n_layers engine layers, NHWC, via the torch modules (frontend only)."""
    y = x
    li = 0
    mods = list(net.conv)
    i = 0
    while li < n_layers and i < len(mods):
        y = mods[i](y)              # conv
        i += 1
        while i < len(mods) and not isinstance(mods[i], torch.nn.Conv2d):
            y = mods[i](y)
            i += 1
        li += 1
    if li < n_layers:
        y = net.agg_act(net.agg(y))
        li += 1
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else 'lstm'
    n, w = 3, 192
    torch.manual_seed(0)
    net = make_net(kind, 120, seed=0, out_gain=6.0 if kind == 'lstm' else 2.5)
    layers, C = netdesc.describe_line_net(net)
    rng = np.random.default_rng(1)
    crops = rng.integers(0, 256, (n, 40, w, 3), dtype=np.uint8)
    x = torch.from_numpy(crops).float() / 255.0
    x = x.permute(0, 3, 1, 2)
    with torch.no_grad():
        ref_logits = net(x).permute(0, 2, 1).numpy()     # [N,T,C]
    dcrops = torch.from_numpy(crops).cuda()
    n_front = sum(1 for l in layers if l['kind'] in (1, 2))
    for prec in (sys.argv[2].split(',') if len(sys.argv) > 2 else ('fp16x3', 'fp16f8', 'fp16')):
        print(f'===== {kind} precision {prec}', flush=True)
        try:
            eng = LineRecognizer(layers, precision=prec)
        except Exception:
            traceback.print_exc()
            continue
        for L in range(1, len(layers)):
            try:
                eng.use_reference_kernels(True)
                a_ref = eng.debug_forward_prefix(dcrops, L)
                eng.use_reference_kernels(False)
                a_tc = eng.debug_forward_prefix(dcrops, L)
                msg = f'layer {L:2d} kind {layers[L-1]["kind"]} shape {a_tc.shape}: tc-vs-ref max|d| {np.abs(a_tc - a_ref).max():.3e} (ref absmax {np.abs(a_ref).max():.3e})'
                if L <= n_front:
                    with torch.no_grad():
                        o = oracle_prefix(net, x, L, layers)
                    if o.shape == a_ref.shape:
                        msg += f' | ref-vs-oracle {np.abs(a_ref - o).max():.3e} tc-vs-oracle {np.abs(a_tc - o).max():.3e}'
                    else:
                        msg += f' | oracle shape {o.shape}'
                print(msg, flush=True)
                bad = ~np.isfinite(a_tc)
                if bad.any():
                    print('   non-finite values in tc output:', int(bad.sum()))
            except Exception:
                traceback.print_exc()
                break
        for use_ref in (True, False):
            try:
                eng.use_reference_kernels(use_ref)
                torch.cuda.synchronize()
                t0 = time.time()
                o = eng.forward(dcrops, want_logits=True, want_confidence=True, want_best_path=True)
                torch.cuda.synchronize()
                dt = time.time() - t0
                lg = o['logits'].cpu().numpy()
                d = np.abs(lg - ref_logits)
                bp = o['best_path'].cpu().numpy()
                ref_bp = ref_logits.argmax(axis=2)
                gi = greedy_ctc_indices(ref_logits.transpose(0, 2, 1))
                lab, ln = o['labels'].cpu().numpy(), o['lengths'].cpu().numpy()
                same = all(list(lab[i, :ln[i]]) == list(gi[i]) for i in range(n))
                print(f'forward ref_kernels={use_ref}: {dt*1e3:.1f} ms  logits max|d| {d.max():.3e} mean {d.mean():.3e} '
                      f'(absmax {np.abs(ref_logits).max():.2f})  argmax mismatches {(bp != ref_bp).sum()}/{bp.size}  '
                      f'labels equal {same}  conf {o["confidence"].cpu().numpy()}', flush=True)
            except Exception:
                traceback.print_exc()
        eng.close()


if __name__ == '__main__':
    main()
