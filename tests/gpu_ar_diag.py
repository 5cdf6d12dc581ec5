"""Bring-up diagnostics + timing of the autoregressive decoder on a GPU box (not a test; uses the oracle as the
checker).  `python -m tests.gpu_ar_diag [out.json]`: stage-wise errors against the reference golden (encoder memory,
per-step logits) and the time of b200ocr_ar_transcribe on the golden batch and on a larger synthetic batch."""
import json
import sys
import tempfile
import time

import numpy as np
import torch

from oracle import cases
from tests.util import load_golden
from tests.conftest import GOLDEN


def main():
    from pero_ocr_b200.transformer_engine import B200TransformerEngineLineOCR
    out = {}
    net, dec, sd = cases.ar_state_dict()
    with tempfile.TemporaryDirectory() as tmp:
        eng = B200TransformerEngineLineOCR(cases.write_ar_engine_json(tmp), torch.device('cuda', 0), state_dict=sd)
    gold = load_golden(GOLDEN, 'ar_decoder.npz')
    x = cases.ar_inputs()
    try:
        nhwc = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).cuda()
        mem = eng.net.debug_forward_prefix(nhwc, 1 << 20)                        # [n][1][T][D]
        with torch.no_grad():
            xf = torch.from_numpy(x).float() / 255.0
            y = net.agg_act(net.agg(net.conv(xf))).squeeze(2).permute(2, 0, 1)
            ref_mem = net.trans_encoder(net.input_norm(y) + net.pe[:y.size(0)]).numpy()   # [T][n][D]
        out['memory_max_abs_err'] = float(np.abs(mem[:, 0].transpose(1, 0, 2) - ref_mem).max())
        out['memory_max_abs'] = float(np.abs(ref_mem).max())
    except Exception as e:                                                        # noqa: BLE001
        out['memory_error'] = repr(e)
    t0 = time.perf_counter()
    outs, logits = eng.transcribe_batch(x)
    out['golden_batch_first_call_s'] = time.perf_counter() - t0
    steps = min(logits.shape[1], gold['logits'].shape[1])
    diff = np.abs(logits[:, :steps] - gold['logits'][:, :steps]).max(axis=2)     # [n][steps]
    out['steps'] = [int(logits.shape[1]), int(gold['logits'].shape[1])]
    out['logit_err_step0'] = diff[:, 0].tolist()
    out['logit_err_step1'] = diff[:, 1].tolist() if steps > 1 else None
    out['logit_err_max_per_line'] = diff.max(axis=1).tolist()
    out['first_step_over_2e-3'] = [int(np.argmax(d > 2e-3)) if (d > 2e-3).any() else -1 for d in diff]
    out['argmax_equal_steps'] = (logits[:, :steps].argmax(2) == gold['logits'][:, :steps].argmax(2)).sum(axis=1).tolist()
    out['lengths'] = [[len(o) for o in outs], gold['lengths'].tolist()]
    # the split-K variant of the step projections (b200ocr_debug_set_flag 2): same checks against the golden
    eng.net.set_flag(2, 1)
    outs_v, logits_v = eng.transcribe_batch(x)
    steps_v = min(logits_v.shape[1], gold['logits'].shape[1])
    diff_v = np.abs(logits_v[:, :steps_v] - gold['logits'][:, :steps_v]).max(axis=2)
    out['splitk'] = {
        'steps': int(logits_v.shape[1]), 'logit_err_max_per_line': diff_v.max(axis=1).tolist(),
        'argmax_equal_steps': (logits_v[:, :steps_v].argmax(2) == gold['logits'][:, :steps_v].argmax(2)).sum(axis=1).tolist(),
        'lengths': [len(o) for o in outs_v],
        'tokens_equal_default': bool(all(np.array_equal(a, b) for a, b in zip(outs, outs_v))),
        'max_abs_diff_to_default': float(np.abs(logits_v[:, :steps] - logits[:, :steps]).max()) if logits_v.shape == logits.shape else None}
    # timing of both variants, crops resident on the device: golden batch (warm), 64 and 256 lines x 1088 px
    sb = cases.AR_CASE['classes'] - 2
    for variant in (0, 1, 2, 3):
        eng.net.set_flag(2, variant)
        for name, n in (('n3', 3), ('n64', 64), ('n256', 256)):
            rng = np.random.default_rng(5)
            b = rng.integers(0, 256, (n, 3, 40, 1088), dtype=np.uint8) if n != 3 else x
            dev = torch.from_numpy(np.ascontiguousarray(b.transpose(0, 2, 3, 1))).cuda()
            eng.net.transcribe(dev, sb, want_logits=False)
            l0 = eng.net.launch_count
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, _, k = eng.net.transcribe(dev, sb, want_logits=False)      # synchronises (returns the step count)
            dt = time.perf_counter() - t0
            out[f'time_v{variant}_{name}'] = {'seconds': dt, 'lines_per_s': n / dt, 'steps': k,
                                             'ms_per_step': 1e3 * dt / k, 'launches': eng.net.launch_count - l0}
    # second golden case (not yet part of the GPU test-suite): 3 decoder layers, 122 classes (partial 32-wide output
    # tiles, several argmax passes per warp), no line stops -> the loop ends on the length limit after W/4 + 1 steps
    try:
        spec = cases.AR_CASES['wide']
        gold_w = load_golden(GOLDEN, spec['golden'])
        _, _, sd_w = cases.ar_state_dict(spec)
        with tempfile.TemporaryDirectory() as tmp:
            eng_w = B200TransformerEngineLineOCR(cases.write_ar_engine_json(tmp, spec=spec), torch.device('cuda', 0),
                                                 state_dict=sd_w)
        outs_w, logits_w = eng_w.transcribe_batch(cases.ar_inputs(spec))
        k = min(logits_w.shape[1], gold_w['logits'].shape[1])
        d = np.abs(logits_w[:, :k] - gold_w['logits'][:, :k]).max(axis=2)
        same = logits_w[:, :k].argmax(2) == gold_w['logits'][:, :k].argmax(2)
        first_diff = [int(np.argmin(r)) if not r.all() else -1 for r in same]
        out['wide_case'] = {'steps': [int(logits_w.shape[1]), int(gold_w['logits'].shape[1])],
                            'lengths': [[len(o) for o in outs_w], gold_w['lengths'].tolist()],
                            'first_step_with_other_argmax': first_diff,
                            'logit_err_max_before_divergence': [float(d[i, :(f if f >= 0 else k)].max()) if (f != 0) else None
                                                                for i, f in enumerate(first_diff)]}
        eng_w.net.close()
    except Exception as e:                                                        # noqa: BLE001
        out['wide_case_error'] = repr(e)
    # the same work on the host cores: torch-CPU encoder + NumPy decoder loop of the oracle (golden batch, 3 lines)
    from oracle.ar_oracle import greedy_transcribe
    sb = cases.AR_CASE['classes'] - 2
    t0 = time.perf_counter()
    with torch.no_grad():
        xf = torch.from_numpy(x).float() / 255.0
        y = net.agg_act(net.agg(net.conv(xf))).squeeze(2).permute(2, 0, 1)
        m = net.trans_encoder(net.input_norm(y) + net.pe[:y.size(0)]).numpy()
    t1 = time.perf_counter()
    greedy_transcribe(m, dec, cases.AR_CASE['decoder_layers'], 8, sb, x.shape[3])
    t2 = time.perf_counter()
    out['cpu_oracle_n3'] = {'encoder_s': t1 - t0, 'decoder_s': t2 - t1, 'lines_per_s': 3 / (t2 - t0),
                            'threads': torch.get_num_threads()}
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            f.write(text)


if __name__ == '__main__':
    main()
