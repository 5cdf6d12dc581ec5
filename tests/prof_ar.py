"""Short driver for ncu / compute-sanitizer runs of the autoregressive decoder (lives under tests/: the seeded decoder
weights come from oracle/ar_oracle.py).  `python -m tests.prof_ar [lines] [calls]`: builds the AR_CASE engine and
decodes a batch of `lines` random 40 x 1088 crops `calls` times, without logits."""
import sys
import tempfile

import numpy as np
import torch

from oracle import cases


def main():
    from pero_ocr_b200.transformer_engine import B200TransformerEngineLineOCR
    lines = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    _, _, sd = cases.ar_state_dict()
    with tempfile.TemporaryDirectory() as tmp:
        eng = B200TransformerEngineLineOCR(cases.write_ar_engine_json(tmp), torch.device('cuda', 0), state_dict=sd)
    rng = np.random.default_rng(5)
    dev = torch.from_numpy(rng.integers(0, 256, (lines, 40, 1088, 3), dtype=np.uint8)).cuda()
    for _ in range(calls):
        _, _, steps = eng.net.transcribe(dev, eng.sentence_boundary_ind, want_logits=False)
    print('done', steps, 'steps,', eng.net.launch_count, 'launches')


if __name__ == '__main__':
    main()
