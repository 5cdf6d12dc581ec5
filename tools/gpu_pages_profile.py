"""GPU session helper: where the host time of the page pipeline (BASELINE config 4) goes -- cProfile of the main thread
around process_pages / process_baselines, wall time of the preparation stage alone, of the ParseNet call alone.
`python tools/gpu_pages_profile.py > gpurun_out/<tag>_pages_profile.txt`"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
from pero_ocr_b200 import synthetic                        # noqa: E402
from pero_ocr_b200.cropper import B200LineCropper, DevicePage   # noqa: E402
from pero_ocr_b200.engine import B200EngineLineOCR         # noqa: E402
from pero_ocr_b200.parsenet import B200ParseNet            # noqa: E402

dev = torch.device('cuda', 0)
pn = B200ParseNet(None, dev, downsample=4, adaptive_downsample=False, module=synthetic.make_net('parsenet', seed=1))
eng = B200EngineLineOCR(bench.write_engine_json(), dev, batch_size=int(os.environ.get('B200OCR_PAGES_BATCH', '360')),
                        module=bench.make_net('lstm'), replicas=int(os.environ.get('B200OCR_PAGES_REPLICAS', '2')))
print('budget', eng.max_input_horizontal_pixels, 'replicas', len(eng._models))
cropper = B200LineCropper(line_height=40, poly=2, scale=1)
rng = np.random.default_rng(4)
imgs = [rng.integers(0, 256, (3000, 4000, 3), dtype=np.uint8) for _ in range(2)]
lines = []
for i in range(60):
    y = 60 + i * 48
    lines.append(([[100, y], [1400, y + rng.integers(-6, 7)], [2700, y + rng.integers(-6, 7)]], [26, 14]))


def source(n):
    for i in range(n):
        yield imgs[i & 1], lines


def timed(label, fn, reps=8):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f'{label}: {1e3 * (time.perf_counter() - t0) / reps:.2f} ms')


page = DevicePage(imgs[0])
fitted = [cropper.poly_params(b, h) for b, h in lines]
timed('poly_params x 60', lambda: [cropper.poly_params(b, h) for b, h in lines])
timed('DevicePage upload (36 MB pageable)', lambda: DevicePage(imgs[1]))
timed('parsenet.get_maps', lambda: pn.get_maps(imgs[0], 4))
timed('process_baselines prepared, no_logits', lambda: eng.process_baselines(page, lines, cropper, prepared=fitted, no_logits=True))
timed('process_baselines prepared, sparse logits', lambda: eng.process_baselines(page, lines, cropper, prepared=fitted))
for prefetch in (1, 2, 3):
    timed(f'process_pages 16 pages, prefetch {prefetch} (per page)',
          lambda: [0 for _ in eng.process_pages(source(16), cropper, parsenet=pn, parsenet_downsample=4, no_logits=True,
                                                prefetch=prefetch)], reps=2)
eng.host_ms = {k: 0.0 for k in eng.host_ms}
t0 = time.perf_counter()
stamps = []
for _ in eng.process_pages(source(32), cropper, no_logits=True, prefetch=3):
    stamps.append(time.perf_counter() - t0)
print('32 pages without ParseNet: host_ms totals', {k: round(v, 1) for k, v in eng.host_ms.items()}, {k: round(v, 1) for k, v in eng.page_ms.items()},
      'page completion times (ms)', [round(1e3 * x, 1) for x in stamps])
# device-resident pages, prepared up front: the GPU-side ceiling of the page path (two replicas overlapping)
pages_dev = [DevicePage(imgs[i & 1]) for i in range(2)]
torch.cuda.synchronize()
def resident(n):
    jobs = (eng._baseline_job(pages_dev[i & 1], fitted) for i in range(n))
    return [0 for _ in eng._run_jobs(jobs, True, False, True, False)]
timed('device-resident prepared pages through _run_jobs (per 16)', lambda: resident(16), reps=3)
timed('process_pages 16 pages without ParseNet (per 16)', lambda: [0 for _ in eng.process_pages(source(16), cropper, no_logits=True)], reps=2)

pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    eng.process_baselines(page, lines, cropper, prepared=fitted, no_logits=True)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumtime').print_stats(25)
print(s.getvalue())
pr = cProfile.Profile()
pr.enable()
for _ in eng.process_pages(source(16), cropper, parsenet=pn, parsenet_downsample=4, no_logits=True):
    pass
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue())
