import gc, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import cases
from oracle.nets import make_net
from tests.util import write_engine_json
from pero_ocr_b200.engine import B200EngineLineOCR
import tempfile
tmp = tempfile.mkdtemp()
spec = cases.ENGINE_CASES['lstm']
js = write_engine_json(tmp, 'lstm')
net = make_net('lstm', spec['classes'], seed=3, out_gain=0.05, **spec['net_kw'])
rng = np.random.default_rng(5)
lines = [rng.integers(0, 256, size=(40, 600 + 8 * (i % 5), 3), dtype=np.uint8) for i in range(48)]
eng = B200EngineLineOCR(js, torch.device('cuda', 0), batch_size=64, precision='fp16f8', module=net, pinned_logit_bytes=1 << 30)
out = eng.process_lines([l.copy() for l in lines])
pool = eng.pinned_pool
print('stats', pool.stats, 'free', {k: len(v) for k, v in pool.free.items()})
import weakref
m0 = out[1][0]
print('types', type(m0.indices.base), type(getattr(m0.indices.base, 'base', None)))
w_arr = weakref.ref(m0.indices.base) if isinstance(m0.indices.base, np.ndarray) else None
blkobj = m0.indices.base.base if isinstance(m0.indices.base, np.ndarray) else m0.indices.base
w_blk = weakref.ref(blkobj)
print('block object', type(blkobj))
del m0, blkobj, out
for i in range(3):
    gc.collect()
    print('after gc', i, {k: len(v) for k, v in pool.free.items()}, 'array alive', w_arr is not None and w_arr() is not None, 'block alive', w_blk() is not None)
def show(obj, depth=0, seen=None):
    seen = seen or set()
    if depth > 3 or id(obj) in seen:
        return
    seen.add(id(obj))
    for r in gc.get_referrers(obj):
        if r is seen or isinstance(r, type(sys._getframe())):
            continue
        desc = type(r).__name__
        if isinstance(r, dict):
            desc += ' keys=' + str(list(r.keys())[:6])
        elif isinstance(r, (list, tuple)):
            desc += f' len={len(r)}'
        print('  ' * depth + '<- ' + desc[:150])
        if not isinstance(r, dict) or depth < 2:
            show(r, depth + 1, seen)
if w_blk() is not None:
    show(w_blk())
