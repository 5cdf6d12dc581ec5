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
print('slots keys', [sorted(k for k in sl.keys()) for sl in (eng._slots or [])][:1])
print('gc.garbage', len(gc.garbage))
import threading
print('threads', [t.name for t in threading.enumerate()])
eng._slots = None
gc.collect()
print('after dropping slots', {k: len(v) for k, v in pool.free.items()}, w_blk() is not None)
eng._executor = None
gc.collect()
print('after dropping executor', {k: len(v) for k, v in pool.free.items()}, w_blk() is not None)
models = eng._models
del eng
gc.collect()
print('after dropping engine', {k: len(v) for k, v in pool.free.items()}, w_blk() is not None)
del models
gc.collect()
print('after dropping models', {k: len(v) for k, v in pool.free.items()}, w_blk() is not None)
import scipy.sparse as sp
objs = [o for o in gc.get_objects() if isinstance(o, sp.csc_matrix)]
print('live csc matrices', len(objs))
if objs:
    for r in gc.get_referrers(objs[0])[:5]:
        d = type(r).__name__
        if isinstance(r, (list, tuple)): d += f' len={len(r)}'
        if isinstance(r, dict): d += ' keys=' + str(list(r.keys())[:8])
        print('  csc referrer:', d)
        for r2 in gc.get_referrers(r)[:4]:
            d2 = type(r2).__name__
            if isinstance(r2, dict): d2 += ' keys=' + str(list(r2.keys())[:8])
            if hasattr(r2, 'gi_code'): d2 += ' ' + r2.gi_code.co_name
            if hasattr(r2, 'f_code'): d2 += ' ' + r2.f_code.co_name
            print('     <-', d2[:160])
