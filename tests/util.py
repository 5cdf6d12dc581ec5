"""Shared helpers for the test-suite (test infrastructure; may use oracle/)."""
import json
import os

import numpy as np

from oracle import cases
from oracle.nets import make_net


def make_case_net(kind):
    spec = cases.ENGINE_CASES[kind]
    return make_net(spec.get('net', kind), spec['classes'], seed=spec['seed'], out_gain=spec['out_gain'], **spec['net_kw'])


def write_engine_json(tmpdir, kind, checkpoint='ck.pt', embed_id=None):
    spec = cases.ENGINE_CASES[kind]
    path = os.path.join(str(tmpdir), f'{kind}.json')
    with open(path, 'w', encoding='utf8') as f:
        cfg = {'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': checkpoint,
               'characters': cases.json_characters(spec.get('json_chars', spec['classes'] - 2)), 'net_name': 'B200_TEST'}
        if embed_id is not None or 'embed_id' in spec:
            cfg['embed_id'] = str(embed_id if embed_id is not None else spec['embed_id'])
        json.dump(cfg, f)
    return path


def load_golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))
