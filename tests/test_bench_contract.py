"""CPU: the reference arm of bench.py prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '0', '--ref-lines', '1'], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'lines/s' and d['higher_is_better'] is True
    assert d['metric'] == 'text-lines/sec (40x1280 crops)' and d['value'] > 0
    for key in ('n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
                'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_product_arm_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the default arm exits non-zero and prints no result line."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '3'],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode != 0
    assert not [l for l in proc.stdout.splitlines() if l.strip().startswith('{')], proc.stdout
