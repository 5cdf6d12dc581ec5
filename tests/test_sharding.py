"""CPU, world_size 2 over gloo: the multi-GPU shard / gather logic with a stand-in engine."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pero_ocr_b200.sharding import ShardedLineOCR, shard_indices


class _FakeEngine:
    """Label ids of a line = a function of its content only (so sharding cannot change the answer)."""
    characters = [chr(ord('a') + i) for i in range(26)]

    def process_lines(self, lines, no_logits=True, return_ids=True):
        ids = [np.array([int(l[0, j, 0]) % 26 for j in range(0, l.shape[1], 7)], dtype=np.int32) for l in lines]
        return ids, [None] * len(lines), [None] * len(lines)


def _lines():
    rng = np.random.default_rng(4)
    return [rng.integers(0, 256, (40, int(w), 3), dtype=np.uint8) for w in rng.integers(8, 300, 23)] + \
           [np.zeros((40, 0, 3), dtype=np.uint8)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lines = _lines()
        texts, _, _ = ShardedLineOCR(_FakeEngine()).process_lines(lines)
        q.put((rank, texts))
    finally:
        dist.destroy_process_group()


def test_shard_indices_partition():
    widths = [5, 100, 7, 64, 64, 3, 900]
    parts = [shard_indices(widths, 3, r) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(len(widths)))
    assert parts[0][0] == 6                                  # widest line goes to rank 0
    loads = [sum(widths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(widths)


def test_two_rank_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng = _FakeEngine()
    lines = _lines()
    ids, _, _ = eng.process_lines(lines)
    want = [''.join(eng.characters[c] for c in v) for v in ids]
    assert got[0] == want and got[1] == want
