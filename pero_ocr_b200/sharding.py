"""Batch-parallel line recognition across the GPUs of one box.

The reference has no distributed code at all; lines are independent inside
``BaseEngineLineOCR.process_lines`` (pero_ocr/ocr_engine/line_ocr_engine.py:57-177), so the path shards by line:
one process per GPU (``torchrun``), every rank holds a full weight replica, lines are dealt round-robin in
width-sorted order (equal pixel load per rank), and ONE collective per call brings the decoded label ids back --
fixed-shape int32 records over NCCL (NVLink 5 / NVSwitch), or gloo in the CPU tests.  There is no data-path
collective inside the forward.
"""
import numpy as np


def shard_indices(widths, world_size, rank):
    """Indices of the lines rank `rank` processes: widest-first order dealt round-robin."""
    order = sorted(range(len(widths)), key=lambda i: -widths[i])
    return order[rank::world_size]


def _dist():
    import torch.distributed as dist
    return dist


def gather_ids(local_ids, local_index, total_lines, group=None, device=None):
    """Every rank contributes (global line index, int32 label-id array) pairs; returns the full list ordered by
    global index on EVERY rank (all_gather: works on nccl and gloo alike; ~1.4 KB per line)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' \
            else torch.device('cpu')
    n_local = len(local_ids)
    lens = np.fromiter((len(v) for v in local_ids), dtype=np.int64, count=n_local)
    l_max = int(lens.max()) if n_local else 0
    dims = torch.tensor([n_local, l_max], dtype=torch.int64, device=device)
    dist.all_reduce(dims, op=dist.ReduceOp.MAX, group=group)
    n_max, l_max = int(dims[0]), int(dims[1])
    rec = np.full((n_max, 2 + l_max), -1, dtype=np.int32)
    if n_local:
        rec[:n_local, 0] = np.asarray(local_index, dtype=np.int64)
        rec[:n_local, 1] = lens
        total = int(lens.sum())
        if total:                        # one vectorised scatter of all ids (100k lines: a Python loop per line shows)
            rows = np.repeat(np.arange(n_local), lens)
            starts = np.cumsum(lens) - lens
            cols = 2 + np.arange(total) - np.repeat(starts, lens)
            rec[rows, cols] = np.concatenate([np.asarray(v, dtype=np.int32) for v in local_ids if len(v)])
    mine = torch.from_numpy(rec).to(device)
    everyone = torch.empty((world * n_max, 2 + l_max), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(everyone, mine, group=group)
    everyone = everyone.cpu().numpy().reshape(world * n_max, 2 + l_max)
    return GatheredIds(everyone, total_lines), int(everyone.nbytes)


class GatheredIds:
    """The gathered records as a read-only sequence ordered by global line index: item g is the int32 label-id array
    of line g (a view of the gathered block), None for a line nobody contributed.  Built with two vectorised passes --
    at 100k lines a Python loop over the records costs more than the collective."""

    def __init__(self, records, total_lines):
        self._rec = records
        self._row = np.full(total_lines, -1, dtype=np.int64)
        valid = np.flatnonzero(records[:, 0] >= 0)
        self._row[records[valid, 0]] = valid
        self._len = np.zeros(total_lines, dtype=np.int64)
        self._len[records[valid, 0]] = records[valid, 1]

    def __len__(self):
        return len(self._row)

    def __getitem__(self, g):
        if isinstance(g, slice):
            return [self[i] for i in range(*g.indices(len(self)))]
        r = int(self._row[g])
        return None if r < 0 else self._rec[r, 2:2 + int(self._len[g])]

    def __iter__(self):
        return (self[g] for g in range(len(self)))


class ShardedLineOCR:
    """SPMD wrapper: every rank calls ``process_lines`` with the same list of lines; each recognises its shard on
    its own GPU; all ranks return the complete transcription list (rank 0 is the one callers normally use)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.characters = engine.characters
        self.gathered_bytes = 0

    def process_lines(self, lines, **kw):
        dist = _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        mine = shard_indices([l.shape[1] for l in lines], world, rank)
        ids, _, _ = self.engine.process_lines([lines[i] for i in mine], no_logits=True, return_ids=True)
        full, nbytes = gather_ids(ids, mine, len(lines), self.group)
        self.gathered_bytes += nbytes
        chars = self.characters
        return [''.join(chars[c] for c in v) for v in full], [None] * len(lines), [None] * len(lines)
