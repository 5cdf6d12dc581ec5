"""Device-side logit sparsification: the host mirror of ``b200ocr_sparsify_logits``.

Replaces the per-line NumPy pass at the end of ``BaseEngineLineOCR.process_lines``
(pero_ocr/ocr_engine/line_ocr_engine.py:152-156 tight crop, :168-172 softmax threshold 1e-4 + ``scipy.sparse.csc_matrix``)
so that ``TextLine.logits`` (pero_ocr/core/layout.py:41-72) is built from a few surviving entries per frame instead of
a dense [T, C] matrix copied over PCIe and soft-maxed on one host core per line.
"""
import ctypes as C

import numpy as np

from . import _lib


class SparseLogits:
    """CSC parts of a batch of lines, still on the device.  ``fetch()`` copies exactly the used prefix to the host."""

    def __init__(self, torch, n, t, c):
        self.torch, self.n, self.t, self.c = torch, n, t, c
        self.indptr = self.nnz = self.base = self.indices = self.data = None
        self.rows = None          # host int array [n]: rows (frames) of every line's matrix
        self._host = {}

    def prefetch_meta(self, pinned):
        """Starts the copy of the small parts (indptr, base) into the caller's pinned buffers `pinned` (a dict that
        lives as long as the caller's slot) on the current stream: once that stream's work is known to be complete
        the total entry count is on the host without another round trip."""
        torch = self.torch
        for name in ('indptr', 'base'):
            src = getattr(self, name)
            dst = pinned.get(name)
            if dst is None or dst.shape != src.shape:
                dst = pinned[name] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            dst.copy_(src, non_blocking=True)
        self._meta = pinned

    def fetch(self, stream=None, pinned=None, pool=None, blocks=None):
        """-> (indptr int32 [n, C+1], base int64 [n+1], indices int32 [total], data float32 [total]) as NumPy arrays.
        The small parts first (indptr, base), then the used prefix of indices / data.  With `pinned` (the dict given
        to prefetch_meta, whose copies must have completed) the small parts are already on the host and the big ones
        are copied straight into arrays the caller may keep (returned wrapped in _Owned for csc_lines): page-locked
        blocks of `blocks` (a PinnedPool) while it has room, fresh pageable arrays otherwise."""
        torch = self.torch
        stream = stream or torch.cuda.current_stream(self.indptr.device)
        if pinned is not None and getattr(self, '_meta', None) is pinned:
            # the entry count is already on the host: the big parts go device -> their final host arrays in ONE pass
            # -- no staging copy that the host would have to copy again
            indptr, base = pinned['indptr'].numpy(), pinned['base'].numpy()
            total = int(base[self.n])
            if blocks is not None and 4 * total >= PinnedPool.MIN_BYTES:
                # recycled page-locked destinations: the copy is one DMA at PCIe rate and touches no fresh page
                bi = blocks.take(4 * total)
                bd = blocks.take(4 * total) if bi is not None else None
                if bd is not None:
                    indices = np.frombuffer(bi, dtype=np.int32, count=total)
                    data = np.frombuffer(bd, dtype=np.float32, count=total)
                    # straight cudaMemcpyAsync into the registered block: no torch tensor over the caller's arrays (torch
                    # keeps such a tensor alive past the copy when it does not recognise the memory as page-locked)
                    lib = _lib.load_library()
                    for dst, src in ((indices, self.indices), (data, self.data)):
                        _lib.check(lib.b200ocr_memcpy_d2h_async(C.c_void_p(dst.ctypes.data), C.c_void_p(src.data_ptr()),
                                                                4 * total, C.c_void_p(stream.cuda_stream)))
                    stream.synchronize()
                    return indptr, base, _Owned(indices), _Owned(data)
                del bi
            indices, data = fresh_host_array(total, np.int32), fresh_host_array(total, np.float32)
            if total:
                def pull(dst, src, st):
                    with torch.cuda.device(src.device), torch.cuda.stream(st):
                        torch.from_numpy(dst).copy_(src[:total])

                if pool is not None and total > (1 << 20):
                    # a pageable D2H copy is bound by the host side of the driver's staging (memcpy + first-touch
                    # faults of the fresh destination): the two arrays go in parallel, on their own streams
                    if getattr(self, '_stream2', None) is None:
                        self._stream2 = torch.cuda.Stream(self.indices.device)
                    other = pool.submit(pull, indices, self.indices, self._stream2)
                    pull(data, self.data, stream)
                    other.result()
                else:
                    pull(indices, self.indices, stream)
                    pull(data, self.data, stream)
            return indptr, base, _Owned(indices), _Owned(data)
        with torch.cuda.stream(stream):
            indptr = self.indptr.cpu()
            base = self.base.cpu()
            total = int(base[self.n])
            indices = self.indices[:total].cpu()
            data = self.data[:total].cpu()
        return indptr.numpy(), base.numpy(), indices.numpy(), data.numpy()


def sparsify_device(logits, t_lo=None, t_hi=None, out=None):
    """logits: CUDA float32 [N, T, C] (contiguous).  t_lo / t_hi: optional int arrays [N] (host) giving the frame range
    kept per line (the reference's tight crop).  Stream-ordered on torch's current stream; returns SparseLogits."""
    import torch
    lib = _lib.load_library()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 3
    n, t, c = logits.shape
    dev = logits.device
    sp = out if isinstance(out, SparseLogits) and (out.n, out.t, out.c) == (n, t, c) else SparseLogits(torch, n, t, c)
    if sp.indptr is None:
        sp.indptr = torch.empty((n, c + 1), dtype=torch.int32, device=dev)
        sp.nnz = torch.empty((n,), dtype=torch.int32, device=dev)
        sp.base = torch.empty((n + 1,), dtype=torch.int64, device=dev)
        sp.indices = torch.empty((n * t * c,), dtype=torch.int32, device=dev)
        sp.data = torch.empty((n * t * c,), dtype=torch.float32, device=dev)
    lo_d = hi_d = None
    if t_lo is not None:
        lo = np.ascontiguousarray(t_lo, dtype=np.int32)
        hi = np.ascontiguousarray(t_hi, dtype=np.int32)
        # frame ranges through this object's own pinned + device pair, stream-ordered: a synchronous copy here would
        # make the host wait for everything queued on the stream (the previous batch's forward).  A reused object is
        # only handed in again after its previous results were collected, so the pair is free
        if getattr(sp, '_range_pin', None) is None:
            sp._range_pin = torch.empty((2, n), dtype=torch.int32, pin_memory=True)
            sp._range_dev = torch.empty((2, n), dtype=torch.int32, device=dev)
        sp._range_pin[0].numpy()[:] = lo
        sp._range_pin[1].numpy()[:] = hi
        sp._range_dev.copy_(sp._range_pin, non_blocking=True)
        lo_d, hi_d = sp._range_dev[0], sp._range_dev[1]
        sp.rows = np.maximum(np.minimum(hi, t) - np.clip(lo, 0, t), 0)
    else:
        sp.rows = np.full((n,), t, dtype=np.int64)
    sp._keep = (lo_d, hi_d)      # alive until the kernels have run
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.b200ocr_sparsify_logits(
        logits.data_ptr(), n, t, c,
        lo_d.data_ptr() if lo_d is not None else None, hi_d.data_ptr() if hi_d is not None else None,
        sp.indptr.data_ptr(), sp.nnz.data_ptr(), sp.base.data_ptr(), sp.indices.data_ptr(), sp.data.data_ptr(),
        n * t * c, C.c_void_p(stream)))
    return sp


class _Owned:
    """Marks an array fetch() allocated for the caller: csc_lines may slice it without another copy."""

    def __init__(self, array):
        self.array = array


class _Block:
    """One page-locked block on loan from a PinnedPool.  It exports the buffer (PEP 688), so every array made over
    it with np.frombuffer keeps it alive; when the last one is gone the memory goes back to the pool."""

    def __init__(self, pool, mem, nbytes):
        self._pool, self._mem, self.nbytes = pool, mem, nbytes

    def __buffer__(self, flags):
        return memoryview(self._mem)

    def __del__(self):
        pool, mem = self._pool, self._mem
        self._pool = self._mem = None
        if pool is not None:
            pool._give_back(mem, self.nbytes)


class PinnedPool:
    """Page-locked host memory for results the CALLER keeps (the CSC parts of ``TextLine.logits``): anonymous huge-page
    mappings registered with the driver once (cudaHostRegister) and recycled when the arrays made over them are
    garbage-collected.  A device->host copy into such a block runs at PCIe rate; into fresh pageable memory it is
    bound by the driver's staging memcpy and first-touch page faults (~4 GB/s measured).  At most `cap_bytes` are ever
    registered: beyond that take() returns None and the caller falls back to pageable arrays."""
    MIN_BYTES = 2 << 20

    def __init__(self, cap_bytes=4 << 30, register=None, unregister=None):
        import threading
        self.cap = int(cap_bytes)
        self.registered = 0
        self.free = {}                      # block size -> [mmap, ...]
        self.lock = threading.Lock()
        self.stats = {'hits': 0, 'new': 0, 'refused': 0}
        self.closed = False
        self._register, self._unregister = register or _cuda_host_register, unregister or _cuda_host_unregister

    @staticmethod
    def block_size(nbytes):
        """Eight size classes per octave (<= 12.5 % slack), 2 MiB granules."""
        n = max(int(nbytes), PinnedPool.MIN_BYTES)
        step = max(1 << 21, (1 << (n.bit_length() - 1)) >> 3)
        return (n + step - 1) // step * step

    def take(self, nbytes):
        import mmap
        size = self.block_size(nbytes)
        with self.lock:
            if self.closed:
                return None
            have = self.free.get(size)
            if have:
                self.stats['hits'] += 1
                return _Block(self, have.pop(), size)
            if self.registered + size > self.cap:
                self.stats['refused'] += 1
                return None
            self.registered += size
        try:
            mem = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            if hasattr(mmap, 'MADV_HUGEPAGE'):
                try:
                    mem.madvise(mmap.MADV_HUGEPAGE)
                except OSError:
                    pass
            self._register(mem, size)
        except Exception:
            with self.lock:
                self.registered -= size
                self.stats['refused'] += 1
            return None
        self.stats['new'] += 1
        return _Block(self, mem, size)

    def _give_back(self, mem, size):
        with self.lock:
            if not self.closed:
                self.free.setdefault(size, []).append(mem)
                return
            self.registered -= size
        try:                                 # the pool is gone: the block goes back to the system right away
            self._unregister(mem)
            mem.close()
        except Exception:
            pass

    def close(self):
        """No further loans; free blocks are unregistered now, blocks on loan when their last array dies."""
        self.closed = True
        self.trim()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """Unregisters and unmaps every block that is not on loan."""
        with self.lock:
            blocks, self.free = self.free, {}
        for size, mems in blocks.items():
            for mem in mems:
                try:
                    self._unregister(mem)
                    mem.close()
                except Exception:
                    pass
                with self.lock:
                    self.registered -= size


def _address(mem):
    return C.addressof(C.c_char.from_buffer(mem))


def _cuda_host_register(mem, size):
    import torch
    err = torch.cuda.cudart().cudaHostRegister(_address(mem), size, 0)
    if int(err) != 0:
        raise RuntimeError(f'cudaHostRegister failed: {err}')


def _cuda_host_unregister(mem):
    import torch
    torch.cuda.cudart().cudaHostUnregister(_address(mem))


def fresh_host_array(count, dtype):
    """Pageable host array of `count` elements that the caller will fill and keep (a batch's CSC parts: tens of MB of
    FRESH memory per batch -- first-touch page faults, not the copy, dominate at 4 KB pages).  Anonymous mmap advised
    to transparent huge pages where the host allows it (THP "madvise" or "always"); a plain np.empty otherwise."""
    import mmap
    nbytes = int(count) * np.dtype(dtype).itemsize
    if nbytes < (4 << 20) or not hasattr(mmap, 'MADV_HUGEPAGE'):
        return np.empty(int(count), dtype=dtype)
    try:
        m = mmap.mmap(-1, (nbytes + (1 << 21) - 1) >> 21 << 21, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        m.madvise(mmap.MADV_HUGEPAGE)
        return np.frombuffer(m, dtype=dtype, count=int(count))
    except (OSError, ValueError, AttributeError):
        return np.empty(int(count), dtype=dtype)


def parallel_copy(dst, src, pool=None, workers=1):
    """dst[:] = src in `workers` contiguous pieces on `pool` (NumPy releases the GIL inside large copies; the page
    faults of a fresh destination are taken by the copying threads in parallel)."""
    n = len(src)
    if pool is None or workers <= 1 or n < (1 << 20):
        np.copyto(dst, src)
        return
    step = (n + workers - 1) // workers
    for j in [pool.submit(np.copyto, dst[lo:lo + step], src[lo:lo + step]) for lo in range(0, n, step)]:
        j.result()


def csc_lines(sp, fetched=None, pool=None, workers=1):
    """list of scipy.sparse.csc_matrix [rows_i, C] float32 -- the value ``process_lines`` stores in ``TextLine.logits``.
    The batch's entries are copied to fresh host memory ONCE (optionally by `workers` threads of `pool`); each line's
    matrix is built on slices of that copy."""
    from scipy import sparse
    indptr, base, indices, data = fetched if fetched is not None else sp.fetch()
    total = int(base[sp.n])
    indptr = np.array(indptr)
    if isinstance(indices, _Owned):
        own_i, own_d = indices.array, data.array
    else:
        own_i, own_d = fresh_host_array(total, np.int32), fresh_host_array(total, np.float32)
        parallel_copy(own_i, indices[:total], pool, workers)
        parallel_copy(own_d, data[:total], pool, workers)
    # scipy copies an operand that is a view of a much larger ndarray (check_format -> prune -> _prune_array looks at
    # `.base.size`): each line's arrays are therefore created over the batch's BUFFER (their base is the buffer
    # object, not an ndarray), which keeps the single copy single
    buf_i, buf_d = _buffer_of(own_i), _buffer_of(own_d)
    out = []
    for i in range(sp.n):
        b0, b1 = int(base[i]), int(base[i + 1])
        if b1 > b0:
            di = np.frombuffer(buf_i, dtype=np.int32, count=b1 - b0, offset=4 * b0)
            dd = np.frombuffer(buf_d, dtype=np.float32, count=b1 - b0, offset=4 * b0)
        else:
            di, dd = np.empty(0, dtype=np.int32), np.empty(0, dtype=np.float32)
        out.append(sparse.csc_matrix((dd, di, indptr[i]), shape=(int(sp.rows[i]), sp.c), copy=False))
    return out


def _buffer_of(arr):
    import mmap
    return arr.base if isinstance(arr.base, (mmap.mmap, _Block)) else memoryview(arr)
