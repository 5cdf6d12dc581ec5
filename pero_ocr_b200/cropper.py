"""Baseline-following line cropper with the pixel resampling on the GPU.

Host mirror of ``pero_ocr.core.crop_engine.EngineLineCropper`` (crop_engine.py:8-30): the geometry
(``get_crop_inputs``, crop_engine.py:54-99 -- a handful of NumPy / SciPy calls per line) stays on the host and is
restated here call for call, because the resampling map it produces is the *input* of the hot kernel; the resampling
itself (``fast_remap``, crop_engine.py:146-163: ``cv2.remap`` bilinear, constant border) runs on the device for all
lines of a page at once (``b200ocr_remap_lines``) and can write straight into the recogniser's padded batch
(``B200EngineLineOCR.process_line_maps``).
"""
import ctypes as C
import math

import numpy as np

from . import _lib


class DevicePage:
    """A page image resident on the device (uploaded once per page) plus pinned staging for the per-batch maps."""

    def __init__(self, image, device=None):
        import torch
        image = np.ascontiguousarray(image)
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            raise ValueError(f'page image must be [H, W, 3] uint8, got {image.shape} {image.dtype}')
        if image.shape[0] > 32767 or image.shape[1] > 32767:
            raise ValueError('page images beyond 32767 px per side are outside cv2.remap\'s own range')
        if not torch.cuda.is_available():
            raise _lib.B200Error('no CUDA device: the B200 line cropper has no CPU fallback')
        self.torch = torch
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.shape = image.shape
        self.image = torch.from_numpy(image).to(self.device)
        self.h2d_bytes = image.nbytes
        self._stage = [None, None]
        self._turn = 0

    def stage_maps(self, maps):
        """Concatenate float32 [H, w_i, 2] maps in pinned memory and start their upload on the current stream.
        -> (coords CUDA f32, offsets CUDA i64 [n], widths CUDA i32 [n], bytes uploaded)."""
        torch = self.torch
        sizes = [int(m.size) for m in maps]
        total = sum(sizes)
        k = self._turn
        self._turn ^= 1
        st = self._stage[k]
        if st is not None:
            st['free'].synchronize()                       # the previous upload from this buffer has been consumed
        if st is None or st['pin'].numel() < total:
            cap = max(total, 1 << 16)
            st = dict(pin=torch.empty(cap, dtype=torch.float32, pin_memory=True),
                      dev=torch.empty(cap, dtype=torch.float32, device=self.device), free=torch.cuda.Event())
            self._stage[k] = st
        host = st['pin'].numpy()
        offs = np.zeros(len(maps), dtype=np.int64)
        at = 0
        for i, m in enumerate(maps):
            offs[i] = at
            host[at:at + sizes[i]] = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)
            at += sizes[i]
        widths = np.array([m.shape[1] for m in maps], dtype=np.int32)
        coords = st['dev'][:max(total, 1)]
        coords[:total].copy_(st['pin'][:total], non_blocking=True)
        st['free'].record(torch.cuda.current_stream(self.device))
        meta_o = torch.from_numpy(offs).to(self.device)
        meta_w = torch.from_numpy(widths).to(self.device)
        return coords, meta_o, meta_w, total * 4 + offs.nbytes + widths.nbytes


def remap_into(page, maps, out, pad):
    """Resample `maps` (list of float32 [H, w_i, 2]) from `page` into the CUDA uint8 batch `out` [n, H, out_w, 3]:
    line i lands in columns [pad, pad + w_i) (cut at out_w), everything else is zero.  Stream-ordered on torch's
    current stream.  Returns the bytes uploaded."""
    torch = page.torch
    lib = _lib.load_library()
    n, line_h, out_w, ch = out.shape
    assert ch == 3 and out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and n == len(maps)
    if n == 0:
        return 0
    coords, offs, widths, nbytes = page.stage_maps(maps)
    stream = torch.cuda.current_stream(page.device).cuda_stream
    _lib.check(lib.b200ocr_remap_lines(page.image.data_ptr(), page.shape[0], page.shape[1], coords.data_ptr(),
                                       offs.data_ptr(), widths.data_ptr(), n, line_h, out.data_ptr(), out_w, pad,
                                       C.c_void_p(stream)))
    return nbytes


def remap_poly_into(page, params, offsets, out, pad):
    """Like remap_into, with the sampling maps evaluated on the device: `params` is a list of _lib.PolyLine,
    `offsets` a float64 array [n, line_h] (B200LineCropper.poly_params).  Returns the bytes uploaded."""
    torch = page.torch
    lib = _lib.load_library()
    n, line_h, out_w, ch = out.shape
    assert ch == 3 and out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and n == len(params)
    if n == 0:
        return 0
    arr = (_lib.PolyLine * n)(*params)
    raw = np.frombuffer(arr, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.float64)
    # line parameters and offsets go up stream-ordered from page-locked staging (torch's host allocator keeps the block
    # until the copy has run): a synchronous copy would make the caller wait for everything queued on the stream,
    # i.e. for the previous batch's forward
    head = (raw.nbytes + 255) // 256 * 256
    pin = torch.empty(head + off.nbytes, dtype=torch.uint8, pin_memory=True)
    pin_np = pin.numpy()
    pin_np[:raw.nbytes] = raw
    pin_np[head:].view(np.float64)[:] = off.reshape(-1)
    d_all = torch.empty(pin.numel(), dtype=torch.uint8, device=page.device)
    d_all.copy_(pin, non_blocking=True)
    d_par, d_off = d_all[:raw.nbytes], d_all[head:]
    stream = torch.cuda.current_stream(page.device).cuda_stream
    _lib.check(lib.b200ocr_remap_poly_lines(page.image.data_ptr(), page.shape[0], page.shape[1], d_par.data_ptr(),
                                            d_off.data_ptr(), n, line_h, out.data_ptr(), out_w, pad,
                                            C.c_void_p(stream)))
    page._keep = d_all                     # alive until the kernel has run (stream-ordered free on the next call)
    return raw.nbytes + off.nbytes


class B200LineCropper:
    """Drop-in for ``EngineLineCropper(correct_slant, line_height, poly, scale, blend_border)``.

    ``crop(img, baseline, heights)`` returns the same uint8 [line_height, w, 3] array as the reference (bit-exact:
    the device kernel restates OpenCV's 8-bit fixed-point bilinear remap).  ``crop_page`` resamples every line of a
    page with one launch.  The reverse mapping used only by ``blend_in`` rendering (crop_engine.py:32-52, 110-134) is
    outside the recognition path and not provided.
    """

    def __init__(self, correct_slant=False, line_height=32, poly=0, scale=1, blend_border=4, device=None):
        self.correct_slant = correct_slant
        self.line_height = line_height
        self.poly = poly
        self.scale = scale
        self.blend_border = blend_border
        self.device = device
        self._page_key = None
        self._page = None

    # ---- geometry (host) -----------------------------------------------------------------------------------
    def get_crop_inputs(self, baseline, line_heights, target_height):
        """Source coordinates of every pixel of the crop, float32 [target_height, w, 2] (crop_engine.py:54-99).
        Restated step by step with the reference's own NumPy / SciPy calls so that the map is bit-identical:
        integer-truncated baseline, rotation to the first->last direction, polynomial or cubic interpolant, arc-length
        resampling, unit normals by a 0.1 px forward difference, rotation back, cast to float32."""
        from scipy import interpolate
        above, below = line_heights[0] * self.scale, line_heights[1] * self.scale
        pts = np.asarray(baseline).copy().astype(int)
        angle = math.atan2(pts[-1, 1] - pts[0, 1], pts[-1, 0] - pts[0, 0])
        rot = np.array([[np.cos(angle), np.sin(angle)], [-np.sin(angle), np.cos(angle)]])
        pts = np.dot(pts, np.linalg.inv(rot))
        if self.poly:
            degree = self.poly if pts.shape[0] > 2 else 1
            curve = np.poly1d(np.polyfit(pts[:, 0], pts[:, 1], degree))
        else:
            try:
                pts[-1, 0] += 0.1          # keeps the forward difference below inside the interpolant's domain
                curve = interpolate.interp1d(pts[:, 0], pts[:, 1], kind='cubic')
            except Exception:              # too few points for a cubic: straight line (the shift above stays)
                curve = np.poly1d(np.polyfit(pts[:, 0], pts[:, 1], 1))
        xs = np.arange(pts[:, 0].min(), pts[:, 0].max())
        ys = curve(xs)
        seg = ((xs[:-1] - xs[1:]) ** 2 + (ys[:-1] - ys[1:]) ** 2) ** 0.5
        arc = np.concatenate([np.zeros(1), np.cumsum(seg)])
        zoom = target_height / (above + below)
        n_out = int(arc[-1] * zoom)
        samples = np.linspace(0, arc[-1], n_out)
        out_x = self._reverse_line_mapping(arc, samples, xs)
        out_y = curve(out_x)
        dx = np.full_like(out_x, 0.1)
        dy = out_y - curve(out_x + 0.1)
        length = (dx ** 2 + dy ** 2) ** 0.5
        nx = -dy / length
        ny = dx / length
        offsets = np.linspace(-above, below, target_height).reshape(-1, 1)
        map_x = nx.reshape(1, -1) * offsets + out_x.reshape(1, -1)
        map_y = ny.reshape(1, -1) * offsets + out_y.reshape(1, -1)
        return np.dot(np.stack((map_x, map_y), axis=2), rot).astype(np.float32)

    def poly_params(self, baseline, line_heights, target_height=None):
        """The host half of get_crop_inputs for `poly` > 0 (crop_engine.py:54-73): rotated baseline, polynomial fit,
        arc length, crop width.  -> (_lib.PolyLine, float64 offsets [target_height]); the device evaluates the rest
        (b200ocr_remap_poly_lines).  A geometry failure yields the reference's fallback, a 32 px all-zero crop.
        Bit-exactness of this path relies on the host evaluating np.dot(pts, inv(rot)) the way the reference's host
        does (a 2 x 2 product: fma(y, r1, x * r0) in NumPy's BLAS-free small-matrix path, pinned by
        tests/golden/cropper.npz in this container); the rotation of the sampled points themselves is restated on the
        device with explicit rounding (csrc/remap.cu)."""
        if not self.poly:
            raise ValueError('poly_params needs a polynomial baseline fit (poly > 0); use get_crop_inputs otherwise')
        if self.poly > 3:
            # b200ocr_poly_line_t carries four coefficients; the reference (np.polyfit of any degree) would crop such
            # lines, so refuse loudly instead of returning its zero-crop fallback: get_crop_inputs + process_line_maps
            # take any degree
            raise ValueError(f'the device-side baseline evaluation supports poly <= 3 (got {self.poly}); '
                             f'use get_crop_inputs / process_line_maps for higher degrees')
        target_height = target_height or self.line_height
        line = _lib.PolyLine()
        try:
            above, below = line_heights[0] * self.scale, line_heights[1] * self.scale
            pts = np.asarray(baseline).copy().astype(int)
            angle = math.atan2(pts[-1, 1] - pts[0, 1], pts[-1, 0] - pts[0, 0])
            rot = np.array([[np.cos(angle), np.sin(angle)], [-np.sin(angle), np.cos(angle)]])
            pts = np.dot(pts, np.linalg.inv(rot))
            degree = self.poly if pts.shape[0] > 2 else 1
            coef = np.polyfit(pts[:, 0], pts[:, 1], degree)
            xs = np.arange(pts[:, 0].min(), pts[:, 0].max())
            # np.poly1d(coef)(xs) is np.polyval on the coefficients with leading zeros trimmed
            ys = np.polyval(coef, xs) if coef[0] != 0 else np.poly1d(coef)(xs)
            seg = ((xs[:-1] - xs[1:]) ** 2 + (ys[:-1] - ys[1:]) ** 2) ** 0.5
            total = np.cumsum(seg)[-1] if seg.shape[0] else np.zeros(1)[-1]   # last of concatenate([0], cumsum(seg))
            n_out = int(total * (target_height / (above + below)))
            if n_out < 1 or len(coef) > 4 or not np.isfinite(total):
                raise ValueError('empty crop')
            offsets = np.linspace(-above, below, target_height)
            for i, v in enumerate(coef):
                line.coef[i] = float(v)
            line.ncoef, line.n_out = len(coef), n_out
            line.x_first, line.x_last, line.total = float(xs[0]), float(xs[-1]), float(total)
            line.step = float(total / (n_out - 1)) if n_out > 1 else 0.0
            line.rot[0], line.rot[1], line.rot[2], line.rot[3] = (float(rot[0, 0]), float(rot[0, 1]), float(rot[1, 0]),
                                                                  float(rot[1, 1]))
            return line, offsets
        except Exception:
            print('ERROR: line crop failed.', line_heights, baseline)
            line = _lib.PolyLine()
            line.ncoef, line.n_out = 0, 32
            return line, np.zeros(target_height)

    @staticmethod
    def _reverse_line_mapping(forward_mapping, sample_positions, sampled_values):
        """crop_engine.py:101-110, quirk preserved: the reference's search loop advances only while the cumulative
        length EXCEEDS the sample, which never happens from position 0 (forward_mapping[0] == 0 <= sample), so every
        sample is interpolated between index -1 (the last point) and index 0: the result is the straight blend
        (1 - da) * values[-1] + da * values[0] with da = (sample - total) / (0 - total)."""
        if sample_positions.shape[0] and (forward_mapping[0] > sample_positions).any():
            # not reachable with the linspace(0, total, n) samples of get_crop_inputs; keep the reference's loop
            out = np.zeros_like(sample_positions)
            pos = 0
            for i in range(sample_positions.shape[0]):
                while forward_mapping[pos] > sample_positions[i]:
                    pos += 1
                d = forward_mapping[pos] - forward_mapping[pos - 1]
                da = (sample_positions[i] - forward_mapping[pos - 1]) / d
                out[i] = (1 - da) * sampled_values[pos - 1] + da * sampled_values[pos]
            return out
        d = forward_mapping[0] - forward_mapping[-1]
        da = (sample_positions - forward_mapping[-1]) / d
        return (1 - da) * sampled_values[-1] + da * sampled_values[0]

    # ---- resampling (device) -------------------------------------------------------------------------------
    def page(self, img):
        """DevicePage of `img`, cached for consecutive calls on the same array (LineCropper.process_page crops every
        line of one image, page_parser.py:384-393)."""
        key = (id(img), img.shape, img.__array_interface__['data'][0] if isinstance(img, np.ndarray) else None)
        if self._page is None or self._page_key != key:
            self._page = DevicePage(img, self.device)
            self._page_key = key
        return self._page

    def crop_page_poly(self, img, lines):
        """crop_page with the maps evaluated on the device (poly > 0 only): one launch, ~200 bytes uploaded per line."""
        page = img if isinstance(img, DevicePage) else self.page(img)
        params, offs = zip(*[self.poly_params(b, h) for b, h in lines]) if lines else ((), ())
        torch = page.torch
        width = max([p.n_out for p in params] + [1])
        out = torch.empty((len(params), self.line_height, width, 3), dtype=torch.uint8, device=page.device)
        remap_poly_into(page, list(params), np.stack(offs) if offs else np.zeros((0, self.line_height)), out, 0)
        host = out.cpu().numpy()
        return [host[i, :, :p.n_out].copy() for i, p in enumerate(params)]

    def crop_page(self, img, lines):
        """`lines`: iterable of (baseline, heights).  -> list of uint8 [line_height, w, 3] crops (the reference's
        zero [line_height, 32, 3] crop where its geometry fails, crop_engine.py:16-22), one launch for the page."""
        page = img if isinstance(img, DevicePage) else self.page(img)
        maps, failed = [], []
        for baseline, heights in lines:
            try:
                m = self.get_crop_inputs(baseline, heights, self.line_height)
                if m.shape[1] == 0:
                    raise ValueError('empty crop')        # cv2.remap raises on an empty map -> the reference's except
                failed.append(False)
            except Exception:
                print('ERROR: line crop failed.', heights, baseline)
                m = np.zeros((self.line_height, 0, 2), dtype=np.float32)
                failed.append(True)
            maps.append(m)
        torch = page.torch
        width = max([m.shape[1] for m in maps] + [1])
        out = torch.empty((len(maps), self.line_height, width, 3), dtype=torch.uint8, device=page.device)
        remap_into(page, maps, out, 0)
        host = out.cpu().numpy()
        crops = []
        for i, m in enumerate(maps):
            if failed[i]:
                crops.append(np.zeros([self.line_height, 32, page.shape[2]], dtype=np.uint8))
            else:
                crops.append(host[i, :, :m.shape[1]].copy())
        return crops

    def crop(self, img, baseline, heights, return_mapping=False, return_forward_mapping=False):
        """crop_engine.py:16-30 (one line; prefer crop_page / process_line_maps for whole pages)."""
        if return_mapping:
            raise NotImplementedError('the reverse mapping of blend_in rendering is outside the recognition path')
        crop = self.crop_page(img, [(baseline, heights)])[0]
        if return_forward_mapping:
            return crop, self.get_crop_inputs(baseline, heights, self.line_height)
        return crop
