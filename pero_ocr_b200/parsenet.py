"""ParseNet page-detector inference on the B200 conv kernels.

Replaces ``TorchParseNet`` (pero_ocr/layout_engines/torch_parsenet.py:21-104): same constructor arguments, same
``get_maps`` / ``get_maps_with_optimal_resolution`` / ``get_med_height`` results.  The INTER_AREA downscale and the
x64 zero-padded canvas stay on the host exactly as in the reference (:42-47); ``self.net(canvas)`` (:51-53) is
b200ocr_forward_maps.  Everything downstream (CPU geometry in cnn_layout_engine.py) is untouched.
"""
import ctypes as C

import numpy as np

from . import _lib, netdesc
from ._lib import DEFAULT_PRECISION


class B200ParseNet:
    def __init__(self, model_path, device=None, downsample=4, max_mp=5, detection_threshold=0.2,
                 adaptive_downsample=True, precision=DEFAULT_PRECISION, module=None):
        import torch
        import cv2  # noqa: F401  (host-side resize, as in the reference)
        if not torch.cuda.is_available():
            raise _lib.B200Error('no CUDA device: the B200 ParseNet path has no CPU fallback')
        self.torch = torch
        self.device = device if device is not None else torch.device('cuda', 0)
        self.max_megapixels = max_mp if max_mp is not None else 5
        if module is None:
            module = torch.jit.load(model_path, map_location='cpu')
        self._lib = _lib.load_library()
        layers = netdesc.describe_parsenet(module)
        self.out_channels = layers[-1]['cout']
        desc, keep = netdesc.to_ctypes(layers, precision, 64, self.device.index or 0)
        h = C.c_void_p()
        torch.cuda.set_device(self.device)
        _lib.check(self._lib.b200ocr_create(C.byref(desc), C.byref(h)))
        self._h = h
        self._reserved = (0, 0)
        # one native engine = one workspace: forwards are enqueued one after the other on the engine's own stream (the
        # lock covers the enqueue only, callers on several threads wait for their own result outside it)
        import threading
        self._lock = threading.Lock()
        self._stream = torch.cuda.Stream(self.device)
        self.detection_threshold = detection_threshold
        self.adaptive_downsample = adaptive_downsample
        self.init_downsample = downsample
        self.last_downsample = downsample
        self.downsample_line_pixel_adapt_threshold = 100
        self.min_line_processing_height = 9
        self.max_line_processing_height = 15
        self.optimal_line_processing_height = 12
        self.min_downsample = 1
        self.max_downsample = 8

    def close(self):
        if getattr(self, '_h', None):
            self._lib.b200ocr_destroy(self._h)
            self._h = None

    __del__ = close

    def net(self, canvas_u8):
        """uint8 [1,H64,W64,3] host array -> float32 [1,C,H64,W64] CUDA tensor."""
        torch = self.torch
        _, hh, ww, _ = canvas_u8.shape
        if hh > self._reserved[0] or ww > self._reserved[1]:
            r = (max(hh, self._reserved[0]), max(ww, self._reserved[1]))
            _lib.check(self._lib.b200ocr_reserve_maps(self._h, r[0], r[1]), self._h)
            self._reserved = r
        dev = torch.from_numpy(np.ascontiguousarray(canvas_u8)).to(self.device)
        maps = torch.empty((1, self.out_channels, hh, ww), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.b200ocr_forward_maps(self._h, dev.data_ptr(), hh, ww, maps.data_ptr(), C.c_void_p(stream)),
                   self._h)
        return maps

    def _forward_to_host(self, canvas):
        """uint8 canvas -> float32 [1, H64, W64, C] host array (page-locked; a view of it is what get_maps returns)."""
        if getattr(self, '_stream', None) is None:          # host-logic tests replace `net` with a CPU callable
            return self.net(canvas).permute(0, 2, 3, 1).cpu().numpy()
        torch = self.torch
        with self._lock, torch.cuda.device(self.device), torch.cuda.stream(self._stream):
            maps = self.net(canvas).permute(0, 2, 3, 1).contiguous()          # [1, H64, W64, C] on the device
            host = torch.empty(maps.shape, dtype=maps.dtype, pin_memory=True)
            host.copy_(maps, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        done.synchronize()
        return host.numpy()

    def get_maps(self, img, downsample):
        import cv2
        img = cv2.resize(img, (0, 0), fx=1 / downsample, fy=1 / downsample, interpolation=cv2.INTER_AREA)
        rows = int(np.ceil(img.shape[0] / 64) * 64)
        cols = int(np.ceil(img.shape[1] / 64) * 64)
        canvas = np.zeros((1, rows, cols, 3), dtype=np.uint8)
        canvas[0, :img.shape[0], :img.shape[1], :] = img
        out_map = self._forward_to_host(canvas)
        return out_map[0, :img.shape[0], :img.shape[1], :]

    def get_med_height(self, out_map):
        heights = (out_map[:, :, 2] > self.detection_threshold).astype(float) * out_map[:, :, 0]
        return np.median(heights[heights > 0])

    def get_maps_with_optimal_resolution(self, img):
        budget = np.sqrt((img.shape[0] * img.shape[1]) / (self.max_megapixels * 10e5))
        first = max(self.last_downsample, budget)
        used = first
        out_map = self.get_maps(img, used)
        if not self.adaptive_downsample:
            return out_map, used
        if (out_map[:, :, 2] > self.detection_threshold).sum() > self.downsample_line_pixel_adapt_threshold:
            med = self.get_med_height(out_map)
            if med > self.max_line_processing_height or med < self.min_line_processing_height:
                second = first * (med / self.optimal_line_processing_height)
                second = max(min(second, self.max_downsample), self.min_downsample)
                self.last_downsample = second
                second = max(self.last_downsample, budget)
                if second / first < 0.8 or second / first > 1.2:
                    used = second
                    out_map = self.get_maps(img, used)
        return out_map, used
