"""B200-native text-line recognition path behind pero-ocr's ocr_engine / decoding plugin surface.

Python host code (PyTorch tensors are only the buffer currency) drives hand-written sm_100a kernels through the
C ABI of ``libb200_lineocr.so`` (``include/b200_lineocr.h``).  There is no CPU fallback: importing the engine
classes works anywhere, *using* them without the built library or without a B200 raises.
"""
from ._lib import B200Error, library_path, load_library  # noqa: F401

__all__ = ['B200Error', 'library_path', 'load_library']
