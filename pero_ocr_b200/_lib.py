"""ctypes binding of libb200_lineocr.so -- one prototype per entry point of include/b200_lineocr.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, E_INVALID, E_CUDA, E_WORKSPACE, E_NO_DEVICE = 0, 1, 2, 3, 4
CONV_FIRST, CONV, BILSTM, CTC_HEAD, UPSAMPLE, LN_PE, TRANSFORMER_LAYER = 1, 2, 3, 4, 5, 6, 7
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2
PREC_FP16, PREC_FP16X3, PREC_FP16F8, PREC_FP16F8W = 0, 1, 2, 3
PRECISIONS = {'fp16': PREC_FP16, 'fp16x3': PREC_FP16X3, 'fp16f8': PREC_FP16F8, 'fp16f8w': PREC_FP16F8W}
CORR_BOTH, CORR_WEIGHT, CORR_NONE = 0, 1, 2
DEFAULT_PRECISION = 'fp16f8'   # fp16 pass + e5m2 correction pass: logits within 1e-4 of the fp32 oracle at 2 pass-equivalents

_FP = C.POINTER(C.c_float)


class B200Error(RuntimeError):
    """Raised for every non-zero status of the native library (message = b200ocr_last_error)."""


class Layer(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('cin', C.c_int32), ('cout', C.c_int32),
        ('kh', C.c_int32), ('kw', C.c_int32), ('pad_h', C.c_int32), ('pad_w', C.c_int32),
        ('act', C.c_int32), ('pool_h', C.c_int32), ('pool_w', C.c_int32),
        ('weight', _FP), ('bias', _FP), ('post_scale', _FP), ('post_shift', _FP),
        ('hidden', C.c_int32),
        ('w_ih', _FP * 2), ('w_hh', _FP * 2), ('b_ih', _FP * 2), ('b_hh', _FP * 2),
        ('heads', C.c_int32), ('dim_ff', C.c_int32),
        ('in_proj_w', _FP), ('in_proj_b', _FP), ('out_proj_w', _FP), ('out_proj_b', _FP),
        ('lin1_w', _FP), ('lin1_b', _FP), ('lin2_w', _FP), ('lin2_b', _FP),
        ('norm1_w', _FP), ('norm1_b', _FP), ('norm2_w', _FP), ('norm2_b', _FP),
        ('act_slope', C.c_float),
    ]


class PolyLine(C.Structure):
    """b200ocr_poly_line_t (include/b200_lineocr.h)."""
    _fields_ = [('coef', C.c_double * 4), ('x_first', C.c_double), ('x_last', C.c_double), ('total', C.c_double),
                ('step', C.c_double), ('rot', C.c_double * 4), ('ncoef', C.c_int32), ('n_out', C.c_int32)]


class ArLayer(C.Structure):
    """b200ocr_ar_layer_t (include/b200_lineocr.h)."""
    _fields_ = [(k, _FP) for k in (
        'self_in_w', 'self_in_b', 'self_out_w', 'self_out_b', 'cross_in_w', 'cross_in_b', 'cross_out_w', 'cross_out_b',
        'lin1_w', 'lin1_b', 'lin2_w', 'lin2_b', 'norm1_w', 'norm1_b', 'norm2_w', 'norm2_b', 'norm3_w', 'norm3_b')]


class ArDesc(C.Structure):
    """b200ocr_ar_desc_t (include/b200_lineocr.h)."""
    _fields_ = [('n_layers', C.c_int32), ('heads', C.c_int32), ('dim_ff', C.c_int32), ('classes', C.c_int32),
                ('layers', C.POINTER(ArLayer)), ('embed', _FP), ('out_w', _FP), ('out_b', _FP)]


class NetDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('layers', C.POINTER(Layer)), ('precision', C.c_int32),
                ('line_height', C.c_int32), ('device', C.c_int32)]


EXPORTS = {
    # name: (restype, argtypes)
    'b200ocr_create': (C.c_int, [C.POINTER(NetDesc), C.POINTER(C.c_void_p)]),
    'b200ocr_reserve': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'b200ocr_reserve_maps': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'b200ocr_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b200ocr_forward_maps': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'b200ocr_destroy': (None, [C.c_void_p]),
    'b200ocr_last_error': (C.c_char_p, [C.c_void_p]),
    'b200ocr_launch_count': (C.c_int64, [C.c_void_p]),
    'b200ocr_forward_flops': (C.c_double, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    'b200ocr_set_layer_correction': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'b200ocr_executed_passes': (C.c_double, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'b200ocr_run_after': (C.c_int, [C.c_void_p, C.c_void_p]),
    'b200ocr_memcpy_d2h_async': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'b200ocr_set_layer_post_shift': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    'b200ocr_profile_read_since': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
    'b200ocr_ctc_greedy': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b200ocr_force_align': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    'b200ocr_char_confidence': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b200ocr_remap_lines': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    'b200ocr_pad_lines': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    'b200ocr_remap_poly_lines': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    'b200ocr_sparsify_logits': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p]),
    'b200ocr_ctc_prefix_beam': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
   'b200ocr_ctc_prefix_beam_ranges': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b200ocr_full_logprobs': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'b200ocr_ar_attach': (C.c_int, [C.c_void_p, C.POINTER(ArDesc)]),
    'b200ocr_ar_reserve': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    'b200ocr_ar_transcribe': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    'b200ocr_profile': (C.c_int, [C.c_void_p, C.c_int32]),
    'b200ocr_profile_read': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int32)]),
    'b200ocr_debug_set_flag': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'b200ocr_debug_use_reference_kernels': (C.c_int, [C.c_void_p, C.c_int32]),
    'b200ocr_debug_forward_prefix': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                               C.c_void_p]),
}


def library_path():
    # B200OCR_LIB: bring-up override used to A/B two builds of the library inside one GPU session
    return os.environ.get('B200OCR_LIB') or os.path.join(_HERE, 'libb200_lineocr.so')


def load_library():
    """Loads the native library or raises: the product path never degrades to a CPU implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise B200Error(f'{path} is missing: build it with `make -C pero_ocr_b200/csrc` '
                        f'(or __graft_entry__.build()); there is no CPU fallback')
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library ever disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(status, handle=None):
    if status != OK:
        msg = load_library().b200ocr_last_error(handle)
        raise B200Error(f'b200ocr status {status}: {msg.decode("utf8", "replace") if msg else "?"}')
