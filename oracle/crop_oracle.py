"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pixel resampling of pero-ocr's line cropper.

``EngineLineCropper.fast_remap`` (pero_ocr/core/crop_engine.py:146-163) calls ``cv2.remap(img, map_x, map_y,
interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)``.  The arithmetic lives in a third-party dependency
that is absent from /root/reference: opencv-python (pyproject.toml dependency ``opencv-python``, un-pinned; 4.13.0 in
this container).  Its published algorithm for 8-bit images (modules/imgproc/src/imgwarp.cpp: remapBilinear with
FixedPtCast<int, uchar, INTER_REMAP_COEF_BITS>, INTER_BITS = 5, INTER_REMAP_COEF_BITS = 15) is restated below in NumPy.

Pinned: tests/golden/cropper.npz holds crops and coordinate maps produced by the UNMODIFIED reference class in this
container (oracle/make_golden.py: golden_cropper); tests/test_oracle_cropper.py checks this restatement against them
bit for bit, and against cv2.remap itself on random maps when cv2 is importable.
"""
import numpy as np


def remap_bilinear_u8(img: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """img uint8 [H, W, C]; coords float32 [h, w, 2] = (x, y) source position of every output pixel.
    -> uint8 [h, w, C], bit-identical to cv2.remap(..., INTER_LINEAR, BORDER_CONSTANT, borderValue=0)."""
    h, w, _ = img.shape
    x = coords[..., 0].astype(np.float32) * np.float32(32)
    y = coords[..., 1].astype(np.float32) * np.float32(32)

    def cv_round(v):                                   # cvtss2si: half to even; NaN / overflow -> INT_MIN
        bad = ~(np.abs(v) < 2147483648.0)
        r = np.rint(np.where(bad, 0, v)).astype(np.int64)
        r[bad] = -2147483648
        return r

    sx, sy = cv_round(x), cv_round(y)
    ix = np.clip(sx >> 5, -32768, 32767)               # saturate_cast<short>
    iy = np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31

    def px(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
        v[~ok] = 0                                     # BORDER_CONSTANT, borderValue 0
        return v

    w00 = ((32 - fx) * (32 - fy) * 32)[..., None]
    w01 = (fx * (32 - fy) * 32)[..., None]
    w10 = ((32 - fx) * fy * 32)[..., None]
    w11 = (fx * fy * 32)[..., None]
    acc = px(iy, ix) * w00 + px(iy, ix + 1) * w01 + px(iy + 1, ix) * w10 + px(iy + 1, ix + 1) * w11
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


# ---- seeded cases shared by oracle/make_golden.py and the tests ---------------------------------------------------
def page_image(seed=11, h=300, w=420):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


# (name, cropper kwargs, baseline, heights)
CROP_CASES = [
    ('poly2_curved', dict(line_height=40, poly=2, scale=1), [[20, 130], [150, 138], [390, 124]], [30, 10]),
    ('cubic_4pts', dict(line_height=40, poly=0, scale=1), [[30, 60], [100, 67], [210, 58], [380, 64]], [26, 9]),
    ('two_points_fallback', dict(line_height=40, poly=0, scale=1), [[50, 230], [300, 218]], [22, 8]),
    ('border_left_top', dict(line_height=40, poly=2, scale=1), [[-15, 12], [120, 6], [260, 18]], [30, 10]),
    ('border_right', dict(line_height=40, poly=1, scale=1), [[250, 170], [440, 176]], [28, 12]),
    ('slanted_scaled', dict(line_height=40, poly=2, scale=1.2), [[40, 280], [160, 258], [300, 226]], [24, 9]),
    ('tall_line_height48', dict(line_height=48, poly=2, scale=1), [[25, 100], [190, 103], [400, 97]], [35, 14]),
    ('degenerate_single_point', dict(line_height=40, poly=2, scale=1), [[150, 150], [150, 150]], [20, 8]),
]


def poly_map_columns(line, offsets, columns):
    """What b200ocr_remap_poly_lines evaluates per output pixel (pero_ocr_b200/csrc/remap.cu: remap_poly_lines_kernel),
    restated operation by operation in IEEE float64 (Python floats; the fused multiply-add of the final rotation by
    exact rational arithmetic) for the given output columns.  `line` is the _lib.PolyLine produced by
    B200LineCropper.poly_params.  -> float32 [len(offsets), len(columns), 2]; must equal the reference's
    get_crop_inputs map (crop_engine.py:74-99) bit for bit."""
    import math
    from fractions import Fraction

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    def poly(x):
        y = 0.0
        for i in range(line.ncoef):
            y = y * x + line.coef[i]
        return y

    out = np.zeros((len(offsets), len(columns), 2), dtype=np.float32)
    for k, cx in enumerate(columns):
        smp = 0.0 if line.n_out == 1 else (line.total if cx == line.n_out - 1 else cx * line.step)
        da = (smp - line.total) / (0.0 - line.total)
        ox = (1.0 - da) * line.x_last + da * line.x_first
        oy = poly(ox)
        dy = oy - poly(ox + 0.1)
        length = math.sqrt(0.1 * 0.1 + dy * dy)
        nx, ny = -dy / length, 0.1 / length
        for i, off in enumerate(offsets):
            mx, my = nx * off + ox, ny * off + oy
            out[i, k, 0] = np.float32(fma(my, line.rot[2], mx * line.rot[0]))
            out[i, k, 1] = np.float32(fma(my, line.rot[3], mx * line.rot[1]))
    return out
