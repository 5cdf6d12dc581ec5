// Baseline-following line crops on the device.
//
// Replaces EngineLineCropper.fast_remap (pero_ocr/core/crop_engine.py:146-163), i.e. cv2.remap(img, map_x, map_y,
// INTER_LINEAR, BORDER_CONSTANT) per text line on one host core, and the zero padding + stacking of the crops into
// the recogniser's batch (BaseEngineLineOCR.process_lines, pero_ocr/ocr_engine/line_ocr_engine.py:121-123): the page
// image is uploaded once, every line of the page is resampled by one launch and lands directly in the padded
// [n][line_h][out_w][3] batch that b200ocr_forward reads.
//
// The arithmetic is OpenCV's 8-bit bilinear remap (opencv-python, un-pinned dependency of the reference; pinned here
// against 4.13.0 outputs, tests/golden/cropper.npz), restated:
//   sx = cvRound(x * 32), sy = cvRound(y * 32)            (INTER_BITS = 5; round half to even)
//   ix = saturate<short>(sx >> 5), fx = sx & 31            (same for y)
//   w00 = (32-fx)(32-fy)*32, w01 = fx(32-fy)*32, w10 = (32-fx)fy*32, w11 = fx*fy*32     (sum = 2^15)
//   dst = (p00*w00 + p01*w01 + p10*w10 + p11*w11 + 2^14) >> 15, a neighbour outside the image counts as 0.
// (OpenCV's table holds 32767 / 1 instead of 32768 / 0 for fx = fy = 0; the result is the same byte for every input.)
// Both branches of fast_remap -- whole image with a constant border, or the bounding-box crop with shifted
// coordinates -- produce exactly these bytes: subtracting the integer box origin from a float32 coordinate is exact.
#include "kernels.cuh"

namespace {

__device__ __forceinline__ int cv_round_x32(float v) {
    // cvRound(v * 32): SSE cvtss2si semantics -- NaN and out-of-range give INT_MIN ("integer indefinite")
    const float s = v * 32.0f;
    if (!(fabsf(s) < 2147483648.0f)) return INT_MIN;
    return __float2int_rn(s);
}

// One output pixel of cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) at source position (x, y).
__device__ __forceinline__ void remap_pixel(const uint8_t* __restrict__ img, int img_h, int img_w, float x, float y,
                                            uint8_t* __restrict__ dst) {
    const int sx = cv_round_x32(x), sy = cv_round_x32(y);
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
    const int fx = sx & 31, fy = sy & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const bool x0 = ix >= 0 && ix < img_w, x1 = ix + 1 >= 0 && ix + 1 < img_w;
    const bool y0 = iy >= 0 && iy < img_h, y1 = iy + 1 >= 0 && iy + 1 < img_h;
    const uint8_t* r0 = img + (static_cast<size_t>(y0 ? iy : 0) * img_w) * 3;
    const uint8_t* r1 = img + (static_cast<size_t>(y1 ? iy + 1 : 0) * img_w) * 3;
    const int o0 = (x0 ? ix : 0) * 3, o1 = (x1 ? ix + 1 : 0) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int p00 = (y0 && x0) ? r0[o0 + c] : 0, p01 = (y0 && x1) ? r0[o1 + c] : 0;
        const int p10 = (y1 && x0) ? r1[o0 + c] : 0, p11 = (y1 && x1) ? r1[o1 + c] : 0;
        dst[c] = static_cast<uint8_t>((p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15);
    }
}

__global__ void __launch_bounds__(256) remap_lines_kernel(const uint8_t* __restrict__ img, int img_h, int img_w,
                                                          const float* __restrict__ coords,
                                                          const int64_t* __restrict__ coord_off,
                                                          const int32_t* __restrict__ widths, int line_h,
                                                          uint8_t* __restrict__ out, int out_w, int pad) {
    const int line = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= out_w) return;
    uint8_t* dst = out + ((static_cast<size_t>(line) * line_h + y) * out_w + x) * 3;
    const int w = widths[line];
    const int cx = x - pad;                       // column inside the line's own crop
    if (cx < 0 || cx >= w) {                      // padding (and lines cut at the batch width keep their left part)
        dst[0] = dst[1] = dst[2] = 0;
        return;
    }
    const float2 xy = *reinterpret_cast<const float2*>(coords + coord_off[line] + (static_cast<size_t>(y) * w + cx) * 2);
    remap_pixel(img, img_h, img_w, xy.x, xy.y, dst);
}

// The same resampling with the sampling map computed in flight from the fitted baseline polynomial (the tail of
// get_crop_inputs, crop_engine.py:74-99, for the `poly` > 0 configurations): per output column the arc-length sample,
// the baseline point, the unit normal from a 0.1 px forward difference, per pixel the offset along the normal and the
// rotation back to page coordinates.  Every float64 operation is the one NumPy performs, in NumPy's order, spelled
// with explicit rounding intrinsics so that nvcc cannot contract them -- except the final rotation, where np.dot's
// BLAS kernel computes fma(y, r1, x * r0) (pinned by tests/test_oracle_cropper.py on this container's BLAS): the
// float32 maps are bit-identical to the reference's, so are the crops.  25 MB of maps per page shrink to ~10 KB of
// line parameters, and the host no longer evaluates 40 x w float64 maps in NumPy (5 of its 8 ms per line).
__device__ __forceinline__ double poly_eval(const b200ocr_poly_line_t& L, double x) {
    double y = 0.0;
    for (int i = 0; i < L.ncoef; ++i) y = __dadd_rn(__dmul_rn(y, x), L.coef[i]);   // np.polyval: y = y * x + p[i]
    return y;
}

__global__ void __launch_bounds__(256) remap_poly_lines_kernel(const uint8_t* __restrict__ img, int img_h, int img_w,
                                                               const b200ocr_poly_line_t* __restrict__ lines,
                                                               const double* __restrict__ offsets, int line_h,
                                                               uint8_t* __restrict__ out, int out_w, int pad) {
    const int line = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= out_w) return;
    uint8_t* dst = out + ((static_cast<size_t>(line) * line_h + y) * out_w + x) * 3;
    const b200ocr_poly_line_t& L = lines[line];
    const int cx = x - pad;
    if (cx < 0 || cx >= L.n_out || L.ncoef <= 0) {        // padding; ncoef == 0: the reference's all-zero fallback crop
        dst[0] = dst[1] = dst[2] = 0;
        return;
    }
    // np.linspace(0, total, n_out)[cx]
    const double smp = L.n_out == 1 ? 0.0 : (cx == L.n_out - 1 ? L.total : __dmul_rn(static_cast<double>(cx), L.step));
    // reverse_line_mapping (its search loop never advances): blend between xs[-1] and xs[0]
    const double da = __ddiv_rn(__dsub_rn(smp, L.total), __dsub_rn(0.0, L.total));
    const double ox = __dadd_rn(__dmul_rn(__dsub_rn(1.0, da), L.x_last), __dmul_rn(da, L.x_first));
    const double oy = poly_eval(L, ox);
    const double dy = __dsub_rn(oy, poly_eval(L, __dadd_rn(ox, 0.1)));
    const double len = __dsqrt_rn(__dadd_rn(__dmul_rn(0.1, 0.1), __dmul_rn(dy, dy)));
    const double nx = __ddiv_rn(-dy, len), ny = __ddiv_rn(0.1, len);
    const double off = offsets[static_cast<size_t>(line) * line_h + y];
    const double mx = __dadd_rn(__dmul_rn(nx, off), ox), my = __dadd_rn(__dmul_rn(ny, off), oy);
    const float fx = __double2float_rn(__fma_rn(my, L.rot[2], __dmul_rn(mx, L.rot[0])));
    const float fy = __double2float_rn(__fma_rn(my, L.rot[3], __dmul_rn(mx, L.rot[1])));
    remap_pixel(img, img_h, img_w, fx, fy, dst);
}

}  // namespace

cudaError_t launch_remap_poly_lines(const uint8_t* img, int img_h, int img_w, const b200ocr_poly_line_t* lines,
                                    const double* offsets, int n, int line_h, uint8_t* out, int out_w, int pad,
                                    cudaStream_t stream) {
    if (n <= 0 || out_w <= 0) return cudaSuccess;
    if (n > 65535 || line_h > 65535) return cudaErrorInvalidValue;
    dim3 grid((out_w + 255) / 256, line_h, n);
    remap_poly_lines_kernel<<<grid, 256, 0, stream>>>(img, img_h, img_w, lines, offsets, line_h, out, out_w, pad);
    return cudaGetLastError();
}

// Zero padding + stacking of host crops (line_ocr_engine.py:121-127) on the device: the crops arrive packed back to
// back ([line_h][w_i][3] each, one contiguous host copy per line), and this kernel spreads them into the padded
// batch.  One thread per 4 output bytes of a row (rows are out_w * 3 bytes, out_w a multiple of 8: 4-byte aligned).
__global__ void __launch_bounds__(256) pad_lines_kernel(const uint8_t* __restrict__ packed,
                                                        const int64_t* __restrict__ line_off,
                                                        const int32_t* __restrict__ widths, int line_h,
                                                        uint8_t* __restrict__ out, int out_w, int pad) {
    const int line = blockIdx.z, y = blockIdx.y;
    const int row_bytes = out_w * 3;
    const int b0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (b0 >= row_bytes) return;
    const int w = min(widths[line], out_w - pad);           // lines cut at the batch width keep their left part
    const int lo = pad * 3, hi = (pad + max(w, 0)) * 3;     // byte range of the row that holds crop pixels
    const uint8_t* src = packed + line_off[line] + static_cast<size_t>(y) * widths[line] * 3 - lo;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = b0 + k;
        if (b >= lo && b < hi) v |= static_cast<uint32_t>(__ldg(src + b)) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(line) * line_h + y) * row_bytes + b0) = v;
}

cudaError_t launch_pad_lines(const uint8_t* packed, const int64_t* line_off, const int32_t* widths, int n, int line_h,
                             uint8_t* out, int out_w, int pad, cudaStream_t stream) {
    if (n <= 0 || out_w <= 0) return cudaSuccess;
    if (n > 65535 || line_h > 65535 || (out_w & 3)) return cudaErrorInvalidValue;
    dim3 grid((out_w * 3 / 4 + 255) / 256, line_h, n);
    pad_lines_kernel<<<grid, 256, 0, stream>>>(packed, line_off, widths, line_h, out, out_w, pad);
    return cudaGetLastError();
}

cudaError_t launch_remap_lines(const uint8_t* img, int img_h, int img_w, const float* coords, const int64_t* coord_off,
                               const int32_t* widths, int n, int line_h, uint8_t* out, int out_w, int pad,
                               cudaStream_t stream) {
    if (n <= 0 || out_w <= 0) return cudaSuccess;
    if (n > 65535 || line_h > 65535) return cudaErrorInvalidValue;
    dim3 grid((out_w + 255) / 256, line_h, n);
    remap_lines_kernel<<<grid, 256, 0, stream>>>(img, img_h, img_w, coords, coord_off, widths, line_h, out, out_w, pad);
    return cudaGetLastError();
}
