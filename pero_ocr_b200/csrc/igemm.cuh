// Shared declarations for the implicit-GEMM convolution / GEMM kernels.
//
// One kernel family serves every dense contraction on the hot path:
//   out[pixel][cout] = epilogue( sum_{tap, cin} in[pixel + tap][cin] * w[tap][cout][cin] )
// A "pixel" is a position of an NHWC fp16 activation tensor; a plain GEMM is the 1x1 case on a [1][1][M][K] image.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "actfmt.cuh"

enum IgemmCorrection : int { CORR_BOTH = 0, CORR_WEIGHT = 1, CORR_NONE = 2 };

enum IgemmEpilogue : int {
    EPI_ACT_F16 = 0,  // bias + act (+pool) (+affine) -> fp16 NHWC (hi plane, optional lo plane)
    EPI_F32 = 1,      // bias (+act) -> fp32 [pixel][cout]
    EPI_CTC = 2,      // bias -> optional fp32 logits [pixel][cout] + per-pixel argmax / max / logsumexp
    EPI_RES_F32 = 3   // bias + residual(fp32 [pixel][cout]) -> fp32 [pixel][cout]  (transformer sub-layer output)
};

struct IgemmParams {
    // input activation: NHWC records [n_img][h_in][w_in][cin * planes fp16-sized slots] (actfmt.cuh): plane 0 = fp16
    // hi, plane 1 = fp16 lo (npass 3) or the 2*cin e5m2 bytes lo' | hi8 (npass 2)
    int n_img, h_in, w_in, cin;
    int cout;      // real output channels
    int cout_pad;  // rows per tap in the packed weight matrix (= tiles_n * BN)
    int kh, kw, pad_h, pad_w;
    int h_out, w_out;  // conv output geometry before pooling
    int pool_h, pool_w;
    int act;
    float act_slope;          // LeakyReLU negative slope (act == 2)
    int npass;  // 1 (fp16), 3 (fp16x3: hi*hi + hi*lo + lo*hi) or 2 (fp16 hi*hi + one e5m2 correction pass, actfmt.cuh)
    float acc_scale;  // epilogue multiplies the accumulator by this before the bias (2^-11 when npass == 2, else 1)
    // npass == 2 only: which first-order correction terms the e5m2 pass evaluates.  The pass contracts the
    // K-concatenated operands [al*2^11 | ah] . [wh | wl*2^11] (actfmt.cuh); CORR_BOTH walks all 2*cin bytes,
    // CORR_WEIGHT only the second half (ah . wl: the rounding of the WEIGHTS, the larger and the static error term;
    // 1.5 pass-equivalents instead of 2), CORR_NONE skips the pass (the layer runs plain fp16).
    int corr_mode;
    int out_fmt;      // ActFormat of the EPI_ACT_F16 output
    int th;     // rows per segment (1 or 2); a segment is th x (32/th) output pixels = one TMEM lane quarter
    int segs_per_row, row_groups, total_segs, m_tiles, tiles_n;
    int epi;
    const float* bias;        // [cout] or null
    const float* post_scale;  // [cout] or null
    const float* post_shift;
    const float* residual;    // EPI_RES_F32
    __half* out_h;            // EPI_ACT_F16
    int out_cstride;          // fp16 elements per output pixel (cout * planes)
    int out_lo_off;           // element offset of the lo plane inside a pixel, or -1
    float* out_f32;           // EPI_F32 / EPI_RES_F32 / EPI_CTC logits (may be null for CTC)
    int32_t* best;            // EPI_CTC
    float* fmax;
    float* flse;
    float* fprob;             // best-class probability under the reference's sparsified softmax (or null)
    int* tile_counter;        // dynamic tile scheduler: zeroed device counter of this launch (null = static striding)
    int out_skip_lo;          // ACT_F16_F8 output: the consumer never reads the lo' plane (weight-side correction only)
    int dbg;                  // bring-up switches (B200OCR_IGEMM_DBG, flag 9): 1 = skip epilogue stores, 2 = skip MMA issue, 4 = 16-byte instead of 256-bit epilogue stores, 8 = 4 weight stages in the BN = 128 halo kernel
};

inline void igemm_fill_geometry(IgemmParams& p, int bn) {
    p.th = (p.pool_h == 2 || (p.h_out % 2 == 0)) ? 2 : 1;
    const int segw = 32 / p.th;
    p.segs_per_row = (p.w_out + segw - 1) / segw;
    p.row_groups = (p.h_out + p.th - 1) / p.th;
    p.total_segs = p.n_img * p.row_groups * p.segs_per_row;
    p.m_tiles = (p.total_segs + 3) / 4;
    p.tiles_n = (p.cout + bn - 1) / bn;
    p.cout_pad = p.tiles_n * bn;
}

inline int igemm_pick_bn(int cout) { return cout > 128 ? 256 : (cout > 64 ? 128 : 64); }

// Launches the tcgen05 kernel.  tmA: 4D map over the input activation (C, W, H, N), box (64, 32/th, th, 1),
// 128B swizzle; tmB: 2D map over the packed weights (cin, rows), box (64, bn).  Returns cudaGetLastError().
cudaError_t launch_igemm_tc(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, int num_sms,
                            cudaStream_t stream);

// Naive CUDA-core kernel with identical semantics (debug / cross-check only).
cudaError_t launch_igemm_ref(const IgemmParams& p, const __half* in, const __half* w_packed, cudaStream_t stream);

// Halo-reuse variant for 3x3 / pad 1 convolutions with cin <= 128 (igemm_halo.cu).  tmA_halo: 4-D map over the
// input activation with box (64, 130, 4, 1); tmB: 2-D weight map with box (64, bn), bn in {64, 128}.
bool igemm_halo_supported(const IgemmParams& p, int bn);
cudaError_t launch_igemm_halo(const IgemmParams& p, const CUtensorMap& tmA_halo, const CUtensorMap& tmB, int bn,
                              int num_sms, cudaStream_t stream);
