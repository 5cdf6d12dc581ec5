// Shared declarations for the implicit-GEMM convolution / GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

enum IgemmEpilogue : int {
    EPI_ACT_F16 = 0,  // bias + act (+pool) (+affine) -> fp16 NHWC (hi plane, optional lo plane)
    EPI_F32 = 1,      // bias (+act) -> fp32 [pixel][cout]
    EPI_CTC = 2,      // bias -> optional fp32 logits [pixel][cout] + per-pixel argmax / max / logsumexp
    EPI_RES_F32 = 3   // bias + residual(fp32 [pixel][cout]) -> fp32 [pixel][cout]  (transformer sub-layer output)
};

struct IgemmParams {
    // input activation: fp16 NHWC [n_img][h_in][w_in][cin * planes]; plane 0 = hi, plane 1 = lo (x3 mode)
    int n_img, h_in, w_in, cin;
    int cout;      // real output channels
    int cout_pad;  // rows per tap in the packed weight matrix (= tiles_n * BN)
    int kh, kw, pad_h, pad_w;
    int h_out, w_out;  // conv output geometry before pooling
    int pool_h, pool_w;
    int act;
    int npass;  // 1 (fp16) or 3 (fp16x3: hi*hi + hi*lo + lo*hi)
    int tiles_x, tiles_y, tiles_n;
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
};

// Launches the tcgen05 kernel.  tmA / tmB are built by make_tmaps_for_igemm().  Returns cudaGetLastError().
cudaError_t launch_igemm_tc(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, int th,
                            int num_sms, cudaStream_t stream);

// Naive CUDA-core kernel with identical semantics (debug / cross-check only).
cudaError_t launch_igemm_ref(const IgemmParams& p, const __half* in, const __half* w_packed, int w_ld,
                             cudaStream_t stream);

// Tile geometry decisions shared by the launcher and the host planner.
inline int igemm_pick_bn(int cout) { return cout > 128 ? 256 : (cout > 64 ? 128 : 64); }
inline int igemm_pick_th(int pool_h, int h_out) { return (pool_h == 2 || (h_out >= 2 && (h_out % 2) == 0)) ? 2 : 1; }
