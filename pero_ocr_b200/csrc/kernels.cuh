// Non-GEMM kernels of the line-recognition path + CUDA-core cross-check kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "actfmt.cuh"   // `fmt` arguments below are ActFormat values
#include "../../include/b200_lineocr.h"   // b200ocr_poly_line_t

// u8 NHWC [n][h][w][3] -> x/255 -> 3x3 conv (pad 1) + bias + act -> fp16 NHWC [n][h][w][planes*cout]
// (replaces the `/255` + permute of run_ocr, pytorch_ocr_engine.py:61-62, and the first conv of the blob).
// w_t: fp32 [27][cout] (tap-major: (r*3+s)*3+c), cout <= 64.
cudaError_t launch_conv_first(const uint8_t* in, int n, int h, int w, const float* w_t, const float* bias, int cout,
                              int act, float slope, int fmt, __half* out, cudaStream_t stream);

// The same layer on warp-level tensor cores (conv_first.cu; the product path -- the kernel above is the fp32
// cross-check).  conv_first_pack turns the PyTorch weight [cout][3][3][3] into per-lane mma.sync fragments
// (conv_first_wfrag_words(cout) 32-bit words) and the per-channel epilogue scale (power-of-two weight scale / 255).
// staging: how the uint8 patch reaches shared memory -- 0 plain loads, 1 cp.async, 2 TMA (tm_in: 3-D uint32 tensor map
// {W*3/4, H, N}, box {104, 6, 1}, no swizzle); the bulk variants need W % 16 == 0 and fall back to 0 otherwise.
cudaError_t launch_conv_first_mma(const uint8_t* in, int n, int h, int w, const uint32_t* wfrag, const float* oscale,
                                  const float* bias, int cout, int act, float slope, int fmt, __half* out, int skip_lo,
                                  int staging,
                                  const CUtensorMap* tm_in, cudaStream_t stream);
size_t conv_first_wfrag_words(int cout);
// tcgen05 variant of the first conv (conv_first.cu): weights as [plane hi | lo][cout][32 K slots] fp16
size_t conv_first_tc_words(int cout);
void conv_first_pack_tc(const float* weight, int cout, uint32_t* wk);
cudaError_t launch_conv_first_tc(const uint8_t* in, int n, int h, int w, const uint32_t* wk, const float* oscale,
                                 const float* bias, int cout, int act, float slope, int fmt, __half* out, int skip_lo,
                                 const CUtensorMap* tm_in, cudaStream_t stream);
void conv_first_pack(const float* weight, int cout, uint32_t* wfrag, float* oscale);

// Per-frame argmax (first maximal index, NaN maximal) / max / logsumexp / sparsified-softmax best prob over
// materialised scores.  layout 0 = [n][t][c], 1 = [n][c][t].
cudaError_t launch_frame_stats(const float* scores, int n, int t, int c, int layout, int32_t* best, float* fmax,
                               float* flse, float* fprob, cudaStream_t stream);

// CTC forced alignment (force_align.cu; core/force_alignment.py).  bp_ws: force_align_workspace_bytes(n, t, l_max).
cudaError_t launch_force_align(const void* neg, int is_f64, int n, int t_max, int C, const int32_t* n_frames,
                               const int32_t* labels, int l_max, const int32_t* lengths, int blank, uint8_t* bp_ws,
                               int32_t* out_symbols, int32_t* out_positions, int32_t* char_pos, int32_t* status,
                               cudaStream_t stream);
size_t force_align_workspace_bytes(int n, int t_max, int l_max);

// Autoregressive decoder step kernels (ar_step.cu; transformer_ocr_engine.py:49-89, transformer.py:183-305, 418-462).
// embed_pe: out[line][:] = table[tokens ? tokens[line] : start_token][:] + sinusoid(pos)   (fp32 [n][d])
// pos_dev (optional, all four launchers below): the decoded position lives in device memory, so that the launch
// arguments of a position are the same for every position and the token loop can be replayed as a CUDA graph.
cudaError_t launch_embed_pe(const float* table, const int32_t* tokens, int start_token, int n, int d, int pos,
                            float* out, cudaStream_t stream, const int32_t* pos_dev = nullptr);
// out[m][o] = act(x[m][:] . w[o][:] + bias[o]) (+ res[m][o]); fp32, w = Linear weight [O][K], K % 32 == 0.
// variant 0: one 64-thread group walks K; 1: four groups split K (register prefetch, fixed-order reduction).
cudaError_t launch_linear_f32(const float* x, long ldx, const float* w, const float* bias, const float* res, long ldr,
                              float* out, long ldo, int M, int O, int K, int relu, int variant, cudaStream_t stream);
// The split-K kernel with (a) output columns >= split_o redirected to out2 (rows ldo2 apart) and (b) the K chunks
// divided over `ksplit` CTAs per tile, CTA z storing its raw partial tile at out + z * part_stride (no bias / relu /
// residual; needs split_o >= O).  launch_sum_layernorm finishes (b): LayerNorm(sum_z part[z] + bias + res).
cudaError_t launch_linear_f32_ex(const float* x, long ldx, const float* w, const float* bias, const float* res, long ldr,
                                 float* out, long ldo, int M, int O, int K, int relu, int split_o, float* out2,
                                 long ldo2, int ksplit, long part_stride, cudaStream_t stream,
                                 const int32_t* pos_dev = nullptr, long out_pos_stride = 0, long out2_pos_stride = 0);
bool sum_layernorm_supported(int d);
cudaError_t launch_sum_layernorm(const float* part, int Z, long part_stride, const float* bias, const float* res,
                                 int rows, int d, const float* gamma, const float* beta, float eps, float* out,
                                 cudaStream_t stream);
// one query position per (line, head) against S key / value positions (position p of line l at + p*ps + l*ls).
// variant 0: one warp per (line, head); 1: one CTA per (line, head), coalesced rows (heads of 32 / 64 / 128).
cudaError_t launch_step_attention(const float* q, long q_ls, const float* k, const float* v, long ps, long ls, int n,
                                  int S, int d, int heads, float* out, int variant, cudaStream_t stream,
                                  const int32_t* pos_dev = nullptr);
// greedy choice + alive mask + stop detection; state = {alive lines, first step after which none was alive or -1}.
cudaError_t launch_argmax_alive(const float* logits, long ld, int n, int C, int stop_token, int step,
                                int32_t* tokens_out, int32_t* alive, int32_t* state, cudaStream_t stream, int use_pos = 0,
                                long logits_pos_stride = 0);
cudaError_t launch_ar_init(int32_t* alive, int n, int32_t* state, cudaStream_t stream);

// Per-character confidences (char_conf.cu; core/confidence_estimation.py:73-104).
cudaError_t launch_char_conf(const float* logp, int n, int t_max, int C, const int32_t* n_frames, const int32_t* labels,
                             int l_max, const int32_t* lengths, const int32_t* char_pos, float* conf,
                             cudaStream_t stream);

// Packed host crops -> zero-padded recogniser batch (remap.cu; line_ocr_engine.py:121-127).
cudaError_t launch_pad_lines(const uint8_t* packed, const int64_t* line_off, const int32_t* widths, int n, int line_h,
                             uint8_t* out, int out_w, int pad, cudaStream_t stream);
// Bilinear 8-bit remap of all lines of a page into the padded recogniser batch (remap.cu; crop_engine.py:146-163).
cudaError_t launch_remap_lines(const uint8_t* img, int img_h, int img_w, const float* coords, const int64_t* coord_off,
                               const int32_t* widths, int n, int line_h, uint8_t* out, int out_w, int pad,
                               cudaStream_t stream);

cudaError_t launch_remap_poly_lines(const uint8_t* img, int img_h, int img_w, const b200ocr_poly_line_t* lines,
                                    const double* offsets, int n, int line_h, uint8_t* out, int out_w, int pad,
                                    cudaStream_t stream);

// Logit sparsification (sparsify.cu): softmax threshold 1e-4 + CSC of every line (line_ocr_engine.py:168-172).
cudaError_t launch_sparsify(const float* logits, int n, int T, int C, const int32_t* t_lo, const int32_t* t_hi,
                            int32_t* indptr, int32_t* nnz, int64_t* base, int32_t* indices, float* data,
                            int64_t capacity, cudaStream_t stream);

// Dense log-probabilities as seen through the sparsify -> get_full_logprobs round trip (sparsify.cu).
cudaError_t launch_full_logprobs(const float* logits, int n, int T, int C, double* out, cudaStream_t stream);

// greedy_decode_ctc's collapse (pytorch_ocr_engine.py:19-27) on per-frame argmax ids: drop repeats (frame 0 is
// compared with a virtual blank), drop blanks; left-packed labels (-1 padded) + lengths; optional line
// confidence = get_prob (page_parser.py:437-450) over fprob.
cudaError_t launch_ctc_collapse(const int32_t* best, const float* fprob, int n, int t, int blank, int32_t* labels,
                                int32_t* lengths, float* confidence, cudaStream_t stream);

// CUDA-core bidirectional LSTM layer (cross-check path).  pre: fp32 [n*T][2*4H] (x W_ih^T + b_ih + b_hh, forward
// gates then reverse gates, PyTorch order i,f,g,o); w_hh_t: fp32 [2][H][4H]; out: fp16 [n*T][planes*2H].
cudaError_t launch_lstm_ref(const float* pre, const float* w_hh_t, int n, int T, int H, int fmt, int round_fp16,
                            __half* out, cudaStream_t stream);

// nearest-neighbour upsampling fp32 [n][h][w][c] (NHWC) -> fp32 [n][c][h*f][w*f] (NCHW)  (ParseNet tail)
cudaError_t launch_upsample_nchw(const float* in, int n, int h, int w, int c, int f, float* out, cudaStream_t stream);

// LayerNorm over channels (+ optional sinusoidal positional encoding indexed by t = row % T) of fp32 rows
// [rows][d] -> fp32 [rows][d] and fp16 hi/lo [rows][planes*d]  (transformer.py:316-332, 378-381).
cudaError_t launch_layernorm(const float* in, int rows, int d, const float* gamma, const float* beta, float eps,
                             int pe_T, float* out_f32, __half* out_h, int fmt, cudaStream_t stream);

// fp16 hi/lo NHWC [rows][planes*d] -> fp32 [rows][d]
cudaError_t launch_h2f(const __half* in, int rows, int d, int fmt, float* out, cudaStream_t stream);

// Multi-head self-attention over one line: qkv fp32 [n*T][3*D] (q | k | v), heads of D/heads -> fp16 hi/lo
// [n*T][planes*D].  softmax(q k^T / sqrt(dh)) v, fp32 math  (nn.MultiheadAttention inside
// nn.TransformerEncoderLayer, transformer.py:371-373).
cudaError_t launch_attention(const float* qkv, int n, int T, int D, int heads, __half* out, int fmt,
                             cudaStream_t stream);
// The same on tcgen05 for 64-wide heads and lines of up to 384 frames (attention_tc.cu): S = Q K^T and O = P V with
// hi / lo operand splits, P kept in tensor memory between the two products.
bool attention_tc_supported(int T, int D, int heads);
cudaError_t launch_attention_tc(const float* qkv, int n, int T, int D, int heads, __half* out, int fmt,
                                cudaStream_t stream);
