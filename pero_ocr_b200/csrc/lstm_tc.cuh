// Persistent tcgen05 BiLSTM recurrence (see lstm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

// w_rec: packed recurrent weights, fp16 [2 dirs][planes][8 CTAs][128 rows = gate*32 + unit][K = 256].
// pre: fp32 [n_lines*T][2*4H] pre-gates; out: activation records [n_lines*T] of 2H channels in format `out_fmt`
// (actfmt.cuh).  `planes` (of W_hh) and `hplanes` (of the exchanged h_t, <= planes) select the recurrence arithmetic:
// (1, 1) fp16, (2, 2) fp16x3, (2, 1) split weights x fp16-rounded h_t; independent of out_fmt.
cudaError_t launch_lstm_tc(const __half* w_rec, const float* pre, __half* out, int n_lines, int T, int H, int planes,
                           int hplanes, int out_fmt, cudaStream_t stream);
size_t lstm_tc_smem_bytes(int hplanes);
