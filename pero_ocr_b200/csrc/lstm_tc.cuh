// Persistent tcgen05 BiLSTM recurrence (see lstm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

// w_rec: packed recurrent weights, fp16 [2 dirs][planes][8 CTAs][128 rows = gate*32 + unit][K = 256].
// pre: fp32 [n_lines*T][2*4H] pre-gates; out: fp16 [n_lines*T][planes*2H].
cudaError_t launch_lstm_tc(const __half* w_rec, const float* pre, __half* out, int n_lines, int T, int H, int planes,
                           cudaStream_t stream);
size_t lstm_tc_smem_bytes(int planes);
