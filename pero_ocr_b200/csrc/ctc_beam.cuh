// CTC prefix beam search (log domain, fp64) -- replaces CTCPrefixLogRawNumpyDecoder.__call__ without LM
// (pero_ocr/decoding/decoders.py:220-299).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// 0 = configuration not supported (k > 64, c > 1024 or the candidate table does not fit shared memory)
size_t ctc_beam_workspace_bytes(int n, int t, int c, int k);

// t_lo / t_hi: optional device i32 [n], the frame range [lo, hi) decoded per line (NULL = all t frames).
cudaError_t launch_ctc_prefix_beam(const double* logprobs, int n, int t, int c, int k, const int32_t* t_lo,
                                   const int32_t* t_hi, int32_t* out_labels, int32_t* out_lengths, double* out_scores,
                                   int32_t* status, void* workspace, cudaStream_t stream);
