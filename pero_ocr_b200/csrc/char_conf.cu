// Per-character confidences of a CTC transcription on the device (second half).
//
// Replaces the loop of get_line_confidence (pero_ocr/core/confidence_estimation.py:73-104): with a_i the frame that
// align_text picked for character i (b200ocr_force_align's char_positions) and borders b_i = (a_{i-1} + 1 + a_i) // 2
// (b_0 = 0, the border after the last character is (a_last + 1 + 1000) // 2, i.e. the end of the line),
//     conf_i = max(0, p[a_i, label_i] - max_{t in [b_i, b_{i+1}), c not in {label_{i-1}, label_i, label_{i+1}, blank}} p[t, c])
// where p = exp(log_probs) and blank is the LAST class (`masked_probs[:, :-1]`).  exp is monotone, so the inner maximum
// is taken over the log-probabilities and exponentiated once.  One CTA per line, one warp per character, lanes over
// classes.  Values differ from NumPy's only by the last-ulp difference between CUDA's expf and NumPy's float32 exp.
#include "kernels.cuh"

#include <math.h>

namespace {

constexpr int CC_THREADS = 256;

__global__ void __launch_bounds__(CC_THREADS) char_conf_kernel(const float* __restrict__ logp, int t_max, int C,
                                                               const int32_t* __restrict__ n_frames,
                                                               const int32_t* __restrict__ labels, int l_max,
                                                               const int32_t* __restrict__ lengths,
                                                               const int32_t* __restrict__ char_pos,
                                                               float* __restrict__ conf) {
    const int line = blockIdx.x;
    const int L = min(lengths[line], l_max);
    const int T = n_frames ? min(n_frames[line], t_max) : t_max;
    const int32_t* lab = labels + static_cast<size_t>(line) * l_max;
    const int32_t* pos = char_pos + static_cast<size_t>(line) * l_max;
    const float* X = logp + static_cast<size_t>(line) * t_max * C;
    float* out = conf + static_cast<size_t>(line) * l_max;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = CC_THREADS / 32;
    for (int i = L + threadIdx.x; i < l_max; i += CC_THREADS) out[i] = 0.f;
    for (int i = warp; i < L; i += warps) {
        const int a = pos[i];
        if (a < 0 || a >= T) {                       // no alignment for this line (status != 0)
            if (lane == 0) out[i] = 0.f;
            continue;
        }
        const int lo = i == 0 ? 0 : (pos[i - 1] + 1 + a) / 2;
        const int nxt = i + 1 < L ? pos[i + 1] : 1000;
        const int hi = min(T, (a + 1 + nxt) / 2);
        const int me = lab[i], prev = i > 0 ? lab[i - 1] : -1, next = i + 1 < L ? lab[i + 1] : -1;
        float m = -INFINITY;
        for (int t = lo; t < hi; ++t) {
            const float* row = X + static_cast<size_t>(t) * C;
            for (int c = lane; c < C - 1; c += 32) {
                if (c == me || c == prev || c == next) continue;
                m = fmaxf(m, row[c]);
            }
        }
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
            // a fully masked window (every non-blank class is one of the three labels) has maximum 0 in the reference
            const float other = m == -INFINITY ? 0.f : expf(m);
            out[i] = fmaxf(0.f, expf(X[static_cast<size_t>(a) * C + me]) - other);
        }
    }
}

}  // namespace

cudaError_t launch_char_conf(const float* logp, int n, int t_max, int C, const int32_t* n_frames, const int32_t* labels,
                             int l_max, const int32_t* lengths, const int32_t* char_pos, float* conf,
                             cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    char_conf_kernel<<<n, CC_THREADS, 0, stream>>>(logp, t_max, C, n_frames, labels, l_max, lengths, char_pos, conf);
    return cudaGetLastError();
}
