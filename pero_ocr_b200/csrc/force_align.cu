// CTC forced alignment on the device.
//
// Replaces force_align / viterbi_align / align_text of pero_ocr/core/force_alignment.py:13-165 -- a per-line Python
// Viterbi over the 2L+1 states (blank, c1, blank, c2, ..., blank) of a transcription against the [T][C] negative
// log-probabilities of the line -- consumed by the ALTO export (core/layout.py:489-519) and by the per-character
// confidences (core/confidence_estimation.py:73-110).  One CTA per line, states across threads, frames sequential.
//
// Semantics mirrored exactly:
//   * transitions (hmm_trans_from_string, :38-60): stay; advance by one; skip the blank between two DIFFERENT symbols;
//   * start in state 0 or 1, end in state S-2 or S-1 (initial_cost / final_cost, :78-101);
//   * costs accumulate in float64 whatever the input type (float64 initial cost + input), frame by frame;
//   * ties: compute_update (:118-130) walks the transitions ordered by source state and replaces only on a strictly
//     smaller cost, so the LOWEST source state wins (skip before advance before stay); a state with no finite
//     predecessor keeps back-pointer 0; the final state is np.argmin over the two end states (first minimum);
//   * the path has infinite cost -> status 1 (the reference raises ValueError, :146-147); the blank symbol inside
//     the transcription or an empty transcription -> status 2 (:41-43, :64-68);
//   * align_text (:152-165): for every character the frame, among those aligned to it, with the largest per-frame
//     maximum probability (first such frame).
#include "once.cuh"
#include "kernels.cuh"

#include <math.h>

namespace {

constexpr int FA_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(FA_THREADS) force_align_kernel(const T* __restrict__ neg, int t_max, int C,
                                                                 const int32_t* __restrict__ n_frames,
                                                                 const int32_t* __restrict__ labels, int l_max,
                                                                 const int32_t* __restrict__ lengths, int blank,
                                                                 uint8_t* __restrict__ bp_ws,
                                                                 int32_t* __restrict__ out_symbols,
                                                                 int32_t* __restrict__ out_positions,
                                                                 int32_t* __restrict__ char_pos,
                                                                 int32_t* __restrict__ status) {
    extern __shared__ double s_cost[];              // [2][S]
    const int line = blockIdx.x;
    const int L = lengths[line];
    const int T_ = n_frames ? min(n_frames[line], t_max) : t_max;
    const int S = 2 * L + 1;
    const int32_t* lab = labels + static_cast<size_t>(line) * l_max;
    const T* X = neg + static_cast<size_t>(line) * t_max * C;
    uint8_t* bp = bp_ws + static_cast<size_t>(line) * t_max * (2 * l_max + 1);
    int32_t* o_sym = out_symbols ? out_symbols + static_cast<size_t>(line) * t_max : nullptr;
    int32_t* o_pos = out_positions ? out_positions + static_cast<size_t>(line) * t_max : nullptr;
    int32_t* o_chr = char_pos ? char_pos + static_cast<size_t>(line) * l_max : nullptr;
    __shared__ int s_bad;

    if (threadIdx.x == 0) s_bad = (L < 1 || L > l_max || T_ < 1) ? 2 : 0;
    for (int t = threadIdx.x; t < t_max; t += FA_THREADS) {
        if (o_sym) o_sym[t] = -1;
        if (o_pos) o_pos[t] = -1;
    }
    if (o_chr) for (int i = threadIdx.x; i < l_max; i += FA_THREADS) o_chr[i] = -1;
    __syncthreads();
    const bool shape_ok = s_bad == 0;
    bool sym_bad = false;
    if (shape_ok)
        for (int i = threadIdx.x; i < L; i += FA_THREADS) {
            const int c = lab[i];
            sym_bad |= (c == blank || c < 0 || c >= C);
        }
    if (__syncthreads_or(sym_bad ? 1 : 0) || !shape_ok) {
        if (threadIdx.x == 0) status[line] = 2;
        return;
    }
    double* cur = s_cost;
    double* nxt = s_cost + S;
    const double inf = INFINITY;
    // frame 0
    for (int i = threadIdx.x; i < S; i += FA_THREADS) {
        const int sym = (i & 1) ? lab[i >> 1] : blank;
        cur[i] = (i < 2 ? 0.0 : inf) + static_cast<double>(X[sym]);
    }
    __syncthreads();
    for (int t = 1; t < T_; ++t) {
        const T* row = X + static_cast<size_t>(t) * C;
        uint8_t* bpt = bp + static_cast<size_t>(t) * S;
        for (int i = threadIdx.x; i < S; i += FA_THREADS) {
            const int sym = (i & 1) ? lab[i >> 1] : blank;
            const double x = static_cast<double>(row[sym]);
            double best = inf;
            int from = 0;                           // np.zeros back-pointer when nothing finite arrives
            if ((i & 1) && i >= 3 && lab[i >> 1] != lab[(i >> 1) - 1]) {
                const double c = cur[i - 2] + x;
                if (c < best) { best = c; from = i - 2; }
            }
            if (i >= 1) {
                const double c = cur[i - 1] + x;
                if (c < best) { best = c; from = i - 1; }
            }
            {
                const double c = cur[i] + x;
                if (c < best) { best = c; from = i; }
            }
            nxt[i] = best;
            bpt[i] = static_cast<uint8_t>(best < inf ? i - from : 255);   // 255: the zero back-pointer of np.zeros
        }
        __syncthreads();
        double* tmp = cur; cur = nxt; nxt = tmp;
    }
    // final state + backtrack (one thread; T steps)
    __shared__ int s_final;
    if (threadIdx.x == 0) {
        const double a = cur[S - 2], b = cur[S - 1];
        const double m = fmin(a, b);
        if (m == inf) {
            s_final = -1;
            status[line] = 1;
        } else {
            s_final = (a <= b) ? S - 2 : S - 1;     // np.argmin: first minimum
            status[line] = 0;
        }
    }
    __syncthreads();
    if (s_final < 0) return;
    if (threadIdx.x == 0) {
        int st = s_final;
        for (int t = T_ - 1; t >= 0; --t) {
            if (o_sym) o_sym[t] = (st & 1) ? lab[st >> 1] : blank;
            if (o_pos) o_pos[t] = (st & 1) ? (st >> 1) : -1;
            if (t > 0) {
                const uint8_t d = bp[static_cast<size_t>(t) * S + st];
                st = d == 255 ? 0 : st - d;
            }
        }
    }
    if (!o_chr) return;
    __syncthreads();
    // align_text: per character, the aligned frame with the largest per-frame max probability (first maximum)
    double* fmaxp = s_cost;                          // reuse (the launcher sizes the buffer as max(2S, T) doubles)
    for (int t = threadIdx.x; t < T_; t += FA_THREADS) {
        const T* row = X + static_cast<size_t>(t) * C;
        T m = -row[0];
        bool nan = m != m;
        for (int c = 1; c < C; ++c) {
            const T v = -row[c];
            if (v != v) nan = true;
            if (v > m) m = v;
        }
        fmaxp[t] = nan ? static_cast<double>(NAN) : static_cast<double>(m);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += FA_THREADS) {
        int best_t = -1;
        double best = 0.0;
        bool best_nan = false;
        for (int t = 0; t < T_; ++t) {
            if (o_pos[t] != i) continue;
            const double v = fmaxp[t];
            if (best_t < 0) { best_t = t; best = v; best_nan = v != v; }
            else if (!best_nan && (v != v || v > best)) { best_t = t; best = v; best_nan = v != v; }   // np.argmax: NaN wins
        }
        o_chr[i] = best_t;
    }
}

}  // namespace

size_t force_align_workspace_bytes(int n, int t_max, int l_max) {
    return static_cast<size_t>(n) * t_max * (2 * static_cast<size_t>(l_max) + 1);
}

cudaError_t launch_force_align(const void* neg, int is_f64, int n, int t_max, int C, const int32_t* n_frames,
                               const int32_t* labels, int l_max, const int32_t* lengths, int blank, uint8_t* bp_ws,
                               int32_t* out_symbols, int32_t* out_positions, int32_t* char_pos, int32_t* status,
                               cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const size_t S = 2 * static_cast<size_t>(l_max) + 1;
    const size_t dyn = sizeof(double) * (2 * S > static_cast<size_t>(t_max) ? 2 * S : static_cast<size_t>(t_max));
    if (dyn > 200 * 1024) return cudaErrorInvalidValue;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaFuncSetAttribute(force_align_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(force_align_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done.mark();
    }
    if (is_f64)
        force_align_kernel<double><<<n, FA_THREADS, dyn, stream>>>(static_cast<const double*>(neg), t_max, C, n_frames,
                                                                   labels, l_max, lengths, blank, bp_ws, out_symbols,
                                                                   out_positions, char_pos, status);
    else
        force_align_kernel<float><<<n, FA_THREADS, dyn, stream>>>(static_cast<const float*>(neg), t_max, C, n_frames,
                                                                  labels, l_max, lengths, blank, bp_ws, out_symbols,
                                                                  out_positions, char_pos, status);
    return cudaGetLastError();
}
