// CTC prefix beam search on the GPU: one CTA per line, beam state and the per-frame candidate table in shared
// memory (fp64, like the reference's NumPy arrays), prefixes as a back-pointer trie in global memory.
//
// Follows CTCPrefixLogRawNumpyDecoder.__call__ (pero_ocr/decoding/decoders.py:220-299) frame by frame:
//   select_relevant_logits (:166-167)  -> order-preserving block compaction of the classes with logprob > -10
//   compute_Pnb (:193-201)             -> candidate table [beam][S+2] (S extend columns, a -inf dummy, "keep")
//   adjust_for_prefix_joining (:138-155) -> trie walk finds the parent prefix inside the beam
//   compute_Pb (:207-208), top_k (multisort.py:4-15) -> k rounds of block arg-max (ties: lower flat index)
//   find_new_prefixes (:116-131)       -> new trie nodes
// and BagOfHypotheses.sort (bag_of_hypotheses.py:19-20) for the output order.
#include "once.cuh"
#include "ctc_beam.cuh"

#include <math.h>
#include <math_constants.h>

namespace {

constexpr int kThreads = 128;
constexpr int kMaxK = 64;
constexpr int kMaxC = 1024;
constexpr int kSelPerLane = 16;                     // candidates a lane keeps in registers for the warp-level top-k
constexpr int kSurvMax = 512;                       // survivors of the top-k prefilter kept in shared memory
constexpr int kSelMax = kThreads * kSelPerLane;   // 2048: beyond that the block-wide rounds are used
#define kNegInf (-CUDART_INF)

__device__ __forceinline__ double lae(double a, double b) {  // logaddexp
    if (a == kNegInf) return b;
    if (b == kNegInf) return a;
    const double d = a - b;
    return d > 0 ? a + log1p(exp(-d)) : b + log1p(exp(d));
}

struct Beam {
    double pb[kMaxK], pnb[kMaxK];
    int last[kMaxK], node[kMaxK], len[kMaxK];
};

__device__ bool same_string(const int* parent, const int* chr, int a, int b) {
    while (a != b) {
        if (a == 0 || b == 0) return false;
        if (chr[a] != chr[b]) return false;
        a = parent[a];
        b = parent[b];
    }
    return true;
}

__global__ void __launch_bounds__(kThreads)
prefix_beam_kernel(const double* __restrict__ lp_all, int T, int C, int K, int32_t* out_labels, int32_t* out_lengths,
                   double* out_scores, int32_t* status, int* ws_parent, int* ws_char, int nodes_per_line,
                   const int32_t* __restrict__ t_lo, const int32_t* __restrict__ t_hi) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int cols_max = C + 1;  // S + 2 <= (C - 1) + 2
    double* table = reinterpret_cast<double*>(dyn);              // [K][cols_max]
    double* score = table + static_cast<size_t>(K) * cols_max;   // [K][cols_max]
    __shared__ Beam beams[2];
    __shared__ short pos[kMaxC];
    __shared__ short sel[kMaxC];
    __shared__ double pc[kMaxC + 1];
    __shared__ double new_pb[kMaxK];
    __shared__ int rlast[kMaxK];
    __shared__ int joinq[kMaxK];
    __shared__ int warp_cnt[kThreads / 32];
    __shared__ double red_v[kThreads / 32];
    __shared__ int red_i[kThreads / 32];
    __shared__ int pick_r[kMaxK], pick_c[kMaxK];
    __shared__ double s_lmax[kThreads], surv_v[kSurvMax], s_thr;
    __shared__ int surv_i[kSurvMax], s_m;
    __shared__ double cand_v[(kThreads / 32) * kMaxK];
    __shared__ int cand_i[(kThreads / 32) * kMaxK];
    __shared__ int s_S, s_flag, s_nkeep;

    const int line = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // frames [f_lo, f_hi) of the line are decoded (PageDecoder slices the log-probs by logit_coords,
    // document_ocr/page_parser.py:133-135); the whole matrix when no ranges are given
    const int f_lo = t_lo ? max(0, min(T, t_lo[line])) : 0;
    const int f_hi = t_hi ? max(f_lo, min(T, t_hi[line])) : T;
    const double* lp = lp_all + static_cast<size_t>(line) * T * C;
    int* parent = ws_parent + static_cast<size_t>(line) * nodes_per_line;
    int* chr = ws_char + static_cast<size_t>(line) * nodes_per_line;
    const int blank = C - 1;

    // normalisation gate (decoders.py:223-224): max_t |sum_c exp(lp) - 1| <= 1e-5
    if (tid == 0) s_flag = 0;
    __syncthreads();
    for (int t = f_lo + tid; t < f_hi; t += kThreads) {
        double s = 0;
        for (int c = 0; c < C; ++c) s += exp(lp[static_cast<size_t>(t) * C + c]);
        if (!(fabs(s - 1.0) <= 1e-5)) s_flag = 1;
    }
    __syncthreads();
    if (s_flag) {
        if (tid == 0) status[line] = 1;
        for (int i = tid; i < K; i += kThreads) {
            out_lengths[static_cast<size_t>(line) * K + i] = -1;
            out_scores[static_cast<size_t>(line) * K + i] = kNegInf;
        }
        return;
    }
    if (tid == 0) {
        status[line] = 0;
        beams[0].pb[0] = 0.0;
        beams[0].pnb[0] = kNegInf;
        beams[0].last[0] = 0;  // decoders.py:246 (zeros)
        beams[0].node[0] = 0;  // trie root = empty prefix
        beams[0].len[0] = 0;
        parent[0] = 0;
        chr[0] = -1;
    }
    int nb = 1, cur = 0, next_node = 1;
    __syncthreads();

    for (int t = f_lo; t < f_hi; ++t) {
        const double* row = lp + static_cast<size_t>(t) * C;
        Beam& B = beams[cur];
        Beam& N = beams[cur ^ 1];
        const double p_blank = row[blank];
        // ---- relevant characters, in increasing class order
        const int per = (blank + kThreads - 1) / kThreads;
        const int c_lo = tid * per, c_hi = min(blank, c_lo + per);
        int cnt = 0;
        for (int c = c_lo; c < c_hi; ++c) cnt += row[c] > -10.0;
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_cnt[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int w2 = 0; w2 < warp; ++w2) base += warp_cnt[w2];
        int at = base + incl - cnt;
        for (int c = c_lo; c < c_hi; ++c) {
            if (row[c] > -10.0) {
                sel[at] = static_cast<short>(c);
                pc[at] = row[c];
                pos[c] = static_cast<short>(at);
                ++at;
            } else {
                pos[c] = -1;
            }
        }
        if (tid == kThreads - 1) s_S = base + incl;
        __syncthreads();
        const int S = s_S;
        if (S == 0) {  // decoders.py:252-255
            if (tid < nb) {
                B.pb[tid] = lae(B.pb[tid], B.pnb[tid]) + p_blank;
                B.pnb[tid] = kNegInf;
            }
            __syncthreads();
            continue;
        }
        const int cols = S + 2;
        if (tid == 0) pc[S] = kNegInf;
        if (tid < nb) {
            const int lc = B.last[tid];
            const int p = (lc >= 0 && lc < blank) ? pos[lc] : -1;
            rlast[tid] = p >= 0 ? p : S;
            new_pb[tid] = lae(B.pb[tid], B.pnb[tid]) + p_blank;
            joinq[tid] = -1;
        }
        __syncthreads();
        // ---- candidate table
        for (int i = tid; i < nb * cols; i += kThreads) {
            const int p = i / cols, j = i - p * cols;
            double v;
            if (j == S + 1) v = B.pnb[p] + pc[rlast[p]];
            else v = lae(B.pb[p] + pc[j], j == rlast[p] ? kNegInf : B.pnb[p] + pc[j]);
            table[p * cols_max + j] = v;
        }
        // ---- prefix joining: q = the beam entry whose prefix is prefix[p] minus its last character
        for (int i = tid; i < nb * nb; i += kThreads) {
            const int p = i / nb, q = i - p * nb;
            if (B.len[p] >= 1 && B.len[q] == B.len[p] - 1 && same_string(parent, chr, parent[B.node[p]], B.node[q]))
                joinq[p] = q;
        }
        __syncthreads();
        if (tid < nb && joinq[tid] >= 0 && rlast[tid] < S) {   // column S is the all -inf dummy: nothing to move
            const int q = joinq[tid];
            double& mine = table[tid * cols_max + S + 1];
            double& theirs = table[q * cols_max + rlast[tid]];
            mine = lae(mine, theirs);
            theirs = kNegInf;
        }
        __syncthreads();
        // ---- scores + number of finite candidates
        int fin = 0;
        for (int i = tid; i < nb * cols; i += kThreads) {
            const int p = i / cols, j = i - p * cols;
            double v = table[p * cols_max + j];
            if (j == S + 1) v = lae(new_pb[p], v);
            score[p * cols_max + j] = v;
            fin += isfinite(v) ? 1 : 0;
        }
        for (int o = 16; o; o >>= 1) fin += __shfl_xor_sync(0xffffffffu, fin, o);
        if (lane == 0) warp_cnt[warp] = fin;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w2 = 0; w2 < kThreads / 32; ++w2) total += warp_cnt[w2];
            s_nkeep = min(K, total);
        }
        __syncthreads();
        const int n_keep = s_nkeep;
        // ---- top-k (value desc, flat index asc)
        const int M = nb * cols;
        if (M <= kSelMax) {
            // two-level selection without block barriers in the pick loop: every lane keeps its <= 16 candidates
            // (flat index tid + 128 e) in registers, each warp extracts the n_keep best of its 32 x 16 by n_keep
            // shuffle arg-max rounds, then warp 0 merges the 4 x n_keep survivors the same way.  (n_keep block-wide
            // arg-max rounds with two barriers each were half of the per-frame time at k = 16.)
            double lv[kSelPerLane];
#pragma unroll
            for (int e = 0; e < kSelPerLane; ++e) {
                const int i = tid + e * kThreads;
                double v = kNegInf;
                if (i < M) {
                    const int p = i / cols, j = i - p * cols;
                    v = score[p * cols_max + j];
                    if (v != v) v = kNegInf;
                }
                lv[e] = v;
            }
            // prefilter: the n_keep-th largest of the 128 per-thread maxima is a lower bound of the n_keep-th best
            // candidate, so only candidates at or above it can be picked (typically 16-40 of ~1500); they are
            // compacted into shared memory and ranked exactly (value desc, flat index asc)
            double lmax = kNegInf;
#pragma unroll
            for (int e = 0; e < kSelPerLane; ++e) lmax = fmax(lmax, lv[e]);
            s_lmax[tid] = lmax;
            if (tid == 0) s_m = 0;
            __syncthreads();
            {
                int rank = 0;
                for (int j2 = 0; j2 < kThreads; ++j2) {
                    const double o = s_lmax[j2];
                    rank += (o > lmax || (o == lmax && j2 < tid)) ? 1 : 0;
                }
                if (rank == n_keep - 1) s_thr = lmax;
            }
            __syncthreads();
            const double thr = s_thr;
#pragma unroll
            for (int e = 0; e < kSelPerLane; ++e)
                if (lv[e] != kNegInf && lv[e] >= thr) {
                    const int at = atomicAdd(&s_m, 1);
                    if (at < kSurvMax) {
                        surv_v[at] = lv[e];
                        surv_i[at] = tid + e * kThreads;
                    }
                }
            __syncthreads();
            const int m = s_m;
            if (m <= kSurvMax) {
                for (int q = tid; q < m; q += kThreads) {
                    const double v = surv_v[q];
                    const int i = surv_i[q];
                    int rank = 0;
                    for (int q2 = 0; q2 < m; ++q2) {
                        const double v2 = surv_v[q2];
                        rank += (v2 > v || (v2 == v && surv_i[q2] < i)) ? 1 : 0;
                    }
                    if (rank < n_keep) {
                        const int p = i / cols;
                        pick_r[rank] = p;
                        pick_c[rank] = i - p * cols;
                    }
                }
                __syncthreads();
            } else {
            // (ties at the threshold overflowed the survivor list: exact two-level warp selection)
            unsigned taken = 0;
            for (int r = 0; r < n_keep; ++r) {
                double bv = kNegInf;
                int be = -1;
#pragma unroll
                for (int e = 0; e < kSelPerLane; ++e)
                    if (!((taken >> e) & 1u) && lv[e] > bv) {
                        bv = lv[e];
                        be = e;
                    }
                const int mine = be >= 0 ? tid + be * kThreads : 0x7fffffff;
                int bi = mine;
                for (int o = 16; o; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) {
                        bv = ov;
                        bi = oi;
                    }
                }
                if (be >= 0 && bi == mine) taken |= 1u << be;
                if (lane == 0) {
                    cand_v[warp * kMaxK + r] = bv;
                    cand_i[warp * kMaxK + r] = bi;
                }
            }
            __syncthreads();
            if (warp == 0) {
                constexpr int kPer = (kThreads / 32) * kMaxK / 32;      // merged candidates per lane
                double mv[kPer];
                int mi[kPer];
#pragma unroll
                for (int e = 0; e < kPer; ++e) {
                    const int slot = lane + e * 32;
                    const int w2 = slot / kMaxK, r2 = slot - w2 * kMaxK;
                    const bool ok = r2 < n_keep;
                    mv[e] = ok ? cand_v[slot] : kNegInf;
                    mi[e] = ok ? cand_i[slot] : 0x7fffffff;
                }
                unsigned tk = 0;
                for (int r = 0; r < n_keep; ++r) {
                    double bv = kNegInf;
                    int bi = 0x7fffffff, be = -1;
#pragma unroll
                    for (int e = 0; e < kPer; ++e)
                        if (!((tk >> e) & 1u) && (mv[e] > bv || (mv[e] == bv && mv[e] != kNegInf && mi[e] < bi))) {
                            bv = mv[e];
                            bi = mi[e];
                            be = e;
                        }
                    const int mine = bi;
                    for (int o = 16; o; o >>= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) {
                            bv = ov;
                            bi = oi;
                        }
                    }
                    if (be >= 0 && bi == mine) tk |= 1u << be;
                    if (lane == 0) {
                        const int p = bi / cols;
                        pick_r[r] = p;
                        pick_c[r] = bi - p * cols;
                    }
                }
            }
            __syncthreads();
            }
        } else
        for (int round = 0; round < n_keep; ++round) {
            double bv = kNegInf;
            int bi = 0x7fffffff;
            for (int i = tid; i < nb * cols; i += kThreads) {
                const int p = i / cols, j = i - p * cols;
                const double v = score[p * cols_max + j];
                if (v > bv || (v == bv && v != kNegInf && i < bi)) {
                    bv = v;
                    bi = i;
                }
            }
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) {
                    bv = ov;
                    bi = oi;
                }
            }
            if (lane == 0) {
                red_v[warp] = bv;
                red_i[warp] = bi;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w2 = 1; w2 < kThreads / 32; ++w2)
                    if (red_v[w2] > bv || (red_v[w2] == bv && red_i[w2] < bi)) {
                        bv = red_v[w2];
                        bi = red_i[w2];
                    }
                const int p = bi / cols, j = bi - p * cols;
                pick_r[round] = p;
                pick_c[round] = j;
                score[p * cols_max + j] = kNegInf;
            }
            __syncthreads();
        }
        // ---- new beam (find_new_prefixes)
        if (tid < n_keep) {
            const int r = pick_r[tid], c = pick_c[tid];
            if (c == S + 1) {
                N.pb[tid] = new_pb[r];
                N.pnb[tid] = table[r * cols_max + c];
                N.last[tid] = B.last[r];
                N.node[tid] = B.node[r];
                N.len[tid] = B.len[r];
            } else {
                const int ch = sel[c];
                const int id = next_node + tid;
                parent[id] = B.node[r];
                chr[id] = ch;
                N.pb[tid] = kNegInf;
                N.pnb[tid] = table[r * cols_max + c];
                N.last[tid] = ch;
                N.node[tid] = id;
                N.len[tid] = B.len[r] + 1;
            }
        }
        next_node += n_keep;
        nb = n_keep;
        cur ^= 1;
        __syncthreads();
    }

    // ---- output, sorted by logaddexp(Pb, Pnb) descending (stable)
    Beam& B = beams[cur];
    if (tid < nb) new_pb[tid] = lae(B.pb[tid], B.pnb[tid]);
    __syncthreads();
    if (tid < nb) {
        int rank = 0;
        for (int j = 0; j < nb; ++j)
            if (new_pb[j] > new_pb[tid] || (new_pb[j] == new_pb[tid] && j < tid)) ++rank;
        const size_t o = static_cast<size_t>(line) * K + rank;
        out_scores[o] = new_pb[tid];
        out_lengths[o] = B.len[tid];
        int32_t* lab = out_labels + o * T;
        int node = B.node[tid];
        for (int i = B.len[tid] - 1; i >= 0; --i) {
            lab[i] = chr[node];
            node = parent[node];
        }
        for (int i = B.len[tid]; i < T; ++i) lab[i] = -1;
    }
    for (int i = nb + tid; i < K; i += kThreads) {
        out_lengths[static_cast<size_t>(line) * K + i] = -1;
        out_scores[static_cast<size_t>(line) * K + i] = kNegInf;
    }
}

size_t dyn_bytes(int c, int k) { return 2 * static_cast<size_t>(k) * (c + 1) * sizeof(double); }

}  // namespace

size_t ctc_beam_workspace_bytes(int n, int t, int c, int k) {
    if (k > kMaxK || c > kMaxC || dyn_bytes(c, k) > 160 * 1024) return 0;
    const size_t nodes = static_cast<size_t>(t) * k + 1;
    return static_cast<size_t>(n) * nodes * 2 * sizeof(int);
}

cudaError_t launch_ctc_prefix_beam(const double* logprobs, int n, int t, int c, int k, const int32_t* t_lo,
                                   const int32_t* t_hi, int32_t* out_labels, int32_t* out_lengths, double* out_scores,
                                   int32_t* status, void* workspace, cudaStream_t stream) {
    const int nodes = t * k + 1;
    int* ws_parent = static_cast<int*>(workspace);
    int* ws_char = ws_parent + static_cast<size_t>(n) * nodes;
    const size_t dyn = dyn_bytes(c, k);
    static PerDeviceOnce attr_done;   // static + dynamic shared memory exceeds the 48 KB default already at k = 16, c = 120
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    prefix_beam_kernel<<<n, kThreads, dyn, stream>>>(logprobs, t, c, k, out_labels, out_lengths, out_scores, status,
                                                     ws_parent, ws_char, nodes, t_lo, t_hi);
    return cudaGetLastError();
}
