// Step kernels of the autoregressive Transformer decoder: the per-token work of
// TransformerEngineLineOCR.transcribe_batch (pero_ocr/ocr_engine/transformer_ocr_engine.py:49-89) that is not a GEMM.
//   * embed_pe:        dec_embeder(token) + PositionalEncoding row of the position   (transformer.py:316-332, 512-513)
//   * step_attention:  one query position per (line, head) against S cached key / value positions -- the cached
//                      self-attention (S = tokens so far) and the encoder-decoder attention (S = T memory frames) of
//                      CustomMultiheadAttention.cached_forward (transformer.py:183-305); fp32, softmax over all S
//   * linear_f32:      the per-step projections (M = lines in the batch <= a few hundred rows: in-proj, out-proj,
//                      feed-forward, class projection) as a tiled fp32 CUDA-core GEMM with fused bias / ReLU /
//                      residual; these are weight-streaming bound at M <= 256 and keep the greedy argmax in fp32.
//                      The large contraction of the decoder (memory K / V projection, M = lines x frames) goes
//                      through the tcgen05 kernel instead (engine.cu)
//   * argmax_alive:    torch.argmax over the classes (first maximum) + alive-mask update + stop detection
//                      (transformer_ocr_engine.py:72-77)
#include "once.cuh"
#include "kernels.cuh"

#include <math.h>

#include <algorithm>

namespace {

// tokens == nullptr: every line starts from `start_token` (position 0).  pos_dev != nullptr: the position is read from
// device memory and `tokens` is the BASE of the [position][line] token matrix (the previous position's row is the
// input; position 0 starts from `start_token`) -- every argument is then the same for all positions (CUDA graph).
__global__ void embed_pe_kernel(const float* __restrict__ table, const int32_t* __restrict__ tokens, int start_token,
                                int n, int d, int pos, const int32_t* __restrict__ pos_dev, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d) return;
    const int line = i / d, c = i - line * d;
    if (pos_dev) {
        pos = *pos_dev;
        tokens = pos > 0 ? tokens + static_cast<size_t>(pos - 1) * n : nullptr;
    }
    const int tok = tokens ? tokens[line] : start_token;
    // pe[pos, 2m] = sin(pos * exp(2m * -ln(1e4)/d)), pe[pos, 2m+1] = cos(same)   (transformer.py:321-328)
    const int m2 = c & ~1;
    const float div = expf(static_cast<float>(m2) * (-logf(10000.0f) / d));
    const float ang = static_cast<float>(pos) * div;
    out[i] = table[static_cast<size_t>(tok) * d + c] + ((c & 1) ? cosf(ang) : sinf(ang));
}

// one warp per (line, head); head width a multiple of 4, <= 128.  q: [n][q_ls]; k, v: position p of line l at
// + p * ps + l * ls (16-byte aligned rows).  Scores of the S positions live in shared memory.
__global__ void step_attention_kernel(const float* __restrict__ q, long q_ls, const float* __restrict__ k,
                                      const float* __restrict__ v, long ps, long ls, int n, int S, int d, int heads,
                                      float* __restrict__ out) {
    extern __shared__ float s_w[];                    // [warps][S] scores, then [warps][128] scaled query
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * warps + warp;
    if (unit >= n * heads) return;
    const int line = unit / heads, head = unit - line * heads;
    const int hd = d / heads;
    const float scale = powf(static_cast<float>(hd), -0.5f);
    float* w = s_w + static_cast<size_t>(warp) * S;
    float* qs = s_w + static_cast<size_t>(warps) * S + warp * 128;
    for (int e = lane; e < hd; e += 32) qs[e] = q[line * q_ls + head * hd + e] * scale;   // q scaled first (:268-271)
    __syncwarp();
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    float mx = -INFINITY;
    for (int p = lane; p < S; p += 32) {
        const float4* kp = reinterpret_cast<const float4*>(k + p * ps + line * ls + head * hd);
        float s = 0.f;
        for (int e = 0; e < hd / 4; ++e) {
            const float4 kk = kp[e], qq = q4[e];
            s = fmaf(qq.x, kk.x, s); s = fmaf(qq.y, kk.y, s); s = fmaf(qq.z, kk.z, s); s = fmaf(qq.w, kk.w, s);
        }
        w[p] = s;
        mx = fmaxf(mx, s);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int p = lane; p < S; p += 32) {
        const float e = expf(w[p] - mx);
        w[p] = e;
        sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.f / sum;
    for (int e = lane; e < hd; e += 32) {
        const float* vp = v + line * ls + head * hd + e;
        float acc = 0.f;
        for (int p = 0; p < S; ++p) acc = fmaf(w[p], vp[p * ps], acc);
        out[static_cast<size_t>(line) * d + head * hd + e] = acc * inv;
    }
}

// out[m][o] = act(sum_k x[m][k] * w[o][k] + bias[o]) (+ res[m][o]);  x rows ldx apart, w = PyTorch Linear weight
// [O][K] (K contiguous), out rows ldo apart.  32 x 32 output tile per CTA, 64 threads x (4 x 4), K in chunks of 32.
constexpr int LBM = 32, LBN = 32, LBK = 32, LLD = 33;

__global__ void __launch_bounds__(64) linear_f32_kernel(const float* __restrict__ x, long ldx,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ res, long ldr,
                                                        float* __restrict__ out, long ldo, int M, int O, int K,
                                                        int relu) {
    __shared__ float xs[LBK][LLD];   // [k][m]
    __shared__ float ws[LBK][LLD];   // [k][o]
    const int m0 = blockIdx.y * LBM, o0 = blockIdx.x * LBN;
    const int tid = threadIdx.x, tm = tid >> 3, tn = tid & 7;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += LBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 64;          // 256 float4 per operand tile
            const int r = idx >> 3, k4 = (idx & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
            if (m0 + r < M) v = *reinterpret_cast<const float4*>(x + static_cast<long>(m0 + r) * ldx + k0 + k4);
            if (o0 + r < O) u = *reinterpret_cast<const float4*>(w + static_cast<size_t>(o0 + r) * K + k0 + k4);
            xs[k4 + 0][r] = v.x; xs[k4 + 1][r] = v.y; xs[k4 + 2][r] = v.z; xs[k4 + 3][r] = v.w;
            ws[k4 + 0][r] = u.x; ws[k4 + 1][r] = u.y; ws[k4 + 2][r] = u.z; ws[k4 + 3][r] = u.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < LBK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[k][tm * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ws[k][tn * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + tn * 4 + j;
            if (o >= O) continue;
            float v = acc[i][j] + (bias ? bias[o] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            if (res) v += res[static_cast<long>(m) * ldr + o];
            out[static_cast<long>(m) * ldo + o] = v;
        }
    }
}

// Latency-oriented variant of the same contraction for the token loop, where M (lines in the batch) is small and the
// K loop of the kernel above is a serial chain of global-load round trips: 256 threads = KG groups of 64, group g
// takes the K chunks c = g (mod KG) into its own shared-memory tiles (64-thread named barriers), the next chunk's
// global loads are issued into registers before the current chunk is multiplied, and the KG partial tiles are summed
// in a fixed order (deterministic) by group 0 before the same epilogue.
constexpr int LKG = 4;

__device__ __forceinline__ void group_barrier(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(64) : "memory");
}

//
// Two extensions for the token loop (launch_linear_f32_ex):
//   * split_o / out2: output columns >= split_o go to out2 (rows ldo2 apart) -- the self-attention in-projection
//     writes q into the step buffer and k | v straight into the cache slot of the position with ONE launch;
//   * gridDim.z = Z > 1: CTA z contracts the z-th share of the K chunks and stores its raw partial tile at
//     out + z * part_stride (no bias / activation / residual): the chain of dependent chunk round trips shrinks by Z
//     and the Z partials are summed, in fixed order, by the kernel that consumes them (sum_layernorm_kernel).
__global__ void __launch_bounds__(64 * LKG) linear_f32_splitk_kernel(const float* __restrict__ x, long ldx,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias,
                                                                     const float* __restrict__ res, long ldr,
                                                                     float* __restrict__ out, long ldo, int M, int O,
                                                                     int K, int relu, int split_o,
                                                                     float* __restrict__ out2, long ldo2,
                                                                     long part_stride, const int32_t* __restrict__ pos_dev,
                                                                     long out_pos_stride, long out2_pos_stride) {
    __shared__ float xs[LKG][LBK][LLD];   // [group][k][m]; afterwards [group][m][o] partial sums
    __shared__ float ws[LKG][LBK][LLD];   // [group][k][o]
    const int m0 = blockIdx.y * LBM, o0 = blockIdx.x * LBN;
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63, tm = t >> 3, tn = t & 7;
    const int all_chunks = K / LBK;
    const int Z = gridDim.z, z = blockIdx.z;
    const int c_lo = static_cast<int>(static_cast<long>(all_chunks) * z / Z);
    const int chunks = static_cast<int>(static_cast<long>(all_chunks) * (z + 1) / Z);
    if (Z > 1) { out += z * part_stride; bias = nullptr; res = nullptr; relu = 0; }
    if (pos_dev) {                       // outputs that move with the decoded position (cache slot, logits row)
        const long pos = *pos_dev;
        out += pos * out_pos_stride;
        if (out2) out2 += pos * out2_pos_stride;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 xv[4], wv[4];
    auto fetch = [&](int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = t + i * 64;
            const int r = idx >> 3, k4 = (idx & 7) * 4;
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            wv[i] = xv[i];
            if (m0 + r < M) xv[i] = *reinterpret_cast<const float4*>(x + static_cast<long>(m0 + r) * ldx + c * LBK + k4);
            if (o0 + r < O) wv[i] = *reinterpret_cast<const float4*>(w + static_cast<size_t>(o0 + r) * K + c * LBK + k4);
        }
    };
    int c = c_lo + g;
    if (c < chunks) fetch(c);
    for (; c < chunks; c += LKG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = t + i * 64;
            const int r = idx >> 3, k4 = (idx & 7) * 4;
            xs[g][k4 + 0][r] = xv[i].x; xs[g][k4 + 1][r] = xv[i].y; xs[g][k4 + 2][r] = xv[i].z; xs[g][k4 + 3][r] = xv[i].w;
            ws[g][k4 + 0][r] = wv[i].x; ws[g][k4 + 1][r] = wv[i].y; ws[g][k4 + 2][r] = wv[i].z; ws[g][k4 + 3][r] = wv[i].w;
        }
        group_barrier(g);
        if (c + LKG < chunks) fetch(c + LKG);      // in flight while this chunk is multiplied
#pragma unroll 8
        for (int k = 0; k < LBK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[g][k][tm * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ws[g][k][tn * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        group_barrier(g);
    }
    // partial tiles -> shared memory (each group's own region: its last reads are behind its last barrier)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[g][tm * 4 + i][tn * 4 + j] = acc[i][j];
    __syncthreads();
    if (g != 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + tn * 4 + j;
            if (o >= O) continue;
            float v = xs[0][tm * 4 + i][tn * 4 + j];
#pragma unroll
            for (int q = 1; q < LKG; ++q) v += xs[q][tm * 4 + i][tn * 4 + j];
            v += bias ? bias[o] : 0.f;
            if (relu) v = fmaxf(v, 0.f);
            if (res) v += res[static_cast<long>(m) * ldr + o];
            if (o < split_o) out[static_cast<long>(m) * ldo + o] = v;
            else out2[static_cast<long>(m) * ldo2 + (o - split_o)] = v;
        }
    }
}

// LayerNorm over rows of x = sum_z part[z] + bias + res (the out-projection / feed-forward partials of the split-K
// kernel above, DecoderLayer.infer's  norm(x + sublayer(x)), transformer.py:435, 447, 450).  One warp per row, the row
// in registers: lane l owns the float4s l + 32 j, j < NV (d = 128 NV).
template <int NV>
__global__ void __launch_bounds__(128) sum_layernorm_kernel(const float* __restrict__ part, int Z, long part_stride,
                                                            const float* __restrict__ bias,
                                                            const float* res, int rows,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* out) {             // out may be res (in place)
    constexpr int D = 128 * NV;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = lane + 32 * j;
        v[j] = reinterpret_cast<const float4*>(part + static_cast<size_t>(row) * D)[i4];
        for (int zz = 1; zz < Z; ++zz) {
            const float4 u = reinterpret_cast<const float4*>(part + zz * part_stride + static_cast<size_t>(row) * D)[i4];
            v[j].x += u.x; v[j].y += u.y; v[j].z += u.z; v[j].w += u.w;
        }
        const float4 b = reinterpret_cast<const float4*>(bias)[i4];
        const float4 r = reinterpret_cast<const float4*>(res + static_cast<size_t>(row) * D)[i4];
        v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;      // linear: acc + bias ...
        v[j].x += r.x; v[j].y += r.y; v[j].z += r.z; v[j].w += r.w;      // ... then the residual, as the fused epilogue
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / D + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i4 = lane + 32 * j;
        const float4 g = reinterpret_cast<const float4*>(gamma)[i4], b = reinterpret_cast<const float4*>(beta)[i4];
        float4 y;
        y.x = (v[j].x - mean) * rstd * g.x + b.x;
        y.y = (v[j].y - mean) * rstd * g.y + b.y;
        y.z = (v[j].z - mean) * rstd * g.z + b.z;
        y.w = (v[j].w - mean) * rstd * g.w + b.w;
        reinterpret_cast<float4*>(out + static_cast<size_t>(row) * D)[i4] = y;
    }
}

// One CTA (4 warps) per (line, head): the same arithmetic as step_attention_kernel with every global access
// coalesced and the S positions spread over the CTA.  Scores: 8 lanes share one K row (HD / 8 consecutive floats
// each), 4 rows per warp instruction, 16 per CTA pass, two passes in flight; values: HD / 4 lanes share one V row
// (a float4 each), the 4 warps take interleaved rows and their partial sums meet in shared memory in fixed order.
template <int HD>
__global__ void __launch_bounds__(128) step_attention_cta_kernel(const float* __restrict__ q, long q_ls,
                                                                 const float* __restrict__ k,
                                                                 const float* __restrict__ v, long ps, long ls, int S,
                                                                 int d, int heads, float* __restrict__ out,
                                                                 const int32_t* __restrict__ pos_dev) {
    extern __shared__ float s_w[];                  // [S] scores / weights, [4][HD] partial outputs, [8] reductions
    constexpr int SEG = HD / 8, LPR = HD / 4, RPW = 32 / LPR;
    float* s_part = s_w + ((S + 3) & ~3);           // 16-byte aligned (S = the launch's upper bound)
    if (pos_dev) S = min(S, *pos_dev + 1);          // cached self-attention: positions 0 .. pos
    float* s_red = s_part + 4 * HD;
    const int line = blockIdx.x / heads, head = blockIdx.x - line * heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane >> 3, sg = lane & 7;
    const float scale = powf(static_cast<float>(HD), -0.5f);
    float qr[SEG];
#pragma unroll
    for (int i = 0; i < SEG; ++i) qr[i] = q[line * q_ls + head * HD + sg * SEG + i] * scale;   // q scaled first (:268-271)
    const float* kb = k + line * ls + head * HD + sg * SEG;
    float mx = -INFINITY;
    for (int p0 = warp * 4 + rg; p0 < S; p0 += 32) {
        const int p1 = p0 + 16;
        float4 ka[SEG / 4], kc[SEG / 4];
#pragma unroll
        for (int e = 0; e < SEG / 4; ++e) {
            ka[e] = reinterpret_cast<const float4*>(kb + p0 * ps)[e];
            kc[e] = p1 < S ? reinterpret_cast<const float4*>(kb + p1 * ps)[e] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < SEG / 4; ++e) {
            s0 = fmaf(qr[4 * e], ka[e].x, s0); s0 = fmaf(qr[4 * e + 1], ka[e].y, s0);
            s0 = fmaf(qr[4 * e + 2], ka[e].z, s0); s0 = fmaf(qr[4 * e + 3], ka[e].w, s0);
            s1 = fmaf(qr[4 * e], kc[e].x, s1); s1 = fmaf(qr[4 * e + 1], kc[e].y, s1);
            s1 = fmaf(qr[4 * e + 2], kc[e].z, s1); s1 = fmaf(qr[4 * e + 3], kc[e].w, s1);
        }
        // rows whose p0 is out of range leave the loop by themselves: the 8 lanes of a row always travel together
        const unsigned grp = 0xffu << (lane & 24);
#pragma unroll
        for (int o = 4; o; o >>= 1) {
            s0 += __shfl_xor_sync(grp, s0, o);
            s1 += __shfl_xor_sync(grp, s1, o);
        }
        if (sg == 0) {
            s_w[p0] = s0;
            if (p1 < S) s_w[p1] = s1;
        }
        mx = fmaxf(mx, s0);
        if (p1 < S) mx = fmaxf(mx, s1);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    float sum = 0.f;
    for (int p = threadIdx.x; p < S; p += 128) {
        const float e = expf(s_w[p] - mx);
        s_w[p] = e;
        sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_red[4 + warp] = sum;
    __syncthreads();
    const float inv = 1.f / ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
    const int vr = lane / LPR, vc = lane - vr * LPR;
    const float* vb = v + line * ls + head * HD + vc * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = warp * RPW + vr;
    for (; p + 12 * RPW < S; p += 16 * RPW) {          // four rows of this lane in flight
        const float4 a0 = *reinterpret_cast<const float4*>(vb + p * ps);
        const float4 a1 = *reinterpret_cast<const float4*>(vb + (p + 4 * RPW) * ps);
        const float4 a2 = *reinterpret_cast<const float4*>(vb + (p + 8 * RPW) * ps);
        const float4 a3 = *reinterpret_cast<const float4*>(vb + (p + 12 * RPW) * ps);
        const float w0 = s_w[p], w1 = s_w[p + 4 * RPW], w2 = s_w[p + 8 * RPW], w3 = s_w[p + 12 * RPW];
        acc.x = fmaf(w0, a0.x, acc.x); acc.y = fmaf(w0, a0.y, acc.y); acc.z = fmaf(w0, a0.z, acc.z); acc.w = fmaf(w0, a0.w, acc.w);
        acc.x = fmaf(w1, a1.x, acc.x); acc.y = fmaf(w1, a1.y, acc.y); acc.z = fmaf(w1, a1.z, acc.z); acc.w = fmaf(w1, a1.w, acc.w);
        acc.x = fmaf(w2, a2.x, acc.x); acc.y = fmaf(w2, a2.y, acc.y); acc.z = fmaf(w2, a2.z, acc.z); acc.w = fmaf(w2, a2.w, acc.w);
        acc.x = fmaf(w3, a3.x, acc.x); acc.y = fmaf(w3, a3.y, acc.y); acc.z = fmaf(w3, a3.z, acc.z); acc.w = fmaf(w3, a3.w, acc.w);
    }
    for (; p < S; p += 4 * RPW) {
        const float4 a0 = *reinterpret_cast<const float4*>(vb + p * ps);
        const float w0 = s_w[p];
        acc.x = fmaf(w0, a0.x, acc.x); acc.y = fmaf(w0, a0.y, acc.y); acc.z = fmaf(w0, a0.z, acc.z); acc.w = fmaf(w0, a0.w, acc.w);
    }
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) {               // the RPW rows a warp walks side by side
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane < LPR) *reinterpret_cast<float4*>(s_part + warp * HD + lane * 4) = acc;
    __syncthreads();
    if (threadIdx.x < HD) {
        const int e = threadIdx.x;
        const float o = (s_part[e] + s_part[HD + e]) + (s_part[2 * HD + e] + s_part[3 * HD + e]);
        out[static_cast<size_t>(line) * d + head * HD + e] = o * inv;
    }
}

__device__ __forceinline__ bool score_better(float v, int i, float bv, int bi) {
    // torch.argmax: NaN counts as the maximum; among equal values the lowest index wins
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (!vn && v != bv) return v > bv;
    return i < bi;
}

// One CTA for the whole batch (a few hundred lines x a few hundred classes): warp per line.
// state[0] = lines still alive after this step, state[1] = first step after which none was (-1 until then).
// use_pos: the step is state[2]; `logits` / `tokens_out` are bases that advance by logits_pos_stride / n per position, and
// state[2] is incremented at the end (the token loop as a replayed CUDA graph).
__global__ void argmax_alive_kernel(const float* __restrict__ logits, long ld, int n, int C, int stop_token, int step,
                                    int32_t* __restrict__ tokens_out, int32_t* __restrict__ alive,
                                    int32_t* __restrict__ state, int use_pos, long logits_pos_stride) {
    __shared__ int total;
    if (use_pos) {
        step = state[2];
        logits += step * logits_pos_stride;
        tokens_out += static_cast<size_t>(step) * n;
    }
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    int mine = 0;
    for (int line = warp; line < n; line += warps) {
        const float* row = logits + static_cast<long>(line) * ld;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = row[c];
            if (bi == 0x7fffffff || score_better(v, c, bv, bi)) { bv = v; bi = c; }
        }
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || score_better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            tokens_out[line] = bi;
            const int a = (alive[line] != 0 && bi != stop_token) ? 1 : 0;
            alive[line] = a;
            mine += a;
        }
    }
    if (lane == 0 && mine) atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        state[0] = total;
        if (total == 0 && state[1] < 0) state[1] = step;
        if (use_pos) state[2] = step + 1;
    }
}

__global__ void ar_init_kernel(int32_t* alive, int n, int32_t* state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) alive[i] = 1;
    if (i == 0) { state[0] = n; state[1] = -1; state[2] = 0; }
}

}  // namespace

cudaError_t launch_embed_pe(const float* table, const int32_t* tokens, int start_token, int n, int d, int pos,
                            float* out, cudaStream_t stream, const int32_t* pos_dev) {
    if (n <= 0) return cudaSuccess;
    embed_pe_kernel<<<(n * d + 255) / 256, 256, 0, stream>>>(table, tokens, start_token, n, d, pos, pos_dev, out);
    return cudaGetLastError();
}

cudaError_t launch_linear_f32(const float* x, long ldx, const float* w, const float* bias, const float* res, long ldr,
                              float* out, long ldo, int M, int O, int K, int relu, int variant, cudaStream_t stream) {
    if (M <= 0 || O <= 0) return cudaSuccess;
    if (K <= 0 || (K % LBK) || (ldx % 4) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15))
        return cudaErrorInvalidValue;
    const dim3 grid((O + LBN - 1) / LBN, (M + LBM - 1) / LBM);
    if (variant == 1)
        linear_f32_splitk_kernel<<<grid, 64 * LKG, 0, stream>>>(x, ldx, w, bias, res, ldr, out, ldo, M, O, K, relu, O,
                                                                nullptr, 0, 0, nullptr, 0, 0);
    else
        linear_f32_kernel<<<grid, 64, 0, stream>>>(x, ldx, w, bias, res, ldr, out, ldo, M, O, K, relu);
    return cudaGetLastError();
}

cudaError_t launch_linear_f32_ex(const float* x, long ldx, const float* w, const float* bias, const float* res, long ldr,
                                 float* out, long ldo, int M, int O, int K, int relu, int split_o, float* out2,
                                 long ldo2, int ksplit, long part_stride, cudaStream_t stream, const int32_t* pos_dev,
                                 long out_pos_stride, long out2_pos_stride) {
    if (M <= 0 || O <= 0) return cudaSuccess;
    if (K <= 0 || (K % LBK) || (ldx % 4) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
        ksplit < 1 || ksplit > K / LBK || (split_o < O && !out2) || (ksplit > 1 && (split_o < O || part_stride < static_cast<long>(M) * ldo)))
        return cudaErrorInvalidValue;
    const dim3 grid((O + LBN - 1) / LBN, (M + LBM - 1) / LBM, ksplit);
    linear_f32_splitk_kernel<<<grid, 64 * LKG, 0, stream>>>(x, ldx, w, bias, res, ldr, out, ldo, M, O, K, relu,
                                                            std::min(split_o, O), out2, ldo2, part_stride, pos_dev,
                                                            out_pos_stride, out2_pos_stride);
    return cudaGetLastError();
}

bool sum_layernorm_supported(int d) { return d == 128 || d == 256 || d == 512 || d == 1024; }

cudaError_t launch_sum_layernorm(const float* part, int Z, long part_stride, const float* bias, const float* res,
                                 int rows, int d, const float* gamma, const float* beta, float eps, float* out,
                                 cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    if (!sum_layernorm_supported(d) || Z < 1 || !bias || !res) return cudaErrorInvalidValue;
    const int grid = (rows + 3) / 4;
    switch (d) {
        case 128: sum_layernorm_kernel<1><<<grid, 128, 0, stream>>>(part, Z, part_stride, bias, res, rows, gamma, beta, eps, out); break;
        case 256: sum_layernorm_kernel<2><<<grid, 128, 0, stream>>>(part, Z, part_stride, bias, res, rows, gamma, beta, eps, out); break;
        case 512: sum_layernorm_kernel<4><<<grid, 128, 0, stream>>>(part, Z, part_stride, bias, res, rows, gamma, beta, eps, out); break;
        default: sum_layernorm_kernel<8><<<grid, 128, 0, stream>>>(part, Z, part_stride, bias, res, rows, gamma, beta, eps, out); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_argmax_alive(const float* logits, long ld, int n, int C, int stop_token, int step,
                                int32_t* tokens_out, int32_t* alive, int32_t* state, cudaStream_t stream, int use_pos,
                                long logits_pos_stride) {
    if (n <= 0) return cudaSuccess;
    argmax_alive_kernel<<<1, 256, 0, stream>>>(logits, ld, n, C, stop_token, step, tokens_out, alive, state, use_pos,
                                               logits_pos_stride);
    return cudaGetLastError();
}

cudaError_t launch_ar_init(int32_t* alive, int n, int32_t* state, cudaStream_t stream) {
    ar_init_kernel<<<(std::max(n, 1) + 255) / 256, 256, 0, stream>>>(alive, n, state);
    return cudaGetLastError();
}

cudaError_t launch_step_attention(const float* q, long q_ls, const float* k, const float* v, long ps, long ls, int n,
                                  int S, int d, int heads, float* out, int variant, cudaStream_t stream,
                                  const int32_t* pos_dev) {
    if (n <= 0 || S <= 0) return cudaSuccess;
    const int hd = heads > 0 ? d / heads : 0;
    if (heads <= 0 || hd * heads != d || (hd % 4) || hd > 128 || (ps % 4) || (ls % 4) ||
        (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15))
        return cudaErrorInvalidValue;
    if (variant == 1 && (hd == 32 || hd == 64 || hd == 128) && S <= 8192) {
        const size_t sm = (((static_cast<size_t>(S) + 3) & ~size_t(3)) + 4 * hd + 8) * sizeof(float);      // < 48 KB
        if (hd == 32) step_attention_cta_kernel<32><<<n * heads, 128, sm, stream>>>(q, q_ls, k, v, ps, ls, S, d, heads, out, pos_dev);
        else if (hd == 64) step_attention_cta_kernel<64><<<n * heads, 128, sm, stream>>>(q, q_ls, k, v, ps, ls, S, d, heads, out, pos_dev);
        else step_attention_cta_kernel<128><<<n * heads, 128, sm, stream>>>(q, q_ls, k, v, ps, ls, S, d, heads, out, pos_dev);
        return cudaGetLastError();
    }
    if (pos_dev) return cudaErrorInvalidValue;      // the device-side position needs the CTA-per-head kernel
    const int warps = 4;
    const size_t smem = static_cast<size_t>(warps) * (S + 128) * sizeof(float);
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(step_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    const int units = n * heads;
    step_attention_kernel<<<(units + warps - 1) / warps, warps * 32, smem, stream>>>(q, q_ls, k, v, ps, ls, n, S, d, heads,
                                                                                    out);
    return cudaGetLastError();
}
