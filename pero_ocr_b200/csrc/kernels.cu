// Non-GEMM kernels of the line-recognition path (sm_100a) + CUDA-core cross-check kernels.
#include "once.cuh"
#include "kernels.cuh"
#include "igemm.cuh"

#include <math.h>

namespace {

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : slope * v;
    return v;
}

// ------------------------------------------------------------------------------------------------ first conv
// One CTA = CF_ROWS image rows x 128 consecutive output pixels.  The (CF_ROWS+2) x 130 x 3 input patch is staged in
// shared memory as fp32 through a 256-entry table of k / 255.0f (the reference's `/255`, pytorch_ocr_engine.py:61,
// bit-exact and without a division per byte; zero outside the image); each thread computes one pixel x COUT
// channels per row in fp32 (weights broadcast from shared memory as float4) and stores its fp16 (hi | lo) pixel
// record with 16-byte stores.
constexpr int CF_PX = 128;
constexpr int CF_ROWS = 4;

template <int COUT>
__global__ void __launch_bounds__(CF_PX) conv_first_kernel(const uint8_t* __restrict__ in, int n, int h, int w,
                                                          const float* __restrict__ w_t, const float* __restrict__ bias,
                                                          int act, float slope, int fmt, __half* __restrict__ out) {
    const int planes = act_planes(fmt);
    __shared__ float s_in[CF_ROWS + 2][CF_PX + 2][3];
    __shared__ __align__(16) float s_w[27 * COUT];
    __shared__ __align__(16) float s_b[COUT];
    __shared__ float s_lut[256];
    extern __shared__ uint4 s_stage[];   // [CF_PX][planes * COUT / 8]: one row of pixel records (hi | lo)

    const int tiles_w = (w + CF_PX - 1) / CF_PX;
    const int tiles_h = (h + CF_ROWS - 1) / CF_ROWS;
    const int tw = blockIdx.x % tiles_w;
    const int row0 = ((blockIdx.x / tiles_w) % tiles_h) * CF_ROWS;
    const int img = blockIdx.x / (tiles_w * tiles_h);
    const int w0 = tw * CF_PX;

    for (int i = threadIdx.x; i < 256; i += CF_PX) s_lut[i] = static_cast<float>(i) / 255.0f;
    for (int i = threadIdx.x; i < 27 * COUT; i += CF_PX) s_w[i] = w_t[i];
    for (int i = threadIdx.x; i < COUT; i += CF_PX) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    constexpr int kRowBytes = (CF_PX + 2) * 3;
    for (int i = threadIdx.x; i < (CF_ROWS + 2) * kRowBytes; i += CF_PX) {
        const int r = i / kRowBytes;
        const int b = i - r * kRowBytes;  // byte inside the staged row: pixel b / 3, channel b % 3
        const int yy = row0 + r - 1;
        const int xb = (w0 - 1) * 3 + b;  // byte offset inside the image row
        float v = 0.f;
        if (yy >= 0 && yy < h && xb >= 0 && xb < w * 3)
            v = s_lut[in[(static_cast<size_t>(img) * h + yy) * w * 3 + xb]];
        (&s_in[r][0][0])[b] = v;
    }
    __syncthreads();

    const int rec = planes * COUT;
    for (int rr = 0; rr < CF_ROWS; ++rr) {
        const int row = row0 + rr;
        if (row >= h) break;
        float acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
        // PyTorch weight [cout][c][r][s]; w_t index ((r*3+s)*3+c)*COUT + o
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float x = s_in[rr + r][threadIdx.x + s][c];
                    const float4* wr = reinterpret_cast<const float4*>(&s_w[((r * 3 + s) * 3 + c) * COUT]);
#pragma unroll
                    for (int o = 0; o < COUT / 4; ++o) {
                        const float4 wv = wr[o];
                        acc[4 * o + 0] = fmaf(x, wv.x, acc[4 * o + 0]);
                        acc[4 * o + 1] = fmaf(x, wv.y, acc[4 * o + 1]);
                        acc[4 * o + 2] = fmaf(x, wv.z, acc[4 * o + 2]);
                        acc[4 * o + 3] = fmaf(x, wv.w, acc[4 * o + 3]);
                    }
                }
        // pixel records -> shared memory (16-byte chunks, XOR-swizzled by pixel so that both the per-pixel writes
        // and the per-record reads are bank-conflict free) -> fully coalesced 16-byte global stores
        const int chunks = rec / 8;                 // 16-byte chunks per pixel record
        __syncthreads();                            // previous row's staging has been read
        {
            uint4* my = s_stage + threadIdx.x * chunks;
            const int sw = threadIdx.x & (chunks - 1);
#pragma unroll
            for (int o = 0; o < COUT; o += 16) {
                uint32_t ph[8], pl[8];   // pl: 8 words fp16 lo (HILO) or 4 words lo' + 4 words hi8 (F8)
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = act_fn(acc[o + e + q] + s_b[o + e + q], act, slope);
                    const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    ph[e >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
                    ph[(e >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    if (fmt == ACT_F16_HILO) {
                        const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y);
                        const __half2 l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
                        pl[e >> 1] = *reinterpret_cast<const uint32_t*>(&l01);
                        pl[(e >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    } else if (fmt == ACT_F16_F8) {
                        pl[e >> 2] = pack_e5m2x4((v[0] - f01.x) * kF8Scale, (v[1] - f01.y) * kF8Scale,
                                                 (v[2] - f23.x) * kF8Scale, (v[3] - f23.y) * kF8Scale);
                        pl[4 + (e >> 2)] = pack_e5m2x4(f01.x, f01.y, f23.x, f23.y);
                    }
                }
                my[(o / 8) ^ sw] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                my[(o / 8 + 1) ^ sw] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                if (fmt == ACT_F16_HILO) {
                    my[(COUT / 8 + o / 8) ^ sw] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    my[(COUT / 8 + o / 8 + 1) ^ sw] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                } else if (fmt == ACT_F16_F8) {
                    my[(COUT / 8 + o / 16) ^ sw] = make_uint4(pl[0], pl[1], pl[2], pl[3]);              // lo' bytes
                    my[(COUT / 8 + COUT / 16 + o / 16) ^ sw] = make_uint4(pl[4], pl[5], pl[6], pl[7]);  // hi8 bytes
                }
            }
        }
        __syncthreads();
        {
            const int npx = min(CF_PX, w - w0);
            uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * h + row) * w + w0) * rec);
            for (int i = threadIdx.x; i < npx * chunks; i += CF_PX) {
                const int px = i / chunks, c = i - px * chunks;
                dst[i] = s_stage[px * chunks + (c ^ (px & (chunks - 1)))];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ frame stats
__device__ __forceinline__ void stat_update(float v, int idx, float& bv, int& bi, bool& bnan) {
    if (bnan) return;
    if (v != v) {
        bnan = true;
        bi = idx;
        bv = v;
    } else if (idx == 0 || v > bv) {
        bv = v;
        bi = idx;
    }
}

// one thread per frame; works for both layouts through (stride_c, stride_t)
__global__ void frame_stats_kernel(const float* __restrict__ s, int n, int t, int c, long stride_n, long stride_t,
                                   long stride_c, int32_t* best, float* fmax, float* flse, float* fprob) {
    const long f = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (f >= static_cast<long>(n) * t) return;
    const int line = static_cast<int>(f / t), fr = static_cast<int>(f % t);
    const float* p = s + line * stride_n + fr * stride_t;
    float bv = -INFINITY, m = -INFINITY, sum = 0.f;
    int bi = 0;
    bool bnan = false;
    for (int j = 0; j < c; ++j) {
        const float v = p[j * stride_c];
        stat_update(v, j, bv, bi, bnan);
        const float m2 = fmaxf(m, v);
        sum = sum * __expf(m - m2) + __expf(v - m2);
        m = m2;
    }
    best[f] = bi;
    if (fmax) fmax[f] = bv;
    if (flse) flse[f] = m + __logf(sum);
    if (fprob) {
        // best-class probability under the reference's sparsify -> dense(-80) -> log-softmax chain
        // (line_ocr_engine.py:168-171, core/layout.py:65-68, page_parser.py:486-490)
        float kept = 0.f;
        int dropped = 0;
        const float thr = 1e-4f * sum;
        for (int j = 0; j < c; ++j) {
            const float v = p[j * stride_c];
            const float e = __expf(v - m);
            if (e < thr || v == 0.f) ++dropped;
            else kept += e;
        }
        kept += dropped * __expf(-80.f - m);
        fprob[f] = __expf(bv - m) / kept;
    }
}

// ------------------------------------------------------------------------------------------------ CTC collapse
// one warp per line
__global__ void ctc_collapse_kernel(const int32_t* __restrict__ best, const float* __restrict__ fprob, int n, int t,
                                    int blank, int32_t* labels, int32_t* lengths, float* confidence) {
    const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (line >= n) return;
    const int32_t* b = best + static_cast<size_t>(line) * t;
    int32_t* out = labels + static_cast<size_t>(line) * t;
    int count = 0;
    for (int base = 0; base < t; base += 32) {
        const int i = base + lane;
        int cur = blank, prev = blank;
        if (i < t) {
            cur = b[i];
            prev = (i == 0) ? blank : b[i - 1];
        }
        const bool keep = (i < t) && cur != prev && cur != blank;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) out[count + __popc(mask & ((1u << lane) - 1))] = cur;
        count += __popc(mask);
    }
    for (int i = count + lane; i < t; i += 32) out[i] = -1;
    if (lane == 0) {
        lengths[line] = count;
        if (confidence) {
            const float* pr = fprob + static_cast<size_t>(line) * t;
            float worst = 1.f, run_p = 1.f;
            int run_id = -1;
            for (int i = 0; i < t; ++i) {
                const int id = b[i];
                const float p = pr[i];
                if (id != run_id) {
                    worst = fminf(worst, run_p);
                    run_p = p;
                    run_id = id;
                } else {
                    run_p = fmaxf(run_p, p);
                }
            }
            confidence[line] = fminf(worst, run_p);
        }
    }
}

// ------------------------------------------------------------------------------------------------ LSTM (ref)
// grid (ceil(n/4), 2 directions), 4H/4 = H threads: thread j owns hidden unit j of 4 lines.
constexpr int LR_LINES = 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void lstm_ref_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh_t, int n, int T, int H,
                                int fmt, int round_fp16, __half* __restrict__ out) {
    const int planes = act_planes(fmt);
    extern __shared__ float s_h[];  // [LR_LINES][H]
    const int dir = blockIdx.y;
    const int line0 = blockIdx.x * LR_LINES;
    const int j = threadIdx.x;
    const float* wt = w_hh_t + static_cast<size_t>(dir) * H * 4 * H;
    float c[LR_LINES];
    for (int l = 0; l < LR_LINES; ++l) {
        c[l] = 0.f;
        s_h[l * H + j] = 0.f;
    }
    __syncthreads();
    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        float a[4][LR_LINES];
        for (int g = 0; g < 4; ++g)
            for (int l = 0; l < LR_LINES; ++l) a[g][l] = 0.f;
        for (int k = 0; k < H; ++k) {
            float hk[LR_LINES];
            for (int l = 0; l < LR_LINES; ++l) hk[l] = s_h[l * H + k];
            for (int g = 0; g < 4; ++g) {
                const float wv = wt[static_cast<size_t>(k) * 4 * H + g * H + j];
                for (int l = 0; l < LR_LINES; ++l) a[g][l] = fmaf(wv, hk[l], a[g][l]);
            }
        }
        __syncthreads();
        for (int l = 0; l < LR_LINES; ++l) {
            const int line = line0 + l;
            if (line >= n) continue;
            const size_t row = static_cast<size_t>(line) * T + t;
            const float* pr = pre + row * (8 * H) + dir * 4 * H;
            const float gi = sigmoidf_(a[0][l] + pr[j]);
            const float gf = sigmoidf_(a[1][l] + pr[H + j]);
            const float gg = tanhf(a[2][l] + pr[2 * H + j]);
            const float go = sigmoidf_(a[3][l] + pr[3 * H + j]);
            c[l] = gf * c[l] + gi * gg;
            const float hv = go * tanhf(c[l]);
            act_store(out + row * (planes * 2 * H), 2 * H, dir * H + j, hv, fmt);
            s_h[l * H + j] = round_fp16 ? __half2float(__float2half_rn(hv)) : hv;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ igemm (ref)
// one thread per (pooled output pixel, cout)
__global__ void igemm_ref_kernel(const IgemmParams p, const __half* __restrict__ in, const __half* __restrict__ wp) {
    const int Hp = p.h_out / p.pool_h, Wp = p.w_out / p.pool_w;
    const long total = static_cast<long>(p.n_img) * Hp * Wp * p.cout;
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int co = static_cast<int>(idx % p.cout);
    long pix = idx / p.cout;
    const int wq = static_cast<int>(pix % Wp);
    const int hq = static_cast<int>((pix / Wp) % Hp);
    const int img = static_cast<int>(pix / (static_cast<long>(Wp) * Hp));
    const int taps = p.kh * p.kw;
    const int planes = p.npass == 1 ? 1 : 2;
    const int cstride = planes * p.cin;
    float result = -INFINITY;
    for (int ph = 0; ph < p.pool_h; ++ph)
        for (int pw = 0; pw < p.pool_w; ++pw) {
            const int ho = hq * p.pool_h + ph, wo = wq * p.pool_w + pw;
            float acc = 0.f;
            for (int pass = 0; pass < p.npass; ++pass) {
                const int pa = p.npass == 2 ? pass : (pass == 2 ? 1 : 0), pb = p.npass == 2 ? pass : (pass == 1 ? 1 : 0);
                for (int tap = 0; tap < taps; ++tap) {
                    const int r = tap / p.kw, s = tap % p.kw;
                    const int hi_ = ho + r - p.pad_h, wi = wo + s - p.pad_w;
                    if (hi_ < 0 || hi_ >= p.h_in || wi < 0 || wi >= p.w_in) continue;
                    const __half* a = in + ((static_cast<size_t>(img) * p.h_in + hi_) * p.w_in + wi) * cstride +
                                      pa * p.cin;
                    const __half* b = wp + (static_cast<size_t>(pb * taps + tap) * p.cout_pad + co) * p.cin;
                    if (p.npass == 2 && pass == 1) {   // e5m2 correction operands: 2 * cin bytes each
                        const uint8_t* a8 = reinterpret_cast<const uint8_t*>(a);
                        const uint8_t* b8 = reinterpret_cast<const uint8_t*>(b);
                        const int k0 = p.corr_mode == CORR_BOTH ? 0 : (p.corr_mode == CORR_WEIGHT ? p.cin : 2 * p.cin);
                        for (int k = k0; k < 2 * p.cin; ++k) acc = fmaf(e5m2_to_f32(a8[k]), e5m2_to_f32(b8[k]), acc);
                    } else {
                        for (int k = 0; k < p.cin; ++k) acc = fmaf(__half2float(a[k]), __half2float(b[k]), acc);
                    }
                }
            }
            float v = fmaf(acc, p.acc_scale, p.bias ? p.bias[co] : 0.f);
            if (p.epi != EPI_RES_F32) v = act_fn(v, p.act, p.act_slope);
            result = fmaxf(result, v);
        }
    const size_t opix = (static_cast<size_t>(img) * Hp + hq) * Wp + wq;
    if (p.epi == EPI_ACT_F16) {
        if (p.post_scale) result = result * p.post_scale[co] + p.post_shift[co];
        act_store(p.out_h + opix * p.out_cstride, p.cout, co, result, p.out_fmt);
    } else {
        if (p.epi == EPI_RES_F32) result += p.residual[opix * p.cout + co];
        p.out_f32[opix * p.cout + co] = result;
    }
}

// ------------------------------------------------------------------------------------------------ misc
__global__ void upsample_nchw_kernel(const float* __restrict__ in, int n, int h, int w, int c, int f, float* out) {
    const long total = static_cast<long>(n) * c * h * f * w * f;
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int W = w * f, Hh = h * f;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % Hh);
    const int ch = static_cast<int>((idx / (static_cast<long>(W) * Hh)) % c);
    const int img = static_cast<int>(idx / (static_cast<long>(W) * Hh * c));
    out[idx] = in[((static_cast<size_t>(img) * h + y / f) * w + x / f) * c + ch];
}

// one warp per row
__global__ void layernorm_kernel(const float* __restrict__ in, int rows, int d, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int pe_T, float* out_f32, __half* out_h,
                                 int fmt) {
    const int planes = act_planes(fmt);
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = in + static_cast<size_t>(row) * d;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s += x[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / d;
    float v = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float dlt = x[i] - mean;
        v += dlt * dlt;
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / d + eps);
    const int t = pe_T > 0 ? row % pe_T : 0;
    for (int i = lane; i < d; i += 32) {
        float y = (x[i] - mean) * rstd * gamma[i] + beta[i];
        if (pe_T > 0) {
            // pe[t, 2m] = sin(t * exp(2m * -ln(1e4)/d)), pe[t, 2m+1] = cos(same)   (transformer.py:323-327)
            const int m2 = i & ~1;
            const float div = expf(static_cast<float>(m2) * (-logf(10000.0f) / d));
            const float ang = static_cast<float>(t) * div;
            y += (i & 1) ? cosf(ang) : sinf(ang);
        }
        if (out_f32) out_f32[static_cast<size_t>(row) * d + i] = y;
        if (out_h) act_store(out_h + static_cast<size_t>(row) * planes * d, d, i, y, fmt);
    }
}

__global__ void h2f_kernel(const __half* __restrict__ in, long rows, int d, int fmt, float* out) {
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (idx >= rows * d) return;
    const long r = idx / d;
    const int i = static_cast<int>(idx % d);
    out[idx] = act_load(in + r * act_planes(fmt) * d, d, i, fmt);
}

// One CTA per (line, head), one warp per query row.  SMEM_KV: K and V of the head are staged in shared memory as
// fp32 (T <= ~400 frames for 64-wide heads); otherwise (very long lines) they are read in place from the L2-resident
// qkv tensor with the same arithmetic in the same order.
template <bool SMEM_KV>
__global__ void attention_kernel(const float* __restrict__ qkv, int n, int T, int D, int heads, __half* out, int fmt) {
    const int planes = act_planes(fmt);
    extern __shared__ float s_att[];
    const int dh = D / heads;
    const int line = blockIdx.x / heads, head = blockIdx.x % heads;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = qkv + static_cast<size_t>(line) * T * 3 * D;
    const float* sK;                   // row tk of the head's K at sK + tk * kv_stride
    const float* sV;
    float* sP;                         // [warps][T]
    int kv_stride;
    if (SMEM_KV) {
        float* k = s_att;              // [T][dh+1]
        float* v = k + T * (dh + 1);   // [T][dh+1]
        sP = v + T * (dh + 1);
        for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
            const int t = i / dh, e = i % dh;
            k[t * (dh + 1) + e] = base[static_cast<size_t>(t) * 3 * D + D + head * dh + e];
            v[t * (dh + 1) + e] = base[static_cast<size_t>(t) * 3 * D + 2 * D + head * dh + e];
        }
        sK = k; sV = v; kv_stride = dh + 1;
    } else {
        sP = s_att;
        sK = base + D + head * dh;
        sV = base + 2 * D + head * dh;
        kv_stride = 3 * D;
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(dh));
    float* p = sP + warp * T;
    for (int tq = warp; tq < T; tq += warps) {
        const float* q = base + static_cast<size_t>(tq) * 3 * D + head * dh;
        float m = -INFINITY;
        for (int tk = lane; tk < T; tk += 32) {
            float s = 0.f;
            for (int e = 0; e < dh; ++e) s = fmaf(q[e] * scale, sK[static_cast<size_t>(tk) * kv_stride + e], s);
            p[tk] = s;
            m = fmaxf(m, s);
        }
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int tk = lane; tk < T; tk += 32) {
            const float e = expf(p[tk] - m);
            p[tk] = e;
            sum += e;
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.f / sum;
        for (int e = lane; e < dh; e += 32) {
            float acc = 0.f;
            for (int tk = 0; tk < T; ++tk) acc = fmaf(p[tk], sV[static_cast<size_t>(tk) * kv_stride + e], acc);
            acc *= inv;
            const size_t row = static_cast<size_t>(line) * T + tq;
            act_store(out + row * planes * D, D, head * dh + e, acc, fmt);
        }
        __syncwarp();
    }
}

// Tiled self-attention for 64-wide heads: one CTA per (line, head, 64-query tile), 8 x 16 threads each owning an
// 8 x 4 block; K / V stream through shared memory in 64-key tiles with the running-maximum ("online") softmax, so the
// line length is unbounded and every value read from shared memory feeds 4 (S = q k^T) or 16 (O += P v) FMAs.  Same
// fp32 arithmetic as attention_kernel up to the order of the softmax sums (the row-per-warp kernel above spent
// 12.5 ms per encoder layer at config 3, this one is bound by the FMA pipe).
constexpr int FA_T = 64;          // queries per CTA = keys per tile = head width
constexpr int FA_LD = FA_T + 4;   // padded row (keeps float4 alignment, spreads banks)

__global__ void __launch_bounds__(128) attention_tiled_kernel(const float* __restrict__ qkv, int n, int T, int D,
                                                              int heads, __half* __restrict__ out, int fmt) {
    // 128 threads = 8 (ty) x 16 (tx); a thread owns 8 queries (ty + 8 i) x 4 keys / head dims (4 tx .. 4 tx + 3):
    // 32 FMAs per three 16-byte shared-memory reads in both products.
    const int planes = act_planes(fmt);
    extern __shared__ __align__(16) float s_fa[];
    float (*sQt)[FA_LD] = reinterpret_cast<float (*)[FA_LD]>(s_fa);                         // [d][query]  (scaled)
    float (*sKt)[FA_LD] = reinterpret_cast<float (*)[FA_LD]>(s_fa + FA_T * FA_LD);          // [d][key]
    float (*sV)[FA_LD] = reinterpret_cast<float (*)[FA_LD]>(s_fa + 2 * FA_T * FA_LD);       // [key][d]
    float (*sPt)[FA_LD] = reinterpret_cast<float (*)[FA_LD]>(s_fa + 3 * FA_T * FA_LD);      // [key][query]
    const int line = blockIdx.x / heads, head = blockIdx.x % heads;
    const int q0 = blockIdx.y * FA_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // ty in [0, 8): queries 8 ty .. 8 ty + 7
    const float* base = qkv + static_cast<size_t>(line) * T * 3 * D + head * FA_T;
    const float scale = rsqrtf(static_cast<float>(FA_T));
    for (int i = threadIdx.x; i < FA_T * FA_T; i += 128) {
        const int r = i >> 6, d = i & 63;
        sQt[d][r] = q0 + r < T ? base[static_cast<size_t>(q0 + r) * 3 * D + d] * scale : 0.f;
    }
    float m_run[8], l_run[8], o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < T; k0 += FA_T) {
        __syncthreads();                              // previous tile fully consumed (and sQt written)
        for (int i = threadIdx.x; i < FA_T * FA_T; i += 128) {
            const int r = i >> 6, d = i & 63;
            const bool ok = k0 + r < T;
            const float* row = base + static_cast<size_t>(k0 + r) * 3 * D;
            sKt[d][r] = ok ? row[D + d] : 0.f;
            sV[r][d] = ok ? row[2 * D + d] : 0.f;
        }
        __syncthreads();
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < FA_T; ++d) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sQt[d][8 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sQt[d][8 * ty + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&sKt[d][4 * tx]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
        }
        // running softmax per query row (the 16 threads of a row are 16 consecutive lanes)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + 4 * tx + j >= T) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int w = 1; w < 16; w <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
            const float m_new = fmaxf(m_run[i], mx);
            const float corr = expf(m_run[i] - m_new);       // 0 on the first tile (m_run = -inf)
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float pv = expf(s[i][j] - m_new);
                ps += pv;
                sPt[4 * tx + j][((8 * ty) ^ (8 * (tx & 7))) + i] = pv;   // 8-query blocks XOR-swizzled by key group
            }
#pragma unroll
            for (int w = 1; w < 16; w <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, w);
            l_run[i] = l_run[i] * corr + ps;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= corr;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < FA_T; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(&sV[k][4 * tx]);
            const int qb = (8 * ty) ^ (8 * ((k >> 2) & 7));
            const float4 p0 = *reinterpret_cast<const float4*>(&sPt[k][qb]);
            const float4 p1 = *reinterpret_cast<const float4*>(&sPt[k][qb + 4]);
            const float vv[4] = {v.x, v.y, v.z, v.w}, pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pp[i], vv[j], o[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int tq = q0 + 8 * ty + i;
        if (tq >= T) continue;
        const float inv = 1.f / l_run[i];
        const size_t row = static_cast<size_t>(line) * T + tq;
#pragma unroll
        for (int j = 0; j < 4; ++j) act_store(out + row * planes * D, D, head * FA_T + 4 * tx + j, o[i][j] * inv, fmt);
    }
}

}  // namespace

cudaError_t launch_conv_first(const uint8_t* in, int n, int h, int w, const float* w_t, const float* bias, int cout,
                              int act, float slope, int fmt, __half* out, cudaStream_t stream) {
    const int planes = act_planes(fmt);
    const int tiles_w = (w + CF_PX - 1) / CF_PX;
    const int grid = n * ((h + CF_ROWS - 1) / CF_ROWS) * tiles_w;
    const size_t dyn = static_cast<size_t>(CF_PX) * planes * cout * sizeof(__half);
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaFuncSetAttribute(conv_first_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(conv_first_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(conv_first_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done.mark();
    }
    switch (cout) {
        case 64: conv_first_kernel<64><<<grid, CF_PX, dyn, stream>>>(in, n, h, w, w_t, bias, act, slope, fmt, out); break;
        case 32: conv_first_kernel<32><<<grid, CF_PX, dyn, stream>>>(in, n, h, w, w_t, bias, act, slope, fmt, out); break;
        case 16: conv_first_kernel<16><<<grid, CF_PX, dyn, stream>>>(in, n, h, w, w_t, bias, act, slope, fmt, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_frame_stats(const float* scores, int n, int t, int c, int layout, int32_t* best, float* fmax,
                               float* flse, float* fprob, cudaStream_t stream) {
    const long frames = static_cast<long>(n) * t;
    if (frames == 0) return cudaSuccess;
    const long sn = static_cast<long>(t) * c;
    const long st = layout == 0 ? c : 1;
    const long sc = layout == 0 ? 1 : t;
    frame_stats_kernel<<<static_cast<unsigned>((frames + 127) / 128), 128, 0, stream>>>(scores, n, t, c, sn, st, sc,
                                                                                         best, fmax, flse, fprob);
    return cudaGetLastError();
}

cudaError_t launch_ctc_collapse(const int32_t* best, const float* fprob, int n, int t, int blank, int32_t* labels,
                                int32_t* lengths, float* confidence, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ctc_collapse_kernel<<<(n * 32 + 127) / 128, 128, 0, stream>>>(best, fprob, n, t, blank, labels, lengths,
                                                                  fprob ? confidence : nullptr);
    return cudaGetLastError();
}

cudaError_t launch_lstm_ref(const float* pre, const float* w_hh_t, int n, int T, int H, int fmt, int round_fp16,
                            __half* out, cudaStream_t stream) {
    dim3 grid((n + LR_LINES - 1) / LR_LINES, 2);
    lstm_ref_kernel<<<grid, H, LR_LINES * H * sizeof(float), stream>>>(pre, w_hh_t, n, T, H, fmt, round_fp16, out);
    return cudaGetLastError();
}

cudaError_t launch_igemm_ref(const IgemmParams& p, const __half* in, const __half* w_packed, cudaStream_t stream) {
    const long total = static_cast<long>(p.n_img) * (p.h_out / p.pool_h) * (p.w_out / p.pool_w) * p.cout;
    if (total == 0) return cudaSuccess;
    igemm_ref_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p, in, w_packed);
    return cudaGetLastError();
}

cudaError_t launch_upsample_nchw(const float* in, int n, int h, int w, int c, int f, float* out, cudaStream_t stream) {
    const long total = static_cast<long>(n) * c * h * f * w * f;
    upsample_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, n, h, w, c, f, out);
    return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* in, int rows, int d, const float* gamma, const float* beta, float eps,
                             int pe_T, float* out_f32, __half* out_h, int fmt, cudaStream_t stream) {
    layernorm_kernel<<<(rows * 32 + 255) / 256, 256, 0, stream>>>(in, rows, d, gamma, beta, eps, pe_T, out_f32, out_h,
                                                                  fmt);
    return cudaGetLastError();
}

cudaError_t launch_h2f(const __half* in, int rows, int d, int fmt, float* out, cudaStream_t stream) {
    const long total = static_cast<long>(rows) * d;
    h2f_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, rows, d, fmt, out);
    return cudaGetLastError();
}

cudaError_t launch_attention(const float* qkv, int n, int T, int D, int heads, __half* out, int fmt,
                             cudaStream_t stream) {
    const int dh = D / heads;
    if (dh == FA_T && heads * dh == D) {
        const size_t smem = 4 * static_cast<size_t>(FA_T) * FA_LD * sizeof(float);
        static PerDeviceOnce fa_attr;
        if (fa_attr.pending()) {
            cudaError_t e = cudaFuncSetAttribute(attention_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
            fa_attr.mark();
        }
        attention_tiled_kernel<<<dim3(n * heads, (T + FA_T - 1) / FA_T), 128, smem, stream>>>(qkv, n, T, D, heads, out, fmt);
        return cudaGetLastError();
    }
    const int warps = 8;
    const size_t smem_kv = (2 * static_cast<size_t>(T) * (dh + 1) + static_cast<size_t>(warps) * T) * sizeof(float);
    const size_t smem_p = static_cast<size_t>(warps) * T * sizeof(float);
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    if (smem_kv <= 220 * 1024)
        attention_kernel<true><<<n * heads, warps * 32, smem_kv, stream>>>(qkv, n, T, D, heads, out, fmt);
    else if (smem_p <= 220 * 1024)
        attention_kernel<false><<<n * heads, warps * 32, smem_p, stream>>>(qkv, n, T, D, heads, out, fmt);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
