// Self-attention of the Transformer-encoder variant on tcgen05 (sm_100a).
//
// Replaces the multi-head attention inside nn.TransformerEncoderLayer (pero_ocr/ocr_engine/transformer.py:366-385:
// LineSelfAttentionEncoder -> nn.TransformerEncoder; no mask, dropout 0) for 64-wide heads and lines of up to
// kMaxKeys frames (T = 336 at BASELINE config 3); longer lines and other head widths keep the CUDA-core kernels of
// kernels.cu.  One CTA per (line, head):
//   * all 256 threads bring K and V of the head from the fp32 QKV rows (igemm_tc's EPI_F32 output) into shared memory
//     as fp16 hi + lo planes in the tcgen05 K-major 128-byte-swizzled layout -- K as [key][64] (operand B of S = Q K^T),
//     V transposed as [64][key] in chunks of 64 keys (operand B of O = P V);
//   * per tile of 128 queries: Q (pre-scaled by 1/8, exact) is staged the same way; one thread issues
//     S = Qh Kh^T + Qh Kl^T + Ql Kh^T (3 x 4 tcgen05.mma of K = 16 per chunk of <= 128 keys) into TMEM columns [0, Tp);
//   * warps 0-3 (one query row per thread = one TMEM lane) make two sweeps over their row: maximum, then
//     p = exp(s - max) with the row sum, and write P back IN PLACE as packed fp16 pairs -- the 32 fp32 columns of a
//     chunk become 16 columns of P_hi and 16 of P_lo (tcgen05.st) -- so P never touches shared memory;
//   * O = Ph Vh + Ph Vl + Pl Vh with P as the TMEM-resident A operand (TS-mode MMAs, K = 16 keys each) into 64 more
//     TMEM columns; the same four warps divide by the row sum and write the head's 64 channels of the activation
//     record (actfmt.cuh) that the out-projection GEMM reads.
// The operand splits keep the products fp32-grade (hi * hi + hi * lo + lo * hi: ~2^-22), which the 1e-3 logit bar of
// the Transformer variant needs (|logit| <= 8.8; single-pass fp16 attention alone would spend the budget).
// Phases are separated by __syncthreads and two mbarriers for MMA completion: the work of a (line, head) is tiny
// (29 MFLOP x 3 passes), the kernel is bound by staging Q / K / V (258 KB of fp32 per CTA) and by the exponentials.
#include "once.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace {

constexpr int kDh = 64;            // head width
constexpr int kMaxKeys = 384;      // frames per line this kernel takes (6 chunks of 64 keys)
constexpr int kQTile = 128;        // queries per tile = TMEM lanes
constexpr int kThreads = 256;
constexpr int kQBytes = kQTile * 128;            // one plane of a Q tile
constexpr int kKBytes = kMaxKeys * 128;          // one plane of K
constexpr int kVChunk = kDh * 128;               // one plane of one 64-key chunk of V^T
constexpr int kVBytes = (kMaxKeys / 64) * kVChunk;
constexpr int kOCol = 448;                       // TMEM columns [448, 512): O accumulator
constexpr int kSmemBytes = 2 * kQBytes + 2 * kKBytes + 2 * kVBytes + 32 + 2 * kQTile * 4 + 1024;

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile (rows of 128 B, 8-row groups)
__device__ __forceinline__ uint32_t sw128(int r, int c) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// 8 consecutive fp32 -> 8 fp16 hi (one uint4) and 8 fp16 lo (one uint4)
__device__ __forceinline__ void split8(const float4& a, const float4& b, float scale, uint4& hi, uint4& lo) {
    const float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&h2);
        l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const float* __restrict__ qkv, int T, int D, int heads, __half* __restrict__ out, int fmt) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // [plane][128 rows x 128 B]
    uint8_t* sK = sQ + 2 * kQBytes;           // [plane][384 rows x 128 B]
    uint8_t* sV = sK + 2 * kKBytes;           // [plane][6 chunks][64 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kVBytes);   // [0] S done, [1] O done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_red = reinterpret_cast<float*>(bars + 4);   // [side 0 | side 1][128 rows]: partial row maxima, then sums

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int line = blockIdx.x / heads, head = blockIdx.x - line * heads;
    const int Tp = (T + 31) & ~31;            // keys padded to whole 32-column softmax chunks (zero K rows / V columns)
    const float* base = qkv + static_cast<size_t>(line) * T * 3 * D + head * kDh;
    const int planes = act_planes(fmt);

    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 4) ptx::tmem_alloc<512>(tmem_slot);

    // Staging is bound by global-load latency, not bandwidth (ncu: half of all stall samples sat behind the two loads of
    // a piece when every piece was loaded, converted and stored in turn): the loads of kBatch pieces are issued
    // together, then converted.
    constexpr int kBatch = 4;
    // ---- K: [key][64] hi | lo.  A piece = (key, 8 channels): 8 lanes read one 256-byte row, a warp 4 rows
    for (int i0 = tid; i0 < Tp * 8; i0 += kBatch * kThreads) {
        float4 a[kBatch], b[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int i = i0 + k * kThreads, r = i >> 3, c = i & 7;
            a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < T) {
                const float4* src = reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * 3 * D + D + c * 8);
                a[k] = __ldg(src);
                b[k] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int i = i0 + k * kThreads, r = i >> 3, c = i & 7;
            if (i < Tp * 8) {
                uint4 hi, lo;
                split8(a[k], b[k], 1.f, hi, lo);
                *reinterpret_cast<uint4*>(sK + sw128(r, c)) = hi;
                *reinterpret_cast<uint4*>(sK + kKBytes + sw128(r, c)) = lo;
            }
        }
    }
    // ---- V^T: [d][key] hi | lo in chunks of 64 keys.  A piece = (key, 8 channels) again, written as 8 two-byte
    // elements down a column; the 32 lanes of a warp take 32 consecutive keys of the same channel group, so one store
    // instruction fills 64 contiguous (swizzled) bytes of a row
    for (int i0 = tid; i0 < Tp * 8; i0 += kBatch * kThreads) {
        float4 a[kBatch], b[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int i = i0 + k * kThreads;
            const int key = (i & 31) + ((i >> 8) << 5), c = (i >> 5) & 7;   // i = (key / 32) * 256 + c * 32 + key % 32
            a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (key < T) {
                const float4* src = reinterpret_cast<const float4*>(base + static_cast<size_t>(key) * 3 * D + 2 * D + c * 8);
                a[k] = __ldg(src);
                b[k] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int i = i0 + k * kThreads;
            if (i >= Tp * 8) break;
            const int key = (i & 31) + ((i >> 8) << 5), c = (i >> 5) & 7;
            const float v[8] = {a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y, b[k].z, b[k].w};
            uint8_t* chunk = sV + (key >> 6) * kVChunk;
            const int kk = key & 63;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int d = c * 8 + e;
                const __half hi = __float2half_rn(v[e]);
                const __half lo = __float2half_rn(v[e] - __half2float(hi));
                const uint32_t off = sw128(d, kk >> 3) + (kk & 7) * 2;
                *reinterpret_cast<__half*>(chunk + off) = hi;
                *reinterpret_cast<__half*>(chunk + kVBytes + off) = lo;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();   // the CTA owns all 512 columns: literal column addresses below

    const uint32_t sQ_u = ptx::smem_u32(sQ), sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);
    const uint64_t desc_hi = ptx::smem_desc_sw128(0);
    uint32_t ph_s = 0, ph_o = 0;

    // Q tile (scaled by 1/sqrt(64) = 1/8: exact), hi | lo, staged by the threads [t0, t0 + nt) of the CTA
    auto stage_q = [&](int q0, int t0, int nt) {
        for (int i0 = tid - t0; i0 < kQTile * 8; i0 += kBatch * nt) {
            float4 a[kBatch], b[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int i = i0 + k * nt, r = i >> 3, c = i & 7;
                a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < kQTile * 8 && q0 + r < T) {
                    const float4* src = reinterpret_cast<const float4*>(base + static_cast<size_t>(q0 + r) * 3 * D + c * 8);
                    a[k] = __ldg(src);
                    b[k] = __ldg(src + 1);
                }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int i = i0 + k * nt, r = i >> 3, c = i & 7;
                if (i < kQTile * 8) {
                    uint4 hi, lo;
                    split8(a[k], b[k], 0.125f, hi, lo);
                    *reinterpret_cast<uint4*>(sQ + sw128(r, c)) = hi;
                    *reinterpret_cast<uint4*>(sQ + kQBytes + sw128(r, c)) = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // st.shared -> tensor-core (async proxy) reads
    };
    stage_q(0, 0, kThreads);
    __syncthreads();

    for (int q0 = 0; q0 < T; q0 += kQTile) {
        // ---- S = Q K^T into TMEM columns [0, Tp): chunks of <= 128 keys, three operand-split passes each
        if (warp == 4) {
            if (ptx::elect_one()) {
                ptx::tc_fence_after();
                for (int k0 = 0; k0 < Tp; k0 += 128) {
                    const int n = min(128, Tp - k0);
                    const uint32_t idesc = ptx::idesc_f16_f32(128, n);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t qa = sQ_u + (pass == 2 ? kQBytes : 0);
                        const uint32_t kb = sK_u + (pass == 1 ? kKBytes : 0) + k0 * 128;
                        const uint64_t a_desc = desc_hi + (qa >> 4), b_desc = desc_hi + (kb >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::mma_f16_ss(static_cast<uint32_t>(k0), a_desc + 2 * k, b_desc + 2 * k, idesc, (pass | k) != 0);
                    }
                }
                ptx::mma_commit(&bars[0]);
            }
            __syncwarp();
        }
        // ---- softmax: all 8 warps.  Warps w and w + 4 share TMEM lane quarter w (one query row per lane) and split
        // the row's 32-column chunks between them (even / odd); row maximum and row sum meet in shared memory.  P is
        // written back in place: the 32 fp32 columns of a chunk become 16 columns of P_hi pairs + 16 of P_lo pairs.
        const int lq = warp & 3, side = warp >> 2;
        const int rowi = lq * 32 + lane;
        const uint32_t row_addr = static_cast<uint32_t>(lq * 32) << 16;
        ptx::mbar_wait(&bars[0], ph_s);
        ptx::tc_fence_after();
        float mx = -INFINITY;
        for (int c0 = side * 32; c0 < Tp; c0 += 64) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(row_addr + c0, r);
            ptx::tmem_ld_wait();
            if (c0 + 32 <= T) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < T) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
        }
        s_red[side * kQTile + rowi] = mx;
        __syncthreads();
        mx = fmaxf(s_red[rowi], s_red[kQTile + rowi]);
        __syncthreads();                 // the two slots per row are reused for the row sums below
        float sum = 0.f;
        for (int c0 = side * 32; c0 < Tp; c0 += 64) {
            uint32_t r[32], w[32];
            ptx::tmem_ld_32x32b_x32(row_addr + c0, r);
            ptx::tmem_ld_wait();
            const bool full = c0 + 32 <= T;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float p0 = __expf(__uint_as_float(r[j]) - mx), p1 = __expf(__uint_as_float(r[j + 1]) - mx);
                if (!full) {
                    if (c0 + j >= T) p0 = 0.f;
                    if (c0 + j + 1 >= T) p1 = 0.f;
                }
                sum += p0 + p1;
                const __half2 h2 = __floats2half2_rn(p0, p1);
                const float2 f = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(p0 - f.x, p1 - f.y);
                w[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);          // columns c0 .. c0+15: P_hi pairs
                w[16 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);   // columns c0+16 .. c0+31: P_lo pairs
            }
            ptx::tmem_st_32x32b_x32(row_addr + c0, w);
        }
        ptx::tmem_st_wait();
        s_red[side * kQTile + rowi] = sum;
        // the Q tile is dead once S is complete: everybody stages the next one before the barrier that releases P
        if (q0 + kQTile < T) stage_q(q0 + kQTile, 0, kThreads);
        ptx::tc_fence_before();
        __syncthreads();
        const float inv_sum = 1.f / (s_red[rowi] + s_red[kQTile + rowi]);

        // ---- O = P V: A = P from TMEM (8 packed columns per K = 16 keys), B = V^T chunk rows from shared memory
        if (warp == 4) {
            if (ptx::elect_one()) {
                ptx::tc_fence_after();
                constexpr uint32_t idesc_o = ptx::idesc_f16_f32(128, kDh);
                const int ksteps = Tp >> 4;
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t p_off = pass == 2 ? 16u : 0u;                  // P_lo columns of a 32-column chunk
                    const uint32_t vb = sV_u + (pass == 1 ? kVBytes : 0);
                    for (int j = 0; j < ksteps; ++j) {
                        const uint32_t a_tmem = static_cast<uint32_t>((j >> 1) * 32 + (j & 1) * 8) + p_off;
                        const uint64_t b_desc = desc_hi + ((vb + (j >> 2) * kVChunk + (j & 3) * 32) >> 4);
                        ptx::mma_f16_ts(kOCol, a_tmem, b_desc, idesc_o, (pass | j) != 0);
                    }
                }
                ptx::mma_commit(&bars[1]);
            }
            __syncwarp();
        }
        if (warp < 4) {
            // ---- O / sum -> the head's 64 channels of the activation record of this query (warps 0-3: lane quarter = warp)
            ptx::mbar_wait(&bars[1], ph_o);
            ptx::tc_fence_after();
            const int tq = q0 + warp * 32 + lane;
            const uint32_t o_addr = (static_cast<uint32_t>(warp * 32) << 16) + kOCol;
            __half* rec = out + (static_cast<size_t>(line) * T + (tq < T ? tq : 0)) * planes * D;
#pragma unroll
            for (int c0 = 0; c0 < kDh; c0 += 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(o_addr + c0, r);
                ptx::tmem_ld_wait();
                if (tq < T) {
                    uint32_t ph[16], pl[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float v0 = __uint_as_float(r[j]) * inv_sum, v1 = __uint_as_float(r[j + 1]) * inv_sum;
                        const float v2 = __uint_as_float(r[j + 2]) * inv_sum, v3 = __uint_as_float(r[j + 3]) * inv_sum;
                        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                        ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
                        ph[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        if (fmt != ACT_F16) {
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            if (fmt == ACT_F16_HILO) {
                                const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
                                const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
                                pl[j >> 1] = *reinterpret_cast<const uint32_t*>(&l01);
                                pl[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                            } else {
                                pl[j >> 2] = pack_e5m2x4((v0 - f01.x) * kF8Scale, (v1 - f01.y) * kF8Scale,
                                                         (v2 - f23.x) * kF8Scale, (v3 - f23.y) * kF8Scale);
                                pl[8 + (j >> 2)] = pack_e5m2x4(f01.x, f01.y, f23.x, f23.y);
                            }
                        }
                    }
                    const int n0 = head * kDh + c0;
                    uint4* dst = reinterpret_cast<uint4*>(rec + n0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                    if (fmt == ACT_F16_HILO) {
                        uint4* dl = reinterpret_cast<uint4*>(rec + D + n0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dl[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
                    } else if (fmt == ACT_F16_F8) {
                        uint8_t* b = reinterpret_cast<uint8_t*>(rec + D);
                        uint4* dlo = reinterpret_cast<uint4*>(b + n0);
                        uint4* dh8 = reinterpret_cast<uint4*>(b + D + n0);
                        dlo[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        dlo[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                        dh8[0] = make_uint4(pl[8], pl[9], pl[10], pl[11]);
                        dh8[1] = make_uint4(pl[12], pl[13], pl[14], pl[15]);
                    }
                }
            }
            ptx::tc_fence_before();
        }
        ph_s ^= 1;
        ph_o ^= 1;
        __syncthreads();   // O is free, the next Q tile is in place
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(0u);
    }
}

}  // namespace

bool attention_tc_supported(int T, int D, int heads) {
    return heads > 0 && D == heads * kDh && T >= 1 && T <= kMaxKeys && (D % 8) == 0;
}

cudaError_t launch_attention_tc(const float* qkv, int n, int T, int D, int heads, __half* out, int fmt,
                                cudaStream_t stream) {
    if (!attention_tc_supported(T, D, heads)) return cudaErrorInvalidValue;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    attention_tc_kernel<<<n * heads, kThreads, kSmemBytes, stream>>>(qkv, T, D, heads, out, fmt);
    return cudaGetLastError();
}
