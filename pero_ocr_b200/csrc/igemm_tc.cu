// tcgen05 implicit-GEMM convolution / GEMM for sm_100a.
//
// Replaces the cuDNN / cuBLAS calls hidden inside the reference's TorchScript recogniser blob
// (pero_ocr/ocr_engine/pytorch_ocr_engine.py:64-69; layer list = pero_ocr/ocr_engine/transformer.py:75-148,335-363).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0     draws tiles (global counter, tilesched.cuh) and is the TMA producer: per (pass, tap, 64-channel chunk)
//              four 4-D box loads of the NHWC activation (shifted by the tap; out-of-bounds = conv zero padding,
//              filled by TMA) + one 2-D load of the weights
//   warp 1     allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16; kind::f16 for the fp16 pass, kind::f8f6f4 with
//              K=32 for the e5m2 correction pass) into one of two TMEM accumulators
//   warps 2-9  epilogue, two warps per TMEM lane quarter (each takes half of the BN columns): tcgen05.ld (one
//              output pixel per thread, 32 channels per load), bias + activation, max-pool by warp shuffles,
//              BatchNorm affine, record planes, 256-bit NHWC stores -- or the fused CTC epilogue (per-frame argmax /
//              max / logsumexp; logits optional)
// A tile is 4 segments of 32 output pixels (th x 32/th); segment q feeds TMEM lanes [32q, 32q+32), so a 2x2 or
// 2x1 max-pool never leaves the warp.
#include "once.cuh"
#include "igemm.cuh"
#include "ptx.cuh"
#include "tilesched.cuh"

namespace {

constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 fp16

template <int BN>
struct Cfg {
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (192 * 1024) / kStageBytes;
    static constexpr int kTmemCols = 512;  // whole TMEM (1 CTA/SM): base address is 0 => warp-uniform operands
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers + tile ring*/;
};

struct SegCoord {
    int img, h0, w0;
};

__device__ __forceinline__ SegCoord seg_coord(const IgemmParams& p, int seg) {
    SegCoord c;
    if (seg >= p.total_segs) {  // past the end: every load is out of bounds (zeros), every store masked
        c.img = p.n_img;
        c.h0 = 0;
        c.w0 = 0;
        return c;
    }
    const int per_img = p.row_groups * p.segs_per_row;
    c.img = seg / per_img;
    const int r = seg - c.img * per_img;
    const int rg = r / p.segs_per_row;
    c.h0 = rg * p.th;
    c.w0 = (r - rg * p.segs_per_row) * (32 / p.th);
    return c;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : slope * v;
    return v;
}

template <int BN, int FMT>
__global__ void __launch_bounds__(kThreads, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const IgemmParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full = bars;                      // [kStages]
    uint64_t* empty = bars + C::kStages;        // [kStages]
    uint64_t* tfull = bars + 2 * C::kStages;    // [2]
    uint64_t* tempty = bars + 2 * C::kStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
    // dynamic tile scheduler (IgemmParams::tile_counter): the producer warp draws tile indices from a global counter
    // and hands them to the MMA and epilogue warps through a 4-deep ring -- a CTA that starts late (its SM was busy
    // with another stream's kernel) simply draws fewer tiles, instead of owning a fixed 1/gridDim share
    uint64_t* tq_full = bars + 2 * C::kStages + 5;   // [4]
    uint64_t* tq_empty = tq_full + 4;                // [4]
    int* tq_tile = reinterpret_cast<int*>(tq_empty + 4);   // [4]
    // per-channel epilogue vectors (bias | post_scale | post_shift), cout_pad floats each, 16-byte aligned
    float* s_vec = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int taps = p.kh * p.kw;
    const int kchunks = p.cin >> 6;
    // FMT (compile time) = operand format of the input AND record format of an EPI_ACT_F16 output
    constexpr int npass = FMT == ACT_F16 ? 1 : (FMT == ACT_F16_HILO ? 3 : 2);
    // first 64-slot chunk of the e5m2 pass (IgemmParams::corr_mode): 0 = both corrections, kchunks/2 = weight side only
    const int f8_kc0 = (npass != 2 || p.corr_mode == CORR_BOTH) ? 0 : (p.corr_mode == CORR_WEIGHT ? (kchunks >> 1) : kchunks);
    const int k_iters = (npass == 2) ? taps * (2 * kchunks - f8_kc0) : npass * taps * kchunks;
    const int total_tiles = p.m_tiles * p.tiles_n;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < C::kStages; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], 8);  // one arrive per epilogue warp
        }
        for (int i = 0; i < 4; ++i) {
            ptx::mbar_init(&tq_full[i], 1);
            ptx::mbar_init(&tq_empty[i], 9);   // MMA warp + 8 epilogue warps
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
        s_vec[i] = p.bias ? p.bias[i] : 0.f;
        s_vec[p.cout_pad + i] = p.post_scale ? p.post_scale[i] : 1.f;
        s_vec[2 * p.cout_pad + i] = p.post_shift ? p.post_shift[i] : 0.f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (*tmem_slot != 0) __trap();
    constexpr uint32_t tmem_base = 0;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // The whole warp runs the (warp-uniform) loop nest so that coordinates live in uniform registers; one
        // elected lane issues.
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        TileFeed feed(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
        for (int tile = feed.next(lane); tile < total_tiles; tile = feed.next(lane)) {
            const int mt = tile / p.tiles_n;
            const int nt = tile - mt * p.tiles_n;
            SegCoord sc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) sc[q] = seg_coord(p, mt * 4 + q);
            for (int pass = 0; pass < npass; ++pass) {
                // fp16x3: (A hi, B hi), (A hi, B lo), (A lo, B hi); fp16+fp8: (A hi, B hi*2^11), (A e5m2, B e5m2)
                const int pa = (npass == 2) ? pass : ((pass == 2) ? 1 : 0);  // activation plane
                const int pb = (npass == 2) ? pass : ((pass == 1) ? 1 : 0);  // weight plane
                for (int tap = 0; tap < taps; ++tap) {
                    const int r = tap / p.kw;
                    const int s = tap - r * p.kw;
                    const int brow = (pb * taps + tap) * p.cout_pad + nt * BN;
                    for (int kc = (npass == 2 && pass == 1) ? f8_kc0 : 0; kc < kchunks; ++kc) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * C::kStageBytes;
                        uint8_t* sB = sA + kABytes;
                        ptx::mbar_expect_tx_pred(&full[stage], C::kStageBytes, leader);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            ptx::tma_load_4d_pred(sA + q * 4096, &tmA, &full[stage], pa * p.cin + kc * 64,
                                                  sc[q].w0 + s - p.pad_w, sc[q].h0 + r - p.pad_h, sc[q].img, leader);
                        ptx::tma_load_2d_pred(sB, &tmB, &full[stage], kc * 64, brow, leader);
                        if (++stage == C::kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: one elected thread runs the role
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::idesc_f16_f32(128, BN);
            constexpr uint32_t idesc8 = ptx::idesc_e5m2_f32(128, BN);
            const uint64_t desc_hi = ptx::smem_desc_sw128(0);
            const uint32_t smem_base = ptx::smem_u32(smem);
            // iterations [0, k16_iters) are fp16 passes; the rest (npass == 2 only) is the e5m2 correction pass
            const int k16_iters = (npass == 2) ? taps * kchunks : k_iters;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase0 = 0, acc_phase1 = 0;
            TileTake take(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
            for (int tile = take.next(); tile < total_tiles; tile = take.next()) {
                ptx::mbar_wait(&tempty[acc], (acc ? acc_phase1 : acc_phase0) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int it = 0; it < k16_iters; ++it) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * C::kStageBytes;
                    const uint64_t a_desc = desc_hi + (a_addr >> 4);
                    const uint64_t b_desc = desc_hi + ((a_addr + kABytes) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 4 x K=16 inside the 128-byte swizzle atom: +32 B per step
                        ptx::mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it | k) != 0);
                    ptx::mma_commit(&empty[stage]);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                for (int it = k16_iters; it < k_iters; ++it) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * C::kStageBytes;
                    const uint64_t a_desc = desc_hi + (a_addr >> 4);
                    const uint64_t b_desc = desc_hi + ((a_addr + kABytes) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 4 x K=32 e5m2 inside the 128-byte swizzle atom: +32 B per step
                        ptx::mma_f8_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc8, 1u);
                    ptx::mma_commit(&empty[stage]);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit(&tfull[acc]);
                if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
                acc ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;     // which half of the tile's columns this warp handles (fp16 / f32 paths)
        const int segw = 32 / p.th;
        const int dh = lane / segw;
        const int dw = lane - dh * segw;
        const float* s_bias = s_vec;
        const float* s_scale = s_vec + p.cout_pad;
        const float* s_shift = s_vec + 2 * p.cout_pad;
        const bool has_affine = p.post_scale != nullptr;
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        TileTake take(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
        for (int tile = take.next_warp(lane); tile < total_tiles; tile = take.next_warp(lane)) {
            const int mt = tile / p.tiles_n;
            const int nt = tile - mt * p.tiles_n;
            const SegCoord sc = seg_coord(p, mt * 4 + q);
            const int ho = sc.h0 + dh, wo = sc.w0 + dw;
            const bool valid = sc.img < p.n_img && ho < p.h_out && wo < p.w_out;
            ptx::mbar_wait(&tfull[acc], acc_phase[acc]);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
            const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;

            if (p.epi == EPI_ACT_F16) {
                const bool writer = valid && (dh % p.pool_h == 0) && (dw % p.pool_w == 0);
                const int hp = ho / p.pool_h, wp = wo / p.pool_w;
                const int Hp = p.h_out / p.pool_h, Wp = p.w_out / p.pool_w;
                __half* orow = p.out_h + (static_cast<size_t>(sc.img) * Hp * Wp + static_cast<size_t>(hp) * Wp + wp) *
                                             p.out_cstride;
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    const int n0 = nt * BN + c0;
                    if (n0 >= p.cout) break;  // warp-uniform
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(t_addr + c0, r);
                    ptx::tmem_ld_wait();
                    // max-pool first (bias + monotone activation commute with max), then bias/act once
                    if (p.pool_w == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]),
                                                         __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), 1)));
                    }
                    if (p.pool_h == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]),
                                                         __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), segw)));
                    }
                    if (writer) {
                        if (p.dbg & 4)                  // A/B: 16-byte stores
                            epi_store32(r, n0, p.acc_scale, s_bias, s_scale, s_shift, has_affine, p.act, p.act_slope, orow, p.cout, FMT);
                        else
                            epi_store32_v8(r, n0, p.acc_scale, s_bias, s_scale, s_shift, has_affine, p.act, p.act_slope, orow, p.cout, FMT, p.out_skip_lo != 0);
                    }
                }
            } else if (p.epi == EPI_CTC) {
              if (half == 0) {
                // one frame per thread: running first-max argmax (torch.argmax: NaN counts as maximal),
                // online logsumexp
                const size_t pix = (static_cast<size_t>(sc.img) * p.h_out + ho) * p.w_out + wo;
                float best_v = -INFINITY, run_m = -INFINITY, run_s = 0.f;
                int best_i = 0;
                bool best_nan = false;
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    if (c0 >= p.cout) break;
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(t_addr + c0, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c0 + j;
                        if (n < p.cout) {
                            const float v = fmaf(__uint_as_float(r[j]), p.acc_scale, s_bias[n]);
                            r[j] = __float_as_uint(v);
                            if (!best_nan) {
                                if (v != v) {
                                    best_nan = true;
                                    best_i = n;
                                    best_v = v;
                                } else if (n == 0 || v > best_v) {
                                    best_v = v;
                                    best_i = n;
                                }
                            }
                            const float m2 = fmaxf(run_m, v);
                            run_s = run_s * __expf(run_m - m2) + __expf(v - m2);
                            run_m = m2;
                        }
                    }
                    if (valid && p.out_f32) {
                        float* dst = p.out_f32 + pix * p.cout + c0;
                        if ((p.cout & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (c0 + j < p.cout)
                                    *reinterpret_cast<float4*>(dst + j) =
                                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < p.cout) dst[j] = __uint_as_float(r[j]);
                        }
                    }
                }
                if (valid) {
                    p.best[pix] = best_i;
                    if (p.fmax) p.fmax[pix] = best_v;
                    if (p.flse) p.flse[pix] = run_m + __logf(run_s);
                }
                if (p.fprob) {
                    // second sweep over the accumulator: softmax mass of the classes the reference keeps when it
                    // sparsifies (p >= 1e-4, line_ocr_engine.py:168-171); dropped ones re-enter as logit -80
                    // (core/layout.py:65-68) before the confidence log-softmax (page_parser.py:486-490)
                    float kept = 0.f;
                    int dropped = 0;
                    const float thr = 1e-4f * run_s;
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        if (c0 >= p.cout) break;
                        uint32_t r[32];
                        ptx::tmem_ld_32x32b_x32(t_addr + c0, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = c0 + j;
                            if (n < p.cout) {
                                const float v = fmaf(__uint_as_float(r[j]), p.acc_scale, s_bias[n]);
                                const float e = __expf(v - run_m);
                                if (e < thr || v == 0.f) ++dropped;
                                else kept += e;
                            }
                        }
                    }
                    kept += dropped * __expf(-80.f - run_m);
                    if (valid) p.fprob[pix] = __expf(best_v - run_m) / kept;
                }
              }
            } else {  // EPI_F32 / EPI_RES_F32
                const size_t pix = (static_cast<size_t>(sc.img) * p.h_out + ho) * p.w_out + wo;
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    const int n0 = nt * BN + c0;
                    if (n0 >= p.cout) break;
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(t_addr + c0, r);
                    ptx::tmem_ld_wait();
                    if (valid) {
                        float* dst = p.out_f32 + pix * p.cout + n0;
                        const float* res = (p.epi == EPI_RES_F32) ? p.residual + pix * p.cout + n0 : nullptr;
                        if (!(p.dbg & 4) && (p.cout & 31) == 0 && !res) {
                            // 256-bit stores: whole sectors of this thread's 128-byte piece of the fp32 row
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint32_t w[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    w[e] = __float_as_uint(apply_act(
                                        fmaf(__uint_as_float(r[j + e]), p.acc_scale, s_bias[n0 + j + e]), p.act, p.act_slope));
                                st_global_v8(dst + j, w);
                            }
                        } else if ((p.cout & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (n0 + j < p.cout) {
                                    float4 v;
                                    {
                                        const float4 b = *reinterpret_cast<const float4*>(s_bias + n0 + j);
                                        v.x = fmaf(__uint_as_float(r[j]), p.acc_scale, b.x);
                                        v.y = fmaf(__uint_as_float(r[j + 1]), p.acc_scale, b.y);
                                        v.z = fmaf(__uint_as_float(r[j + 2]), p.acc_scale, b.z);
                                        v.w = fmaf(__uint_as_float(r[j + 3]), p.acc_scale, b.w);
                                    }
                                    v.x = apply_act(v.x, p.act, p.act_slope); v.y = apply_act(v.y, p.act, p.act_slope);
                                    v.z = apply_act(v.z, p.act, p.act_slope); v.w = apply_act(v.w, p.act, p.act_slope);
                                    if (res) {
                                        const float4 e = __ldg(reinterpret_cast<const float4*>(res + j));
                                        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
                                    }
                                    *reinterpret_cast<float4*>(dst + j) = v;
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (n0 + j < p.cout) {
                                    float v = fmaf(__uint_as_float(r[j]), p.acc_scale, s_bias[n0 + j]);
                                    v = apply_act(v, p.act, p.act_slope);
                                    if (res) v += __ldg(res + j);
                                    dst[j] = v;
                                }
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            acc_phase[acc] ^= 1;
            acc ^= 1;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

template <int BN, int FMT>
cudaError_t launch_bn(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                      cudaStream_t stream) {
    using C = Cfg<BN>;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    const size_t smem_bytes = C::kSmemBytes + 3 * static_cast<size_t>(p.cout_pad) * sizeof(float);
    if (smem_bytes > 227 * 1024) return cudaErrorInvalidValue;
    const int total_tiles = p.m_tiles * p.tiles_n;
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    igemm_tc_kernel<BN, FMT><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, p);
    return cudaGetLastError();
}

template <int BN>
cudaError_t launch_fmt(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                       cudaStream_t stream) {
    switch (p.npass) {
        case 1: return launch_bn<BN, ACT_F16>(p, tmA, tmB, num_sms, stream);
        case 3: return launch_bn<BN, ACT_F16_HILO>(p, tmA, tmB, num_sms, stream);
        case 2: return launch_bn<BN, ACT_F16_F8>(p, tmA, tmB, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t launch_igemm_tc(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, int num_sms,
                            cudaStream_t stream) {
    switch (bn) {
        case 64: return launch_fmt<64>(p, tmA, tmB, num_sms, stream);
        case 128: return launch_fmt<128>(p, tmA, tmB, num_sms, stream);
        case 256: return launch_fmt<256>(p, tmA, tmB, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}
