// tcgen05 3x3 convolution with shared-memory halo reuse, for the wide / shallow layers of the frontend
// (conv1_2, conv2_1, conv2_2, conv3_1 of the layer list in pero_ocr/ocr_engine/transformer.py:75-148).
//
// igemm_tc.cu reloads the activation tile once per filter tap; for 64/128-channel layers that makes the kernel
// L2->SM-bandwidth bound (profiles/: 9 x re-reads, ~10 TB/s).  Here one CTA computes 2 output rows x 128 output
// columns and loads the (2+2) x (128+2) pixel halo of a 64-channel chunk ONCE (one 4-D TMA box, 128B swizzle); the
// nine taps are nine shifted views of that buffer: a K-major swizzle-128B operand is 128 consecutive pixel rows of
// 128 B, so a tap (r, s) is just a different descriptor start address ((r*130 + s) * 128 B further).  Weights stream
// through their own ring.  Two accumulators (one per output row) make the 2x2 / 2x1 max-pool a per-thread max
// plus one shuffle.
//
// Warps: 0 = halo (A) producer, 1 = weight (B) producer, 2 = MMA issuer, 3-10 = epilogue (two per TMEM lane quarter).
#include "once.cuh"
#include "igemm.cuh"
#include "ptx.cuh"
#include "tilesched.cuh"

namespace {

constexpr int kThreads = 352;
constexpr int kHaloW = 130, kHaloH = 4;
constexpr int kHaloBytes = kHaloH * kHaloW * 128;             // 66,560
constexpr int kHaloStage = ((kHaloBytes + 1023) / 1024) * 1024;  // 66,560 -> 66,560 (65 KB) multiple of 1024

template <int BN, int NB>
struct Cfg {
    static constexpr int kBBytes = BN * 128;
    // An N = 64 MMA lasts 32 clocks, so with one filter tap per weight stage (8 MMAs) the single issuing thread spent
    // longer on the per-stage wait / commit / ring bookkeeping than the tensor pipe on the MMAs (ncu: pipe 41 %
    // busy, issuer never blocked on a barrier).  BN = 64 therefore streams a whole filter ROW (3 taps, 24 MMAs) per
    // stage.  (A third halo stage instead made no difference: the halo loads are not the limiter.)
    static constexpr int kTapsPerStage = BN == 64 ? 3 : 1;
    // BN = 64 additionally PAIRS the two output rows of a tile in the N dimension.  A weight stage then holds one
    // filter COLUMN s as three consecutive 64-row blocks W(2,s) | W(1,s) | W(0,s); halo row i feeds output row 0 with
    // filter row i and output row 1 with filter row i-1, so the two middle halo rows take ONE N = 128 MMA over two
    // adjacent blocks (accumulator columns [0,64) = row 0, [64,128) = row 1) and the outer halo rows one N = 64 MMA
    // each: 4 instead of 6 MMAs per (s, K step) and 84 instead of 108 KB of operand reads -- an M = 128 MMA reads
    // 4 KB of A whatever N is, which is what bounds this layer (profiles/r01d_halo_limits.md).
    static constexpr bool kPairRows = BN == 64;
    static constexpr int kAStages = 2;
    // The epilogue stores 32-byte sectors straight from registers (actfmt.cuh: epi_store32_v8, 256-bit stores); the
    // shared-memory staged variant of round 1 (whole sectors through a 32 KB detour) measured slower on every layer
    // (profiles/r02u_store_ab.json).  The 32 KB it used would hold a fourth weight stage at BN = 128: measured
    // slower (conv2_1 / conv2_2 +0.1 ms each, profiles/r02w_store_ab.json), so the ring stays at 3
    static constexpr int kBStages = NB;        // 3 (A/B of 4 at BN = 128: IgemmParams::dbg bit 8)
    static constexpr int kBStageBytes = kTapsPerStage * kBBytes;
    static constexpr int kTmemCols = 512;
    static constexpr int kBarOff = kAStages * kHaloStage + kBStages * kBStageBytes;
    static constexpr int kSmemBytes = kBarOff + 256 + 1024;
};

struct Tile {
    int img, h0, w0, nt;
};

__device__ __forceinline__ Tile tile_coord(const IgemmParams& p, int tile, int tiles_w, int tiles_h) {
    Tile t;
    t.nt = tile % p.tiles_n;
    int m = tile / p.tiles_n;
    const int tw = m % tiles_w;
    m /= tiles_w;
    t.h0 = (m % tiles_h) * 2;
    t.img = m / tiles_h;
    t.w0 = tw * 128;
    return t;
}

template <int BN, int FMT, int NB>
__global__ void __launch_bounds__(kThreads, 1)
igemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const IgemmParams p, int tiles_w, int tiles_h, int total_tiles) {
    using C = Cfg<BN, NB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::kAStages * kHaloStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
    uint64_t* a_full = bars;                              // [kAStages]
    uint64_t* a_empty = a_full + C::kAStages;             // [kAStages]
    uint64_t* b_full = a_empty + C::kAStages;                // [kBStages]
    uint64_t* b_empty = b_full + C::kBStages;             // [kBStages]
    uint64_t* tfull = b_empty + C::kBStages;              // [2]
    uint64_t* tempty = tfull + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* tq_full = tempty + 3;                       // [4] dynamic tile scheduler ring (tilesched.cuh)
    uint64_t* tq_empty = tq_full + kTileRing;             // [4]
    int* tq_tile = reinterpret_cast<int*>(tq_empty + kTileRing);
    float* s_vec = reinterpret_cast<float*>(smem + C::kBarOff + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kchunks = p.cin >> 6;
    // FMT (compile time) = operand format of the input AND record format of the output (one format per engine)
    constexpr int planes = FMT == ACT_F16 ? 1 : 2;
    constexpr bool f8 = FMT == ACT_F16_F8;   // plane 1 = e5m2 correction operands: A plane ap meets B plane ap only
    // IgemmParams::corr_mode: the e5m2 plane of a pixel is [lo'(cin) | hi8(cin)] bytes = kchunks 128-byte chunks.
    // CORR_WEIGHT keeps the hi8 half only: with cin = 128 that is chunk 1, with cin = 64 the upper two K = 32 steps
    // of the single chunk; CORR_NONE skips the plane.
    const int corr = p.corr_mode;
    auto skip_f8 = [&](int kc) { return corr == CORR_NONE || (corr == CORR_WEIGHT && kchunks == 2 && kc == 0); };
    const int f8_k0 = (corr == CORR_WEIGHT && kchunks == 1) ? 2 : 0;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < C::kAStages; ++i) {
            ptx::mbar_init(&a_full[i], 1);
            ptx::mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < C::kBStages; ++i) {
            ptx::mbar_init(&b_full[i], 1);
            ptx::mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], 8);
        }
        for (int i = 0; i < kTileRing; ++i) {
            ptx::mbar_init(&tq_full[i], 1);
            ptx::mbar_init(&tq_empty[i], 10);   // weight producer + MMA issuer + 8 epilogue warps
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
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
        // ------------------------------------------------------------------ halo (A) producer
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        TileFeed feed(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
        for (int tile = feed.next(lane); tile < total_tiles; tile = feed.next(lane)) {
            const Tile t = tile_coord(p, tile, tiles_w, tiles_h);
            for (int kc = 0; kc < kchunks; ++kc)
                for (int ap = 0; ap < planes; ++ap) {
                    if (f8 && ap == 1 && skip_f8(kc)) continue;
                    ptx::mbar_wait(&a_empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx_pred(&a_full[stage], kHaloBytes, leader);
                    ptx::tma_load_4d_pred(sA + stage * kHaloStage, &tmA, &a_full[stage], ap * p.cin + kc * 64,
                                          t.w0 - 1, t.h0 - 1, t.img, leader);
                    if (++stage == C::kAStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ weight (B) producer
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        TileTake take(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
        for (int tile = take.next_warp(lane); tile < total_tiles; tile = take.next_warp(lane)) {
            const Tile t = tile_coord(p, tile, tiles_w, tiles_h);
            for (int kc = 0; kc < kchunks; ++kc)
                for (int ap = 0; ap < planes; ++ap) {
                    if (f8 && ap == 1 && skip_f8(kc)) continue;
                    // fp16x3: A_hi meets B_hi and B_lo, A_lo meets B_hi only; fp16+fp8: A plane ap meets B plane ap
                    const int nbp = (ap == 0 && !f8) ? planes : 1;
                    for (int bp = 0; bp < nbp; ++bp)
                        for (int tap = 0; tap < 9; tap += C::kTapsPerStage) {
                            ptx::mbar_wait(&b_empty[stage], phase ^ 1);
                            ptx::mbar_expect_tx_pred(&b_full[stage], C::kBStageBytes, leader);
#pragma unroll
                            for (int tt = 0; tt < C::kTapsPerStage; ++tt)
                                ptx::tma_load_2d_pred(sB + stage * C::kBStageBytes + tt * C::kBBytes, &tmB, &b_full[stage],
                                                      kc * 64,
                                                      ((f8 ? ap : bp) * 9 + (C::kPairRows ? (2 - tt) * 3 + tap / 3 : tap + tt)) *
                                                              p.cout_pad + t.nt * BN,
                                                      leader);
                            if (++stage == C::kBStages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ MMA issuer
        // One elected thread runs the whole role (waits included): no per-stage reconvergence; the plane and tap loops
        // are unrolled, so the operand kind is a compile-time fact and the shifted-view offsets are immediates.
        // (Running the loop nest warp-wide with lane-predicated MMAs was tried: ptxas then emits an R2UR + VOTEU pair
        // per operand of every MMA.)
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::idesc_f16_f32(128, BN);
            constexpr uint32_t idesc8 = ptx::idesc_e5m2_f32(128, BN);
            const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
            const uint64_t desc_hi = ptx::smem_desc_sw128(0);   // descriptor with a zero start-address field
            const bool skip_mma = (p.dbg & 2) != 0;             // bring-up: time the pipeline without tensor work
            int as = 0, bs = 0, acc = 0;
            uint32_t aph = 0, bph = 0, acc_phase0 = 0, acc_phase1 = 0;
            TileTake take(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
            for (int tile = take.next(); tile < total_tiles; tile = take.next()) {
                ptx::mbar_wait(&tempty[acc], (acc ? acc_phase1 : acc_phase0) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d0 = tmem_base + acc * 2 * BN;
                for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
                    for (int ap = 0; ap < planes; ++ap) {
                        if (f8 && ap == 1 && skip_f8(kc)) continue;
                        ptx::mbar_wait(&a_full[as], aph);
                        const uint64_t a_base = desc_hi + ((sA_u + as * kHaloStage) >> 4);
                        const int nbp = (ap == 0 && !f8) ? planes : 1;
                        const bool e5m2 = f8 && ap == 1;
#pragma unroll
                        for (int bp = 0; bp < nbp; ++bp) {
                            // the very first MMA of a tile overwrites the accumulators
                            const uint32_t started = (ap | bp) != 0 ? 1u : (kc != 0 ? 1u : 0u);
#pragma unroll
                            for (int tap0 = 0; tap0 < 9; tap0 += C::kTapsPerStage) {
                                ptx::mbar_wait(&b_full[bs], bph);
                                ptx::tc_fence_after();
                                if (!skip_mma && C::kPairRows) {
                                    // stage = filter column s: blocks W(2,s) | W(1,s) | W(0,s).  Outer halo rows first:
                                    // they are the MMAs that may have to overwrite a row's accumulator
                                    constexpr uint32_t idesc_n2 = ptx::idesc_f16_f32(128, 2 * BN);
                                    constexpr uint32_t idesc8_n2 = ptx::idesc_e5m2_f32(128, 2 * BN);
                                    const int s = tap0 / 3;
                                    const uint64_t b0 = desc_hi + ((sB_u + bs * C::kBStageBytes) >> 4);
#pragma unroll
                                    for (int step = 0; step < 4; ++step) {
                                        const int hrow = step == 0 ? 3 : (step == 1 ? 0 : step - 1);   // 3, 0, 1, 2
                                        const int blk = step == 0 ? 0 : (step == 1 ? 2 : (step == 2 ? 1 : 0));
                                        const bool wide = step >= 2;
                                        const uint32_t d = d0 + (step == 0 ? BN : 0);
                                        const uint64_t a_desc = a_base + ((hrow * kHaloW + s) * 8);
                                        const uint64_t b_desc = b0 + ((blk * C::kBBytes) >> 4);
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (e5m2) {
                                                if (k >= f8_k0)
                                                    ptx::mma_f8_ss(d, a_desc + 2 * k, b_desc + 2 * k, wide ? idesc8_n2 : idesc8, 1u);
                                            } else
                                                ptx::mma_f16_ss(d, a_desc + 2 * k, b_desc + 2 * k, wide ? idesc_n2 : idesc,
                                                                (wide || s != 0 || k != 0) ? 1u : started);
                                        }
                                    }
                                } else if (!skip_mma) {
#pragma unroll
                                    for (int tt = 0; tt < C::kTapsPerStage; ++tt) {
                                        const int tap = tap0 + tt;
                                        const uint64_t b_desc =
                                            desc_hi + ((sB_u + bs * C::kBStageBytes + tt * C::kBBytes) >> 4);
#pragma unroll
                                        for (int r = 0; r < 2; ++r) {
                                            // pixel ((r + tap/3) * 130 + tap%3) of the halo, 128 B per pixel, >>4 encoded
                                            const uint64_t a_desc = a_base + (((r + tap / 3) * kHaloW + tap % 3) * 8);
#pragma unroll
                                            for (int k = 0; k < 4; ++k) {
                                                if (e5m2) {
                                                    if (k >= f8_k0)
                                                        ptx::mma_f8_ss(d0 + r * BN, a_desc + 2 * k, b_desc + 2 * k, idesc8, 1u);
                                                } else
                                                    ptx::mma_f16_ss(d0 + r * BN, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                                                    (tap | k) != 0 ? 1u : started);
                                            }
                                        }
                                    }
                                }
                                ptx::mma_commit(&b_empty[bs]);
                                if (++bs == C::kBStages) {
                                    bs = 0;
                                    bph ^= 1;
                                }
                            }
                        }
                        ptx::mma_commit(&a_empty[as]);
                        if (++as == C::kAStages) {
                            as = 0;
                            aph ^= 1;
                        }
                    }
                }
                ptx::mma_commit(&tfull[acc]);
                if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
                acc ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 3..10)
        const int q = warp & 3;
        const int half = (warp - 3) >> 2;
        const float* s_bias = s_vec;
        const float* s_scale = s_vec + p.cout_pad;
        const float* s_shift = s_vec + 2 * p.cout_pad;
        const bool has_affine = p.post_scale != nullptr;
        const int Hp = p.h_out / p.pool_h, Wp = p.w_out / p.pool_w;
        int acc = 0;
        uint32_t acc_phase0 = 0, acc_phase1 = 0;
        TileTake take(p.tile_counter, total_tiles, tq_full, tq_empty, tq_tile);
        for (int tile = take.next_warp(lane); tile < total_tiles; tile = take.next_warp(lane)) {
            const Tile t = tile_coord(p, tile, tiles_w, tiles_h);
            const int wo = t.w0 + q * 32 + lane;
            ptx::mbar_wait(&tfull[acc], acc ? acc_phase1 : acc_phase0);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * 2 * BN + (static_cast<uint32_t>(q * 32) << 16);
            const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
            const bool col_ok = wo < p.w_out;
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                const int n0 = t.nt * BN + c0;
                if (n0 >= p.cout) break;
                uint32_t r0[32], r1[32];
                ptx::tmem_ld_32x32b_x32(t_addr + c0, r0);
                ptx::tmem_ld_32x32b_x32(t_addr + BN + c0, r1);
                ptx::tmem_ld_wait();
                if (p.pool_h == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        r0[j] = __float_as_uint(fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
                }
                if (p.pool_w == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        r0[j] = __float_as_uint(fmaxf(__uint_as_float(r0[j]),
                                                      __shfl_xor_sync(0xffffffffu, __uint_as_float(r0[j]), 1)));
                    if (p.pool_h == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r1[j] = __float_as_uint(fmaxf(__uint_as_float(r1[j]),
                                                          __shfl_xor_sync(0xffffffffu, __uint_as_float(r1[j]), 1)));
                    }
                }
                const bool writer = col_ok && (p.pool_w == 1 || (lane & 1) == 0);
                const int rows = p.pool_h == 2 ? 1 : 2;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    if (rr >= rows) break;
                    const int ho = t.h0 + rr;
                    if (ho >= p.h_out) continue;            // warp-uniform
                    __half* orow = p.out_h + (static_cast<size_t>(t.img) * Hp * Wp +
                                              static_cast<size_t>(ho / p.pool_h) * Wp + wo / p.pool_w) * p.out_cstride;
                    if (p.dbg & 1) continue;                // bring-up: time the pipeline without the global stores
                    if (!writer) continue;
                    if (p.dbg & 4)                          // A/B: 16-byte stores
                        epi_store32(rr == 0 ? r0 : r1, n0, p.acc_scale, s_bias, s_scale, s_shift, has_affine, p.act, p.act_slope, orow,
                                    p.cout, FMT);
                    else
                        epi_store32_v8(rr == 0 ? r0 : r1, n0, p.acc_scale, s_bias, s_scale, s_shift, has_affine, p.act, p.act_slope,
                                       orow, p.cout, FMT, p.out_skip_lo != 0);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
            acc ^= 1;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

template <int BN, int FMT, int NB>
cudaError_t launch_nb(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                      cudaStream_t stream) {
    using C = Cfg<BN, NB>;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel<BN, FMT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.mark();
    }
    const size_t smem_bytes = C::kSmemBytes + 3 * static_cast<size_t>(p.cout_pad) * sizeof(float);
    if (smem_bytes > 227 * 1024) return cudaErrorInvalidValue;
    const int tiles_w = (p.w_out + 127) / 128, tiles_h = (p.h_out + 1) / 2;
    const int total_tiles = p.n_img * tiles_h * tiles_w * p.tiles_n;
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    igemm_halo_kernel<BN, FMT, NB><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, p, tiles_w, tiles_h, total_tiles);
    return cudaGetLastError();
}

template <int BN, int FMT>
cudaError_t launch_bn(const IgemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                      cudaStream_t stream) {
    if (BN == 128 && (p.dbg & 8)) return launch_nb<BN, FMT, BN == 128 ? 4 : 3>(p, tmA, tmB, num_sms, stream);
    return launch_nb<BN, FMT, 3>(p, tmA, tmB, num_sms, stream);
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

bool igemm_halo_supported(const IgemmParams& p, int bn) {
    return p.epi == EPI_ACT_F16 && p.kh == 3 && p.kw == 3 && p.pad_h == 1 && p.pad_w == 1 && (bn == 64 || bn == 128) &&
           p.cin <= 128 && (p.h_out % 2) == 0 && (p.cout % 32) == 0;
}

cudaError_t launch_igemm_halo(const IgemmParams& p, const CUtensorMap& tmA_halo, const CUtensorMap& tmB, int bn,
                              int num_sms, cudaStream_t stream) {
    if (bn == 64) return launch_fmt<64>(p, tmA_halo, tmB, num_sms, stream);
    if (bn == 128) return launch_fmt<128>(p, tmA_halo, tmB, num_sms, stream);
    return cudaErrorInvalidValue;
}
