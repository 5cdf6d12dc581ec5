// First convolution of the recogniser (3 -> COUT channels, 3x3, pad 1) fused with the reference's `/255` and
// NHWC -> NCHW permute (pero_ocr/ocr_engine/pytorch_ocr_engine.py:61-62) -- the bytes of the padded crop batch are
// the operand.
//
// K = 27 is far too short for a tcgen05 pipeline (one 128 x 64 x 32 MMA per 128 pixels against 32 KB of output),
// and on CUDA cores the layer was issue-bound at ~30 % of the FMA peak (1728 FMAs per pixel fed by shared-memory
// weight broadcasts).  Here it runs on warp-level mma.sync (m16n8k16, fp16 operands, fp32 accumulate):
//   * the uint8 pixels are exact in fp16, so the A operand is the raw byte value and the 1/255 moves to the epilogue;
//   * the weights are normalised per output channel by a power of two s_n (largest |w| in [0.5, 1)) and split
//     w / s_n = hi + lo into two fp16 operands -- both products accumulate into the same fp32 registers, so the layer
//     keeps fp32-grade accuracy (|lo| error <= 2^-25 of the channel's largest weight);
//   * K is laid out as k = 10 r + (3 s + c) (slots 9, 19, 29, 30, 31 are zero weights), so that an A-fragment
//     register (two consecutive k) is two consecutive fp16 of one row of the staged input patch;
//   * epilogue: acc * (s_n / 255) + bias -> activation -> activation record (actfmt.cuh) staged in shared memory
//     (16-byte chunks XOR-swizzled by pixel: conflict-free fragment writes and conflict-free record reads) ->
//     fully coalesced 16-byte stores.  The kernel is bound by that store stream (the 64-channel hi|lo records of
//     a 256 x 40 x 1344 batch are 3.5 GB).
//
// Staging of the uint8 crop patch (6 image rows x 130 pixels x 3 bytes per CTA) from HBM into shared memory comes in
// three variants, selectable for A/B (b200ocr_debug_set_flag 4; profiles/r02*_conv_first_staging.md):
//   STAGE_TMA      one cp.async.bulk.tensor box (104 x 6 32-bit words) per CTA through a 3-D tensor map over the crop
//                  batch, completing on an mbarrier; rows / columns outside the image are the TMA unit's zero fill
//                  -- the convolution's zero padding costs no instruction (default whenever W is a multiple of 16)
//   STAGE_CPASYNC  16-byte cp.async (LDGSTS) chunks, out-of-image chunks zero-filled through the src-size operand
//   STAGE_PLAIN    one byte per thread and iteration with ordinary loads (any W; round 1's path)
// The staged bytes are then widened to the fp16 operand patch in shared memory by all threads.
#include "once.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace {

constexpr int CFM_PX = 128;      // output pixels per CTA row (4 warps x 32)
constexpr int CFM_ROWS = 4;      // image rows per CTA
constexpr int CFM_PSTRIDE = 392; // fp16 per staged patch row: 130 pixels x 3 channels + 2 pad
constexpr int CFM_U8ROW = 416;   // bytes per staged uint8 row: [w0 * 3 - 16, w0 * 3 + 400), 16-byte aligned when W % 16 == 0
constexpr int CFM_U8OFF = 13;    // patch byte b (pixel w0 - 1, channel 0 = byte 0) sits at CFM_U8OFF + b of its uint8 row
enum { STAGE_PLAIN = 0, STAGE_CPASYNC = 1, STAGE_TMA = 2 };

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float cfm_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : slope * v;
    return v;
}

__device__ __forceinline__ uint32_t ld_pair(const __half* p) {   // two consecutive fp16 at any 2-byte alignment
    const uint32_t lo = *reinterpret_cast<const unsigned short*>(p);
    const uint32_t hi = *reinterpret_cast<const unsigned short*>(p + 1);
    return lo | (hi << 16);
}

// bulk copy shared::cta -> global (bytes a multiple of 16, both addresses 16-byte aligned), bulk-group completion
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async_smem_cf() {   // generic-proxy smem writes -> visible to tcgen05.mma
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

template <int COUT, int STAGE>
__global__ void __launch_bounds__(CFM_PX, 5) conv_first_mma_kernel(const uint8_t* __restrict__ in, int n, int h, int w,
                                                              const uint2* __restrict__ wfrag,
                                                              const float* __restrict__ oscale,
                                                              const float* __restrict__ bias, int act, float slope,
                                                              int fmt, __half* __restrict__ out, int skip_lo,
                                                              const __grid_constant__ CUtensorMap tm_in) {
    constexpr int NT = COUT / 8;
    const int planes = act_planes(fmt);
    __shared__ __align__(16) __half s_p[(CFM_ROWS + 2) * CFM_PSTRIDE];
    __shared__ __align__(128) uint8_t s_u8[STAGE == STAGE_PLAIN ? 16 : (CFM_ROWS + 2) * CFM_U8ROW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_sc[COUT], s_b[COUT];
    extern __shared__ uint4 s_stage[];   // [4 warps][16 px][planes * COUT / 8]: one m-tile of pixel records per warp

    const int tiles_w = (w + CFM_PX - 1) / CFM_PX;
    const int tiles_h = (h + CFM_ROWS - 1) / CFM_ROWS;
    const int tw = blockIdx.x % tiles_w;
    const int row0 = ((blockIdx.x / tiles_w) % tiles_h) * CFM_ROWS;
    const int img = blockIdx.x / (tiles_w * tiles_h);
    const int w0 = tw * CFM_PX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;

    // weight fragments ([plane][k step][n tile][lane] x {b0b1, b2b3}) live in shared memory: in registers they cost
    // 64 of 168 and held the kernel at 3 warps per scheduler (ncu: latency-bound, 7 cycles per issued instruction);
    // a lane reads its own 8 bytes, consecutive lanes consecutive words: conflict-free
    __shared__ uint2 s_bf[2 * 2 * NT * 32];
    for (int i = threadIdx.x; i < 2 * 2 * NT * 32; i += CFM_PX) s_bf[i] = __ldg(wfrag + i);
    for (int i = threadIdx.x; i < COUT; i += CFM_PX) {
        s_sc[i] = oscale[i];
        s_b[i] = bias ? bias[i] : 0.f;
    }
    // input patch (rows row0-1 .. row0+CFM_ROWS, pixels w0-1 .. w0+128) as fp16 byte values, zero outside the image
    constexpr int kRowVals = (CFM_PX + 2) * 3;
    if (STAGE == STAGE_PLAIN) {
        for (int i = threadIdx.x; i < (CFM_ROWS + 2) * CFM_PSTRIDE; i += CFM_PX) {
            const int r = i / CFM_PSTRIDE;
            const int b = i - r * CFM_PSTRIDE;
            const int yy = row0 + r - 1;
            const int xb = (w0 - 1) * 3 + b;
            unsigned short v = 0;
            if (b < kRowVals && yy >= 0 && yy < h && xb >= 0 && xb < w * 3)
                v = in[(static_cast<size_t>(img) * h + yy) * w * 3 + xb];
            s_p[i] = __ushort2half_rn(v);
        }
    } else {
        if (STAGE == STAGE_TMA) {
            // one box: 104 words x 6 rows of image `img`, starting 16 bytes left of pixel w0 and one row above row0
            if (threadIdx.x == 0) {
                ptx::mbar_init(&s_bar, 1);
                ptx::fence_mbar_init();
                ptx::mbar_expect_tx(&s_bar, (CFM_ROWS + 2) * CFM_U8ROW);
                ptx::tma_load_3d(s_u8, &tm_in, &s_bar, w0 * 3 / 4 - 4, row0 - 1, img);
            }
        } else {
            const int row_bytes = w * 3;
            for (int i = threadIdx.x; i < (CFM_ROWS + 2) * (CFM_U8ROW / 16); i += CFM_PX) {
                const int r = i / (CFM_U8ROW / 16), c = i - r * (CFM_U8ROW / 16);
                const int yy = row0 + r - 1;
                const int xb = w0 * 3 - 16 + c * 16;
                const bool ok = yy >= 0 && yy < h && xb >= 0 && xb < row_bytes;   // W % 16 == 0: chunks never straddle
                const uint8_t* src = in + (static_cast<size_t>(img) * h + (ok ? yy : 0)) * row_bytes + (ok ? xb : 0);
                ptx::cp_async_16(s_u8 + r * CFM_U8ROW + c * 16, src, ok ? 16u : 0u);
            }
            ptx::cp_async_commit_wait_all();
        }
        __syncthreads();   // TMA: the barrier init is visible to the waiters; cp.async: every thread's chunks landed
        if (STAGE == STAGE_TMA) ptx::mbar_wait(&s_bar, 0);
        for (int i = threadIdx.x; i < (CFM_ROWS + 2) * CFM_PSTRIDE; i += CFM_PX) {
            const int r = i / CFM_PSTRIDE;
            const int b = i - r * CFM_PSTRIDE;
            const unsigned short v = b < kRowVals ? s_u8[r * CFM_U8ROW + CFM_U8OFF + b] : 0;
            s_p[i] = __ushort2half_rn(v);
        }
    }
    // patch offsets of this thread's A-fragment columns: k = 16 ks + 2 tig (+ 8) -> row r = k / 10, value j = k % 10
    int aoff[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int k = ks * 16 + tig * 2 + 8 * hh;
            aoff[ks][hh] = k < 30 ? (k / 10) * CFM_PSTRIDE + (k % 10) : 0;   // k >= 30: zero weights, any finite value
        }
    __syncthreads();

    const int rec = planes * COUT;
    const int chunks = rec / 8;                 // 16-byte chunks per pixel record
    // Each warp stages ONE 16-pixel m-tile of records at a time in its private slice of shared memory and writes it
    // back itself (16 px x `rec` fp16 contiguous in HBM: whole 128-byte lines): no block-wide barrier in the row loop
    // and 16 KB instead of 32 KB of staging per CTA -- the version that staged a whole 128-pixel row per CTA was
    // limited to 4 CTAs (16 warps) per SM by shared memory and ran at 36 % of the copy bandwidth
    // (profiles/r02d_conv_first_staging_ab.md).
    uint4* w_stage = s_stage + warp * 16 * chunks;
    for (int rr = 0; rr < CFM_ROWS; ++rr) {
        const int row = row0 + rr;
        if (row >= h) break;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int pb = warp * 32 + mt * 16;
            const __half* p0 = s_p + rr * CFM_PSTRIDE + 3 * (pb + gid);
            const __half* p1 = p0 + 3 * 8;
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                a[0] = ld_pair(p0 + aoff[ks][0]);
                a[1] = ld_pair(p1 + aoff[ks][0]);
                a[2] = ld_pair(p0 + aoff[ks][1]);
                a[3] = ld_pair(p1 + aoff[ks][1]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint2 bh = s_bf[((0 * 2 + ks) * NT + nt) * 32 + lane];
                    const uint2 bl = s_bf[((1 * 2 + ks) * NT + nt) * 32 + lane];
                    mma16816(acc[nt], a, bh.x, bh.y);
                    mma16816(acc[nt], a, bl.x, bl.y);
                }
            }
            __syncwarp();                       // the previous m-tile's records have been read out of the slice
            // fragment -> staged records: rows (pixels) gid and gid + 8 of the m-tile, channels 8 nt + 2 tig + {0, 1}
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int px = gid + 8 * half;
                uint4* my = w_stage + px * chunks;
                const int sw = px & (chunks - 1);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int ch = nt * 8 + tig * 2;
                    const float v0 = cfm_act(fmaf(acc[nt][2 * half], s_sc[ch], s_b[ch]), act, slope);
                    const float v1 = cfm_act(fmaf(acc[nt][2 * half + 1], s_sc[ch + 1], s_b[ch + 1]), act, slope);
                    const __half2 h2 = __floats2half2_rn(v0, v1);
                    reinterpret_cast<uint32_t*>(my + (nt ^ sw))[tig] = *reinterpret_cast<const uint32_t*>(&h2);
                    if (fmt != ACT_F16) {
                        const float2 hf = __half22float2(h2);
                        if (fmt == ACT_F16_HILO) {
                            const __half2 l2 = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                            reinterpret_cast<uint32_t*>(my + ((NT + nt) ^ sw))[tig] = *reinterpret_cast<const uint32_t*>(&l2);
                        } else {
                            const unsigned short lo8 =
                                static_cast<unsigned short>(pack_e5m2x2((v0 - hf.x) * kF8Scale, (v1 - hf.y) * kF8Scale));
                            const unsigned short hi8 = static_cast<unsigned short>(pack_e5m2x2(hf.x, hf.y));
                            const int sub = (nt & 1) * 4 + tig;   // 16-bit slot inside the 16-byte chunk
                            reinterpret_cast<unsigned short*>(my + ((NT + (nt >> 1)) ^ sw))[sub] = lo8;
                            reinterpret_cast<unsigned short*>(my + ((NT + NT / 2 + (nt >> 1)) ^ sw))[sub] = hi8;
                        }
                    }
                }
            }
            __syncwarp();
            {
                const int npx = min(16, w - (w0 + pb));     // pixels of this m-tile inside the image (<= 0: none)
                uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * h + row) * w + w0 + pb) * rec);
                for (int i = lane; i < npx * chunks; i += 32) {
                    const int px = i / chunks, c = i - px * chunks;
                    // skip_lo (ACT_F16_F8): the consumer multiplies with weight-side correction only and never reads
                    // the lo' plane (chunks [NT, NT + NT/2) of the record) -- a quarter of the bytes stays unwritten
                    if (skip_lo && c >= NT && c < NT + NT / 2) continue;
                    dst[i] = w_stage[px * chunks + (c ^ (px & (chunks - 1)))];
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------- tcgen05 variant
// The same arithmetic on the 5th-generation tensor cores (staging value 3; b200ocr_debug_set_flag 4).  The mma.sync
// kernel above is bound by L1 / shared-memory wavefronts (ncu: l1tex 85 %): per 16-pixel m-tile a warp moves ~190
// wavefronts of fragments and staged records for 64 MMAs' worth of work.  Here a CTA of 128 threads owns 128 pixels
// of CFM_ROWS rows; per row
//   * thread p gathers the 30 patch values of pixel p (three runs of 10 consecutive fp16 of the staged patch) and
//     writes its 64-byte K row of the A tile [128 px][K = 32] in the K-major core-matrix layout (no swizzle: 8 rows x
//     16 bytes per core matrix; consecutive threads write consecutive 16 bytes -- conflict-free);
//   * one thread issues 4 tcgen05.mma (M = 128, N = COUT, K = 16; two K steps x hi / lo weight planes) into one of two
//     TMEM accumulators and commits to an mbarrier: the MMAs of row r + 1 run under the epilogue of row r;
//   * epilogue: tcgen05.ld of the thread's own TMEM lane (its pixel), scale / bias / activation, record planes; the
//     record is assembled in the thread's slot of a shared-memory stage (conflict-free 16-byte stores) and leaves as
//     one cp.async.bulk shared -> global per pixel.  (Measured on the way: each thread storing its record straight
//     from registers -- 32-byte pieces 256 bytes apart -- 1.55-1.70 ms; parking the records and reading them back for
//     coalesced 16-byte stores 1.32 ms; the mma.sync kernel 1.45-1.55 ms.)
// Shared-memory traffic per 128-pixel row: 8 KB A written + 4 x (4 KB A + COUT x 32 B) read by the tensor core, against
// ~190 KB of wavefronts for the same pixels above.
template <int COUT>
struct CftCfg {
    static constexpr int kABytes = CFM_PX * 64;             // [128 px][32 fp16]
    static constexpr int kALbo = (CFM_PX / 8) * 128;        // K-direction core-matrix stride
    static constexpr int kBPlane = COUT * 64;               // [COUT][32 fp16]
    static constexpr int kBLbo = (COUT / 8) * 128;
    static constexpr int kTmemCols = 2 * COUT < 32 ? 32 : 2 * COUT;   // two accumulators
};

__device__ __forceinline__ uint64_t cft_desc(uint32_t addr, uint32_t lbo) {   // K-major, no swizzle, SBO = 128
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(128 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// 32 consecutive channels [n0, n0 + 32) of the thread's pixel -> packed record words (ph: 16 words hi; pl: HILO 16
// words lo, F8 8 words lo' + 8 words hi8), per-channel scale folded into the fma.
__device__ __forceinline__ void cft_pack32(const uint32_t (&r)[32], int n0, const float* s_sc, const float* s_b, int act,
                                           float slope, int fmt, uint32_t (&ph)[16], uint32_t (&pl)[16]) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            v[e] = cfm_act(fmaf(__uint_as_float(r[j + e]), s_sc[n0 + j + e], s_b[n0 + j + e]), act, slope);
        const __half2 h01 = __floats2half2_rn(v[0], v[1]);
        const __half2 h23 = __floats2half2_rn(v[2], v[3]);
        ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
        ph[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        if (fmt != ACT_F16) {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            if (fmt == ACT_F16_HILO) {
                const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y);
                const __half2 l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
                pl[j >> 1] = *reinterpret_cast<const uint32_t*>(&l01);
                pl[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l23);
            } else {
                pl[j >> 2] = pack_e5m2x4((v[0] - f01.x) * kF8Scale, (v[1] - f01.y) * kF8Scale,
                                         (v[2] - f23.x) * kF8Scale, (v[3] - f23.y) * kF8Scale);
                pl[8 + (j >> 2)] = pack_e5m2x4(f01.x, f01.y, f23.x, f23.y);
            }
        }
    }
}

template <int COUT>
__global__ void __launch_bounds__(CFM_PX, 4) conv_first_tc_kernel(const uint8_t* __restrict__ in, int n, int h, int w,
                                                                  const uint4* __restrict__ wk,
                                                                  const float* __restrict__ oscale,
                                                                  const float* __restrict__ bias, int act, float slope,
                                                                  int fmt, __half* __restrict__ out, int skip_lo,
                                                                  const __grid_constant__ CUtensorMap tm_in) {
    using C = CftCfg<COUT>;
    const int planes = act_planes(fmt);
    __shared__ __align__(16) __half s_p[(CFM_ROWS + 2) * CFM_PSTRIDE];
    __shared__ __align__(128) uint8_t s_a[C::kABytes];
    uint8_t* s_u8 = s_a;       // the staged uint8 patch (2.5 KB) lives in the A tile's space until it is widened
    static_assert((CFM_ROWS + 2) * CFM_U8ROW <= C::kABytes, "patch must fit the A tile");
    __shared__ __align__(128) uint8_t s_bw[2 * C::kBPlane];
    __shared__ __align__(8) uint64_t s_bar, s_done[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_sc[COUT];
    __shared__ __align__(16) float s_b[COUT];
    extern __shared__ uint4 s_stage_tc[];        // [4 warps][32 px][planes * COUT / 8 + 1 chunks of 16 bytes]

    const int tiles_w = (w + CFM_PX - 1) / CFM_PX;
    const int tiles_h = (h + CFM_ROWS - 1) / CFM_ROWS;
    const int tw = blockIdx.x % tiles_w;
    const int row0 = ((blockIdx.x / tiles_w) % tiles_h) * CFM_ROWS;
    const int img = blockIdx.x / (tiles_w * tiles_h);
    const int w0 = tw * CFM_PX;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        ptx::mbar_init(&s_bar, 1);
        ptx::mbar_init(&s_done[0], 1);
        ptx::mbar_init(&s_done[1], 1);
        ptx::fence_mbar_init();
        ptx::mbar_expect_tx(&s_bar, (CFM_ROWS + 2) * CFM_U8ROW);
        ptx::tma_load_3d(s_u8, &tm_in, &s_bar, w0 * 3 / 4 - 4, row0 - 1, img);
    }
    if (warp == 0) ptx::tmem_alloc<C::kTmemCols>(&s_tmem);
    // weights [plane][COUT][32] (k contiguous) -> [plane][k chunk][n][8 fp16]
    for (int i = tid; i < 2 * COUT * 4; i += CFM_PX) {
        const int pl = i / (COUT * 4), rest = i - pl * COUT * 4;
        const int nn = rest >> 2, j = rest & 3;
        *reinterpret_cast<uint4*>(s_bw + pl * C::kBPlane + j * C::kBLbo + nn * 16) = __ldg(wk + i);
    }
    for (int i = tid; i < COUT; i += CFM_PX) {
        s_sc[i] = oscale[i];
        s_b[i] = bias ? bias[i] : 0.f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    ptx::mbar_wait(&s_bar, 0);
    constexpr int kRowVals = (CFM_PX + 2) * 3;
    for (int i = tid; i < (CFM_ROWS + 2) * CFM_PSTRIDE; i += CFM_PX) {
        const int r = i / CFM_PSTRIDE;
        const int b = i - r * CFM_PSTRIDE;
        const unsigned short v = b < kRowVals ? s_u8[r * CFM_U8ROW + CFM_U8OFF + b] : 0;
        s_p[i] = __ushort2half_rn(v);
    }
    __syncthreads();

    const int rows = min(CFM_ROWS, h - row0);
    const uint32_t a_base = ptx::smem_u32(s_a), b_base = ptx::smem_u32(s_bw);
    constexpr uint32_t idesc = ptx::idesc_f16_f32(CFM_PX, COUT);

    auto build = [&](int rr) {                     // A tile of output row rr: the K row of pixel `tid`
        uint32_t wd[16];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const __half* q = s_p + (rr + r) * CFM_PSTRIDE + 3 * tid;
#pragma unroll
            for (int m = 0; m < 5; ++m) wd[5 * r + m] = ld_pair(q + 2 * m);
        }
        wd[15] = 0u;
        uint8_t* dst = s_a + tid * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + j * C::kALbo) = make_uint4(wd[4 * j], wd[4 * j + 1], wd[4 * j + 2], wd[4 * j + 3]);
    };
    auto issue = [&](int buf) {                    // thread 0: 2 K steps x (hi, lo) planes into accumulator `buf`
        const uint32_t acc = tmem_base + buf * COUT;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                ptx::mma_f16_ss(acc, cft_desc(a_base + ks * 2 * C::kALbo, C::kALbo),
                                cft_desc(b_base + pl * C::kBPlane + ks * 2 * C::kBLbo, C::kBLbo), idesc,
                                (pl | ks) ? 1u : 0u);
        ptx::mma_commit(&s_done[buf]);
    };

    build(0);
    fence_proxy_async_smem_cf();
    __syncthreads();
    if (tid == 0) {
        ptx::tc_fence_after();
        issue(0);
    }
    const int rec = planes * COUT;               // fp16 units per pixel record
    const int chunks = rec / 8;                  // 16-byte chunks per record (16 for the two-plane formats at COUT = 64)
    uint4* stage = s_stage_tc + warp * 32 * (chunks + 1);
    const int px0 = w0 + warp * 32;              // first pixel of this warp's 32
    for (int rr = 0; rr < rows; ++rr) {
        const int buf = rr & 1;
        ptx::mbar_wait(&s_done[buf], (rr >> 1) & 1);          // MMAs of row rr done: A is free, accumulator buf is full
        ptx::tc_fence_after();
        if (rr + 1 < rows) {
            build(rr + 1);
            fence_proxy_async_smem_cf();
            ptx::tc_fence_before();                           // the tcgen05.ld of row rr - 1 (accumulator buf ^ 1) are done
            __syncthreads();
            if (tid == 0) {
                ptx::tc_fence_after();
                issue(buf ^ 1);                               // runs under this row's epilogue
            }
        }
        // epilogue: this thread's pixel = its TMEM lane.  The record is assembled in the thread's own slot of the stage
        // (slots 16 bytes longer than a record: consecutive lanes start in consecutive bank groups, the 16-byte
        // stores are conflict-free) and leaves as ONE bulk copy shared -> global per pixel (cp.async.bulk: the record
        // is contiguous on both sides); no read-back through the LSU, no warp-level synchronisation -- a thread only
        // waits for its own previous copy before it reuses its slot
        const uint32_t t_addr = tmem_base + buf * COUT + (static_cast<uint32_t>(warp * 32) << 16);
        bulk_wait_read_all();
        uint4* my = stage + lane * (chunks + 1);
#pragma unroll
        for (int n0 = 0; n0 < COUT; n0 += 32) {
            uint32_t r[32], ph[16], pl[16];
            ptx::tmem_ld_32x32b_x32(t_addr + n0, r);
            ptx::tmem_ld_wait();
            cft_pack32(r, n0, s_sc, s_b, act, slope, fmt, ph, pl);
            const int c_hi = n0 / 8;                          // chunk of channel n0 in the hi plane
#pragma unroll
            for (int j = 0; j < 4; ++j) my[c_hi + j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
            if (fmt == ACT_F16_HILO) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    my[COUT / 8 + c_hi + j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            } else if (fmt == ACT_F16_F8) {
                const int c_lo = COUT / 8 + n0 / 16, c_h8 = COUT / 8 + COUT / 16 + n0 / 16;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    my[c_lo + j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
                    my[c_h8 + j] = make_uint4(pl[8 + 4 * j], pl[8 + 4 * j + 1], pl[8 + 4 * j + 2], pl[8 + 4 * j + 3]);
                }
            }
        }
        fence_proxy_async_smem_cf();
        if (px0 + lane < w) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(out + ((static_cast<size_t>(img) * h + row0 + rr) * w + px0 + lane) * rec);
            const uint32_t src = ptx::smem_u32(my);
            if (skip_lo && fmt == ACT_F16_F8) {               // the lo' plane stays unwritten: hi, then hi8
                bulk_store(dst, src, COUT * 2);
                bulk_store(dst + COUT * 3, src + COUT * 3, COUT);
            } else {
                bulk_store(dst, src, rec * 2);
            }
        }
        bulk_commit();
    }
    bulk_wait_all();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

}  // namespace

size_t conv_first_wfrag_words(int cout) { return static_cast<size_t>(2) * 2 * (cout / 8) * 32 * 2; }

void conv_first_pack(const float* weight, int cout, uint32_t* wfrag, float* oscale) {
    // weight: PyTorch [cout][3][3][3] (o, c, r, s).  K slot k = 10 r + 3 s + c.
    for (int o = 0; o < cout; ++o) {
        float m = 0.f;
        for (int i = 0; i < 27; ++i) m = fmaxf(m, fabsf(weight[o * 27 + i]));
        int ex = 0;
        if (m > 0.f) frexpf(m, &ex);             // m = f * 2^ex, f in [0.5, 1)
        oscale[o] = ldexpf(1.f, ex) / 255.0f;    // folded with the reference's /255
    }
    auto wk = [&](int o, int k, int plane) -> __half {
        if (k >= 30 || k % 10 == 9) return __float2half_rn(0.f);
        const int r = k / 10, j = k % 10, s = j / 3, c = j % 3;
        int ex = 0;
        float m = 0.f;
        for (int i = 0; i < 27; ++i) m = fmaxf(m, fabsf(weight[o * 27 + i]));
        if (m > 0.f) frexpf(m, &ex);
        const float v = ldexpf(weight[((o * 3 + c) * 3 + r) * 3 + s], -ex);
        const __half hi = __float2half_rn(v);
        return plane == 0 ? hi : __float2half_rn(v - __half2float(hi));
    };
    const int NT = cout / 8;
    for (int p = 0; p < 2; ++p)
        for (int ks = 0; ks < 2; ++ks)
            for (int nt = 0; nt < NT; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int gid = lane >> 2, tig = lane & 3, o = nt * 8 + gid;
                    for (int q = 0; q < 2; ++q) {
                        const int k = ks * 16 + tig * 2 + 8 * q;
                        const __half a = wk(o, k, p), b = wk(o, k + 1, p);
                        const uint32_t lo = *reinterpret_cast<const unsigned short*>(&a);
                        const uint32_t hi = *reinterpret_cast<const unsigned short*>(&b);
                        wfrag[((((p * 2 + ks) * NT + nt) * 32) + lane) * 2 + q] = lo | (hi << 16);
                    }
                }
}

size_t conv_first_tc_words(int cout) { return static_cast<size_t>(2) * cout * 32 / 2; }   // 32-bit words

void conv_first_pack_tc(const float* weight, int cout, uint32_t* wk) {
    // [plane][cout][k = 10 r + 3 s + c, 32 slots] fp16, weights scaled per output channel like conv_first_pack
    __half* dst = reinterpret_cast<__half*>(wk);
    for (int o = 0; o < cout; ++o) {
        float m = 0.f;
        for (int i = 0; i < 27; ++i) m = fmaxf(m, fabsf(weight[o * 27 + i]));
        int ex = 0;
        if (m > 0.f) frexpf(m, &ex);
        for (int k = 0; k < 32; ++k) {
            float v = 0.f;
            if (k < 30 && k % 10 != 9) {
                const int r = k / 10, j = k % 10, sx = j / 3, c = j % 3;
                v = ldexpf(weight[((o * 3 + c) * 3 + r) * 3 + sx], -ex);
            }
            const __half hi = __float2half_rn(v);
            dst[(static_cast<size_t>(0) * cout + o) * 32 + k] = hi;
            dst[(static_cast<size_t>(1) * cout + o) * 32 + k] = __float2half_rn(v - __half2float(hi));
        }
    }
}

cudaError_t launch_conv_first_tc(const uint8_t* in, int n, int h, int w, const uint32_t* wk, const float* oscale,
                                 const float* bias, int cout, int act, float slope, int fmt, __half* out, int skip_lo,
                                 const CUtensorMap* tm_in, cudaStream_t stream) {
    if (!tm_in || (w % 16) || (cout != 64 && cout != 32)) return cudaErrorInvalidValue;
    if (fmt != ACT_F16_F8) skip_lo = 0;
    const int tiles_w = (w + CFM_PX - 1) / CFM_PX;
    const int grid = n * ((h + CFM_ROWS - 1) / CFM_ROWS) * tiles_w;
    const uint4* wv = reinterpret_cast<const uint4*>(wk);
    // dynamic shared memory: each warp's stage of 32 pixel records
    const int kPad = 4 * 32 * (act_planes(fmt) * cout * static_cast<int>(sizeof(__half)) + 16);
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaFuncSetAttribute(conv_first_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(conv_first_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done.mark();
    }
    if (cout == 64)
        conv_first_tc_kernel<64><<<grid, CFM_PX, kPad, stream>>>(in, n, h, w, wv, oscale, bias, act, slope, fmt, out, skip_lo, *tm_in);
    else
        conv_first_tc_kernel<32><<<grid, CFM_PX, kPad, stream>>>(in, n, h, w, wv, oscale, bias, act, slope, fmt, out, skip_lo, *tm_in);
    return cudaGetLastError();
}

cudaError_t launch_conv_first_mma(const uint8_t* in, int n, int h, int w, const uint32_t* wfrag, const float* oscale,
                                  const float* bias, int cout, int act, float slope, int fmt, __half* out, int skip_lo,
                                  int staging,
                                  const CUtensorMap* tm_in, cudaStream_t stream) {
    if (fmt != ACT_F16_F8) skip_lo = 0;
    const int planes = act_planes(fmt);
    const int tiles_w = (w + CFM_PX - 1) / CFM_PX;
    const int grid = n * ((h + CFM_ROWS - 1) / CFM_ROWS) * tiles_w;
    const size_t dyn = static_cast<size_t>(CFM_PX / 32) * 16 * planes * cout * sizeof(__half);
    if ((w % 16) || (staging == STAGE_TMA && !tm_in)) staging = STAGE_PLAIN;   // bulk variants need 16-byte aligned rows
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
#define CFM_ATTR(C, S) cudaFuncSetAttribute(conv_first_mma_kernel<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)
        CFM_ATTR(64, 0); CFM_ATTR(64, 1); CFM_ATTR(64, 2); CFM_ATTR(32, 0); CFM_ATTR(32, 1); CFM_ATTR(32, 2);
        CFM_ATTR(16, 0); CFM_ATTR(16, 1); CFM_ATTR(16, 2);
#undef CFM_ATTR
        attr_done.mark();
    }
    const uint2* wf = reinterpret_cast<const uint2*>(wfrag);
    static const CUtensorMap no_map = {};
    const CUtensorMap& tm = tm_in ? *tm_in : no_map;
#define CFM_LAUNCH(C, S) \
    conv_first_mma_kernel<C, S><<<grid, CFM_PX, dyn, stream>>>(in, n, h, w, wf, oscale, bias, act, slope, fmt, out, skip_lo, tm)
#define CFM_STAGE(C)                                        \
    switch (staging) {                                      \
        case STAGE_TMA: CFM_LAUNCH(C, STAGE_TMA); break;    \
        case STAGE_CPASYNC: CFM_LAUNCH(C, STAGE_CPASYNC); break; \
        default: CFM_LAUNCH(C, STAGE_PLAIN); break;         \
    }
    switch (cout) {
        case 64: CFM_STAGE(64); break;
        case 32: CFM_STAGE(32); break;
        case 16: CFM_STAGE(16); break;
        default: return cudaErrorInvalidValue;
    }
#undef CFM_STAGE
#undef CFM_LAUNCH
    return cudaGetLastError();
}
