// Persistent bidirectional-LSTM recurrence on tcgen05 (sm_100a).
//
// Replaces the cuDNN RNN call inside the reference's TorchScript recogniser blob (SURVEY.md 2b K3;
// pero_ocr/ocr_engine/pytorch_ocr_engine.py:64-69).  The input projection x W_ih^T + b is one big GEMM done
// beforehand (igemm_tc.cu, fp32 "pre-gates"); this kernel runs the T strictly sequential steps.
//
// One thread-block cluster of 8 CTAs owns 32 lines of one direction for all T steps:
//   * CTA j keeps the W_hh rows of hidden units [32j, 32j+32) (4 gates x 32 units = 128 rows x K=256, fp16 hi
//     (+lo)) resident in TENSOR MEMORY for the whole kernel (tcgen05.st once; 128 columns per plane) as operand A
//     of TS-mode MMAs -- with N = 32 an A operand in shared memory would make every MMA pay a 4 KB smem fetch;
//   * h_{t-1} of the 32 lines (N=32 x K=256, fp16 hi (+lo)) is operand B in the no-swizzle core-matrix layout,
//     double-buffered; one elected thread issues the 16 (x3) tcgen05.mma (M=128,N=32,K=16) of the step into TMEM;
//   * epilogue: tcgen05.ld -> shared-memory transpose so that one thread holds i,f,g,o of (line, 8 units),
//     gates + cell update in fp32 (cell state lives in registers for all T steps), h_t is written to HBM (fp16
//     hi|lo, next layer's GEMM operand) and into the CTA's own slice of the next B buffer (hi and lo planes of a
//     slice contiguous), which is then pushed to the 7 peer CTAs with ONE cp.async.bulk (shared::cta ->
//     shared::cluster) each, completing on the peers' mbarriers: no cluster-wide barrier inside the time loop.
//     (Measured alternatives: plain st.shared::cluster stores + remote mbarrier arrives were 2x slower; separate
//     copies per plane doubled the step time -- a bulk push costs ~0.5 us and pushes serialise.)
#include "lstm_tc.cuh"
#include "actfmt.cuh"
#include "ptx.cuh"

namespace {

constexpr int kH = 256;          // hidden units per direction
constexpr int kCl = 8;           // CTAs per cluster
constexpr int kUnits = kH / kCl; // 32 hidden units per CTA
constexpr int kLines = 32;       // lines per cluster (MMA N)
constexpr int kThreads = 256;
constexpr int kHPlane = kLines * kH * 2;     // 16 KB per plane; a B buffer is [slice j][plane][k/8 - 4j][n/8][n%8][k%8]
constexpr int kSliceBytes = kUnits * kLines * 2;  // 2 KB: one CTA's k range of one plane
constexpr int kGStride = 33;
constexpr int kAccs = 4;          // K split over independent TMEM accumulators (summed in the epilogue)
constexpr int kWCol0 = 128;       // TMEM column of W plane 0 (plane p at kWCol0 + 128 p); accumulator at column 0
constexpr int kTmemCols = 512;

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void bulk_copy_to_peer_pred(uint32_t dst_cluster_addr, uint32_t src_cta_addr,
                                                       uint32_t bytes, uint32_t mbar_cluster_addr, uint32_t pred) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
        "@q cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n" ::"r"(
            dst_cluster_addr),
        "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr), "r"(pred)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = K-direction core-matrix stride, SBO = 8-row-group stride
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
    // tanh(x) = 1 - 2 / (exp(2x) + 1): full fp32 accuracy up to the __expf error (2 ulp), no cancellation blow-up
    // for |x| >~ 1e-2; below that the relative error of the result is still < 1e-5.
    const float e = __expf(2.f * x);
    return 1.f - 2.f * __fdividef(1.f, e + 1.f);
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_kernel(const __half* __restrict__ w_rec, const float* __restrict__ pre, __half* __restrict__ out,
               int n_lines, int T, int planes, int out_fmt, int line_groups) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sH = smem;                                  // 2 buffers * planes * 16 KB
    float* sG = reinterpret_cast<float*>(sH + 2 * planes * kHPlane);   // [128][33]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 128 * kGStride);
    uint64_t* wfull = bars;
    uint64_t* hfull = bars + 1;   // [2]
    uint64_t* mma_done = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = cluster_id_x();
    const int dir = cluster / line_groups;
    const int line0 = (cluster - dir * line_groups) * kLines;
    const int npass = planes == 2 ? 3 : 1;

    if (tid == 0) {
        ptx::mbar_init(wfull, 1);
        ptx::mbar_init(&hfull[0], 1);
        ptx::mbar_init(&hfull[1], 1);
        ptx::mbar_init(mma_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc<kTmemCols>(tmem_slot);
    // h_{-1} = 0: zero both B buffers
    for (int i = tid; i < 2 * planes * kHPlane / 16; i += kThreads)
        reinterpret_cast<uint4*>(sH)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    cluster_sync_all();  // peers' barriers are initialised before anybody signals them
    const uint32_t tmem_base = *tmem_slot;

    // W_hh slice -> TMEM: thread (row = lane quarter * 32 + lane, plane = warp / 4) copies its 256-element row
    // (512 B, packed k-pairs = 128 32-bit columns).
    {
        const int pl = warp >> 2;
        if (pl < planes) {
            const int rowi = (warp & 3) * 32 + lane;
            const uint4* src = reinterpret_cast<const uint4*>(
                w_rec + ((static_cast<size_t>(dir * planes + pl) * kCl + rank) * 128 + rowi) * kH);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kWCol0 + pl * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const uint4 x = __ldg(src + c * 8 + v);
                    r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                }
                ptx::tmem_st_32x32b_x32(taddr + c * 32, r);
            }
            ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncthreads();
        ptx::tc_fence_after();
    }

    // epilogue-2 role of this thread: line nl, units [4*ug, 4*ug+4) of this CTA's 32
    const int nl = tid >> 3, ug = tid & 7;
    const int line = line0 + nl;
    const bool line_ok = line < n_lines;
    const int unit0 = rank * kUnits + ug * 4;  // hidden-unit index within the direction
    float c_state[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c_state[e] = 0.f;
    // byte offset of (line nl, units unit0..unit0+3) of plane 0 inside a B buffer (plane 1 at + kSliceBytes)
    const uint32_t slice_off = rank * planes * kSliceBytes + (ug >> 1) * 512 + (nl >> 3) * 128 + (nl & 7) * 16 + (ug & 1) * 8;
    const uint32_t leader = (warp == 0 && ptx::elect_one()) ? 1u : 0u;
    // The kernel owns all 512 TMEM columns, so the allocation starts at address 0; using the literal keeps every
    // tcgen05 operand warp-uniform (no per-MMA R2UR / BRA.U.ANY sequences in the issue loop).
    if (tmem_base != 0) __trap();
    constexpr uint32_t tmem_u = 0;

    uint32_t hphase0 = 0, hphase1 = 0;
    uint32_t mphase = 0;

    for (int s = 0; s < T; ++s) {
        const int t = dir ? (T - 1 - s) : s;
        const int b = s & 1, nb = b ^ 1;
        // prefetch this step's pre-gates (i,f,g,o x 4 units) while the MMA runs
        float pg[4][4];
        const size_t row = static_cast<size_t>(line_ok ? line : 0) * T + t;
        {
            const float* pr = pre + row * (8 * kH) + dir * 4 * kH + unit0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (line_ok) v0 = __ldg(reinterpret_cast<const float4*>(pr + g * kH));
                pg[g][0] = v0.x; pg[g][1] = v0.y; pg[g][2] = v0.z; pg[g][3] = v0.w;
            }
        }
        if (s > 0) {
            if (warp == 0) {  // whole warp runs the uniform issue code; one elected lane issues
                ptx::mbar_wait(b ? &hfull[1] : &hfull[0], b ? hphase1 : hphase0);
                if (b) hphase1 ^= 1; else hphase0 ^= 1;
                ptx::tc_fence_after();
                constexpr uint32_t idesc = ptx::idesc_f16_f32(128, kLines);
                const uint32_t h_base = ptx::smem_u32(sH) + b * planes * kHPlane;
                if (ptx::elect_one()) {
                    for (int pass = 0; pass < npass; ++pass) {
                        const uint32_t wa = tmem_u + kWCol0 + ((pass == 2) ? 128 : 0);   // W plane (TMEM columns)
                        const uint32_t ha = h_base + ((pass == 1) ? kSliceBytes : 0);    // h plane inside each slice block
#pragma unroll
                        for (int k16 = 0; k16 < 16; ++k16) {
                            const uint64_t b_desc =
                                smem_desc_nosw(ha + (k16 >> 1) * planes * kSliceBytes + (k16 & 1) * 1024, 512, 128);
                            ptx::mma_f16_ts(tmem_u + (k16 & (kAccs - 1)) * kLines, wa + k16 * 8, b_desc, idesc,
                                            (pass != 0 || k16 >= kAccs) ? 1u : 0u);
                        }
                    }
                    ptx::mma_commit(mma_done);
                }
                __syncwarp();
            }
            ptx::mbar_wait(mma_done, mphase);
            mphase ^= 1;
            ptx::tc_fence_after();
            // phase 1: TMEM lane = gate row (lane quarter = gate type, lane = unit), column = line;
            // warps 0-3 take lines 0-15, warps 4-7 lines 16-31
            const int q = warp & 3, half = warp >> 2;
            float sum[16];
            {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(tmem_u + (static_cast<uint32_t>(q * 32) << 16) + half * 16, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(r[j]);
            }
#pragma unroll
            for (int a = 1; a < kAccs; ++a) {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(tmem_u + (static_cast<uint32_t>(q * 32) << 16) + a * kLines + half * 16, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(r[j]);
            }
            float* g_row = sG + (q * 32 + lane) * kGStride + half * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) g_row[j] = sum[j];
            ptx::tc_fence_before();
        }
        __syncthreads();
        // phase 2: gates for (line nl, units 4ug..4ug+3)
        float hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int u = ug * 4 + e;
            float gi = pg[0][e], gf = pg[1][e], gg = pg[2][e], go = pg[3][e];
            if (s > 0) {
                gi += sG[(0 * 32 + u) * kGStride + nl];
                gf += sG[(1 * 32 + u) * kGStride + nl];
                gg += sG[(2 * 32 + u) * kGStride + nl];
                go += sG[(3 * 32 + u) * kGStride + nl];
            }
            const float ig = sigm(gi), fg = sigm(gf), cg = tanh_(gg), og = sigm(go);
            c_state[e] = fg * c_state[e] + ig * cg;
            hv[e] = og * tanh_(c_state[e]);
        }
        uint32_t hi_w[2], lo_w[2];
        float hf4[4];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const __half2 h2 = __floats2half2_rn(hv[e], hv[e + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(hv[e] - hf.x, hv[e + 1] - hf.y);
            hi_w[e >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            lo_w[e >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
            hf4[e] = hf.x;
            hf4[e + 1] = hf.y;
        }
        if (line_ok) {
            // next layer's GEMM operand record (actfmt.cuh): [2H fp16 hi][second plane]
            __half* rec = out + row * (act_planes(out_fmt) * 2 * kH);
            __half* o = rec + dir * kH + unit0;
            *reinterpret_cast<uint2*>(o) = make_uint2(hi_w[0], hi_w[1]);
            if (out_fmt == ACT_F16_HILO) {
                *reinterpret_cast<uint2*>(o + 2 * kH) = make_uint2(lo_w[0], lo_w[1]);
            } else if (out_fmt == ACT_F16_F8) {
                uint8_t* b = reinterpret_cast<uint8_t*>(rec + 2 * kH) + dir * kH + unit0;
                *reinterpret_cast<uint32_t*>(b) = pack_e5m2x4((hv[0] - hf4[0]) * kF8Scale, (hv[1] - hf4[1]) * kF8Scale,
                                                              (hv[2] - hf4[2]) * kF8Scale, (hv[3] - hf4[3]) * kF8Scale);
                *reinterpret_cast<uint32_t*>(b + 2 * kH) = pack_e5m2x4(hf4[0], hf4[1], hf4[2], hf4[3]);
            }
        }
        if (s + 1 < T) {
            // own slice (hi | lo contiguous) of the next B buffer, then ONE bulk push per peer: a cp.async.bulk
            // shared::cta -> shared::cluster costs ~0.5 us and they serialise, so planes share a copy
            uint8_t* hb = sH + nb * planes * kHPlane;
            *reinterpret_cast<uint2*>(hb + slice_off) = make_uint2(hi_w[0], hi_w[1]);
            if (planes == 2) *reinterpret_cast<uint2*>(hb + slice_off + kSliceBytes) = make_uint2(lo_w[0], lo_w[1]);
            fence_proxy_async_smem();
            __syncthreads();
            if (warp == 0) {
                uint64_t* hbar = nb ? &hfull[1] : &hfull[0];
                const uint32_t bytes = planes * kSliceBytes;
                ptx::mbar_expect_tx_pred(hbar, (kCl - 1) * bytes, leader);
                const uint32_t bar = ptx::smem_u32(hbar);
                const uint32_t src = ptx::smem_u32(hb + rank * bytes);
#pragma unroll
                for (uint32_t d = 1; d < kCl; ++d) {
                    const uint32_t peer = (rank + d) & (kCl - 1);
                    bulk_copy_to_peer_pred(mapa(src, peer), src, bytes, mapa(bar, peer), leader);
                }
            }
        }
    }

    ptx::tc_fence_before();
    cluster_sync_all();  // nobody leaves while a peer may still push into its shared memory
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_u);
    }
}

}  // namespace

size_t lstm_tc_smem_bytes(int planes) {
    return 2 * static_cast<size_t>(planes) * kHPlane +
           128 * kGStride * sizeof(float) + 64 + 1024;
}

cudaError_t launch_lstm_tc(const __half* w_rec, const float* pre, __half* out, int n_lines, int T, int H, int planes,
                           int out_fmt, cudaStream_t stream) {
    if (H != kH) return cudaErrorInvalidValue;
    const size_t smem = lstm_tc_smem_bytes(planes);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int line_groups = (n_lines + kLines - 1) / kLines;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * line_groups * kCl);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lstm_tc_kernel, w_rec, pre, out, n_lines, T, planes, out_fmt, line_groups);
}
