// Persistent bidirectional-LSTM recurrence on tcgen05 (sm_100a).
//
// Replaces the cuDNN RNN call inside the reference's TorchScript recogniser blob (SURVEY.md 2b K3;
// pero_ocr/ocr_engine/pytorch_ocr_engine.py:64-69).  The input projection x W_ih^T + b is one big GEMM done
// beforehand (igemm_tc.cu, fp32 "pre-gates"); this kernel runs the T strictly sequential steps.
//
// One thread-block cluster of 8 CTAs owns TWO independent groups of 32 lines of one direction for all T steps; each
// 256-thread half of a CTA runs one group's recurrence with its own shared-memory buffers, mbarriers and TMEM
// accumulators, so that one group's exchange latency hides behind the other group's MMA / gate work, and 256 lines
// x 2 directions need only 8 clusters (16 clusters of 8 do not fit the 8 GPCs of a B200 at once: the second wave
// doubled the kernel time in the one-group-per-cluster version, profiles/r01c_lstm_phases.log):
//   * CTA j keeps the W_hh rows of hidden units [32j, 32j+32) (4 gates x 32 units = 128 rows x K=256, fp16 hi
//     (+lo)) resident in TENSOR MEMORY for the whole kernel (tcgen05.st once; 128 columns per plane) as operand A
//     of TS-mode MMAs, shared by both groups -- with N = 32 an A operand in shared memory would make every MMA pay
//     a 4 KB smem fetch;
//   * h_{t-1} of a group's 32 lines (N=32 x K=256, fp16 hi (+lo)) is operand B in the no-swizzle core-matrix
//     layout, double-buffered; one elected thread per group issues the 16 (x3) tcgen05.mma (M=128,N=32,K=16) of the
//     step into the group's TMEM accumulators;
//   * epilogue: tcgen05.ld -> shared-memory transpose so that one thread holds i,f,g,o of (line, 4 units),
//     gates + cell update in fp32 (cell state lives in registers for all T steps), h_t is written to HBM (fp16
//     hi|lo, next layer's GEMM operand) and into the CTA's own slice of the next B buffer (hi and lo planes of a
//     slice contiguous), which is then pushed to the 7 peer CTAs with ONE cp.async.bulk (shared::cta ->
//     shared::cluster) each, completing on the peers' mbarriers: no cluster-wide barrier inside the time loop.
//     (Measured alternatives: plain st.shared::cluster stores + remote mbarrier arrives were 2x slower; separate
//     copies per plane doubled the step time; staging h_t in L2 and bringing it back with one multicast
//     cp.async.bulk per CTA cost 370 cycles/step more than the pushes -- the generic->async proxy fence on global
//     memory alone is ~900 cycles.)
#include "once.cuh"
#include "lstm_tc.cuh"
#include "actfmt.cuh"
#include "ptx.cuh"

namespace {

constexpr int kH = 256;          // hidden units per direction
constexpr int kCl = 8;           // CTAs per cluster
constexpr int kUnits = kH / kCl; // 32 hidden units per CTA
constexpr int kLines = 32;       // lines per group (MMA N)
constexpr int kGroups = 2;       // independent line groups per cluster, one per 256-thread half of a CTA
constexpr int kGThreads = 256;
constexpr int kThreads = kGroups * kGThreads;
constexpr int kHPlane = kLines * kH * 2;     // 16 KB per plane; a B buffer is [slice j][plane][k/8 - 4j][n/8][n%8][k%8]
constexpr int kSliceBytes = kUnits * kLines * 2;  // 2 KB: one CTA's k range of one plane
constexpr int kGStride = 33;
constexpr int kAccs = 4;          // K split over independent TMEM accumulators (summed in the epilogue)
constexpr int kWCol0 = 128;       // TMEM columns: group-0 accumulators [0,128), W planes [128,384), group-1 acc [384,512)
constexpr int kAccCol1 = 384;
constexpr int kTmemCols = 512;
constexpr int kBarsPerGroup = 4;   // hfull[2 buffers], mma_done, pad

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

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
    // volatile: keeps the prefetch where it is written (ahead of the mbarrier wait), not sunk to its first use
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <bool DBG>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_kernel(const __half* __restrict__ w_rec, const float* __restrict__ pre, __half* __restrict__ out,
               int n_lines, int T, int planes, int hplanes, int out_fmt, int pair_groups, long long* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 8;              // which line group this half of the CTA serves
    const int gt = tid & (kGThreads - 1);  // thread index within the half
    const int gwarp = gt >> 5;
    const int group_bytes = 2 * hplanes * kHPlane;     // two B buffers
    uint8_t* sH = smem + grp * group_bytes;
    float* sG = reinterpret_cast<float*>(smem + kGroups * group_bytes) + grp * 128 * kGStride;   // [128][33]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGroups * group_bytes + kGroups * 128 * kGStride * 4);
    uint64_t* hfull = bars + grp * kBarsPerGroup;      // [2 buffers]
    uint64_t* mma_done = hfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBarsPerGroup * kGroups);

    const uint32_t rank = cluster_ctarank();
    const int cluster = cluster_id_x();
    const int dir = cluster / pair_groups;
    const int line0 = ((cluster - dir * pair_groups) * kGroups + grp) * kLines;
    const bool active = line0 < n_lines;   // the same in every CTA of the cluster
    // planes = fp16 planes of W_hh in TMEM (hi, lo), hplanes = planes of h_t that are exchanged and multiplied:
    //   (2, 2) W_hi h_hi + W_hi h_lo + W_lo h_hi   (fp16x3)
    //   (2, 1) W_hi h_hi + W_lo h_hi               h_t rounded to fp16 once per step, weights exact to ~2^-22: the
    //          logits move by 3e-5 (tools/emulate_mixed_precision.py) and the exchange -- the kernel's bound -- halves
    //   (1, 1) W_hi h_hi                           (fp16)
    const int npass = planes == 2 ? (hplanes == 2 ? 3 : 2) : 1;

    if (tid == 0) {
        for (int i = 0; i < kGroups * kBarsPerGroup; ++i) ptx::mbar_init(&bars[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc<kTmemCols>(tmem_slot);
    // h_{-1} = 0: zero every B buffer
    for (int i = tid; i < kGroups * group_bytes / 16; i += kThreads)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    cluster_sync_all();  // peers' barriers are initialised before anybody signals them
    const uint32_t tmem_base = *tmem_slot;

    // W_hh slice -> TMEM: thread (row = lane quarter * 32 + lane, plane = warp / 4) of the first half copies its
    // 256-element row (512 B, packed k-pairs = 128 32-bit columns).
    {
        const int pl = warp >> 2;
        if (pl < planes) {
            const int rowi = (warp & 3) * 32 + lane;
            const uint4* src = reinterpret_cast<const uint4*>(
                w_rec + ((static_cast<size_t>(dir * planes + pl) * kCl + rank) * 128 + rowi) * kH);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kWCol0 + pl * 128;
#pragma unroll 1
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
    // The kernel owns all 512 TMEM columns, so the allocation starts at address 0; literals keep every tcgen05
    // operand warp-uniform (no per-MMA R2UR / BRA.U.ANY sequences in the issue loop).
    if (tmem_base != 0) __trap();
    const uint32_t acc_u = grp ? kAccCol1 : 0;

    // epilogue-2 role of this thread: line nl, units [4*ug, 4*ug+4) of this CTA's 32
    const int nl = gt >> 3, ug = gt & 7;
    const int line = line0 + nl;
    const bool line_ok = line < n_lines;
    const int unit0 = rank * kUnits + ug * 4;  // hidden-unit index within the direction
    float c_state[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c_state[e] = 0.f;
    // byte offset of (line nl, units unit0..unit0+3) of plane 0 inside a B buffer (plane 1 at + kSliceBytes)
    const uint32_t slice_off = rank * hplanes * kSliceBytes + (ug >> 1) * 512 + (nl >> 3) * 128 + (nl & 7) * 16 + (ug & 1) * 8;
    uint32_t mphase = 0;

    // optional per-phase cycle counters of CTA 0 / thread 0 (bring-up: B200OCR_LSTM_DBG=1)
    long long tacc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0;
#define PROBE(i)                                   \
    if (DBG && tid == 0) {                         \
        const long long now_ = clock64();          \
        tacc[i] += now_ - tprev;                   \
        tprev = now_;                              \
    }
    if (DBG && tid == 0) tprev = clock64();

    const float* pre_base = pre + static_cast<size_t>(line_ok ? line : 0) * T * (8 * kH) + dir * 4 * kH + unit0;
    for (int s = 0; active && s < T; ++s) {
        const int t = dir ? (T - 1 - s) : s;
        const int b = s & 1, nb = b ^ 1;
        // prefetch this step's pre-gates (i,f,g,o x 4 units) while the peers' h_{t-1} arrives and the MMA runs
        float4 pg[4];
        {
            const float* pr = pre_base + static_cast<size_t>(t) * (8 * kH);
#pragma unroll
            for (int g = 0; g < 4; ++g) pg[g] = line_ok ? ld_nc_f4(pr + g * kH) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (s > 0) {
            if (gwarp == 0) {  // whole warp runs the uniform issue code; one elected lane issues
                ptx::mbar_wait(&hfull[b], ((s - 1) >> 1) & 1);
                PROBE(0);
                ptx::tc_fence_after();
                constexpr uint32_t idesc = ptx::idesc_f16_f32(128, kLines);
                const uint32_t h_base = ptx::smem_u32(sH) + b * hplanes * kHPlane;
                if (ptx::elect_one()) {
                    // fully unrolled with compile-time offsets: runtime-indexed descriptors cost ~100 cycles per MMA in
                    // uniform-datapath address arithmetic (profiles/r01c_lstm_phases.log)
                    for (int pass = 0; pass < npass; ++pass) {
                        const bool w_lo = hplanes == 2 ? pass == 2 : pass == 1;
                        const uint32_t wa = kWCol0 + (w_lo ? 128 : 0);                   // W plane (TMEM columns)
                        const uint32_t ha = h_base + ((hplanes == 2 && pass == 1) ? kSliceBytes : 0);   // h plane inside each slice block
#pragma unroll
                        for (int k16 = 0; k16 < 16; ++k16) {
                            const uint64_t b_desc =
                                smem_desc_nosw(ha + (k16 >> 1) * hplanes * kSliceBytes + (k16 & 1) * 1024, 512, 128);
                            ptx::mma_f16_ts(acc_u + (k16 & (kAccs - 1)) * kLines, wa + k16 * 8, b_desc, idesc,
                                            (pass != 0 || k16 >= kAccs) ? 1u : 0u);
                        }
                    }
                    ptx::mma_commit(mma_done);
                }
                __syncwarp();
            }
            ptx::mbar_wait(mma_done, mphase);
            mphase ^= 1;
            PROBE(1);
            ptx::tc_fence_after();
            // phase 1: TMEM lane = gate row (lane quarter = gate type, lane = unit), column = line;
            // warps 0-3 of the half take lines 0-15, warps 4-7 lines 16-31
            const int q = gwarp & 3, half = gwarp >> 2;
            float sum[16];
            {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(acc_u + (static_cast<uint32_t>(q * 32) << 16) + half * 16, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(r[j]);
            }
#pragma unroll
            for (int a = 1; a < kAccs; ++a) {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(acc_u + (static_cast<uint32_t>(q * 32) << 16) + a * kLines + half * 16, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(r[j]);
            }
            float* g_row = sG + (q * 32 + lane) * kGStride + half * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) g_row[j] = sum[j];
            ptx::tc_fence_before();
        }
        ptx::named_bar_sync(1 + grp, kGThreads);
        PROBE(2);
        // phase 2: gates for (line nl, units 4ug..4ug+3)
        float hv[4];
        {
            const float pgi[4] = {pg[0].x, pg[0].y, pg[0].z, pg[0].w};
            const float pgf[4] = {pg[1].x, pg[1].y, pg[1].z, pg[1].w};
            const float pgg[4] = {pg[2].x, pg[2].y, pg[2].z, pg[2].w};
            const float pgo[4] = {pg[3].x, pg[3].y, pg[3].z, pg[3].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = ug * 4 + e;
                float gi = pgi[e], gf = pgf[e], gg = pgg[e], go = pgo[e];
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
        if (s + 1 < T) {
            // own slice (hi | lo contiguous) of the next B buffer, then ONE bulk push per peer: a cp.async.bulk
            // shared::cta -> shared::cluster costs ~0.5 us and they serialise, so planes share a copy
            uint8_t* hb = sH + nb * hplanes * kHPlane;
            *reinterpret_cast<uint2*>(hb + slice_off) = make_uint2(hi_w[0], hi_w[1]);
            if (hplanes == 2) *reinterpret_cast<uint2*>(hb + slice_off + kSliceBytes) = make_uint2(lo_w[0], lo_w[1]);
            fence_proxy_async_smem();
            ptx::named_bar_sync(1 + grp, kGThreads);
            PROBE(3);
            if (gwarp == 0) {
                // one thread issues the seven pushes back to back; issued from seven warps at once, or completing on
                // seven per-slice mbarriers so that the MMAs of a slice can start as it lands, they were slower: with
                // two groups per SM the exchange runs at the SM-to-SM network's ~20 B/clk per SM
                // (profiles/r01c_lstm_phases.log)
                const uint32_t bytes = hplanes * kSliceBytes;
                const uint32_t lead = ptx::elect_one() ? 1u : 0u;
                ptx::mbar_expect_tx_pred(&hfull[nb], (kCl - 1) * bytes, lead);
                const uint32_t bar = ptx::smem_u32(&hfull[nb]);
                const uint32_t src = ptx::smem_u32(hb + rank * bytes);
#pragma unroll
                for (uint32_t d = 1; d < kCl; ++d) {
                    const uint32_t peer = (rank + d) & (kCl - 1);
                    bulk_copy_to_peer_pred(mapa(src, peer), src, bytes, mapa(bar, peer), lead);
                }
            }
            PROBE(4);
        }
        if (line_ok) {
            // next layer's GEMM operand record (actfmt.cuh): [2H fp16 hi][second plane]; off the critical path
            // (after the pushes are on their way)
            const size_t row = static_cast<size_t>(line) * T + t;
            __half* rec = out + row * (act_planes(out_fmt) * 2 * kH);
            __half* o = rec + dir * kH + unit0;
            *reinterpret_cast<uint2*>(o) = make_uint2(hi_w[0], hi_w[1]);
            if (out_fmt == ACT_F16_HILO) {
                *reinterpret_cast<uint2*>(o + 2 * kH) = make_uint2(lo_w[0], lo_w[1]);
            } else if (out_fmt == ACT_F16_F8) {
                uint8_t* b8 = reinterpret_cast<uint8_t*>(rec + 2 * kH) + dir * kH + unit0;
                *reinterpret_cast<uint32_t*>(b8) = pack_e5m2x4((hv[0] - hf4[0]) * kF8Scale, (hv[1] - hf4[1]) * kF8Scale,
                                                               (hv[2] - hf4[2]) * kF8Scale, (hv[3] - hf4[3]) * kF8Scale);
                *reinterpret_cast<uint32_t*>(b8 + 2 * kH) = pack_e5m2x4(hf4[0], hf4[1], hf4[2], hf4[3]);
            }
        }
        PROBE(5);
    }
    if (DBG && tid == 0 && blockIdx.x == 0)
        for (int i = 0; i < 9; ++i) dbg[i] = tacc[i];
#undef PROBE

    ptx::tc_fence_before();
    cluster_sync_all();  // nobody leaves while a peer may still push into its shared memory
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(0u);
    }
}

}  // namespace

size_t lstm_tc_smem_bytes(int hplanes) {
    return kGroups * (2 * static_cast<size_t>(hplanes) * kHPlane + 128 * kGStride * sizeof(float)) + 128 + 1024;
}

cudaError_t launch_lstm_tc(const __half* w_rec, const float* pre, __half* out, int n_lines, int T, int H, int planes,
                           int hplanes, int out_fmt, cudaStream_t stream) {
    if (H != kH || hplanes < 1 || hplanes > planes) return cudaErrorInvalidValue;
    const size_t smem = lstm_tc_smem_bytes(hplanes);
    static PerDeviceOnce init;
    static long long* dbg = nullptr;   // bring-up only (B200OCR_LSTM_DBG=1): per-phase cycle counters, synchronises
    if (init.pending()) {
        cudaError_t e = cudaFuncSetAttribute(lstm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lstm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (getenv("B200OCR_LSTM_DBG") && !dbg) cudaMalloc(reinterpret_cast<void**>(&dbg), 9 * sizeof(long long));
        init.mark();
    }
    const int pair_groups = (n_lines + kGroups * kLines - 1) / (kGroups * kLines);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pair_groups * kCl);
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
    if (!dbg)
        return cudaLaunchKernelEx(&cfg, lstm_tc_kernel<false>, w_rec, pre, out, n_lines, T, planes, hplanes, out_fmt, pair_groups,
                                  dbg);
    int max_clusters = -1;
    cudaOccupancyMaxActiveClusters(&max_clusters, lstm_tc_kernel<true>, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc_kernel<true>, w_rec, pre, out, n_lines, T, planes, hplanes, out_fmt,
                                       pair_groups, dbg);
    if (e == cudaSuccess) {
        long long h[9];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr,
                "lstm_tc T=%d clusters %d (max co-resident %d) cycles/step: wait_h %.0f mma %.0f tmem_ld %.0f "
                "gates+fence+sync %.0f push_issue %.0f store %.0f\n",
                T, 2 * pair_groups, max_clusters, (double)h[0] / T, (double)h[1] / T, (double)h[2] / T,
                (double)h[3] / T, (double)h[4] / T, (double)h[5] / T);
    }
    return e;
}
