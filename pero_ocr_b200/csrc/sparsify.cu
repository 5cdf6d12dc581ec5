// Logit sparsification on the device.
//
// Replaces the per-line NumPy pass of BaseEngineLineOCR.process_lines
//     line_probs = softmax(line_logits, axis=1); line_logits[line_probs < 0.0001] = 0
//     line_logits = sparse.csc_matrix(line_logits)            (pero_ocr/ocr_engine/line_ocr_engine.py:168-172)
// including the optional tight crop of the frame range (:152-156), so that PageOCR.process_page
// (document_ocr/page_parser.py:423-430), which always asks for logits, ships a few CSC entries per frame instead of
// the dense [T][C] fp32 matrix and runs no per-line softmax on the host.
//
// Output is scipy's canonical CSC of the [hi-lo][C] matrix of every line, concatenated:
//   indptr  i32 [n][C+1]   per-line column pointers (relative to the line's first entry)
//   base    i64 [n+1]      first entry of each line in indices / data (base[n] = total entries)
//   indices i32 [total]    frame index (relative to the line's lo), ascending inside a column
//   data    f32 [total]    the raw logit (bit pattern of the input)
// Semantics mirrored exactly: softmax in fp32 as p = exp(v - max) / sum; an entry is dropped when
// (double)p < 0.0001 (NumPy compares the fp32 probability with a Python float) or when the raw logit is 0.0 (CSC
// never stores zeros -- the reference loses genuine zero logits by design, core/layout.py:65-68); NaN survives.
// exp and the order of the sum are CUDA's, not NumPy's: an entry whose probability is within a few fp32 ulps of
// 1e-4 may fall on the other side (tests/test_gpu_sparsify.py bounds that band).
//
// Tiled path (the default): one CTA per (line, chunk of 32 frames).
//   mask:   the chunk's [32][C] logits are read ONCE, coalesced, into shared memory; per-frame max / sum from there
//           (same sequential order as before: bit-identical decisions), then thread c builds the 32-bit word whose
//           bit f says "entry (frame 32k + f, class c) survives" -> colmask[line][k][c] (stream-ordered scratch);
//   indptr: CTA per line, thread per class: popcounts over the chunks, exclusive scan over classes;
//   base:   one-CTA scan over lines;
//   fill:   CTA per (line, chunk) again: the tile is re-read coalesced, thread c starts at base + indptr[c] + the
//           popcount of its earlier chunks and writes its surviving entries in frame order.
// HBM traffic: the logits twice, coalesced, + the entries; 2816 CTAs at config 2 instead of 256 CTAs walking 336
// frames serially with 120 of 256 threads (round 1: 0.31 ms per batch = 0.07 of the copy bandwidth).
// The round-1 kernels (one CTA per line) remain as the fallback for class counts whose tile does not fit shared
// memory.
#include "once.cuh"
#include "kernels.cuh"

namespace {

constexpr int SP_THREADS = 256;

__device__ __forceinline__ bool sp_keep(float v, float mx, float sum) {
    const float p = __fdiv_rn(expf(v - mx), sum);
    return !(static_cast<double>(p) < 0.0001) && v != 0.0f;
}

__device__ void sp_frame_stats(const float* __restrict__ L, int lo, int hi, int C, float* s_mx, float* s_sum) {
    for (int t = lo + threadIdx.x; t < hi; t += SP_THREADS) {
        const float* row = L + static_cast<size_t>(t) * C;
        float mx = row[0];
        bool nan = mx != mx;
        for (int c = 1; c < C; ++c) {
            const float v = row[c];
            if (v != v) nan = true;
            mx = fmaxf(mx, v);
        }
        if (nan) mx = __int_as_float(0x7fc00000);   // np.max propagates NaN
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        s_mx[t - lo] = mx;
        s_sum[t - lo] = sum;
    }
}

__global__ void __launch_bounds__(SP_THREADS) sparsify_count_kernel(const float* __restrict__ logits, int T, int C,
                                                                    const int32_t* __restrict__ t_lo,
                                                                    const int32_t* __restrict__ t_hi,
                                                                    int32_t* __restrict__ indptr,
                                                                    int32_t* __restrict__ nnz) {
    extern __shared__ float s_dyn[];
    const int line = blockIdx.x;
    const int lo = t_lo ? max(0, min(T, t_lo[line])) : 0;
    const int hi = t_hi ? max(lo, min(T, t_hi[line])) : T;
    float* s_mx = s_dyn;
    float* s_sum = s_dyn + T;
    int32_t* s_cnt = reinterpret_cast<int32_t*>(s_dyn + 2 * T);   // [C]
    const float* L = logits + static_cast<size_t>(line) * T * C;
    sp_frame_stats(L, lo, hi, C, s_mx, s_sum);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += SP_THREADS) {
        int cnt = 0;
        for (int t = lo; t < hi; ++t) cnt += sp_keep(L[static_cast<size_t>(t) * C + c], s_mx[t - lo], s_sum[t - lo]) ? 1 : 0;
        s_cnt[c] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t* ip = indptr + static_cast<size_t>(line) * (C + 1);
        int run = 0;
        for (int c = 0; c < C; ++c) {
            ip[c] = run;
            run += s_cnt[c];
        }
        ip[C] = run;
        nnz[line] = run;
    }
}

__global__ void sparsify_scan_kernel(const int32_t* __restrict__ nnz, int n, int64_t* __restrict__ base) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < n; ++i) {
            base[i] = run;
            run += nnz[i];
        }
        base[n] = run;
    }
}

__global__ void __launch_bounds__(SP_THREADS) sparsify_fill_kernel(const float* __restrict__ logits, int T, int C,
                                                                   const int32_t* __restrict__ t_lo,
                                                                   const int32_t* __restrict__ t_hi,
                                                                   const int32_t* __restrict__ indptr,
                                                                   const int64_t* __restrict__ base, int64_t capacity,
                                                                   int32_t* __restrict__ indices,
                                                                   float* __restrict__ data) {
    extern __shared__ float s_dyn[];
    const int line = blockIdx.x;
    const int lo = t_lo ? max(0, min(T, t_lo[line])) : 0;
    const int hi = t_hi ? max(lo, min(T, t_hi[line])) : T;
    float* s_mx = s_dyn;
    float* s_sum = s_dyn + T;
    const float* L = logits + static_cast<size_t>(line) * T * C;
    sp_frame_stats(L, lo, hi, C, s_mx, s_sum);
    __syncthreads();
    const int32_t* ip = indptr + static_cast<size_t>(line) * (C + 1);
    const int64_t b0 = base[line];
    for (int c = threadIdx.x; c < C; c += SP_THREADS) {
        int64_t at = b0 + ip[c];
        for (int t = lo; t < hi; ++t) {
            const float v = L[static_cast<size_t>(t) * C + c];
            if (sp_keep(v, s_mx[t - lo], s_sum[t - lo])) {
                if (at < capacity) {
                    indices[at] = t - lo;
                    data[at] = v;
                }
                ++at;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- tiled path
constexpr int ST_THREADS = 128;
constexpr int ST_FRAMES = 32;

// tile[f][c] (row pitch C + 1 + ((C & 1) ^ 1): odd, conflict-free column walks), s_mx / s_sum [32]
__device__ __forceinline__ int st_pitch(int C) { return (C + 1) | 1; }

__device__ __forceinline__ int st_load_tile(const float* __restrict__ logits, int T, int C, const int32_t* t_lo,
                                            const int32_t* t_hi, int line, int k, float* tile, int* t0_out) {
    const int lo = t_lo ? max(0, min(T, t_lo[line])) : 0;
    const int hi = t_hi ? max(lo, min(T, t_hi[line])) : T;
    const int t0 = lo + k * ST_FRAMES;
    const int nf = max(0, min(ST_FRAMES, hi - t0));
    *t0_out = t0 - lo;                           // frame index of the chunk's first row relative to the line's range
    const float* src = logits + (static_cast<size_t>(line) * T + t0) * C;
    const int pitch = st_pitch(C);
    for (int i = threadIdx.x; i < nf * C; i += ST_THREADS) {
        const int f = i / C, c = i - f * C;
        tile[f * pitch + c] = src[i];
    }
    return nf;
}

__global__ void __launch_bounds__(ST_THREADS) sparsify_mask_kernel(const float* __restrict__ logits, int T, int C,
                                                                   const int32_t* __restrict__ t_lo,
                                                                   const int32_t* __restrict__ t_hi, int chunks,
                                                                   uint32_t* __restrict__ colmask) {
    extern __shared__ float s_dyn[];
    const int line = blockIdx.y, k = blockIdx.x;
    const int pitch = st_pitch(C);
    float* tile = s_dyn;
    float* s_mx = s_dyn + ST_FRAMES * pitch;
    float* s_sum = s_mx + ST_FRAMES;
    int rel0;
    const int nf = st_load_tile(logits, T, C, t_lo, t_hi, line, k, tile, &rel0);
    __syncthreads();
    if (threadIdx.x < nf) {
        const float* row = tile + threadIdx.x * pitch;
        float mx = row[0];
        bool nan = mx != mx;
        for (int c = 1; c < C; ++c) {
            const float v = row[c];
            if (v != v) nan = true;
            mx = fmaxf(mx, v);
        }
        if (nan) mx = __int_as_float(0x7fc00000);   // np.max propagates NaN
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        s_mx[threadIdx.x] = mx;
        s_sum[threadIdx.x] = sum;
    }
    __syncthreads();
    uint32_t* out = colmask + (static_cast<size_t>(line) * chunks + k) * C;
    for (int c = threadIdx.x; c < C; c += ST_THREADS) {
        uint32_t word = 0;
        for (int f = 0; f < nf; ++f)
            if (sp_keep(tile[f * pitch + c], s_mx[f], s_sum[f])) word |= 1u << f;
        out[c] = word;
    }
}

__global__ void __launch_bounds__(ST_THREADS) sparsify_indptr_kernel(const uint32_t* __restrict__ colmask, int chunks,
                                                                     int C, int32_t* __restrict__ indptr,
                                                                     int32_t* __restrict__ nnz) {
    extern __shared__ int32_t s_cnt[];               // [C]
    const int line = blockIdx.x;
    const uint32_t* m = colmask + static_cast<size_t>(line) * chunks * C;
    for (int c = threadIdx.x; c < C; c += ST_THREADS) {
        int cnt = 0;
        for (int k = 0; k < chunks; ++k) cnt += __popc(m[static_cast<size_t>(k) * C + c]);
        s_cnt[c] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t* ip = indptr + static_cast<size_t>(line) * (C + 1);
        int run = 0;
        for (int c = 0; c < C; ++c) {
            ip[c] = run;
            run += s_cnt[c];
        }
        ip[C] = run;
        nnz[line] = run;
    }
}

__global__ void __launch_bounds__(ST_THREADS) sparsify_fill_tiled_kernel(const float* __restrict__ logits, int T, int C,
                                                                         const int32_t* __restrict__ t_lo,
                                                                         const int32_t* __restrict__ t_hi, int chunks,
                                                                         const uint32_t* __restrict__ colmask,
                                                                         const int32_t* __restrict__ indptr,
                                                                         const int64_t* __restrict__ base,
                                                                         int64_t capacity, int32_t* __restrict__ indices,
                                                                         float* __restrict__ data) {
    extern __shared__ float s_dyn[];
    const int line = blockIdx.y, k = blockIdx.x;
    const int pitch = st_pitch(C);
    float* tile = s_dyn;
    int rel0;
    const int nf = st_load_tile(logits, T, C, t_lo, t_hi, line, k, tile, &rel0);
    __syncthreads();
    if (nf == 0) return;
    const uint32_t* m = colmask + static_cast<size_t>(line) * chunks * C;
    const int32_t* ip = indptr + static_cast<size_t>(line) * (C + 1);
    const int64_t b0 = base[line];
    // thread per class: where this chunk's run of column c starts; then warp per column: lane f owns frame f, so a
    // column's surviving entries (consecutive in the CSC arrays) leave in ONE coalesced store per array
    uint32_t* s_word = reinterpret_cast<uint32_t*>(tile + ST_FRAMES * pitch);        // [C]
    long long* s_at = reinterpret_cast<long long*>(s_word + ((C + 1) & ~1));          // [C], 8-byte aligned
    for (int c = threadIdx.x; c < C; c += ST_THREADS) {
        int before = 0;
        for (int kk = 0; kk < k; ++kk) before += __popc(m[static_cast<size_t>(kk) * C + c]);
        s_word[c] = m[static_cast<size_t>(k) * C + c];
        s_at[c] = b0 + ip[c] + before;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < C; c += ST_THREADS / 32) {
        const uint32_t word = s_word[c];
        if (!((word >> lane) & 1u)) continue;
        const long long at = s_at[c] + __popc(word & ((1u << lane) - 1u));
        if (at < capacity) {
            indices[at] = rel0 + lane;
            data[at] = tile[lane * pitch + c];
        }
    }
}

// What a consumer of TextLine.logits sees after the sparsify -> CSC -> get_full_logprobs round trip
// (line_ocr_engine.py:168-172, core/layout.py:65-72), computed straight from the dense logits on the device: entries
// the sparsification drops (softmax p < 1e-4, or a raw 0.0) become -80, then a float32 log-softmax per frame (the
// reference's x - np.logaddexp.reduce(x) of core/layout.py:32-34, evaluated as x - max - log(sum(exp(x - max))): equal
// up to float32 rounding, ~1e-6); written as float64, the decoders' working type.
__global__ void full_logprobs_kernel(const float* __restrict__ logits, long frames, int C, double* __restrict__ out) {
    const long f = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    const float* row = logits + f * C;
    double* o = out + f * C;
    float mx = row[0];
    bool nan = mx != mx;
    for (int c = 1; c < C; ++c) {
        const float v = row[c];
        if (v != v) nan = true;
        mx = fmaxf(mx, v);
    }
    if (nan) mx = __int_as_float(0x7fc00000);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
    // second softmax: over the densified values
    float mx2 = -INFINITY;
    bool nan2 = false;
    for (int c = 0; c < C; ++c) {
        const float v = row[c];
        const float d = sp_keep(v, mx, sum) ? v : -80.f;
        if (d != d) nan2 = true;
        mx2 = fmaxf(mx2, d);
    }
    if (nan2) mx2 = __int_as_float(0x7fc00000);
    float sum2 = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = row[c];
        const float d = sp_keep(v, mx, sum) ? v : -80.f;
        sum2 += expf(d - mx2);
    }
    const float lse = logf(sum2);
    for (int c = 0; c < C; ++c) {
        const float v = row[c];
        const float d = sp_keep(v, mx, sum) ? v : -80.f;
        o[c] = static_cast<double>((d - mx2) - lse);
    }
}

// Tiled variant: CTA per chunk of 32 frames of the flattened [frames][C] matrix.  The chunk is read once, coalesced,
// into shared memory; thread f < 32 makes the sequential passes of frame f there (same order as the kernel above: the
// results are bit-identical); then all threads write the chunk's doubles in memory order (coalesced).
__global__ void __launch_bounds__(ST_THREADS) full_logprobs_tiled_kernel(const float* __restrict__ logits, long frames,
                                                                         int C, double* __restrict__ out) {
    extern __shared__ float s_dyn[];
    const int pitch = st_pitch(C);
    float* tile = s_dyn;
    float* s_mx = s_dyn + ST_FRAMES * pitch;      // [32] max of the raw frame
    float* s_sum = s_mx + ST_FRAMES;              // [32] sum of exp(raw - max)
    float* s_mx2 = s_sum + ST_FRAMES;             // [32] max of the densified frame
    float* s_lse = s_mx2 + ST_FRAMES;             // [32] log of the densified frame's sum
    const long f0 = static_cast<long>(blockIdx.x) * ST_FRAMES;
    const int nf = static_cast<int>(min(static_cast<long>(ST_FRAMES), frames - f0));
    const float* src = logits + f0 * C;
    for (int i = threadIdx.x; i < nf * C; i += ST_THREADS) {
        const int f = i / C, c = i - f * C;
        tile[f * pitch + c] = src[i];
    }
    __syncthreads();
    if (threadIdx.x < nf) {
        const float* row = tile + threadIdx.x * pitch;
        float mx = row[0];
        bool nan = mx != mx;
        for (int c = 1; c < C; ++c) {
            const float v = row[c];
            if (v != v) nan = true;
            mx = fmaxf(mx, v);
        }
        if (nan) mx = __int_as_float(0x7fc00000);
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        float mx2 = -INFINITY;
        bool nan2 = false;
        for (int c = 0; c < C; ++c) {
            const float v = row[c];
            const float d = sp_keep(v, mx, sum) ? v : -80.f;
            if (d != d) nan2 = true;
            mx2 = fmaxf(mx2, d);
        }
        if (nan2) mx2 = __int_as_float(0x7fc00000);
        float sum2 = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = row[c];
            const float d = sp_keep(v, mx, sum) ? v : -80.f;
            sum2 += expf(d - mx2);
        }
        s_mx[threadIdx.x] = mx;
        s_sum[threadIdx.x] = sum;
        s_mx2[threadIdx.x] = mx2;
        s_lse[threadIdx.x] = logf(sum2);
    }
    __syncthreads();
    double* dst = out + f0 * C;
    for (int i = threadIdx.x; i < nf * C; i += ST_THREADS) {
        const int f = i / C, c = i - f * C;
        const float v = tile[f * pitch + c];
        const float d = sp_keep(v, s_mx[f], s_sum[f]) ? v : -80.f;
        dst[i] = static_cast<double>((d - s_mx2[f]) - s_lse[f]);
    }
}

}  // namespace

cudaError_t launch_full_logprobs(const float* logits, int n, int T, int C, double* out, cudaStream_t stream) {
    const long frames = static_cast<long>(n) * T;
    if (frames <= 0) return cudaSuccess;
    const size_t dyn = (static_cast<size_t>(ST_FRAMES) * ((C + 1) | 1) + 4 * ST_FRAMES) * sizeof(float);
    if (dyn <= 200 * 1024) {
        static PerDeviceOnce attr_done;
        if (attr_done.pending()) {
            cudaFuncSetAttribute(full_logprobs_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_done.mark();
        }
        full_logprobs_tiled_kernel<<<static_cast<unsigned>((frames + ST_FRAMES - 1) / ST_FRAMES), ST_THREADS, dyn, stream>>>(
            logits, frames, C, out);
        return cudaGetLastError();
    }
    full_logprobs_kernel<<<static_cast<unsigned>((frames + 127) / 128), 128, 0, stream>>>(logits, frames, C, out);
    return cudaGetLastError();
}

cudaError_t launch_sparsify(const float* logits, int n, int T, int C, const int32_t* t_lo, const int32_t* t_hi,
                            int32_t* indptr, int32_t* nnz, int64_t* base, int32_t* indices, float* data,
                            int64_t capacity, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    static PerDeviceOnce attr_done;
    if (attr_done.pending()) {
        cudaFuncSetAttribute(sparsify_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(sparsify_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(sparsify_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(sparsify_fill_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        // the stream-ordered allocator keeps what it has handed out once (no cudaMalloc per call)
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        attr_done.mark();
    }
    const int pitch = (C + 1) | 1;
    // tile + (mask kernel: 64 floats of frame statistics | fill kernel: C mask words + C 64-bit offsets)
    const size_t tile_dyn = (static_cast<size_t>(ST_FRAMES) * pitch + 64 + 3 * static_cast<size_t>(C) + 4) * sizeof(float);
    if (tile_dyn <= 200 * 1024 && T > 0) {
        const int chunks = (T + ST_FRAMES - 1) / ST_FRAMES;
        uint32_t* colmask = nullptr;
        cudaError_t err = cudaMallocAsync(reinterpret_cast<void**>(&colmask),
                                          static_cast<size_t>(n) * chunks * C * sizeof(uint32_t), stream);
        if (err != cudaSuccess) return err;
        const dim3 grid(chunks, n);
        sparsify_mask_kernel<<<grid, ST_THREADS, tile_dyn, stream>>>(logits, T, C, t_lo, t_hi, chunks, colmask);
        sparsify_indptr_kernel<<<n, ST_THREADS, C * sizeof(int32_t), stream>>>(colmask, chunks, C, indptr, nnz);
        sparsify_scan_kernel<<<1, 32, 0, stream>>>(nnz, n, base);
        sparsify_fill_tiled_kernel<<<grid, ST_THREADS, tile_dyn, stream>>>(logits, T, C, t_lo, t_hi, chunks, colmask, indptr,
                                                                          base, capacity, indices, data);
        err = cudaGetLastError();
        const cudaError_t ferr = cudaFreeAsync(colmask, stream);
        return err != cudaSuccess ? err : ferr;
    }
    const size_t dyn = (2 * static_cast<size_t>(T) + C) * sizeof(float);
    if (dyn > 200 * 1024) return cudaErrorInvalidValue;
    sparsify_count_kernel<<<n, SP_THREADS, dyn, stream>>>(logits, T, C, t_lo, t_hi, indptr, nnz);
    sparsify_scan_kernel<<<1, 32, 0, stream>>>(nnz, n, base);
    sparsify_fill_kernel<<<n, SP_THREADS, dyn, stream>>>(logits, T, C, t_lo, t_hi, indptr, base, capacity, indices, data);
    return cudaGetLastError();
}
