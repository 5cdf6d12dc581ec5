// Host side of libb200_lineocr.so: weight packing, workspace planning, layer walk, C ABI (include/b200_lineocr.h).
#include "../../include/b200_lineocr.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ctc_beam.cuh"
#include "igemm.cuh"
#include "kernels.cuh"
#include "lstm_tc.cuh"

namespace {

constexpr int kTileCounters = 1024;
std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct Gemm {  // one dense contraction: packed fp16 weights [plane][tap][cout_pad][cin] + fp32 epilogue vectors
    int cin = 0, cout = 0, kh = 1, kw = 1, pad_h = 0, pad_w = 0, bn = 64, cout_pad = 0, planes = 1;
    __half* w = nullptr;
    float* bias = nullptr;
    float* post_scale = nullptr;
    float* post_shift = nullptr;
    CUtensorMap tmB;
    int bn_halo = 0;       // tile N of the halo-reuse kernel (0 = not applicable)
    CUtensorMap tmB_halo;
    float act_slope = 0.01f;   // LeakyReLU negative slope of the layer's fused activation
    int corr = CORR_BOTH;  // fp16f8 only: correction terms of the e5m2 pass (b200ocr_set_layer_correction)
};

struct LayerRT {
    int kind = 0, act = 0, pool_h = 1, pool_w = 1;
    Gemm g;                  // CONV / CTC_HEAD / BILSTM input projection
    float* w_t = nullptr;    // CONV_FIRST: fp32 [27][cout]
    float* bias0 = nullptr;  // CONV_FIRST
    uint32_t* wfrag0 = nullptr;   // CONV_FIRST: mma.sync weight fragments (conv_first.cu)
    uint32_t* wtc0 = nullptr;     // CONV_FIRST: [hi | lo][cout][32] fp16 weights of the tcgen05 variant
    float* oscale0 = nullptr;     // CONV_FIRST: per-channel power-of-two weight scale / 255
    int cout0 = 0;
    int hidden = 0;          // BILSTM
    __half* w_rec = nullptr;
    float* w_hh_t = nullptr;
    CUtensorMap tmW;
    int heads = 0, dim_ff = 0;  // TRANSFORMER
    Gemm g_in, g_out, g_l1, g_l2;
    float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;
};

struct Shape {
    int n, h, w, c;
};

// Decoder half of TransformerOCR (b200ocr_ar_attach): fp32 weights of the per-step projections, the tcgen05 GEMM of
// the memory K | V projection, and the decode workspace.
struct ArLayer {
    float *self_in_w = nullptr, *self_in_b = nullptr, *self_out_w = nullptr, *self_out_b = nullptr;
    float *cross_q_w = nullptr, *cross_q_b = nullptr, *cross_out_w = nullptr, *cross_out_b = nullptr;
    float *l1w = nullptr, *l1b = nullptr, *l2w = nullptr, *l2b = nullptr;
    float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr, *n3w = nullptr, *n3b = nullptr;
    Gemm g_memkv;   // rows D..3D-1 of the encoder-decoder in_proj: memory -> K | V
};

constexpr int kArSplitMax = 4;

struct ArState {
    bool attached = false;
    // token-loop kernels (b200ocr_debug_set_flag 2): 0 = tiled projections, 1 = split-K projections, 2 = split-K
    // projections with q | k | v in one launch, the K split of the out-projections and of the second feed-forward
    // matrix spread over CTAs and summed inside the LayerNorm, CTA-per-(line, head) step attention; 3 (default) = 2 with
    // the position in device memory and the 25 launches of a position replayed as a CUDA graph
    int linear_variant = 3;
    int heads = 0, dim_ff = 0, classes = 0, D = 0;
    std::vector<ArLayer> layers;
    float *embed = nullptr, *out_w = nullptr, *out_b = nullptr;
    // workspace (b200ocr_ar_reserve)
    int cap_lines = 0, cap_T = 0, cap_steps = 0;
    float* memkv = nullptr;    // [layer][cap_lines * cap_T][2D]   K | V of the memory frames
    float* selfkv = nullptr;   // [layer][cap_steps][lines][2D]    K | V of the decoded positions
    float *x = nullptr, *q = nullptr, *a = nullptr, *t = nullptr, *f = nullptr, *lg = nullptr;
    float* part = nullptr;     // [kArSplitMax][lines][D] partial sums of the K-split projections
    // token loop v3: the launches of one position as a CUDA graph, replayed on the engine's own stream
    struct GraphKey {
        int n, T, max_steps, start_token;
        const void *tokens, *logits, *workspace;
        int variant;
        bool operator==(const GraphKey& o) const {
            return n == o.n && T == o.T && max_steps == o.max_steps && start_token == o.start_token && tokens == o.tokens &&
                   logits == o.logits && workspace == o.workspace && variant == o.variant;
        }
    };
    GraphKey gkey{};
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t loop_stream = nullptr;
    cudaEvent_t loop_fork = nullptr, loop_join = nullptr;
    int32_t *alive = nullptr, *state = nullptr;
    int32_t* h_state = nullptr;   // pinned host copy of `state`
    std::vector<void*> ws;
};

}  // namespace

struct b200ocr_engine {
    int device = 0, num_sms = 148, precision = 0, planes = 1, npass = 1, line_height = 40;
    int fmt = ACT_F16;        // activation record format between layers (actfmt.cuh)
    int lstm_planes = 1;      // fp16 planes of W_hh in the LSTM recurrence (1 = fp16, 2 = hi + lo)
    int lstm_hplanes = 1;     // planes of h_t exchanged and multiplied per step (<= lstm_planes; lstm_tc.cuh)
    bool use_ref = false;
    bool use_halo = true;
    int ref_only_layer = -1;  // debug flag 5: this layer's contraction alone runs on the CUDA-core cross-check kernel
    bool dynamic_tiles = true;   // persistent GEMM kernels draw tiles from a global counter (tilesched.cuh; flag 7)
    int* tile_counters = nullptr;   // [kTileCounters] zeroed at the start of every layer walk
    int tile_counter_next = 0;
    // Option (flag 10, default off): the BiLSTM recurrence on a high-priority side stream (fork / join by events around
    // the launch), so that the block scheduler places its clusters before another engine's next conv layer.  Measured
    // with and without: no difference once the replicas are linked (b200ocr_run_after)
    // b200ocr_run_after: `front_done` is recorded when this engine's walk reaches its first recurrence; a walk waits
    // for the `front_done` of `after` first.  `followers` = engines whose `after` is this one (cleared on destroy)
    cudaEvent_t front_done = nullptr;
    b200ocr_engine* after = nullptr;
    std::vector<b200ocr_engine*> followers;
    bool front_recorded = false;
    bool lstm_priority = false;
    cudaStream_t hi_stream = nullptr;
    cudaEvent_t hi_fork = nullptr, hi_join = nullptr;
    // Option (flag 11, default off): a producer whose consumer multiplies with weight-side correction only does not
    // write the lo' plane of its ACT_F16_F8 records -- a quarter of those layers' write traffic.  Measured: no gain
    // (profiles/r02y_lean_records_ab.json; the layers are not write-bound and holes in the records cost as much as
    // the bytes they save)
    bool lean_records = false;
    int igemm_dbg = 0;           // OR-ed into IgemmParams::dbg (flag 9): 4 = 16-byte epilogue stores
    bool attention_tc = true;    // Transformer variant: tcgen05 attention where it applies (attention_tc.cu; flag 8)
    int l2_chunk_lines = 0;   // first conv + next layer run over chunks of this many lines (0 = whole batch; flag 6)
    int crop_staging = 3;     // first conv: how the uint8 patch is staged (0 plain loads, 1 cp.async, 2 TMA) on the mma.sync
                              // kernel, 3 = TMA staging + the tcgen05 kernel (conv_first.cu)
    std::vector<LayerRT> layers;
    std::vector<void*> owned;
    // workspace
    void* hbuf[3] = {nullptr, nullptr, nullptr};
    size_t hbuf_bytes[3] = {0, 0, 0};
    void* fbuf[3] = {nullptr, nullptr, nullptr};
    size_t fbuf_bytes[3] = {0, 0, 0};
    int32_t* best = nullptr;
    float *fmax = nullptr, *flse = nullptr, *fprob = nullptr;
    size_t frames_cap = 0;
    int64_t launches = 0;
    std::string err;
    bool profiling = false;
    int cur_layer = -1;
    struct ProfRec { int tag, layer; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    ArState ar;
};

namespace {

int fail(b200ocr_engine* e, int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    else g_create_error = buf;
    return status;
}

// Brackets one kernel launch with CUDA events on its own stream when profiling is on (bench.py's roofline leg).
struct ProfScope {
    b200ocr_engine* e;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    int tag;
    bool on;
    ProfScope(b200ocr_engine* e_, cudaStream_t st_, int tag_) : e(e_), st(st_), tag(tag_), on(e_->profiling) {
        if (on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(b, st);
            e->prof.push_back({tag, e->cur_layer, a, b});
        }
    }
};
enum { PROF_CONV_FIRST = 0, PROF_IGEMM = 1, PROF_LSTM = 2, PROF_OTHER = 3 };

#define CU_TRY(e, call)                                                                                   \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess)                                                                         \
            return fail(e, B200OCR_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

template <typename T>
int upload(b200ocr_engine* e, const T* host, size_t count, T** dev) {
    CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(dev), std::max<size_t>(count, 1) * sizeof(T)));
    e->owned.push_back(*dev);
    CU_TRY(e, cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int make_map_2d(b200ocr_engine* e, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
    return 0;
}

int make_map_act(b200ocr_engine* e, CUtensorMap* m, const void* base, int n, int h, int w, int c_total, int box_w,
                 int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)c_total, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c_total * 2, (cuuint64_t)w * c_total * 2, (cuuint64_t)h * w * c_total * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled(4d) failed: %d", (int)r);
    return 0;
}

// weight: PyTorch [cout][cin][kh][kw] fp32 (linear: kh = kw = 1)
int build_gemm(b200ocr_engine* e, Gemm& g, const float* weight, const float* bias, const float* post_scale,
               const float* post_shift, int cin, int cout, int kh, int kw, int pad_h, int pad_w) {
    if (cin % 64) return fail(e, B200OCR_E_INVALID, "implicit GEMM needs cin %% 64 == 0 (got %d)", cin);
    g.cin = cin; g.cout = cout; g.kh = kh; g.kw = kw; g.pad_h = pad_h; g.pad_w = pad_w;
    g.planes = e->planes;
    g.bn = igemm_pick_bn(cout);
    g.cout_pad = ((cout + g.bn - 1) / g.bn) * g.bn;
    const int taps = kh * kw;
    const size_t rows = static_cast<size_t>(g.planes) * taps * g.cout_pad;
    std::vector<__half> packed(rows * cin, __float2half(0.f));
    const bool f8 = e->fmt == ACT_F16_F8;
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < taps; ++t) {
                const float v = weight[(static_cast<size_t>(o) * cin + c) * taps + t];
                const __half hi = __float2half_rn(v);
                const float lo = v - __half2float(hi);
                const size_t row0 = (static_cast<size_t>(t) * g.cout_pad + o) * cin;
                const size_t row1 = ((static_cast<size_t>(taps) + t) * g.cout_pad + o) * cin;
                if (f8) {
                    // plane 0: fp16(w) * 2^11 (accumulators live at scale 2^11); plane 1: 2*cin e5m2 bytes
                    // [e5m2(hi) x cin | e5m2(lo * 2^11) x cin], the K-concatenated partner of [lo' | hi8] (actfmt.cuh)
                    const float scaled = __half2float(hi) * kF8Scale;
                    if (!(std::fabs(scaled) <= 65504.f))
                        return fail(e, B200OCR_E_INVALID, "fp16f8 precision needs |weight| < 32 (got %g)", v);
                    packed[row0 + c] = __float2half_rn(scaled);
                    uint8_t* b = reinterpret_cast<uint8_t*>(&packed[row1]);
                    b[c] = f32_to_e5m2(__half2float(hi));
                    b[cin + c] = f32_to_e5m2(lo * kF8Scale);
                } else {
                    packed[row0 + c] = hi;
                    if (g.planes == 2) packed[row1 + c] = __float2half_rn(lo);
                }
            }
    if (int s = upload(e, packed.data(), packed.size(), &g.w)) return s;
    if (bias) {
        std::vector<float> b(g.cout_pad, 0.f);
        std::copy(bias, bias + cout, b.begin());
        if (int s = upload(e, b.data(), b.size(), &g.bias)) return s;
    }
    if (post_scale) {
        std::vector<float> a(g.cout_pad, 1.f), b(g.cout_pad, 0.f);
        std::copy(post_scale, post_scale + cout, a.begin());
        std::copy(post_shift, post_shift + cout, b.begin());
        if (int s = upload(e, a.data(), a.size(), &g.post_scale)) return s;
        if (int s = upload(e, b.data(), b.size(), &g.post_shift)) return s;
    }
    if (kh == 3 && kw == 3 && pad_h == 1 && pad_w == 1 && cin <= 128 && cout <= 256 && (cout % 32) == 0) {
        g.bn_halo = cout > 64 ? 128 : 64;
        if (int s = make_map_2d(e, &g.tmB_halo, g.w, cin, rows, g.bn_halo)) return s;
    }
    return make_map_2d(e, &g.tmB, g.w, cin, rows, g.bn);
}

struct EpiOut {
    int epi = EPI_ACT_F16;
    __half* out_h = nullptr;
    float* out_f32 = nullptr;
    const float* residual = nullptr;
    int32_t* best = nullptr;
    float *fmax = nullptr, *flse = nullptr, *fprob = nullptr;
    bool skip_lo = false;   // the consumer of these records never reads their lo' plane
};

// input: fp16 NHWC [in.n][in.h][in.w][planes * g.cin]
int run_gemm(b200ocr_engine* e, const Gemm& g, const __half* in, Shape in_s, int act, int pool_h, int pool_w,
             const EpiOut& o, cudaStream_t st, Shape* out_s) {
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.n_img = in_s.n; p.h_in = in_s.h; p.w_in = in_s.w; p.cin = g.cin; p.cout = g.cout;
    p.kh = g.kh; p.kw = g.kw; p.pad_h = g.pad_h; p.pad_w = g.pad_w;
    p.h_out = in_s.h + 2 * g.pad_h - g.kh + 1;
    p.w_out = in_s.w + 2 * g.pad_w - g.kw + 1;
    p.pool_h = pool_h; p.pool_w = pool_w; p.act = act; p.act_slope = g.act_slope; p.npass = e->npass;
    p.corr_mode = e->fmt == ACT_F16_F8 ? g.corr : CORR_BOTH;
    if (p.h_out <= 0 || p.w_out <= 0) return fail(e, B200OCR_E_INVALID, "empty convolution output");
    if (p.h_out % pool_h || p.w_out % pool_w)
        return fail(e, B200OCR_E_INVALID, "pooled layer needs even output (%d x %d)", p.h_out, p.w_out);
    igemm_fill_geometry(p, g.bn);
    p.epi = o.epi;
    p.bias = g.bias; p.post_scale = g.post_scale; p.post_shift = g.post_shift; p.residual = o.residual;
    p.out_h = o.out_h; p.out_cstride = g.cout * e->planes; p.out_lo_off = e->planes == 2 ? g.cout : -1;
    p.out_fmt = e->fmt;
    p.out_skip_lo = (o.skip_lo && e->fmt == ACT_F16_F8 && e->lean_records) ? 1 : 0;
    p.acc_scale = e->fmt == ACT_F16_F8 ? 1.f / kF8Scale : 1.f;
    p.out_f32 = o.out_f32; p.best = o.best; p.fmax = o.fmax; p.flse = o.flse; p.fprob = o.fprob;
    {
        static const int dbg = getenv("B200OCR_IGEMM_DBG") ? atoi(getenv("B200OCR_IGEMM_DBG")) : 0;   // bring-up only
        p.dbg = dbg | e->igemm_dbg;
    }
    p.tile_counter = (e->dynamic_tiles && e->tile_counters && e->tile_counter_next < kTileCounters)
                         ? e->tile_counters + e->tile_counter_next++ : nullptr;
    if (o.epi == EPI_ACT_F16 && (g.cout % 32))
        return fail(e, B200OCR_E_INVALID, "fp16 activation output needs cout %% 32 == 0 (got %d)", g.cout);
    if (o.epi == EPI_CTC && p.tiles_n != 1)
        return fail(e, B200OCR_E_INVALID, "fused CTC head supports at most 256 classes");
    if (out_s) *out_s = Shape{in_s.n, p.h_out / pool_h, p.w_out / pool_w, g.cout};
    if (e->use_ref || (e->ref_only_layer >= 0 && e->ref_only_layer == e->cur_layer && o.epi != EPI_CTC)) {
        if (o.epi == EPI_CTC) return fail(e, B200OCR_E_INVALID, "internal: CTC epilogue has no reference kernel");
        CU_TRY(e, launch_igemm_ref(p, in, g.w, st));
        e->launches++;
        return 0;
    }
    CUtensorMap tmA;
    if (e->use_halo && g.bn_halo) {
        IgemmParams ph = p;
        ph.tiles_n = (g.cout + g.bn_halo - 1) / g.bn_halo;
        if (ph.tiles_n * g.bn_halo == g.cout_pad && igemm_halo_supported(ph, g.bn_halo)) {
            if (int s = make_map_act(e, &tmA, in, in_s.n, in_s.h, in_s.w, e->planes * g.cin, 130, 4)) return s;
            {
                ProfScope ps(e, st, PROF_IGEMM);
                CU_TRY(e, launch_igemm_halo(ph, tmA, g.tmB_halo, g.bn_halo, e->num_sms, st));
            }
            e->launches++;
            return 0;
        }
    }
    if (int s = make_map_act(e, &tmA, in, in_s.n, in_s.h, in_s.w, e->planes * g.cin, 32 / p.th, p.th)) return s;
    {
        ProfScope ps(e, st, PROF_IGEMM);
        CU_TRY(e, launch_igemm_tc(p, tmA, g.tmB, g.bn, e->num_sms, st));
    }
    e->launches++;
    return 0;
}

// 3-D map over the uint8 crop batch [n][h][w*3 bytes] in 32-bit words (a TMA box is at most 256 elements per
// dimension: 416 bytes = 104 words), box {104, 6, 1}: the first conv's patch (conv_first.cu)
int make_map_crops(b200ocr_engine* e, CUtensorMap* m, const void* base, int n, int h, int w) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)w * 3 / 4, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)h * w * 3};
    cuuint32_t box[3] = {104, 6, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, B200OCR_E_CUDA, "cuTensorMapEncodeTiled(crops) failed: %d", (int)r);
    return 0;
}

size_t hbytes(const b200ocr_engine* e, Shape s) {
    return static_cast<size_t>(s.n) * s.h * s.w * s.c * e->planes * sizeof(__half);
}
size_t fbytes(Shape s) { return static_cast<size_t>(s.n) * s.h * s.w * s.c * sizeof(float); }

struct Need {
    size_t h[3] = {0, 0, 0}, f[3] = {0, 0, 0}, frames = 0;
    void hn(int i, size_t b) { h[i] = std::max(h[i], b); }
    void fn(int i, size_t b) { f[i] = std::max(f[i], b); }
};

struct Outputs {
    float* logits = nullptr;
    int32_t *labels = nullptr, *lengths = nullptr, *best_path = nullptr;
    float* confidence = nullptr;
    float* maps = nullptr;
};

// Walks the layer list.  dry != nullptr: only collects workspace needs.  Otherwise launches.
int walk(b200ocr_engine* e, const uint8_t* crops, int n, int h, int w, int n_layers, const Outputs& out,
         cudaStream_t st, Need* dry, Shape* final_shape, const void** final_ptr, bool* final_is_f32) {
    Shape cur{n, h, w, 3};
    const __half* cur_h = nullptr;  // fp16 activation
    const float* cur_f = nullptr;   // fp32 activation (residual stream / maps)
    int hslot = -1;                 // slot of cur_h
    auto next_h = [&](int avoid_a, int avoid_b) {
        for (int i = 0; i < 3; ++i)
            if (i != avoid_a && i != avoid_b) return i;
        return 0;
    };
    const int L = std::min<int>(n_layers, e->layers.size());
    // does layer li + 1 (executed in this walk, as a tensor-core contraction over layer li's records) skip the
    // activation-side correction?  Then layer li need not write the lo' plane.
    auto next_skips_lo = [&](int li) {
        if (e->fmt != ACT_F16_F8 || !e->lean_records || e->use_ref || li + 1 >= L) return false;
        const LayerRT& nx = e->layers[li + 1];
        if (nx.kind != B200OCR_CONV && nx.kind != B200OCR_BILSTM && nx.kind != B200OCR_CTC_HEAD) return false;
        return nx.g.corr != CORR_BOTH && e->ref_only_layer != li + 1;
    };
    if (!dry && crops && e->after && e->after->front_done)
        CU_TRY(e, cudaStreamWaitEvent(st, e->after->front_done, 0));
    e->front_recorded = false;
    if (!dry && e->dynamic_tiles && e->tile_counters) {
        CU_TRY(e, cudaMemsetAsync(e->tile_counters, 0, kTileCounters * sizeof(int), st));
        e->tile_counter_next = 0;
    }
    for (int li = 0; li < L; ++li) {
        LayerRT& ly = e->layers[li];
        e->cur_layer = li;
        const int next_kind = li + 1 < (int)e->layers.size() ? e->layers[li + 1].kind : 0;
        switch (ly.kind) {
            case B200OCR_CONV_FIRST: {
                Shape os{cur.n, cur.h, cur.w, ly.cout0};
                const int slot = 0;
                // L2 chunking: the 64-channel records of the first conv are the largest tensor of the step (256 B per
                // pixel: 3.5 GB at config 2) and are read exactly once, by the next layer.  Run the two layers over
                // chunks of a few lines, the first conv always writing the SAME region: its records stay dirty in the
                // 126 MB L2, are consumed from there and overwritten by the next chunk -- they never travel to HBM.
                LayerRT* nx = li + 1 < L ? &e->layers[li + 1] : nullptr;
                const bool next_f32 = li + 2 < (int)e->layers.size() &&
                                      (e->layers[li + 2].kind == B200OCR_UPSAMPLE || e->layers[li + 2].kind == B200OCR_LN_PE);
                const bool chunked = !dry && !e->use_ref && e->l2_chunk_lines > 0 && cur.n > e->l2_chunk_lines && nx &&
                                     nx->kind == B200OCR_CONV && !next_f32;
                if (dry) dry->hn(slot, hbytes(e, os));
                else {
                    if (hbytes(e, os) > e->hbuf_bytes[slot]) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                    const bool bulk_ok = (cur.w % 16) == 0 && (reinterpret_cast<uintptr_t>(crops) % 16) == 0 &&
                                         ((static_cast<size_t>(cur.h) * cur.w * 3) % 16) == 0;
                    const int staging = bulk_ok ? e->crop_staging : 0;
                    __half* rec = static_cast<__half*>(e->hbuf[slot]);
                    auto first = [&](const uint8_t* src, int n_lines) -> int {
                        CUtensorMap tm_crops;
                        if (!e->use_ref && staging >= 2)
                            if (int s = make_map_crops(e, &tm_crops, src, n_lines, cur.h, cur.w)) return s;
                        ProfScope ps(e, st, PROF_CONV_FIRST);
                        if (e->use_ref)   // CUDA-core fp32 cross-check kernel
                            CU_TRY(e, launch_conv_first(src, n_lines, cur.h, cur.w, ly.w_t, ly.bias0, ly.cout0, ly.act, ly.g.act_slope,
                                                        e->fmt, rec, st));
                        else if (staging == 3 && ly.wtc0)    // tcgen05 variant
                            CU_TRY(e, launch_conv_first_tc(src, n_lines, cur.h, cur.w, ly.wtc0, ly.oscale0, ly.bias0, ly.cout0,
                                                           ly.act, ly.g.act_slope, e->fmt, rec, next_skips_lo(li) ? 1 : 0,
                                                           &tm_crops, st));
                        else
                            CU_TRY(e, launch_conv_first_mma(src, n_lines, cur.h, cur.w, ly.wfrag0, ly.oscale0, ly.bias0,
                                                            ly.cout0, ly.act, ly.g.act_slope, e->fmt, rec, next_skips_lo(li) ? 1 : 0,
                                                            std::min(staging, 2), staging >= 2 ? &tm_crops : nullptr, st));
                        e->launches++;
                        return 0;
                    };
                    if (!chunked) {
                        if (int s = first(crops, cur.n)) return s;
                    } else {
                        Shape os2{cur.n, (cur.h + 2 * nx->g.pad_h - nx->g.kh + 1) / nx->pool_h,
                                  (cur.w + 2 * nx->g.pad_w - nx->g.kw + 1) / nx->pool_w, nx->g.cout};
                        const int slot2 = next_h(slot, -1);
                        if (hbytes(e, os2) > e->hbuf_bytes[slot2]) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                        const size_t in_line = static_cast<size_t>(cur.h) * cur.w * 3;
                        const size_t out_line = hbytes(e, Shape{1, os2.h, os2.w, os2.c}) / sizeof(__half);
                        for (int c0 = 0; c0 < cur.n; c0 += e->l2_chunk_lines) {
                            const int nl = std::min(e->l2_chunk_lines, cur.n - c0);
                            e->cur_layer = li;
                            if (int s = first(crops + c0 * in_line, nl)) return s;
                            e->cur_layer = li + 1;
                            EpiOut eo;
                            eo.out_h = static_cast<__half*>(e->hbuf[slot2]) + c0 * out_line;
                            if (int s = run_gemm(e, nx->g, rec, Shape{nl, cur.h, cur.w, ly.cout0}, nx->act, nx->pool_h,
                                                 nx->pool_w, eo, st, nullptr))
                                return s;
                        }
                        cur_h = static_cast<__half*>(e->hbuf[slot2]);
                        hslot = slot2;
                        cur = os2;
                        cur_f = nullptr;
                        ++li;   // the next layer has been run chunk by chunk
                        break;
                    }
                    cur_h = rec;
                }
                hslot = slot;
                cur = os;
                cur_f = nullptr;
                break;
            }
            case B200OCR_CONV: {
                if (hslot < 0) return fail(e, B200OCR_E_INVALID, "layer %d: conv needs an fp16 activation input", li);
                const bool to_f32 = next_kind == B200OCR_UPSAMPLE || next_kind == B200OCR_LN_PE;
                Shape os{cur.n, (cur.h + 2 * ly.g.pad_h - ly.g.kh + 1) / ly.pool_h,
                         (cur.w + 2 * ly.g.pad_w - ly.g.kw + 1) / ly.pool_w, ly.g.cout};
                EpiOut eo;
                int slot = -1;
                if (to_f32) {
                    eo.epi = EPI_F32;
                    if (dry) dry->fn(0, fbytes(os));
                    else {
                        if (fbytes(os) > e->fbuf_bytes[0]) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                        eo.out_f32 = static_cast<float*>(e->fbuf[0]);
                    }
                } else {
                    slot = next_h(hslot, -1);
                    if (dry) dry->hn(slot, hbytes(e, os));
                    else {
                        if (hbytes(e, os) > e->hbuf_bytes[slot]) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                        eo.out_h = static_cast<__half*>(e->hbuf[slot]);
                    }
                    eo.skip_lo = next_skips_lo(li);
                }
                if (!dry) {
                    Shape chk;
                    if (int s = run_gemm(e, ly.g, cur_h, cur, ly.act, ly.pool_h, ly.pool_w, eo, st, &chk)) return s;
                }
                cur = os;
                if (to_f32) {
                    cur_f = static_cast<float*>(e->fbuf[0]);
                    cur_h = nullptr;
                    hslot = -1;
                } else {
                    cur_h = static_cast<__half*>(e->hbuf[slot]);
                    hslot = slot;
                    cur_f = nullptr;
                }
                break;
            }
            case B200OCR_BILSTM: {
                if (hslot < 0 || cur.h != 1) return fail(e, B200OCR_E_INVALID, "layer %d: BiLSTM needs [n][1][T][D] fp16 input", li);
                const int T = cur.w, H = ly.hidden;
                Shape rows{1, 1, cur.n * T, cur.c};
                Shape pre_s{1, 1, cur.n * T, 8 * H};
                Shape os{cur.n, 1, T, 2 * H};
                const int slot = next_h(hslot, -1);
                if (dry) {
                    dry->fn(1, fbytes(pre_s));
                    dry->hn(slot, hbytes(e, os));
                } else {
                    if (fbytes(pre_s) > e->fbuf_bytes[1] || hbytes(e, os) > e->hbuf_bytes[slot])
                        return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                    EpiOut eo;
                    eo.epi = EPI_F32;
                    eo.out_f32 = static_cast<float*>(e->fbuf[1]);
                    if (int s = run_gemm(e, ly.g, cur_h, rows, 0, 1, 1, eo, st, nullptr)) return s;
                    __half* o = static_cast<__half*>(e->hbuf[slot]);
                    if (e->front_done && !e->followers.empty() && !e->front_recorded) {
                        CU_TRY(e, cudaEventRecord(e->front_done, st));       // the conv front end of this batch is done
                        e->front_recorded = true;
                    }
                    if (e->use_ref || H != 256) {
                        ProfScope ps(e, st, PROF_LSTM);
                        CU_TRY(e, launch_lstm_ref(eo.out_f32, ly.w_hh_t, cur.n, T, H, e->fmt, e->planes == 1, o, st));
                    } else if (e->lstm_priority && e->hi_stream) {
                        CU_TRY(e, cudaEventRecord(e->hi_fork, st));
                        CU_TRY(e, cudaStreamWaitEvent(e->hi_stream, e->hi_fork, 0));
                        {
                            ProfScope ps(e, e->hi_stream, PROF_LSTM);
                            CU_TRY(e, launch_lstm_tc(ly.w_rec, eo.out_f32, o, cur.n, T, H, e->lstm_planes, e->lstm_hplanes,
                                                     e->fmt, e->hi_stream));
                        }
                        CU_TRY(e, cudaEventRecord(e->hi_join, e->hi_stream));
                        CU_TRY(e, cudaStreamWaitEvent(st, e->hi_join, 0));
                    } else {
                        ProfScope ps(e, st, PROF_LSTM);
                        CU_TRY(e, launch_lstm_tc(ly.w_rec, eo.out_f32, o, cur.n, T, H, e->lstm_planes, e->lstm_hplanes, e->fmt, st));
                    }
                    e->launches++;
                    cur_h = o;
                }
                hslot = slot;
                cur = os;
                break;
            }
            case B200OCR_LN_PE: {
                if (!cur_f && !dry) return fail(e, B200OCR_E_INVALID, "layer %d: LayerNorm needs an fp32 input", li);
                // cur: [n][1][T][D] fp32 in fbuf[0] -> residual stream fp32 (fbuf[2]) + fp16 operand (hbuf[0])
                const int T = cur.w;
                if (dry) {
                    dry->fn(2, fbytes(cur));
                    dry->hn(0, hbytes(e, cur));
                } else {
                    if (fbytes(cur) > e->fbuf_bytes[2] || hbytes(e, cur) > e->hbuf_bytes[0])
                        return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                    CU_TRY(e, launch_layernorm(cur_f, cur.n * T, cur.c, ly.n1w, ly.n1b, 1e-5f, T,
                                               static_cast<float*>(e->fbuf[2]), static_cast<__half*>(e->hbuf[0]),
                                               e->fmt, st));
                    e->launches++;
                    cur_f = static_cast<float*>(e->fbuf[2]);
                    cur_h = static_cast<__half*>(e->hbuf[0]);
                }
                hslot = 0;
                break;
            }
            case B200OCR_TRANSFORMER_LAYER: {
                // in: residual x fp32 (fbuf[2]) + x fp16 (hbuf[0]); out: same slots
                if (hslot != 0) return fail(e, B200OCR_E_INVALID, "layer %d: transformer layer must follow LN_PE or another transformer layer", li);
                const int T = cur.w, D = cur.c, R = cur.n * T;
                Shape rows{1, 1, R, D};
                Shape qkv_s{1, 1, R, 3 * D}, ff_s{1, 1, R, ly.dim_ff};
                if (dry) {
                    dry->fn(0, fbytes(qkv_s));
                    dry->fn(1, fbytes(rows));
                    dry->hn(1, hbytes(e, rows));
                    dry->hn(2, hbytes(e, ff_s));
                } else {
                    if (fbytes(qkv_s) > e->fbuf_bytes[0] || fbytes(rows) > e->fbuf_bytes[1] ||
                        hbytes(e, rows) > e->hbuf_bytes[1] || hbytes(e, ff_s) > e->hbuf_bytes[2])
                        return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                    float* x = static_cast<float*>(e->fbuf[2]);
                    float* qkv = static_cast<float*>(e->fbuf[0]);
                    float* tmp = static_cast<float*>(e->fbuf[1]);
                    __half* xh = static_cast<__half*>(e->hbuf[0]);
                    __half* ah = static_cast<__half*>(e->hbuf[1]);
                    __half* fh = static_cast<__half*>(e->hbuf[2]);
                    EpiOut eo;
                    eo.epi = EPI_F32; eo.out_f32 = qkv;
                    if (int s = run_gemm(e, ly.g_in, xh, rows, 0, 1, 1, eo, st, nullptr)) return s;
                    {
                        ProfScope ps(e, st, PROF_OTHER);
                        if (!e->use_ref && e->attention_tc && attention_tc_supported(T, D, ly.heads))
                            CU_TRY(e, launch_attention_tc(qkv, cur.n, T, D, ly.heads, ah, e->fmt, st));
                        else
                            CU_TRY(e, launch_attention(qkv, cur.n, T, D, ly.heads, ah, e->fmt, st));
                    }
                    e->launches++;
                    EpiOut er;
                    er.epi = EPI_RES_F32; er.out_f32 = tmp; er.residual = x;
                    if (int s = run_gemm(e, ly.g_out, ah, rows, 0, 1, 1, er, st, nullptr)) return s;
                    { ProfScope ps(e, st, PROF_OTHER); CU_TRY(e, launch_layernorm(tmp, R, D, ly.n1w, ly.n1b, 1e-5f, 0, x, xh, e->fmt, st)); }
                    e->launches++;
                    EpiOut ef;
                    ef.epi = EPI_ACT_F16; ef.out_h = fh;
                    if (int s = run_gemm(e, ly.g_l1, xh, rows, B200OCR_ACT_RELU, 1, 1, ef, st, nullptr)) return s;
                    Shape ffrows{1, 1, R, ly.dim_ff};
                    er.out_f32 = tmp; er.residual = x;
                    if (int s = run_gemm(e, ly.g_l2, fh, ffrows, 0, 1, 1, er, st, nullptr)) return s;
                    { ProfScope ps(e, st, PROF_OTHER); CU_TRY(e, launch_layernorm(tmp, R, D, ly.n2w, ly.n2b, 1e-5f, 0, x, xh, e->fmt, st)); }
                    e->launches++;
                }
                break;
            }
            case B200OCR_CTC_HEAD: {
                if (hslot < 0 || cur.h != 1) return fail(e, B200OCR_E_INVALID, "layer %d: CTC head needs [n][1][T][D] fp16 input", li);
                const int T = cur.w, C = ly.g.cout;
                const size_t frames = static_cast<size_t>(cur.n) * T;
                Shape rows{1, 1, cur.n * T, cur.c};
                Shape lg{cur.n, 1, T, C};
                if (dry) {
                    dry->frames = std::max(dry->frames, frames);
                    dry->fn(0, fbytes(lg));  // reference-kernel path materialises logits even when not asked for
                } else {
                    if (frames > e->frames_cap) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                    int32_t* best = out.best_path ? out.best_path : e->best;
                    if (e->use_ref) {
                        float* logits = out.logits;
                        if (!logits) {
                            if (fbytes(lg) > e->fbuf_bytes[0]) return fail(e, B200OCR_E_WORKSPACE, "workspace too small");
                            logits = static_cast<float*>(e->fbuf[0]);
                        }
                        EpiOut eo;
                        eo.epi = EPI_F32; eo.out_f32 = logits;
                        if (int s = run_gemm(e, ly.g, cur_h, rows, 0, 1, 1, eo, st, nullptr)) return s;
                        CU_TRY(e, launch_frame_stats(logits, cur.n, T, C, 0, best, e->fmax, e->flse, e->fprob, st));
                        e->launches++;
                    } else {
                        EpiOut eo;
                        eo.epi = EPI_CTC; eo.out_f32 = out.logits; eo.best = best;
                        eo.fmax = e->fmax; eo.flse = e->flse; eo.fprob = out.confidence ? e->fprob : nullptr;
                        if (int s = run_gemm(e, ly.g, cur_h, rows, 0, 1, 1, eo, st, nullptr)) return s;
                    }
                    if (out.labels) {
                        ProfScope ps(e, st, PROF_OTHER);
                        CU_TRY(e, launch_ctc_collapse(best, out.confidence ? e->fprob : nullptr, cur.n, T, C - 1,
                                                      out.labels, out.lengths, out.confidence, st));
                        e->launches++;
                    }
                }
                cur = lg;
                cur_f = out.logits;
                cur_h = nullptr;
                hslot = -1;
                break;
            }
            case B200OCR_UPSAMPLE: {
                if (!dry) {
                    if (!cur_f) return fail(e, B200OCR_E_INVALID, "layer %d: upsample needs an fp32 input", li);
                    if (out.maps) {
                        CU_TRY(e, launch_upsample_nchw(cur_f, cur.n, cur.h, cur.w, cur.c, ly.pool_h, out.maps, st));
                        e->launches++;
                    }
                }
                cur = Shape{cur.n, cur.h * ly.pool_h, cur.w * ly.pool_h, cur.c};
                cur_f = out.maps;
                break;
            }
            default:
                return fail(e, B200OCR_E_INVALID, "layer %d: unknown kind %d", li, ly.kind);
        }
    }
    if (final_shape) *final_shape = cur;
    if (final_ptr) *final_ptr = cur_h ? static_cast<const void*>(cur_h) : static_cast<const void*>(cur_f);
    if (final_is_f32) *final_is_f32 = cur_h == nullptr;
    return 0;
}

// The engine-less entry points (decoders, cropper, sparsification, alignment) launch on the stream they are given; a
// stream and the per-device kernel attributes belong to ONE device, which need not be the calling thread's current
// one (a tensor on cuda:1 while cuda:0 is current).  Makes the device that owns `ptr` current for the call.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const void* ptr) {
        cudaPointerAttributes at;
        int cur = 0;
        if (ptr && cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
            cudaGetDevice(&cur) == cudaSuccess && cur != at.device && cudaSetDevice(at.device) == cudaSuccess)
            prev = cur;
        cudaGetLastError();   // a host pointer is not an error here
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_device(b200ocr_engine* e, int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= device)
        return fail(e, B200OCR_E_NO_DEVICE, "no CUDA device %d visible (there is no CPU fallback)", device);
    cudaDeviceProp prop;
    CU_TRY(e, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(e, B200OCR_E_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device,
                    prop.major, prop.minor);
    return 0;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int b200ocr_create(const b200ocr_net_desc_t* desc, b200ocr_engine_t** out) {
    if (!desc || !out || desc->n_layers <= 0 || !desc->layers) return fail(nullptr, B200OCR_E_INVALID, "bad descriptor");
    *out = nullptr;
    if (int s = check_device(nullptr, desc->device)) return s;
    b200ocr_engine* e = new b200ocr_engine();
    auto bail = [&](int s) {
        g_create_error = e->err;
        b200ocr_destroy(e);
        return s;
    };
    e->device = desc->device;
    if (cudaSetDevice(e->device) != cudaSuccess) return bail(fail(e, B200OCR_E_CUDA, "cudaSetDevice failed"));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, e->device);
    e->num_sms = prop.multiProcessorCount;
    e->precision = desc->precision;
    switch (desc->precision) {
        case B200OCR_PREC_FP16: e->fmt = ACT_F16; e->planes = 1; e->npass = 1; e->lstm_planes = 1; break;
        case B200OCR_PREC_FP16X3: e->fmt = ACT_F16_HILO; e->planes = 2; e->npass = 3; e->lstm_planes = e->lstm_hplanes = 2; break;
        case B200OCR_PREC_FP16F8:
        case B200OCR_PREC_FP16F8W: e->fmt = ACT_F16_F8; e->planes = 2; e->npass = 2; e->lstm_planes = 2; break;
        default: return bail(fail(e, B200OCR_E_INVALID, "unknown precision %d", desc->precision));
    }
    e->line_height = desc->line_height;
    e->layers.resize(desc->n_layers);
    for (int i = 0; i < desc->n_layers; ++i) {
        const b200ocr_layer_t& d = desc->layers[i];
        LayerRT& ly = e->layers[i];
        ly.kind = d.kind; ly.act = d.act;
        ly.g.act_slope = d.act_slope > 0.f ? d.act_slope : 0.01f;
        ly.pool_h = d.pool_h > 0 ? d.pool_h : 1;
        ly.pool_w = d.pool_w > 0 ? d.pool_w : 1;
        int s = 0;
        switch (d.kind) {
            case B200OCR_CONV_FIRST: {
                if (d.cin != 3 || d.kh != 3 || d.kw != 3 || (d.cout != 64 && d.cout != 32 && d.cout != 16))
                    return bail(fail(e, B200OCR_E_INVALID, "layer %d: first conv must be 3x3, 3 -> 16/32/64 channels", i));
                std::vector<float> wt(27 * d.cout);
                for (int o = 0; o < d.cout; ++o)
                    for (int c = 0; c < 3; ++c)
                        for (int r = 0; r < 3; ++r)
                            for (int q = 0; q < 3; ++q)
                                wt[((r * 3 + q) * 3 + c) * d.cout + o] = d.weight[((o * 3 + c) * 3 + r) * 3 + q];
                ly.cout0 = d.cout;
                if ((s = upload(e, wt.data(), wt.size(), &ly.w_t))) return bail(s);
                if (d.bias && (s = upload(e, d.bias, (size_t)d.cout, &ly.bias0))) return bail(s);
                std::vector<uint32_t> wf(conv_first_wfrag_words(d.cout));
                std::vector<float> osc(d.cout);
                conv_first_pack(d.weight, d.cout, wf.data(), osc.data());
                if ((s = upload(e, wf.data(), wf.size(), &ly.wfrag0))) return bail(s);
                if ((s = upload(e, osc.data(), osc.size(), &ly.oscale0))) return bail(s);
                if (d.cout >= 32) {
                    std::vector<uint32_t> wk(conv_first_tc_words(d.cout));
                    conv_first_pack_tc(d.weight, d.cout, wk.data());
                    if ((s = upload(e, wk.data(), wk.size(), &ly.wtc0))) return bail(s);
                }
                break;
            }
            case B200OCR_CONV:
            case B200OCR_CTC_HEAD:
                if ((s = build_gemm(e, ly.g, d.weight, d.bias, d.post_scale, d.post_shift, d.cin, d.cout,
                                    d.kh > 0 ? d.kh : 1, d.kw > 0 ? d.kw : 1, d.pad_h, d.pad_w)))
                    return bail(s);
                break;
            case B200OCR_BILSTM: {
                const int H = d.hidden;
                // H = 256 runs on the tcgen05 cluster kernel (lstm_tc.cu); other sizes on the fp32 CUDA-core recurrence
                // (kernels.cu: lstm_ref_kernel, one thread per hidden unit) -- functional, not tuned
                if (H < 32 || H > 1024 || (H % 32))
                    return bail(fail(e, B200OCR_E_INVALID, "layer %d: BiLSTM hidden size must be a multiple of 32 in [32, 1024] (got %d)", i, H));
                // input projection of both directions as one GEMM: rows [dir][4H], bias = b_ih + b_hh
                std::vector<float> wih(static_cast<size_t>(8) * H * d.cin), bsum(8 * H);
                for (int dir = 0; dir < 2; ++dir) {
                    std::copy(d.w_ih[dir], d.w_ih[dir] + static_cast<size_t>(4) * H * d.cin,
                              wih.begin() + static_cast<size_t>(dir) * 4 * H * d.cin);
                    for (int k = 0; k < 4 * H; ++k) bsum[dir * 4 * H + k] = d.b_ih[dir][k] + d.b_hh[dir][k];
                }
                if ((s = build_gemm(e, ly.g, wih.data(), bsum.data(), nullptr, nullptr, d.cin, 8 * H, 1, 1, 0, 0)))
                    return bail(s);
                ly.hidden = H;
                // recurrent weights for the tcgen05 kernel: [dir][plane][cta j][row = gate*32 + u][k]
                const int P = e->lstm_planes;
                const bool tc = H == 256;
                std::vector<__half> rec(tc ? static_cast<size_t>(2) * P * 8 * 128 * H : 0);
                std::vector<float> wt(static_cast<size_t>(2) * H * 4 * H);
                for (int dir = 0; dir < 2; ++dir)
                    for (int g = 0; g < 4; ++g)
                        for (int u = 0; u < H; ++u)
                            for (int k = 0; k < H; ++k) {
                                const float v = d.w_hh[dir][(static_cast<size_t>(g) * H + u) * H + k];
                                const __half hi = __float2half_rn(v);
                                const int j = u / 32, row = g * 32 + (u % 32);
                                if (tc) {
                                    rec[(((static_cast<size_t>(dir) * P + 0) * 8 + j) * 128 + row) * H + k] = hi;
                                    if (P == 2)
                                        rec[(((static_cast<size_t>(dir) * P + 1) * 8 + j) * 128 + row) * H + k] =
                                            __float2half_rn(v - __half2float(hi));
                                }
                                // cross-check kernel: fp32 (x3) or the fp16-rounded value (fp16 mode), [dir][k][4H]
                                wt[(static_cast<size_t>(dir) * H + k) * 4 * H + g * H + u] = P == 2 ? v : __half2float(hi);
                            }
                if (tc && (s = upload(e, rec.data(), rec.size(), &ly.w_rec))) return bail(s);
                if ((s = upload(e, wt.data(), wt.size(), &ly.w_hh_t))) return bail(s);
                if (tc && (s = make_map_2d(e, &ly.tmW, ly.w_rec, H, static_cast<uint64_t>(2) * P * 8 * 128, 128))) return bail(s);
                break;
            }
            case B200OCR_LN_PE:
                if ((s = upload(e, d.norm1_w, (size_t)d.cin, &ly.n1w))) return bail(s);
                if ((s = upload(e, d.norm1_b, (size_t)d.cin, &ly.n1b))) return bail(s);
                break;
            case B200OCR_TRANSFORMER_LAYER: {
                const int D = d.cin;
                ly.heads = d.heads; ly.dim_ff = d.dim_ff;
                if ((s = build_gemm(e, ly.g_in, d.in_proj_w, d.in_proj_b, nullptr, nullptr, D, 3 * D, 1, 1, 0, 0))) return bail(s);
                if ((s = build_gemm(e, ly.g_out, d.out_proj_w, d.out_proj_b, nullptr, nullptr, D, D, 1, 1, 0, 0))) return bail(s);
                if ((s = build_gemm(e, ly.g_l1, d.lin1_w, d.lin1_b, nullptr, nullptr, D, d.dim_ff, 1, 1, 0, 0))) return bail(s);
                if ((s = build_gemm(e, ly.g_l2, d.lin2_w, d.lin2_b, nullptr, nullptr, d.dim_ff, D, 1, 1, 0, 0))) return bail(s);
                if ((s = upload(e, d.norm1_w, (size_t)D, &ly.n1w))) return bail(s);
                if ((s = upload(e, d.norm1_b, (size_t)D, &ly.n1b))) return bail(s);
                if ((s = upload(e, d.norm2_w, (size_t)D, &ly.n2w))) return bail(s);
                if ((s = upload(e, d.norm2_b, (size_t)D, &ly.n2b))) return bail(s);
                break;
            }
            case B200OCR_UPSAMPLE:
                break;
            default:
                return bail(fail(e, B200OCR_E_INVALID, "layer %d: unknown kind %d", i, d.kind));
        }
    }
    if (cudaMalloc(reinterpret_cast<void**>(&e->tile_counters), kTileCounters * sizeof(int)) != cudaSuccess)
        return bail(fail(e, B200OCR_E_CUDA, "cudaMalloc(tile counters) failed"));
    e->owned.push_back(e->tile_counters);
    {
        int lo = 0, hi = 0;                        // numerically lower = higher priority
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&e->hi_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&e->hi_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&e->hi_join, cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(e, B200OCR_E_CUDA, "side stream for the recurrence could not be created"));
        if (const char* v = getenv("B200OCR_LSTM_PRIORITY")) e->lstm_priority = atoi(v) != 0;   // A/B runs of bench.py
    }
    if (desc->precision == B200OCR_PREC_FP16F8W)   // preset: weight-side correction only in the deep 3x3 layers
        for (LayerRT& ly : e->layers)
            if (ly.kind == B200OCR_CONV && ly.g.kh * ly.g.kw == 9 && ly.g.cin >= 256) ly.g.corr = CORR_WEIGHT;
    *out = e;
    return B200OCR_OK;
}

int b200ocr_reserve(b200ocr_engine_t* e, int32_t max_lines, int32_t max_width_px) {
    if (!e || max_lines <= 0 || max_width_px <= 0) return fail(e, B200OCR_E_INVALID, "bad reserve arguments");
    CU_TRY(e, cudaSetDevice(e->device));
    Need need;
    Outputs none;
    if (int s = walk(e, nullptr, max_lines, e->line_height, max_width_px, 1 << 30, none, nullptr, &need, nullptr,
                     nullptr, nullptr))
        return s;
    for (int i = 0; i < 3; ++i) {
        if (need.h[i] > e->hbuf_bytes[i]) {
            if (e->hbuf[i]) cudaFree(e->hbuf[i]);
            e->hbuf[i] = nullptr; e->hbuf_bytes[i] = 0;
            CU_TRY(e, cudaMalloc(&e->hbuf[i], need.h[i]));
            e->hbuf_bytes[i] = need.h[i];
        }
        if (need.f[i] > e->fbuf_bytes[i]) {
            if (e->fbuf[i]) cudaFree(e->fbuf[i]);
            e->fbuf[i] = nullptr; e->fbuf_bytes[i] = 0;
            CU_TRY(e, cudaMalloc(&e->fbuf[i], need.f[i]));
            e->fbuf_bytes[i] = need.f[i];
        }
    }
    if (need.frames > e->frames_cap) {
        if (e->best) { cudaFree(e->best); cudaFree(e->fmax); cudaFree(e->flse); cudaFree(e->fprob); }
        e->frames_cap = 0;
        CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->best), need.frames * 4));
        CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->fmax), need.frames * 4));
        CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->flse), need.frames * 4));
        CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->fprob), need.frames * 4));
        e->frames_cap = need.frames;
    }
    return B200OCR_OK;
}

int b200ocr_forward(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w, float* logits,
                    int32_t* labels, int32_t* lengths, float* confidence, int32_t* best_path, void* cuda_stream) {
    if (!e) return B200OCR_E_INVALID;
    if (!crops || n <= 0 || h != e->line_height || w <= 0 || (w % 8))
        return fail(e, B200OCR_E_INVALID, "bad forward arguments (n=%d h=%d w=%d; h must be %d, w a multiple of 8)", n, h,
                    w, e->line_height);
    if (labels && !lengths) return fail(e, B200OCR_E_INVALID, "labels requires lengths");
    Outputs o;
    o.logits = logits; o.labels = labels; o.lengths = lengths; o.confidence = confidence; o.best_path = best_path;
    return walk(e, crops, n, h, w, 1 << 30, o, static_cast<cudaStream_t>(cuda_stream), nullptr, nullptr, nullptr, nullptr);
}

int b200ocr_forward_maps(b200ocr_engine_t* e, const uint8_t* image, int32_t h, int32_t w, float* maps, void* cuda_stream) {
    if (!e) return B200OCR_E_INVALID;
    if (!image || !maps || h <= 0 || w <= 0) return fail(e, B200OCR_E_INVALID, "bad forward_maps arguments");
    Need need;
    Outputs none;
    if (int s = walk(e, nullptr, 1, h, w, 1 << 30, none, nullptr, &need, nullptr, nullptr, nullptr)) return s;
    for (int i = 0; i < 3; ++i)
        if (need.h[i] > e->hbuf_bytes[i] || need.f[i] > e->fbuf_bytes[i])
            return fail(e, B200OCR_E_WORKSPACE, "workspace too small: call b200ocr_reserve_maps first");
    Outputs o;
    o.maps = maps;
    return walk(e, image, 1, h, w, 1 << 30, o, static_cast<cudaStream_t>(cuda_stream), nullptr, nullptr, nullptr, nullptr);
}

int b200ocr_reserve_maps(b200ocr_engine_t* e, int32_t max_h, int32_t max_w) {
    if (!e || max_h <= 0 || max_w <= 0) return fail(e, B200OCR_E_INVALID, "bad reserve arguments");
    const int saved = e->line_height;
    e->line_height = max_h;
    const int s = b200ocr_reserve(e, 1, max_w);
    e->line_height = saved;
    return s;
}

void b200ocr_destroy(b200ocr_engine_t* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->owned) cudaFree(p);
    for (int i = 0; i < 3; ++i) {
        if (e->hbuf[i]) cudaFree(e->hbuf[i]);
        if (e->fbuf[i]) cudaFree(e->fbuf[i]);
    }
    if (e->best) { cudaFree(e->best); cudaFree(e->fmax); cudaFree(e->flse); cudaFree(e->fprob); }
    for (void* p : e->ar.ws) cudaFree(p);
    if (e->ar.h_state) cudaFreeHost(e->ar.h_state);
    if (e->ar.gexec) cudaGraphExecDestroy(e->ar.gexec);
    if (e->ar.loop_fork) cudaEventDestroy(e->ar.loop_fork);
    if (e->ar.loop_join) cudaEventDestroy(e->ar.loop_join);
    if (e->ar.loop_stream) cudaStreamDestroy(e->ar.loop_stream);
    if (e->after) {
        auto& f = e->after->followers;
        f.erase(std::remove(f.begin(), f.end(), e), f.end());
    }
    for (b200ocr_engine* f : e->followers) f->after = nullptr;
    if (e->front_done) cudaEventDestroy(e->front_done);
    if (e->hi_fork) cudaEventDestroy(e->hi_fork);
    if (e->hi_join) cudaEventDestroy(e->hi_join);
    if (e->hi_stream) cudaStreamDestroy(e->hi_stream);
    delete e;
}

const char* b200ocr_last_error(const b200ocr_engine_t* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int64_t b200ocr_launch_count(const b200ocr_engine_t* e) { return e ? e->launches : 0; }

double b200ocr_forward_flops(const b200ocr_engine_t* e, int32_t n, int32_t w, double* conv_gemm_flops) {
    if (!e) return 0.0;
    double total = 0.0, gemm = 0.0;
    int h = e->line_height, cw = w;
    for (const LayerRT& ly : e->layers) {
        auto gf = [&](const Gemm& g, double pixels) { return 2.0 * pixels * g.cout * g.cin * g.kh * g.kw; };
        switch (ly.kind) {
            case B200OCR_CONV_FIRST:
                total += 2.0 * n * h * cw * ly.cout0 * 27;
                break;
            case B200OCR_CONV: {
                const int ho = h + 2 * ly.g.pad_h - ly.g.kh + 1, wo = cw + 2 * ly.g.pad_w - ly.g.kw + 1;
                const double f = gf(ly.g, static_cast<double>(n) * ho * wo);
                total += f; gemm += f;
                h = ho / ly.pool_h; cw = wo / ly.pool_w;
                break;
            }
            case B200OCR_BILSTM: {
                const double f = gf(ly.g, static_cast<double>(n) * cw);
                const double r = 2.0 * n * cw * 2 * 4 * ly.hidden * ly.hidden;
                total += f + r; gemm += f;
                break;
            }
            case B200OCR_CTC_HEAD: {
                const double f = gf(ly.g, static_cast<double>(n) * cw);
                total += f; gemm += f;
                break;
            }
            case B200OCR_TRANSFORMER_LAYER: {
                const double px = static_cast<double>(n) * cw;
                const double f = gf(ly.g_in, px) + gf(ly.g_out, px) + gf(ly.g_l1, px) + gf(ly.g_l2, px);
                const double a = 4.0 * n * cw * cw * ly.g_out.cin;
                total += f + a; gemm += f;
                break;
            }
            default: break;
        }
    }
    if (conv_gemm_flops) *conv_gemm_flops = gemm;
    return total;
}

int b200ocr_set_layer_correction(b200ocr_engine_t* e, int32_t layer, int32_t mode) {
    if (!e) return B200OCR_E_INVALID;
    if (layer < 0 || layer >= (int)e->layers.size() || mode < CORR_BOTH || mode > CORR_NONE)
        return fail(e, B200OCR_E_INVALID, "bad layer correction (layer %d, mode %d)", layer, mode);
    LayerRT& ly = e->layers[layer];
    switch (ly.kind) {
        case B200OCR_CONV: case B200OCR_CTC_HEAD: case B200OCR_BILSTM: ly.g.corr = mode; break;
        case B200OCR_TRANSFORMER_LAYER: ly.g_in.corr = ly.g_out.corr = ly.g_l1.corr = ly.g_l2.corr = mode; break;
        default: return fail(e, B200OCR_E_INVALID, "layer %d has no tensor-core contraction", layer);
    }
    return B200OCR_OK;
}

int b200ocr_set_layer_post_shift(b200ocr_engine_t* e, int32_t layer, const float* shift) {
    if (!e) return B200OCR_E_INVALID;
    if (layer < 0 || layer >= static_cast<int>(e->layers.size()) || !shift)
        return fail(e, B200OCR_E_INVALID, "bad set_layer_post_shift arguments (layer %d)", layer);
    LayerRT& ly = e->layers[layer];
    if (ly.kind != B200OCR_CONV || !ly.g.post_shift)
        return fail(e, B200OCR_E_INVALID, "layer %d has no post-activation affine (create it with post_scale / post_shift)", layer);
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev != e->device) cudaSetDevice(e->device);
    cudaError_t err = cudaDeviceSynchronize();                     // no forward may be reading the old shift
    if (err == cudaSuccess) err = cudaMemcpy(ly.g.post_shift, shift, ly.g.cout * sizeof(float), cudaMemcpyHostToDevice);
    if (cur_dev != e->device) cudaSetDevice(cur_dev);
    if (err != cudaSuccess) return fail(e, B200OCR_E_CUDA, "set_layer_post_shift: %s", cudaGetErrorString(err));
    return B200OCR_OK;
}

int b200ocr_run_after(b200ocr_engine_t* e, b200ocr_engine_t* after) {
    if (!e || after == e) return fail(e, B200OCR_E_INVALID, "bad run_after arguments");
    if (after && after->device != e->device) return fail(e, B200OCR_E_INVALID, "engines live on different devices");
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev != e->device) cudaSetDevice(e->device);
    if (e->after) {
        auto& f = e->after->followers;
        f.erase(std::remove(f.begin(), f.end(), e), f.end());
        e->after = nullptr;
    }
    if (!after) {
        if (cur_dev != e->device) cudaSetDevice(cur_dev);
        return B200OCR_OK;
    }
    cudaError_t err = after->front_done ? cudaSuccess : cudaEventCreateWithFlags(&after->front_done, cudaEventDisableTiming);
    if (cur_dev != e->device) cudaSetDevice(cur_dev);
    if (err != cudaSuccess) return fail(e, B200OCR_E_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(err));
    after->followers.push_back(e);
    e->after = after;
    return B200OCR_OK;
}

double b200ocr_executed_passes(const b200ocr_engine_t* e, int32_t n, int32_t w, int32_t capacity, float* per_layer) {
    if (!e) return 0.0;
    // tensor-core pass-equivalents a contraction executes: 1 (fp16), 3 (fp16x3), or 1 + the share of the 2*cin-byte
    // e5m2 operand the correction pass walks (igemm_tc.cu: whole 128-byte chunks, starting at chunk kchunks/2 for the
    // weight-side-only mode; an e5m2 MMA of K = 32 takes as long as an fp16 MMA of K = 16)
    auto passes = [&](const Gemm& g) -> double {
        if (e->fmt == ACT_F16) return 1.0;
        if (e->fmt == ACT_F16_HILO) return 3.0;
        const int kchunks = g.cin / 64;
        if (g.corr == CORR_NONE) return 1.0;
        if (g.corr == CORR_WEIGHT) {
            if (g.bn_halo && kchunks == 1) return 1.5;       // halo kernel, cin = 64: upper two K steps of the chunk
            return 1.0 + static_cast<double>(kchunks - (kchunks >> 1)) / kchunks;
        }
        return 2.0;
    };
    double flops = 0.0, weighted = 0.0;
    int h = e->line_height, cw = w, li = 0;
    for (const LayerRT& ly : e->layers) {
        auto gf = [&](const Gemm& g, double pixels) { return 2.0 * pixels * g.cout * g.cin * g.kh * g.kw; };
        double f = 0.0, fw = 0.0;
        switch (ly.kind) {
            case B200OCR_CONV: {
                const int ho = h + 2 * ly.g.pad_h - ly.g.kh + 1, wo = cw + 2 * ly.g.pad_w - ly.g.kw + 1;
                f = gf(ly.g, static_cast<double>(n) * ho * wo); fw = f * passes(ly.g);
                h = ho / ly.pool_h; cw = wo / ly.pool_w;
                break;
            }
            case B200OCR_BILSTM: case B200OCR_CTC_HEAD:
                f = gf(ly.g, static_cast<double>(n) * cw); fw = f * passes(ly.g);
                break;
            case B200OCR_TRANSFORMER_LAYER: {
                const double px = static_cast<double>(n) * cw;
                for (const Gemm* g : {&ly.g_in, &ly.g_out, &ly.g_l1, &ly.g_l2}) { f += gf(*g, px); fw += gf(*g, px) * passes(*g); }
                break;
            }
            default: break;
        }
        if (per_layer && li < capacity) per_layer[li] = f > 0.0 ? static_cast<float>(fw / f) : 0.f;
        flops += f; weighted += fw;
        ++li;
    }
    return flops > 0.0 ? weighted / flops : 0.0;
}

int b200ocr_ctc_greedy(const float* scores, int32_t n, int32_t t, int32_t c, int32_t layout, int32_t* labels,
                       int32_t* lengths, float* confidence, int32_t* best_path, float* frame_max, float* frame_lse,
                       void* cuda_stream) {
    if (!scores || n < 0 || t <= 0 || c <= 1 || !labels || !lengths || !best_path || (layout != 0 && layout != 1))
        return fail(nullptr, B200OCR_E_INVALID, "bad ctc_greedy arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(scores);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    float* fprob = nullptr;
    if (confidence) CU_TRY(nullptr, cudaMallocAsync(reinterpret_cast<void**>(&fprob), static_cast<size_t>(n) * t * 4, st));
    CU_TRY(nullptr, launch_frame_stats(scores, n, t, c, layout, best_path, frame_max, frame_lse, fprob, st));
    CU_TRY(nullptr, launch_ctc_collapse(best_path, fprob, n, t, c - 1, labels, lengths, confidence, st));
    if (fprob) CU_TRY(nullptr, cudaFreeAsync(fprob, st));
    return B200OCR_OK;
}

int b200ocr_force_align(const void* neg_logprobs, int32_t is_f64, int32_t n, int32_t t, int32_t c,
                        const int32_t* n_frames, const int32_t* labels, int32_t l_max, const int32_t* lengths,
                        int32_t blank, int32_t* out_symbols, int32_t* out_positions, int32_t* char_positions,
                        int32_t* status, void* cuda_stream) {
    if (!neg_logprobs || n < 0 || t <= 0 || c <= 1 || !labels || l_max < 1 || !lengths || blank < 0 || blank >= c ||
        !status || (char_positions && !out_positions))
        return fail(nullptr, B200OCR_E_INVALID, "bad force_align arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(neg_logprobs);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint8_t* ws = nullptr;
    CU_TRY(nullptr, cudaMallocAsync(reinterpret_cast<void**>(&ws), force_align_workspace_bytes(n, t, l_max), st));
    cudaError_t err = launch_force_align(neg_logprobs, is_f64, n, t, c, n_frames, labels, l_max, lengths, blank, ws,
                                         out_symbols, out_positions, char_positions, status, st);
    cudaFreeAsync(ws, st);
    CU_TRY(nullptr, err);
    return B200OCR_OK;
}

int b200ocr_char_confidence(const float* log_probs, int32_t n, int32_t t, int32_t c, const int32_t* n_frames,
                            const int32_t* labels, int32_t l_max, const int32_t* lengths,
                            const int32_t* char_positions, float* confidences, void* cuda_stream) {
    if (!log_probs || n < 0 || t <= 0 || c <= 1 || !labels || l_max < 1 || !lengths || !char_positions || !confidences)
        return fail(nullptr, B200OCR_E_INVALID, "bad char_confidence arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(log_probs);
    CU_TRY(nullptr, launch_char_conf(log_probs, n, t, c, n_frames, labels, l_max, lengths, char_positions, confidences,
                                     static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_remap_lines(const uint8_t* image, int32_t img_h, int32_t img_w, const float* coords,
                        const int64_t* coord_off, const int32_t* widths, int32_t n, int32_t line_h, uint8_t* out,
                        int32_t out_w, int32_t pad, void* cuda_stream) {
    if (!image || img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767 || n < 0 || line_h <= 0 || !out ||
        out_w <= 0 || pad < 0 || (n > 0 && (!coords || !coord_off || !widths)))
        return fail(nullptr, B200OCR_E_INVALID, "bad remap_lines arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(out);
    CU_TRY(nullptr, launch_remap_lines(image, img_h, img_w, coords, coord_off, widths, n, line_h, out, out_w, pad,
                                       static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_memcpy_d2h_async(void* host_dst, const void* device_src, int64_t bytes, void* cuda_stream) {
    if (bytes < 0 || (bytes > 0 && (!host_dst || !device_src)))
        return fail(nullptr, B200OCR_E_INVALID, "bad memcpy_d2h_async arguments");
    if (bytes == 0) return B200OCR_OK;
    DeviceGuard guard(device_src);
    CU_TRY(nullptr, cudaMemcpyAsync(host_dst, device_src, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost,
                                    static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_pad_lines(const uint8_t* packed, const int64_t* line_off, const int32_t* widths, int32_t n, int32_t line_h,
                      uint8_t* out, int32_t out_w, int32_t pad, void* cuda_stream) {
    if (n < 0 || line_h <= 0 || !out || out_w <= 0 || (out_w % 4) || pad < 0 || (n > 0 && (!packed || !line_off || !widths)))
        return fail(nullptr, B200OCR_E_INVALID, "bad pad_lines arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(out);
    CU_TRY(nullptr, launch_pad_lines(packed, line_off, widths, n, line_h, out, out_w, pad,
                                     static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_remap_poly_lines(const uint8_t* image, int32_t img_h, int32_t img_w, const b200ocr_poly_line_t* lines,
                             const double* offsets, int32_t n, int32_t line_h, uint8_t* out, int32_t out_w,
                             int32_t pad, void* cuda_stream) {
    if (!image || img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767 || n < 0 || line_h <= 0 || !out ||
        out_w <= 0 || pad < 0 || (n > 0 && (!lines || !offsets)))
        return fail(nullptr, B200OCR_E_INVALID, "bad remap_poly_lines arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(out);
    CU_TRY(nullptr, launch_remap_poly_lines(image, img_h, img_w, lines, offsets, n, line_h, out, out_w, pad,
                                            static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_sparsify_logits(const float* logits, int32_t n, int32_t t, int32_t c, const int32_t* t_lo,
                            const int32_t* t_hi, int32_t* indptr, int32_t* nnz, int64_t* base, int32_t* indices,
                            float* data, int64_t capacity, void* cuda_stream) {
    if (!logits || n < 0 || t <= 0 || c <= 0 || !indptr || !nnz || !base || !indices || !data || capacity < 0)
        return fail(nullptr, B200OCR_E_INVALID, "bad sparsify_logits arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(logits);
    CU_TRY(nullptr, launch_sparsify(logits, n, t, c, t_lo, t_hi, indptr, nnz, base, indices, data, capacity,
                                    static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_ctc_prefix_beam_ranges(const double* logprobs, int32_t n, int32_t t, int32_t c, int32_t k,
                                   const int32_t* t_lo, const int32_t* t_hi, int32_t* out_labels, int32_t* out_lengths,
                                   double* out_scores, int32_t* status, void* cuda_stream) {
    if (!logprobs || n < 0 || t <= 0 || c <= 1 || k < 1 || !out_labels || !out_lengths || !out_scores || !status ||
        ((t_lo == nullptr) != (t_hi == nullptr)))
        return fail(nullptr, B200OCR_E_INVALID, "bad ctc_prefix_beam arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(logprobs);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    void* ws = nullptr;
    const size_t ws_bytes = ctc_beam_workspace_bytes(n, t, c, k);
    if (ws_bytes == 0) return fail(nullptr, B200OCR_E_INVALID, "beam size / class count not supported (k <= 64, c <= 1024)");
    CU_TRY(nullptr, cudaMallocAsync(&ws, ws_bytes, st));
    cudaError_t err = launch_ctc_prefix_beam(logprobs, n, t, c, k, t_lo, t_hi, out_labels, out_lengths, out_scores, status,
                                             ws, st);
    cudaFreeAsync(ws, st);
    if (err != cudaSuccess) return fail(nullptr, B200OCR_E_CUDA, "prefix beam launch failed: %s", cudaGetErrorString(err));
    return B200OCR_OK;
}

int b200ocr_ctc_prefix_beam(const double* logprobs, int32_t n, int32_t t, int32_t c, int32_t k, int32_t* out_labels,
                            int32_t* out_lengths, double* out_scores, int32_t* status, void* cuda_stream) {
    return b200ocr_ctc_prefix_beam_ranges(logprobs, n, t, c, k, nullptr, nullptr, out_labels, out_lengths, out_scores,
                                          status, cuda_stream);
}

int b200ocr_full_logprobs(const float* logits, int32_t n, int32_t t, int32_t c, double* logprobs, void* cuda_stream) {
    if (!logits || n < 0 || t <= 0 || c <= 0 || !logprobs)
        return fail(nullptr, B200OCR_E_INVALID, "bad full_logprobs arguments");
    if (n == 0) return B200OCR_OK;
    DeviceGuard guard(logits);
    CU_TRY(nullptr, launch_full_logprobs(logits, n, t, c, logprobs, static_cast<cudaStream_t>(cuda_stream)));
    return B200OCR_OK;
}

int b200ocr_ar_attach(b200ocr_engine_t* e, const b200ocr_ar_desc_t* d) {
    if (!e) return B200OCR_E_INVALID;
    if (!d || d->n_layers <= 0 || !d->layers || d->heads <= 0 || d->dim_ff <= 0 || d->classes <= 1 || !d->embed ||
        !d->out_w || !d->out_b)
        return fail(e, B200OCR_E_INVALID, "bad decoder descriptor");
    if (e->ar.attached) return fail(e, B200OCR_E_INVALID, "a decoder is already attached to this engine");
    if (e->layers.empty() || e->layers.back().kind != B200OCR_TRANSFORMER_LAYER)
        return fail(e, B200OCR_E_INVALID, "the decoder attaches to an engine that ends in the Transformer encoder");
    CU_TRY(e, cudaSetDevice(e->device));
    const int D = e->layers.back().g_out.cout;
    const int hd = D / d->heads;
    if (hd * d->heads != D || (hd % 4) || hd > 128 || (D % 32) || (d->dim_ff % 32))
        return fail(e, B200OCR_E_INVALID, "decoder needs head width %% 4 == 0 (<= 128) and D, dim_ff %% 32 == 0");
    ArState& ar = e->ar;
    ar.heads = d->heads; ar.dim_ff = d->dim_ff; ar.classes = d->classes; ar.D = D;
    ar.layers.resize(d->n_layers);
    const size_t DD = static_cast<size_t>(D) * D;
    for (int i = 0; i < d->n_layers; ++i) {
        const b200ocr_ar_layer_t& src = d->layers[i];
        ArLayer& L = ar.layers[i];
        const float* all[] = {src.self_in_w, src.self_in_b, src.self_out_w, src.self_out_b, src.cross_in_w,
                              src.cross_in_b, src.cross_out_w, src.cross_out_b, src.lin1_w, src.lin1_b, src.lin2_w,
                              src.lin2_b, src.norm1_w, src.norm1_b, src.norm2_w, src.norm2_b, src.norm3_w, src.norm3_b};
        for (const float* p : all)
            if (!p) return fail(e, B200OCR_E_INVALID, "decoder layer %d: missing parameter", i);
        int s = 0;
        if ((s = upload(e, src.self_in_w, 3 * DD, &L.self_in_w))) return s;
        if ((s = upload(e, src.self_in_b, (size_t)3 * D, &L.self_in_b))) return s;
        if ((s = upload(e, src.self_out_w, DD, &L.self_out_w))) return s;
        if ((s = upload(e, src.self_out_b, (size_t)D, &L.self_out_b))) return s;
        if ((s = upload(e, src.cross_in_w, DD, &L.cross_q_w))) return s;          // query rows only
        if ((s = upload(e, src.cross_in_b, (size_t)D, &L.cross_q_b))) return s;
        if ((s = upload(e, src.cross_out_w, DD, &L.cross_out_w))) return s;
        if ((s = upload(e, src.cross_out_b, (size_t)D, &L.cross_out_b))) return s;
        if ((s = upload(e, src.lin1_w, (size_t)d->dim_ff * D, &L.l1w))) return s;
        if ((s = upload(e, src.lin1_b, (size_t)d->dim_ff, &L.l1b))) return s;
        if ((s = upload(e, src.lin2_w, (size_t)d->dim_ff * D, &L.l2w))) return s;
        if ((s = upload(e, src.lin2_b, (size_t)D, &L.l2b))) return s;
        if ((s = upload(e, src.norm1_w, (size_t)D, &L.n1w))) return s;
        if ((s = upload(e, src.norm1_b, (size_t)D, &L.n1b))) return s;
        if ((s = upload(e, src.norm2_w, (size_t)D, &L.n2w))) return s;
        if ((s = upload(e, src.norm2_b, (size_t)D, &L.n2b))) return s;
        if ((s = upload(e, src.norm3_w, (size_t)D, &L.n3w))) return s;
        if ((s = upload(e, src.norm3_b, (size_t)D, &L.n3b))) return s;
        if ((s = build_gemm(e, L.g_memkv, src.cross_in_w + DD, src.cross_in_b + D, nullptr, nullptr, D, 2 * D, 1, 1, 0, 0)))
            return s;
    }
    int s = 0;
    if ((s = upload(e, d->embed, (size_t)d->classes * D, &ar.embed))) return s;
    if ((s = upload(e, d->out_w, (size_t)d->classes * D, &ar.out_w))) return s;
    if ((s = upload(e, d->out_b, (size_t)d->classes, &ar.out_b))) return s;
    CU_TRY(e, cudaMallocHost(reinterpret_cast<void**>(&ar.h_state), 2 * sizeof(int32_t)));
    ar.attached = true;
    return B200OCR_OK;
}

int b200ocr_ar_reserve(b200ocr_engine_t* e, int32_t max_lines, int32_t max_width_px, int32_t max_steps) {
    if (!e) return B200OCR_E_INVALID;
    if (!e->ar.attached) return fail(e, B200OCR_E_INVALID, "no decoder attached (b200ocr_ar_attach)");
    if (max_lines <= 0 || max_width_px <= 0 || max_steps <= 0) return fail(e, B200OCR_E_INVALID, "bad reserve arguments");
    if (int s = b200ocr_reserve(e, max_lines, max_width_px)) return s;
    Need need;
    Outputs none;
    Shape fs;
    if (int s = walk(e, nullptr, max_lines, e->line_height, max_width_px, 1 << 30, none, nullptr, &need, &fs, nullptr,
                     nullptr))
        return s;
    ArState& ar = e->ar;
    if (fs.h != 1 || fs.c != ar.D) return fail(e, B200OCR_E_INVALID, "encoder output is not a [n][1][T][%d] sequence", ar.D);
    if (max_lines <= ar.cap_lines && fs.w <= ar.cap_T && max_steps <= ar.cap_steps) return B200OCR_OK;
    const int N = std::max(max_lines, ar.cap_lines), T = std::max(fs.w, ar.cap_T), S = std::max(max_steps, ar.cap_steps);
    for (void* p : ar.ws) cudaFree(p);
    ar.ws.clear();
    ar.cap_lines = ar.cap_T = ar.cap_steps = 0;
    const size_t L = ar.layers.size(), D = ar.D;
    auto grab = [&](size_t bytes, void** out) -> cudaError_t {
        cudaError_t err = cudaMalloc(out, std::max<size_t>(bytes, 16));
        if (err == cudaSuccess) ar.ws.push_back(*out);
        return err;
    };
    CU_TRY(e, grab(L * N * T * 2 * D * sizeof(float), reinterpret_cast<void**>(&ar.memkv)));
    CU_TRY(e, grab(L * S * N * 2 * D * sizeof(float), reinterpret_cast<void**>(&ar.selfkv)));
    CU_TRY(e, grab(static_cast<size_t>(N) * D * sizeof(float), reinterpret_cast<void**>(&ar.x)));
    CU_TRY(e, grab(static_cast<size_t>(N) * D * sizeof(float), reinterpret_cast<void**>(&ar.q)));
    CU_TRY(e, grab(static_cast<size_t>(N) * D * sizeof(float), reinterpret_cast<void**>(&ar.a)));
    CU_TRY(e, grab(static_cast<size_t>(N) * D * sizeof(float), reinterpret_cast<void**>(&ar.t)));
    CU_TRY(e, grab(static_cast<size_t>(kArSplitMax) * N * D * sizeof(float), reinterpret_cast<void**>(&ar.part)));
    CU_TRY(e, grab(static_cast<size_t>(N) * ar.dim_ff * sizeof(float), reinterpret_cast<void**>(&ar.f)));
    CU_TRY(e, grab(static_cast<size_t>(N) * ar.classes * sizeof(float), reinterpret_cast<void**>(&ar.lg)));
    CU_TRY(e, grab(static_cast<size_t>(N) * sizeof(int32_t), reinterpret_cast<void**>(&ar.alive)));
    CU_TRY(e, grab(4 * sizeof(int32_t), reinterpret_cast<void**>(&ar.state)));
    ar.cap_lines = N; ar.cap_T = T; ar.cap_steps = S;
    return B200OCR_OK;
}

int b200ocr_ar_transcribe(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w,
                          int32_t start_token, int32_t max_steps, int32_t check_every, int32_t* tokens, float* logits,
                          int32_t* steps, void* cuda_stream) {
    if (!e) return B200OCR_E_INVALID;
    ArState& ar = e->ar;
    if (!ar.attached) return fail(e, B200OCR_E_INVALID, "no decoder attached (b200ocr_ar_attach)");
    if (!crops || n <= 0 || h != e->line_height || w <= 0 || (w % 8) || !tokens || !steps || max_steps <= 0 ||
        start_token < 0 || start_token >= ar.classes)
        return fail(e, B200OCR_E_INVALID, "bad ar_transcribe arguments (n=%d h=%d w=%d max_steps=%d)", n, h, w, max_steps);
    if (check_every <= 0) check_every = 1;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // TransformerOCR.encode (transformer.py:548-555): the layer walk; memory = records in hbuf[0], rows n*T + t
    Outputs none;
    Shape fs;
    const void* mem = nullptr;
    bool mem_f32 = false;
    if (int s = walk(e, crops, n, h, w, 1 << 30, none, st, nullptr, &fs, &mem, &mem_f32)) return s;
    const int D = ar.D, T = fs.w, C = ar.classes, FF = ar.dim_ff;
    if (mem_f32 || !mem || fs.h != 1 || fs.c != D) return fail(e, B200OCR_E_INVALID, "encoder output is not a [n][1][T][%d] sequence", D);
    if (n > ar.cap_lines || T > ar.cap_T || max_steps > ar.cap_steps)
        return fail(e, B200OCR_E_WORKSPACE, "workspace too small: call b200ocr_ar_reserve first");
    const int L = static_cast<int>(ar.layers.size());
    const size_t memkv_layer = static_cast<size_t>(ar.cap_lines) * ar.cap_T * 2 * D;
    const size_t selfkv_layer = static_cast<size_t>(ar.cap_steps) * ar.cap_lines * 2 * D;
#define AR_LAUNCH(call)      \
    do {                     \
        CU_TRY(e, call);     \
        e->launches++;       \
    } while (0)
    // K | V of the memory for every layer, once per batch (cached_forward's static branch, transformer.py:239-249)
    for (int i = 0; i < L; ++i) {
        EpiOut eo;
        eo.epi = EPI_F32;
        eo.out_f32 = ar.memkv + i * memkv_layer;
        e->cur_layer = static_cast<int>(e->layers.size()) + i;
        if (int s = run_gemm(e, ar.layers[i].g_memkv, static_cast<const __half*>(mem), Shape{1, 1, n * T, D}, 0, 1, 1, eo,
                             st, nullptr))
            return s;
    }
    AR_LAUNCH(launch_ar_init(ar.alive, n, ar.state, st));
    int done_steps = -1;
    const int lv = ar.linear_variant >= 2 ? 1 : ar.linear_variant;      // kernel of the plain projections
    const bool fused = ar.linear_variant >= 2 && sum_layernorm_supported(D);
    const int att = ar.linear_variant >= 2 ? 1 : 0;
    const long part_stride = static_cast<long>(ar.cap_lines) * D;
    auto ksplit = [](int K) { return std::max(1, std::min(kArSplitMax, K / 32 / 8)); };   // >= 8 chunks per CTA
    const int hd = ar.heads > 0 ? D / ar.heads : 0;
    // The launches of one decoded position.  pos_dev == nullptr: position `s` as immediates.  pos_dev != nullptr
    // (token loop v3): the position is the device-side counter state[2] -- cache slot, token row, logits row and the
    // self-attention length are derived from it inside the kernels, so the arguments are the same for every
    // position and the 25 launches are captured ONCE into a CUDA graph and replayed.
    auto position = [&](int s, cudaStream_t ls, const int32_t* pos_dev) -> int {
        // norm(x + W a + b): the projection's K split over CTAs, summed in the LayerNorm (fused), or the fused epilogue
        // of one CTA per tile + the plain LayerNorm
        auto proj_norm = [&](const float* a, int K, const float* w, const float* b, const float* gamma,
                             const float* beta) -> int {
            if (fused) {
                const int Z = ksplit(K);
                AR_LAUNCH(launch_linear_f32_ex(a, K, w, nullptr, nullptr, 0, ar.part, D, n, D, K, 0, D, nullptr, 0, Z,
                                               part_stride, ls));
                AR_LAUNCH(launch_sum_layernorm(ar.part, Z, part_stride, b, ar.x, n, D, gamma, beta, 1e-5f, ar.x, ls));
            } else {
                AR_LAUNCH(launch_linear_f32(a, K, w, b, ar.x, D, ar.t, D, n, D, K, 0, lv, ls));
                AR_LAUNCH(launch_layernorm(ar.t, n, D, gamma, beta, 1e-5f, 0, ar.x, nullptr, e->fmt, ls));
            }
            return B200OCR_OK;
        };
        const int32_t* prev = pos_dev ? tokens : (s == 0 ? nullptr : tokens + static_cast<size_t>(s - 1) * n);
        AR_LAUNCH(launch_embed_pe(ar.embed, prev, start_token, n, D, s, ar.x, ls, pos_dev));
        for (int i = 0; i < L; ++i) {
            const ArLayer& ly = ar.layers[i];
            float* kv = ar.selfkv + i * selfkv_layer;                       // position p at + p * n * 2D
            float* kv_s = pos_dev ? kv : kv + static_cast<size_t>(s) * n * 2 * D;
            const float* mkv = ar.memkv + i * memkv_layer;                  // row (line * T + t) * 2D
            // cached self-attention over positions 0..s (DecoderLayer.infer, transformer.py:431-435)
            if (ar.linear_variant >= 2) {
                AR_LAUNCH(launch_linear_f32_ex(ar.x, D, ly.self_in_w, ly.self_in_b, nullptr, 0, ar.q, D, n, 3 * D, D, 0, D,
                                               kv_s, 2 * D, 1, 0, ls, pos_dev, 0, static_cast<long>(n) * 2 * D));
            } else {
                AR_LAUNCH(launch_linear_f32(ar.x, D, ly.self_in_w, ly.self_in_b, nullptr, 0, ar.q, D, n, D, D, 0, lv, ls));
                AR_LAUNCH(launch_linear_f32(ar.x, D, ly.self_in_w + static_cast<size_t>(D) * D, ly.self_in_b + D, nullptr, 0,
                                            kv_s, 2 * D, n, 2 * D, D, 0, lv, ls));
            }
            AR_LAUNCH(launch_step_attention(ar.q, D, kv, kv + D, static_cast<long>(n) * 2 * D, 2 * D, n,
                                            pos_dev ? max_steps : s + 1, D, ar.heads, ar.a, att, ls, pos_dev));
            if (int r = proj_norm(ar.a, D, ly.self_out_w, ly.self_out_b, ly.n1w, ly.n1b)) return r;
            // encoder-decoder attention over the T memory frames (:438-447)
            AR_LAUNCH(launch_linear_f32(ar.x, D, ly.cross_q_w, ly.cross_q_b, nullptr, 0, ar.q, D, n, D, D, 0, lv, ls));
            AR_LAUNCH(launch_step_attention(ar.q, D, mkv, mkv + D, 2 * D, static_cast<long>(T) * 2 * D, n, T, D, ar.heads,
                                            ar.a, att, ls));
            if (int r = proj_norm(ar.a, D, ly.cross_out_w, ly.cross_out_b, ly.n2w, ly.n2b)) return r;
            // feed-forward (:449-450)
            AR_LAUNCH(launch_linear_f32(ar.x, D, ly.l1w, ly.l1b, nullptr, 0, ar.f, FF, n, FF, D, 1, lv, ls));
            if (int r = proj_norm(ar.f, FF, ly.l2w, ly.l2b, ly.n3w, ly.n3b)) return r;
        }
        // dec_out_proj + argmax + alive mask (transformer_ocr_engine.py:69-75)
        const long lg_ld = logits ? static_cast<long>(max_steps) * C : C;
        const long lg_pos = logits ? C : 0;                                 // the logits row of a position
        if (pos_dev) {
            float* lg = logits ? logits : ar.lg;
            AR_LAUNCH(launch_linear_f32_ex(ar.x, D, ar.out_w, ar.out_b, nullptr, 0, lg, lg_ld, n, C, D, 0, C, nullptr, 0, 1, 0,
                                           ls, pos_dev, lg_pos, 0));
            AR_LAUNCH(launch_argmax_alive(lg, lg_ld, n, C, start_token, 0, tokens, ar.alive, ar.state, ls, 1, lg_pos));
        } else {
            float* lg = logits ? logits + static_cast<size_t>(s) * C : ar.lg;
            AR_LAUNCH(launch_linear_f32(ar.x, D, ar.out_w, ar.out_b, nullptr, 0, lg, lg_ld, n, C, D, 0, lv, ls));
            AR_LAUNCH(launch_argmax_alive(lg, lg_ld, n, C, start_token, s, tokens + static_cast<size_t>(s) * n, ar.alive,
                                          ar.state, ls));
        }
        return B200OCR_OK;
    };
    auto finished = [&](cudaStream_t ls) -> int {          // 1: every line has emitted the stop symbol, 0: not yet, < 0: error
        if (cudaMemcpyAsync(ar.h_state, ar.state, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ls) != cudaSuccess ||
            cudaStreamSynchronize(ls) != cudaSuccess)
            return -1;
        if (ar.h_state[1] >= 0) {
            done_steps = ar.h_state[1] + 1;
            return 1;
        }
        return 0;
    };
    const bool graph = ar.linear_variant >= 3 && fused && (hd == 32 || hd == 64 || hd == 128) && max_steps > 2;
    if (!graph) {
        for (int s = 0; s < max_steps; ++s) {
            if (int r = position(s, st, nullptr)) return r;
            if ((s + 1) % check_every == 0 || s + 1 == max_steps) {
                const int f = finished(st);
                if (f < 0) return fail(e, B200OCR_E_CUDA, "token loop: stop check failed: %s", cudaGetErrorString(cudaGetLastError()));
                if (f) break;
            }
        }
    } else {
        // the loop runs on the engine's own stream (the caller's may be the legacy default stream, which cannot be
        // captured), forked from and joined to the caller's stream by events
        if (!ar.loop_stream) {
            CU_TRY(e, cudaStreamCreateWithFlags(&ar.loop_stream, cudaStreamNonBlocking));
            CU_TRY(e, cudaEventCreateWithFlags(&ar.loop_fork, cudaEventDisableTiming));
            CU_TRY(e, cudaEventCreateWithFlags(&ar.loop_join, cudaEventDisableTiming));
        }
        cudaStream_t ls = ar.loop_stream;
        CU_TRY(e, cudaEventRecord(ar.loop_fork, st));
        CU_TRY(e, cudaStreamWaitEvent(ls, ar.loop_fork, 0));
        const int32_t* pos_dev = ar.state + 2;
        const int64_t before = e->launches;
        if (int r = position(0, ls, pos_dev)) return r;    // position 0 eagerly: one-time kernel attributes are set here
        const int per_position = static_cast<int>(e->launches - before);
        const ArState::GraphKey key{n, T, max_steps, start_token, tokens, logits, ar.x, ar.linear_variant};
        if (!ar.gexec || !(ar.gkey == key)) {
            if (ar.gexec) { cudaGraphExecDestroy(ar.gexec); ar.gexec = nullptr; }
            cudaGraph_t g = nullptr;
            CU_TRY(e, cudaStreamBeginCapture(ls, cudaStreamCaptureModeThreadLocal));
            const int rc = position(1, ls, pos_dev);
            const cudaError_t ce = cudaStreamEndCapture(ls, &g);
            e->launches -= per_position;                   // captured, not launched
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) return fail(e, B200OCR_E_CUDA, "token loop: graph capture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&ar.gexec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) return fail(e, B200OCR_E_CUDA, "token loop: graph instantiation failed: %s", cudaGetErrorString(ie));
            ar.gkey = key;
        }
        for (int s = 1; s < max_steps; ++s) {
            CU_TRY(e, cudaGraphLaunch(ar.gexec, ls));
            e->launches += per_position;
            if ((s + 1) % check_every == 0 || s + 1 == max_steps) {
                const int f = finished(ls);
                if (f < 0) return fail(e, B200OCR_E_CUDA, "token loop: stop check failed: %s", cudaGetErrorString(cudaGetLastError()));
                if (f) break;
            }
        }
        CU_TRY(e, cudaEventRecord(ar.loop_join, ls));
        CU_TRY(e, cudaStreamWaitEvent(st, ar.loop_join, 0));
    }
#undef AR_LAUNCH
    *steps = done_steps >= 0 ? done_steps : max_steps;
    return B200OCR_OK;
}

int b200ocr_profile(b200ocr_engine_t* e, int32_t on) {
    if (!e) return B200OCR_E_INVALID;
    for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
    e->profiling = on != 0;
    return B200OCR_OK;
}

int b200ocr_profile_read(b200ocr_engine_t* e, int32_t capacity, int32_t* tags, int32_t* layers, float* ms,
                         int32_t* count) {
    if (!e || !count) return B200OCR_E_INVALID;
    CU_TRY(e, cudaDeviceSynchronize());
    const int n = static_cast<int>(e->prof.size());
    *count = n;
    for (int i = 0; i < n && i < capacity; ++i) {
        float t = 0.f;
        CU_TRY(e, cudaEventElapsedTime(&t, e->prof[i].a, e->prof[i].b));
        tags[i] = e->prof[i].tag; layers[i] = e->prof[i].layer; ms[i] = t;
    }
    for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
    return B200OCR_OK;
}

int b200ocr_profile_read_since(b200ocr_engine_t* e, void* reference, int32_t capacity, int32_t* tags, int32_t* layers,
                               float* start_ms, float* end_ms, int32_t* count) {
    if (!e || !count || !reference) return B200OCR_E_INVALID;
    CU_TRY(e, cudaDeviceSynchronize());
    cudaEvent_t ref = static_cast<cudaEvent_t>(reference);
    const int n = static_cast<int>(e->prof.size());
    *count = n;
    for (int i = 0; i < n && i < capacity; ++i) {
        float a = 0.f, b = 0.f;
        CU_TRY(e, cudaEventElapsedTime(&a, ref, e->prof[i].a));
        CU_TRY(e, cudaEventElapsedTime(&b, ref, e->prof[i].b));
        tags[i] = e->prof[i].tag; layers[i] = e->prof[i].layer; start_ms[i] = a; end_ms[i] = b;
    }
    for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
    return B200OCR_OK;
}

int b200ocr_debug_set_flag(b200ocr_engine_t* e, int32_t flag, int32_t value) {
    if (!e) return B200OCR_E_INVALID;
    if (flag == 1) e->use_halo = value != 0;
    else if (flag == 2) e->ar.linear_variant = value < 0 ? 0 : (value > 3 ? 3 : value);
    else if (flag == 3) e->lstm_hplanes = (value != 0 && e->lstm_planes == 2) ? 2 : 1;
    else if (flag == 4) e->crop_staging = value < 0 || value > 3 ? 2 : value;
    else if (flag == 5) e->ref_only_layer = value;
    else if (flag == 6) e->l2_chunk_lines = value < 0 ? 0 : value;
    else if (flag == 7) e->dynamic_tiles = value != 0;
    else if (flag == 8) e->attention_tc = value != 0;
    else if (flag == 9) e->igemm_dbg = value;
    else if (flag == 10) e->lstm_priority = value != 0;
    else if (flag == 11) e->lean_records = value != 0;
    else return fail(e, B200OCR_E_INVALID, "unknown debug flag %d", flag);
    return B200OCR_OK;
}

int b200ocr_debug_use_reference_kernels(b200ocr_engine_t* e, int32_t on) {
    if (!e) return B200OCR_E_INVALID;
    e->use_ref = on != 0;
    return B200OCR_OK;
}

int b200ocr_debug_forward_prefix(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w,
                                 int32_t n_layers, float* out, int64_t capacity, int64_t* written, int32_t* shape4,
                                 void* cuda_stream) {
    if (!e || !crops || !out || n_layers <= 0) return fail(e, B200OCR_E_INVALID, "bad debug arguments");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    Outputs o;
    Shape fs;
    const void* ptr = nullptr;
    bool is_f32 = false;
    if (int s = walk(e, crops, n, h, w, n_layers, o, st, nullptr, &fs, &ptr, &is_f32)) return s;
    const int64_t count = static_cast<int64_t>(fs.n) * fs.h * fs.w * fs.c;
    if (shape4) { shape4[0] = fs.n; shape4[1] = fs.h; shape4[2] = fs.w; shape4[3] = fs.c; }
    if (written) *written = count;
    if (count > capacity || !ptr) return fail(e, B200OCR_E_INVALID, "debug buffer too small (%lld needed)", (long long)count);
    if (is_f32) {
        CU_TRY(e, cudaMemcpyAsync(out, ptr, count * 4, cudaMemcpyDeviceToHost, st));
    } else {
        float* tmp = nullptr;
        CU_TRY(e, cudaMalloc(reinterpret_cast<void**>(&tmp), count * 4));
        CU_TRY(e, launch_h2f(static_cast<const __half*>(ptr), fs.n * fs.h * fs.w, fs.c, e->fmt, tmp, st));
        CU_TRY(e, cudaMemcpyAsync(out, tmp, count * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(e, cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    CU_TRY(e, cudaStreamSynchronize(st));
    return B200OCR_OK;
}

}  // extern "C"
#pragma GCC visibility pop
