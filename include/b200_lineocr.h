/* b200_lineocr.h -- C ABI of the B200-native text-line recognition path (libb200_lineocr.so).
 *
 * The reference (DCGM/pero-ocr v0.7.0) is pure Python and has no FFI: its "plugin surface" is the duck-typed
 * engine object held by PageParser.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  All pointers are plain device or host pointers as stated; no
 * framework types cross this boundary.  The library owns only the packed weights and the workspace it
 * allocates in b200ocr_create / b200ocr_reserve; every input/output buffer is caller-owned.
 * Calls are stream-ordered on the given CUDA stream, perform no hidden synchronisation and no allocation
 * (b200ocr_forward fails with B200OCR_E_WORKSPACE if the reserved workspace is too small).
 * One engine per device, not thread-safe -- same as the reference engine objects.
 */
#ifndef B200_LINEOCR_H
#define B200_LINEOCR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200ocr_engine b200ocr_engine_t;

enum b200ocr_status {
    B200OCR_OK = 0,
    B200OCR_E_INVALID = 1,   /* bad argument / unsupported shape */
    B200OCR_E_CUDA = 2,      /* CUDA runtime or driver error (text in b200ocr_last_error) */
    B200OCR_E_WORKSPACE = 3, /* b200ocr_reserve was not called or is too small for this batch */
    B200OCR_E_NO_DEVICE = 4  /* no sm_100 device visible: there is no CPU fallback */
};

enum b200ocr_layer_kind {
    B200OCR_CONV_FIRST = 1, /* u8 NHWC image -> /255 -> 3x3 conv (pad 1) + bias + act; tcgen05 (K = 27 padded to 32), conv_first.cu */
    B200OCR_CONV = 2,       /* kh x kw conv, fp16 NHWC in, tcgen05 implicit GEMM, fused bias+act+affine+max-pool */
    B200OCR_BILSTM = 3,     /* one bidirectional LSTM layer (PyTorch gate order i,f,g,o) */
    B200OCR_CTC_HEAD = 4,   /* per-frame linear -> logits [N,T,C] (+ fused per-frame argmax / max / logsumexp) */
    B200OCR_UPSAMPLE = 5,   /* nearest-neighbour x`pool_h` upsampling to fp32 NCHW maps (ParseNet tail) */
    B200OCR_LN_PE = 6,      /* LayerNorm over channels + sinusoidal positional encoding (transformer.py:316-332,378-381) */
    B200OCR_TRANSFORMER_LAYER = 7 /* post-LN nn.TransformerEncoderLayer, ReLU FFN (transformer.py:371-373) */
};

enum b200ocr_act { B200OCR_ACT_NONE = 0, B200OCR_ACT_RELU = 1, B200OCR_ACT_LEAKY_RELU = 2 /* slope 0.01 */ };

enum b200ocr_precision {
    B200OCR_PREC_FP16 = 0,  /* fp16 operands, fp32 accumulate (same 10-bit mantissa as cuDNN's default TF32 convs) */
    B200OCR_PREC_FP16X3 = 1, /* hi/lo fp16 split of both operands, 3 MMAs per product: ~fp32-faithful */
    B200OCR_PREC_FP16F8 = 2, /* fp16 hi*hi pass + the two first-order correction terms (lo*hi + hi*lo) as ONE e5m2
                                tensor-core pass at twice the fp16 rate: 2 pass-equivalents, logits within ~1e-4 */
    B200OCR_PREC_FP16F8W = 3 /* FP16F8 with b200ocr_set_layer_correction(layer, B200OCR_CORR_WEIGHT) preset on the 3x3
                                convolutions with >= 256 input channels (the tensor-pipe-bound layers) */
};

/* B200OCR_PREC_FP16F8: which first-order correction terms the e5m2 pass of ONE layer evaluates.  The pass contracts
 * [a_lo | a_hi] . [w_hi | w_lo]; WEIGHT walks the second half only (a_hi . w_lo: the fp16 rounding of the weights,
 * the larger and static error term) for 1.5 instead of 2 pass-equivalents, NONE leaves plain fp16. */
enum b200ocr_correction { B200OCR_CORR_BOTH = 0, B200OCR_CORR_WEIGHT = 1, B200OCR_CORR_NONE = 2 };

/* One layer.  All weight pointers are HOST pointers to fp32 arrays in PyTorch's native layouts; the library
 * packs them (fp16, tap-major, K-major) into device memory it owns.  Unused fields are 0 / NULL. */
typedef struct b200ocr_layer {
    int32_t kind;
    int32_t cin, cout;
    int32_t kh, kw, pad_h, pad_w;
    int32_t act;
    int32_t pool_h, pool_w;  /* fused max-pool after the activation (1,1 = none) */
    const float* weight;     /* conv: [cout][cin][kh][kw]; linear: [cout][cin] */
    const float* bias;       /* [cout] */
    const float* post_scale; /* optional eval-mode BatchNorm folded to y*scale+shift, applied AFTER act (+pool) */
    const float* post_shift;
    /* B200OCR_BILSTM: index 0 = forward, 1 = reverse direction; w_ih [4H][cin], w_hh [4H][H], biases [4H] */
    int32_t hidden;
    const float* w_ih[2];
    const float* w_hh[2];
    const float* b_ih[2];
    const float* b_hh[2];
    /* B200OCR_TRANSFORMER_LAYER / B200OCR_LN_PE */
    int32_t heads, dim_ff;
    const float* in_proj_w;  /* [3D][D] */
    const float* in_proj_b;  /* [3D] */
    const float* out_proj_w; /* [D][D] */
    const float* out_proj_b;
    const float* lin1_w;     /* [dim_ff][D] */
    const float* lin1_b;
    const float* lin2_w;     /* [D][dim_ff] */
    const float* lin2_b;
    const float* norm1_w;    /* LN_PE uses norm1_* as its LayerNorm parameters */
    const float* norm1_b;
    const float* norm2_w;
    const float* norm2_b;
    float act_slope;         /* B200OCR_ACT_LEAKY_RELU: negative slope (nn.LeakyReLU.negative_slope); 0 = the default 0.01 */
} b200ocr_layer_t;

typedef struct b200ocr_net_desc {
    int32_t n_layers;
    const b200ocr_layer_t* layers;
    int32_t precision;   /* enum b200ocr_precision */
    int32_t line_height; /* input crop height (40) */
    int32_t device;      /* CUDA device ordinal */
} b200ocr_net_desc_t;

/* Replaces PytorchEngineLineOCR._load_exported_model (pero_ocr/ocr_engine/pytorch_ocr_engine.py:52-57):
 * builds the device-resident network from host weights.  Returns B200OCR_OK and *out, else a status
 * (message via b200ocr_last_error(NULL)). */
int b200ocr_create(const b200ocr_net_desc_t* desc, b200ocr_engine_t** out);

/* Pre-allocates activations/workspace for batches up to max_lines x (line_height x max_width_px).
 * (The reference relies on torch's caching allocator + torch.cuda.empty_cache(), line_ocr_engine.py:174-175.) */
int b200ocr_reserve(b200ocr_engine_t* e, int32_t max_lines, int32_t max_width_px);

/* Replaces the device part of PytorchEngineLineOCR.run_ocr (pytorch_ocr_engine.py:59-74): `/255`, NHWC->net,
 * self.model(batch) (:64-69), greedy_decode_ctc (:13-34) up to (not including) the id->char join.
 *   crops      device u8 [n][h][w][3]  (the padded batch built by BaseEngineLineOCR.process_lines, :121-123)
 *   logits     device f32 [n][T][C] or NULL (T = w/4, C = classes incl. blank LAST) -- run_ocr's 2nd result (:72)
 *   labels     device i32 [n][T]  collapsed label ids, left-packed, rest -1
 *   lengths    device i32 [n]
 *   confidence device f32 [n] or NULL: min over non-blank runs of the run's max softmax prob over ALL T frames
 *              (PageParser.get_prob, document_ocr/page_parser.py:437-450)
 *   best_path  device i32 [n][T] or NULL: raw per-frame argmax ("bit-exact CTC argmax indices")
 */
int b200ocr_forward(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w, float* logits,
                    int32_t* labels, int32_t* lengths, float* confidence, int32_t* best_path, void* cuda_stream);

/* ParseNet-style conv net: replaces `self.net(image)` in TorchParseNet.get_maps
 * (pero_ocr/layout_engines/torch_parsenet.py:51-53).  image: device u8 [1][h][w][3]; maps: device f32
 * [1][cout][h][w] (NCHW, like the TorchScript blob's first output). */
int b200ocr_forward_maps(b200ocr_engine_t* e, const uint8_t* image, int32_t h, int32_t w, float* maps,
                         void* cuda_stream);

/* Workspace for b200ocr_forward_maps on canvases up to max_h x max_w. */
int b200ocr_reserve_maps(b200ocr_engine_t* e, int32_t max_h, int32_t max_w);

void b200ocr_destroy(b200ocr_engine_t* e);

/* Last error text of this engine (or of the last failed b200ocr_create when e == NULL). */
const char* b200ocr_last_error(const b200ocr_engine_t* e);

/* Number of kernels this library launched on behalf of `e` since creation (bench.py's gpu_launches). */
int64_t b200ocr_launch_count(const b200ocr_engine_t* e);

/* Algorithmic FLOPs (2*MACs) of one forward at (n, w) and the share of the implicit-GEMM conv kernel. */
double b200ocr_forward_flops(const b200ocr_engine_t* e, int32_t n, int32_t w, double* conv_gemm_flops);

/* Per-layer arithmetic of a B200OCR_PREC_FP16F8 engine (the reference has one arithmetic: whatever torch dispatches
 * for `self.model`, pytorch_ocr_engine.py:64-69 -- cuDNN TF32 convolutions by default).  `layer` indexes the
 * descriptor's layer list; layers without a tensor-core contraction are rejected.  Takes effect at the next forward.
 * pero_ocr_b200.engine.LineRecognizer.autotune_precision chooses the modes by measured logit error. */
int b200ocr_set_layer_correction(b200ocr_engine_t* e, int32_t layer, int32_t mode);

/* Replaces the post-activation per-channel shift of a convolution layer created with post_scale / post_shift
 * (`shift`: HOST f32 [cout]).  Used for embedding-conditioned recognisers -- `model(batch, ids_embedding)` with one id
 * for the whole batch (pytorch_ocr_engine.py:64-66), whose gathered vector is such a shift -- when the caller changes
 * `embed_id` between runs (user_scripts/select_embed_id.py:79-80).  Synchronises the device. */
int b200ocr_set_layer_post_shift(b200ocr_engine_t* e, int32_t layer, const float* shift);

/* FLOP-weighted tensor-core pass-equivalents the GEMM layers execute per algorithmic FLOP at (n, w) (1 = fp16,
 * 2 = fp16f8, 3 = fp16x3, in between with per-layer corrections); per_layer (HOST f32 [capacity], may be NULL)
 * receives the figure of each layer (0 for layers without a contraction). */
double b200ocr_executed_passes(const b200ocr_engine_t* e, int32_t n, int32_t w, int32_t capacity, float* per_layer);

/* Two (or more) engines on one GPU, each fed on its own stream (the reference has one module and one stream,
 * pytorch_ocr_engine.py:59-74): after this call every forward of `e` starts its layer walk only when the latest forward
 * of `after` that was enqueued before it has finished its convolutional front end (everything up to the first BiLSTM
 * recurrence).  Linked in a ring, the engines hand the whole GPU to one another for their throughput-bound
 * convolutions while the latency-bound recurrence of the previous batch (32-64 of the 148 SMs) runs beside them;
 * unlinked, the hardware interleaves the two conv chains layer by layer, both reach their recurrences together, and
 * 84+ SMs idle for their duration.  `after` = NULL unlinks.  Both engines must live on the same device; an engine
 * may be destroyed while linked. */
int b200ocr_run_after(b200ocr_engine_t* e, b200ocr_engine_t* after);

/* Replaces greedy_decode_ctc (pytorch_ocr_engine.py:13-34) and decoding.decoders.GreedyDecoder.__call__
 * (pero_ocr/decoding/decoders.py:42-62) on materialised scores.
 *   scores  device f32, layout 0 = [n][t][c] (run_ocr / decoder layout), 1 = [n][c][t] (model output layout)
 *   labels/lengths/confidence/best_path as in b200ocr_forward; frame_max: device f32 [n][t] or NULL (per-frame
 *   max score, GreedyDecoder's `maxes`); frame_lse: device f32 [n][t] or NULL (per-frame logsumexp). */
int b200ocr_ctc_greedy(const float* scores, int32_t n, int32_t t, int32_t c, int32_t layout, int32_t* labels,
                       int32_t* lengths, float* confidence, int32_t* best_path, float* frame_max, float* frame_lse,
                       void* cuda_stream);

/* Replaces force_align / align_text (pero_ocr/core/force_alignment.py:13-35, 152-165): Viterbi alignment of a
 * transcription to the frames of a CTC output, for a batch of lines.
 *   neg_logprobs  device f32 (is_f64 = 0) or f64 (1) [n][t][c] NEGATIVE log-probabilities; n_frames device i32 [n] or
 *                 NULL = frames actually used per line (<= t)
 *   labels        device i32 [n][l_max] symbol ids (blank excluded), lengths device i32 [n]; blank = blank class id
 *   out_symbols   device i32 [n][t] or NULL: the most probable path as symbols incl. blanks (force_align's default)
 *   out_positions device i32 [n][t] or NULL: the path as character indices, -1 on blank frames
 *                 (return_seq_positions=True)
 *   char_positions device i32 [n][l_max] or NULL: align_text's result, one frame per character (needs out_positions)
 *   status        device i32 [n]: 0 ok, 1 no finite-cost alignment exists (reference: ValueError, :146-147),
 *                 2 empty transcription / blank or out-of-range symbol in it (reference: ValueError, :41-43, :64-68)
 * Unused tails of the outputs are -1.  Costs accumulate in float64 as in the reference. */
int b200ocr_force_align(const void* neg_logprobs, int32_t is_f64, int32_t n, int32_t t, int32_t c,
                        const int32_t* n_frames, const int32_t* labels, int32_t l_max, const int32_t* lengths,
                        int32_t blank, int32_t* out_symbols, int32_t* out_positions, int32_t* char_positions,
                        int32_t* status, void* cuda_stream);

/* Replaces the per-character loop of get_line_confidence (pero_ocr/core/confidence_estimation.py:73-104) for a batch
 * of lines: confidence of character i = max(0, p[a_i, label_i] - the largest probability of any OTHER non-blank
 * class (neighbouring labels excluded too) between the borders half-way to the neighbouring characters' frames).
 *   log_probs      device f32 [n][t][c] log-probabilities (TextLine.get_full_logprobs, core/layout.py:70-72), blank LAST
 *   n_frames       device i32 [n] or NULL; labels device i32 [n][l_max]; lengths device i32 [n]
 *   char_positions device i32 [n][l_max]  align_text's frames (b200ocr_force_align); a negative entry gives 0
 *   confidences    device f32 [n][l_max]  (entries past a line's length are 0) */
int b200ocr_char_confidence(const float* log_probs, int32_t n, int32_t t, int32_t c, const int32_t* n_frames,
                            const int32_t* labels, int32_t l_max, const int32_t* lengths,
                            const int32_t* char_positions, float* confidences, void* cuda_stream);

/* Replaces EngineLineCropper.fast_remap (pero_ocr/core/crop_engine.py:146-163: cv2.remap, INTER_LINEAR,
 * BORDER_CONSTANT 0, 8-bit fixed-point bilinear) for all lines of a page in one launch, writing straight into the
 * zero-padded batch that BaseEngineLineOCR.process_lines builds on the host (line_ocr_engine.py:121-123).
 *   image     device u8 [img_h][img_w][3]   the page (img_h, img_w <= 32767, OpenCV's own limit for this path)
 *   coords    device f32: per line a [line_h][w_i][2] map of source (x, y), i.e. get_crop_inputs' result
 *             (crop_engine.py:54-99), concatenated; coord_off device i64 [n] = offset of line i in floats
 *   widths    device i32 [n]  w_i
 *   out       device u8 [n][line_h][out_w][3]: line i occupies columns [pad, pad + w_i) (cut at out_w), the rest is 0 */
int b200ocr_remap_lines(const uint8_t* image, int32_t img_h, int32_t img_w, const float* coords,
                        const int64_t* coord_off, const int32_t* widths, int32_t n, int32_t line_h, uint8_t* out,
                        int32_t out_w, int32_t pad, void* cuda_stream);

/* Plumbing for results that go straight into caller-owned host memory (the CSC parts of TextLine.logits, which the
 * reference builds on the host, line_ocr_engine.py:168-172): one stream-ordered device -> host copy.  `host_dst` should
 * be page-locked (cudaHostRegister / cudaHostAlloc) for the copy to be asynchronous; the caller synchronises the
 * stream before reading. */
int b200ocr_memcpy_d2h_async(void* host_dst, const void* device_src, int64_t bytes, void* cuda_stream);

/* Replaces the zero padding + stacking of BaseEngineLineOCR.process_lines (pero_ocr/ocr_engine/line_ocr_engine.py:
 * 121-127: `batch_data[i, :, 32:32 + w_i] = line_i`, lines beyond the batch width cut) on the device, so that the host
 * stages each crop with ONE contiguous copy and no padding bytes cross PCIe.
 *   packed    device u8: the crops back to back, crop i = [line_h][w_i][3] at byte offset line_off[i] (device i64 [n])
 *   widths    device i32 [n]  w_i
 *   out       device u8 [n][line_h][out_w][3] (out_w a multiple of 4): columns [pad, pad + w_i) hold crop i, the rest 0 */
int b200ocr_pad_lines(const uint8_t* packed, const int64_t* line_off, const int32_t* widths, int32_t n, int32_t line_h,
                      uint8_t* out, int32_t out_w, int32_t pad, void* cuda_stream);

/* One text line for b200ocr_remap_poly_lines: what the host keeps of EngineLineCropper.get_crop_inputs
 * (crop_engine.py:54-73) for the `poly` > 0 configurations -- the fitted baseline polynomial in the rotated frame and
 * the arc-length resampling constants; the device evaluates the rest (:74-99). */
typedef struct b200ocr_poly_line {
    double coef[4];   /* np.polyfit coefficients, highest power first; `ncoef` of them are used */
    double x_first;   /* xs[0]  = left end of the rotated baseline */
    double x_last;    /* xs[-1] = last sample of np.arange(left, right) */
    double total;     /* arc length of the sampled baseline (mapping_x_to_line_pos[-1]) */
    double step;      /* total / (n_out - 1): np.linspace's step (unused when n_out == 1) */
    double rot[4];    /* rotation back to page coordinates, row-major r00 r01 r10 r11 */
    int32_t ncoef;    /* 0 = geometry failed: the reference's all-zero crop (crop_engine.py:16-22) */
    int32_t n_out;    /* crop width in pixels */
} b200ocr_poly_line_t;

/* b200ocr_remap_lines with the sampling maps computed on the device from per-line parameters instead of uploaded:
 *   lines    device b200ocr_poly_line_t [n]
 *   offsets  device f64 [n][line_h]: np.linspace(-heights[0]*scale, heights[1]*scale, line_h) (crop_engine.py:91)
 * Output bytes are those of b200ocr_remap_lines on the reference's own float32 maps. */
int b200ocr_remap_poly_lines(const uint8_t* image, int32_t img_h, int32_t img_w, const b200ocr_poly_line_t* lines,
                             const double* offsets, int32_t n, int32_t line_h, uint8_t* out, int32_t out_w,
                             int32_t pad, void* cuda_stream);

/* Replaces the per-line logit sparsification of BaseEngineLineOCR.process_lines
 * (pero_ocr/ocr_engine/line_ocr_engine.py:168-172: softmax, zero raw logits with p < 1e-4, scipy CSC) together with
 * the optional tight crop of the frame range (:152-156), on the device, so that only the surviving entries cross PCIe.
 *   logits   device f32 [n][t][c] (b200ocr_forward's `logits`)
 *   t_lo/hi  device i32 [n] or NULL: frame range [lo, hi) kept per line (NULL = [0, t))
 *   indptr   device i32 [n][c+1]   CSC column pointers of each line's [hi-lo][c] matrix, relative to base[line]
 *   nnz      device i32 [n]        entries per line
 *   base     device i64 [n+1]      offset of each line's first entry in indices/data; base[n] = total
 *   indices  device i32 [capacity] frame index minus lo, ascending inside a column
 *   data     device f32 [capacity] the raw logit
 * Entries beyond `capacity` are counted but not written (n*t*c always suffices).  A raw logit equal to 0.0 is
 * never stored, as in the reference (scipy drops zeros; core/layout.py:65-68 maps them back to -80). */
int b200ocr_sparsify_logits(const float* logits, int32_t n, int32_t t, int32_t c, const int32_t* t_lo,
                            const int32_t* t_hi, int32_t* indptr, int32_t* nnz, int64_t* base, int32_t* indices,
                            float* data, int64_t capacity, void* cuda_stream);

/* Replaces CTCPrefixLogRawNumpyDecoder.__call__ without LM (pero_ocr/decoding/decoders.py:220-299).
 *   logprobs  device f64 [n][t][c] normalised log-probabilities, blank last
 *   out_labels device i32 [n][k][t], out_lengths i32 [n][k] (-1 = unused beam slot), out_scores f64 [n][k]
 *   (logaddexp(Pb, Pnb) per surviving prefix), status i32 [n]: 0 ok, 1 = not normalised (reference raises
 *   ValueError, decoders.py:223-224). */
int b200ocr_ctc_prefix_beam(const double* logprobs, int32_t n, int32_t t, int32_t c, int32_t k, int32_t* out_labels,
                            int32_t* out_lengths, double* out_scores, int32_t* status, void* cuda_stream);

/* b200ocr_ctc_prefix_beam on a frame range per line: t_lo / t_hi device i32 [n] (both or neither), the slice
 * `logprobs[logit_coords[0]:logit_coords[1]]` that PageDecoder.decode_line takes before calling the decoder
 * (pero_ocr/document_ocr/page_parser.py:133-135).  Lets a page's lines of different widths share one launch. */
int b200ocr_ctc_prefix_beam_ranges(const double* logprobs, int32_t n, int32_t t, int32_t c, int32_t k,
                                   const int32_t* t_lo, const int32_t* t_hi, int32_t* out_labels, int32_t* out_lengths,
                                   double* out_scores, int32_t* status, void* cuda_stream);

/* Replaces, on the device, the chain that turns the recogniser's raw logits into the decoders' input: sparsification
 * (line_ocr_engine.py:168-172) -> TextLine.get_dense_logits (zeros -> -80) -> get_full_logprobs (float32 log-softmax)
 * (pero_ocr/core/layout.py:65-72).
 *   logits    device f32 [n][t][c] (b200ocr_forward's `logits`)
 *   logprobs  device f64 [n][t][c] (the decoders' working type) */
int b200ocr_full_logprobs(const float* logits, int32_t n, int32_t t, int32_t c, double* logprobs, void* cuda_stream);

/* ---- autoregressive Transformer decoder (the reference's `pytorch_ocr-transformer` engine) ------------------- */
/* One DecoderLayer (pero_ocr/ocr_engine/transformer.py:388-462): cached self-attention, encoder-decoder attention,
 * ReLU feed-forward, post-LN.  HOST fp32 pointers in PyTorch's layouts (state-dict entries
 * trans_decoder.layers.i.{self_attn,multihead_attn}.{in_proj_weight,in_proj_bias,out_proj.*}, linear1/2, norm1-3). */
typedef struct b200ocr_ar_layer {
    const float *self_in_w, *self_in_b;     /* [3D][D], [3D] */
    const float *self_out_w, *self_out_b;   /* [D][D], [D] */
    const float *cross_in_w, *cross_in_b;   /* [3D][D], [3D]: rows 0..D-1 project the query, D..3D-1 the memory K | V */
    const float *cross_out_w, *cross_out_b;
    const float *lin1_w, *lin1_b;           /* [dim_ff][D] */
    const float *lin2_w, *lin2_b;           /* [D][dim_ff] */
    const float *norm1_w, *norm1_b, *norm2_w, *norm2_b, *norm3_w, *norm3_b;
} b200ocr_ar_layer_t;

typedef struct b200ocr_ar_desc {
    int32_t n_layers, heads, dim_ff;
    int32_t classes;                 /* TransformerOCR.num_classes = output symbols + 2 (transformer.py:39) */
    const b200ocr_ar_layer_t* layers;
    const float* embed;              /* dec_embeder.weight [classes][D] (transformer.py:512) */
    const float* out_w;              /* dec_out_proj.weight [classes][D] (:513) */
    const float* out_b;              /* [classes] */
} b200ocr_ar_desc_t;

/* Attaches the decoder half of TransformerOCR (transformer.py:489-546) to an engine whose layer list is the encoder
 * half (conv frontend, B200OCR_LN_PE, B200OCR_TRANSFORMER_LAYER x L -- no CTC head): TransformerOCR.encode (:548-555)
 * is then b200ocr's layer walk and its output stays in the engine's workspace as the decoder's memory.
 * Replaces the decoder part of TransformerEngineLineOCR.__init__ (transformer_ocr_engine.py:21-31). */
int b200ocr_ar_attach(b200ocr_engine_t* e, const b200ocr_ar_desc_t* desc);

/* Workspace for b200ocr_ar_transcribe: batches up to max_lines x max_width_px, up to max_steps decoded positions
 * (also calls b200ocr_reserve for the encoder). */
int b200ocr_ar_reserve(b200ocr_engine_t* e, int32_t max_lines, int32_t max_width_px, int32_t max_steps);

/* Replaces TransformerEngineLineOCR.transcribe_batch(inputs, is_cached=True) (transformer_ocr_engine.py:49-89) up to
 * (not including) postprocess_decoded: `/255`, encode, then greedy decoding one position per step with cached
 * self-attention keys / values (Decoder.infer / DecoderLayer.infer / CustomMultiheadAttention.cached_forward,
 * transformer.py:183-305, 418-486) -- the whole token loop runs inside this call.
 *   crops        device u8 [n][h][w][3] (run_ocr's batch, already centre-padded to >= 1088 px by the caller, :36-40)
 *   start_token  the sentence-boundary symbol (:17): first input token of every line and the stop symbol
 *   max_steps    bound on decoded positions; the reference stops after w/4 + 1 (:79-82)
 *   check_every  the alive mask is read back (one 8-byte copy + a stream synchronisation) every this many steps; the
 *                reference synchronises at every step (:75).  Positions decoded past the stop are never reported.
 *   tokens       device i32 [max_steps][n]: greedy choice of every line at every executed step (row s = `samples`
 *                of step s; the reference's partial_transcripts[1:] are rows 0 .. *steps - 2)
 *   logits       device f32 [n][max_steps][classes] or NULL: the reference's second result, rows 0 .. *steps - 1 valid
 *   steps        HOST i32: number of steps the reference loop executes on this batch (its `len(logits)`)
 * Unlike b200ocr_forward this call synchronises `cuda_stream` (it returns a host value), as the reference loop does. */
int b200ocr_ar_transcribe(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w,
                          int32_t start_token, int32_t max_steps, int32_t check_every, int32_t* tokens, float* logits,
                          int32_t* steps, void* cuda_stream);

/* Per-launch device timing for bench.py's roofline leg: while on, every kernel launched by b200ocr_forward is
 * bracketed by CUDA events on its stream.  b200ocr_profile_read synchronises the device, returns one record per
 * launch (tag 0 = first conv, 1 = tcgen05 implicit GEMM, 2 = tcgen05 LSTM recurrence, 3 = other; index of the layer;
 * milliseconds) and clears the list.  Never on inside a timed throughput region. */
int b200ocr_profile(b200ocr_engine_t* e, int32_t on);
int b200ocr_profile_read(b200ocr_engine_t* e, int32_t capacity, int32_t* tags, int32_t* layers, float* ms,
                         int32_t* count);
/* Like b200ocr_profile_read, for timelines across engines and streams: `reference` is a recorded cudaEvent_t created
 * with timing enabled; start_ms / end_ms (HOST f32 [capacity]) receive each launch's begin and end relative to it. */
int b200ocr_profile_read_since(b200ocr_engine_t* e, void* reference, int32_t capacity, int32_t* tags, int32_t* layers,
                               float* start_ms, float* end_ms, int32_t* count);

/* ---- debug / test hooks (not used by the product path) ------------------------------------------------- */
/* Route every implicit-GEMM layer through a naive one-thread-per-output CUDA-core kernel (same packed fp16
 * operands, fp32 accumulate) so the tcgen05 path can be checked on a GPU box where the reference is absent. */
int b200ocr_debug_use_reference_kernels(b200ocr_engine_t* e, int32_t on);
/* A/B switches for kernel variants (tests, profiling).  flag 1: halo-reuse 3x3 kernel for cin <= 128 (default on);
 * flag 2: token-loop kernels of b200ocr_ar_transcribe: 0 = one K walker per projection tile, 1 = split-K projections,
 * 2 = 1 + q | k | v in one launch, K split over CTAs summed in the LayerNorm, CTA-per-(line, head) step attention,
 * 3 (default) = 2 with the decoded position in device memory and the launches of a position replayed as a CUDA graph;
 * flag 3: the BiLSTM recurrence also multiplies the fp16 rounding residual of h_t (three passes and twice the SM-to-SM
 * exchange per step; default on only in B200OCR_PREC_FP16X3);
 * flag 4: the first convolution: 3 = TMA box load of the uint8 crop patch + the tcgen05 kernel (default for 32 / 64
 * output channels), 2 = TMA box load + the mma.sync kernel, 1 = 16-byte cp.async, 0 = plain byte loads (the last two
 * also on mma.sync; 0 is the fallback when the batch width is not a multiple of 16);
 * flag 5: index of ONE layer whose tensor-core contraction runs on the CUDA-core cross-check kernel while every other
 * layer stays on its product kernel (-1 = none): isolates a layer on identical inputs;
 * flag 6: lines per chunk of the first-conv + second-conv pair (the first conv's records then live and die in L2
 * instead of making a 3.5 GB round trip through HBM); 0 = whole batch at once (default: on B200 the ~130 extra launches
 * of 4-line chunks cost more than the HBM round trip, profiles/r02e_l2_chunking.md);
 * flag 7: the persistent tensor-core kernels draw their tiles from a global counter (default on) instead of a fixed
 * 1/grid share per CTA -- what lets two engines on two streams share the GPU without serialising;
 * flag 8: tcgen05 self-attention of the Transformer variant (64-wide heads, lines of up to 384 frames; default on),
 * 0 = the fp32 CUDA-core attention kernels for every shape;
 * flag 9: bring-up bits of the tensor-core GEMM kernels: 1 = epilogues do not store, 2 = no MMA issued (both give
 * wrong results on purpose: timing only), 4 = 16-byte epilogue stores instead of the 256-bit st.global.v8.b32, 8 = a
 * fourth weight stage in the BN = 128 halo kernel;
 * flag 10: the BiLSTM recurrence is launched on the engine's high-priority side stream (fork / join by events).
 * Default off: with b200ocr_run_after the recurrence already runs beside the other engine's convolutions, and
 * the priority changed nothing measurable (profiles/r02x_replica_link_ab.md);
 * flag 11: a layer whose consumer multiplies with weight-side correction only (b200ocr_set_layer_correction) does
 * not write the lo' plane of its activation records (default off: measured, no gain -- profiles/r02y_lean_records_ab.json). */
int b200ocr_debug_set_flag(b200ocr_engine_t* e, int32_t flag, int32_t value);
/* Runs only the first `n_layers` layers of the recogniser on `crops` and copies the fp32-expanded (hi + lo)
 * output activation of the last one to HOST memory `out` (NHWC); shape4 receives {n, h, w, c}.  Synchronises. */
int b200ocr_debug_forward_prefix(b200ocr_engine_t* e, const uint8_t* crops, int32_t n, int32_t h, int32_t w,
                                 int32_t n_layers, float* out, int64_t capacity, int64_t* written, int32_t* shape4,
                                 void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_LINEOCR_H */
