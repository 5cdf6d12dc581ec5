"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by oracle/make_golden.py, tests/ and bench.py.

Everything is drawn from numpy's PCG64 (``default_rng``), which is stable across machines and numpy versions,
so the GPU box regenerates bit-identical inputs and weights without shipping them.
"""
import numpy as np
from scipy.special import log_softmax

from pero_ocr_b200.synthetic import bench_crops, config1_logits, json_characters, peaky_logprobs  # noqa: F401

BLANK = '<BLANK>'

# Recogniser cases hosted by the reference engine in make_golden.py.  `engine_batch_size` is the constructor
# argument of PytorchEngineLineOCR (pixel budget = 480 * batch_size, line_ocr_engine.py:55): 1 forces several
# width-sorted batches of different widths, and an over-budget line that is cropped (line_ocr_engine.py:125-127).
ENGINE_CASES = {
    'lstm': dict(classes=120, seed=0, out_gain=6.0, net_kw={}, engine_batch_size=1,
                 widths=[256, 200, 131, 64, 33, 500, 224, 97]),
    'transformer': dict(classes=120, seed=3, out_gain=2.5, net_kw={'layers': 2}, engine_batch_size=1,
                        widths=[256, 180, 77, 40, 224, 500]),
}
# config-2 width under the reference engine: 8 lines of up to 1280 px (T = 336), several reference batches (budget 3840 px)
ENGINE_CASES['lstm_wide'] = dict(net='lstm', classes=120, seed=0, out_gain=6.0, net_kw={}, engine_batch_size=8,
                                 widths=[1280, 1280, 1279, 1100, 1280, 801, 1280, 640], store_logits=4)
# the class-count convention of real pero checkpoints: the net emits len(JSON characters) + 1 classes -- U+200B, appended
# by the engine (pytorch_ocr_engine.py:42), shares the blank's slot, and decoder_factory's letters are the JSON
# characters + '<BLANK>' (decoding_itf.py:49-50)
ENGINE_CASES['lstm_c119'] = dict(net='lstm', classes=120, json_chars=119, seed=7, out_gain=6.0, net_kw={},
                                 engine_batch_size=2, widths=[300, 212, 97, 160])
# a second recogniser family: other module names and nesting, LeakyReLU slopes 0.1-0.3, one BiLSTM layer, Conv1d head
ENGINE_CASES['lstm_alt'] = dict(net='lstm_alt', classes=120, seed=11, out_gain=6.0, net_kw={}, engine_batch_size=2,
                                widths=[320, 300, 211, 96, 450, 40])
# embedding-conditioned recogniser hosted by the reference as model(x, ids) (pytorch_ocr_engine.py:64-66)
ENGINE_CASES['lstm_embed'] = dict(net='lstm_embed', classes=120, seed=13, out_gain=6.0, net_kw={'num_embeddings': 6},
                                  engine_batch_size=2, widths=[280, 264, 150, 72], embed_id=2)
# BiLSTM hidden size other than 256 (the tcgen05 recurrence kernel's size): the generic recurrence kernel
ENGINE_CASES['lstm_h128'] = dict(net='lstm', classes=120, seed=17, out_gain=6.0, net_kw={'hidden': 128},
                                 engine_batch_size=2, widths=[260, 200, 131, 64])
PARSENET_CASE = dict(seed=5, downsample=2, height=250, width=330)
# BASELINE config 4 size: a 3000 x 4000 page at DOWNSAMPLE = 4 -> canvas 768 x 1024, and the adaptive second pass
# (torch_parsenet.py:60-93); the stand-in's head is biased so that the first pass "detects" 20 px text (>15) and the
# second pass runs at downsample 4 * 20 / 12
PARSENET_PAGE_CASE = dict(seed=6, downsample=4, height=3000, width=4000, head_bias=[20.0, 5.0, 1.0, 0.0, 0.0],
                          head_gain=0.02, stride=6)
CONFIG1_BEAM_LINES = 6


def line_crop(rng, width, height=40):
    """One synthetic "gray" crop [H,W,3] u8: equal channels."""
    g = rng.integers(0, 256, (height, width), dtype=np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2)


def engine_lines(kind):
    spec = ENGINE_CASES[kind]
    rng = np.random.default_rng(100 + spec['seed'] + (1000 if kind not in ('lstm', 'transformer') else 0))
    lines = [line_crop(rng, w) for w in spec['widths']]
    # one genuinely coloured line: the net sees 3 distinct channels
    lines[1] = rng.integers(0, 256, lines[1].shape, dtype=np.uint8)
    return lines


def peaky_cases():
    rng = np.random.default_rng(77)
    small = [chr(ord('a') + i) for i in range(5)] + [BLANK]
    big = [chr(0x100 + i) for i in range(119)] + [BLANK]
    return {
        'peaky_small': (peaky_logprobs(rng, 6, 40, 6, sharp=5.0), small),
        'peaky_big': (peaky_logprobs(rng, 4, 96, 120, sharp=12.0), big),
    }


def greedy_edge_cases():
    """[N,C,T] f32 score tensors for greedy_decode_ctc's tie / first-frame / NaN rules."""
    chars = ['a', 'b', 'c', '']
    C = 4

    def frames(rows):
        a = np.array(rows, dtype=np.float32)         # [T,C]
        return a.T[None].copy()                      # [1,C,T]
    return {
        'tie_lowest_index': (frames([[1, 1, 0, 0]]), chars),
        'tie_char_beats_blank': (frames([[0, 2, 0, 2]]), chars),
        'first_frame_kept': (frames([[5, 0, 0, 0], [5, 0, 0, 0], [0, 0, 0, 5]]), chars),
        'a_blank_a': (frames([[5, 0, 0, 0], [0, 0, 0, 5], [5, 0, 0, 0]]), chars),
        'all_blank': (frames([[0, 0, 0, 5]] * 4), chars),
        'nan_wins': (frames([[0, np.nan, 0, 5], [0, 0, 3, 0]]), chars),
        'repeat_then_switch': (frames([[0, 4, 0, 0], [0, 4, 0, 0], [0, 0, 4, 0], [0, 0, 4, 0], [0, 4, 0, 0]]), chars),
        'batch': (np.concatenate([frames([[5, 0, 0, 0], [0, 0, 0, 5], [0, 5, 0, 0]]),
                                  frames([[0, 0, 0, 5], [0, 0, 5, 0], [0, 0, 5, 0]])]), chars),
    }


def confidence_logits():
    """Raw logit matrices [T,C] f32 for the sparsify -> dense -> confidence chain."""
    rng = np.random.default_rng(55)
    out = []
    for t, c, s in ((30, 12, 6.0), (64, 120, 10.0), (17, 120, 3.0), (1, 5, 4.0)):
        lp = peaky_logprobs(rng, 1, t, c, sharp=s)[0]
        out.append((lp + rng.uniform(-3, 3, (t, 1))).astype(np.float32))   # un-normalised, like raw net output
    return out


def parsenet_image(spec=None):
    spec = spec or PARSENET_CASE
    rng = np.random.default_rng(500 + spec['seed'])
    return rng.integers(0, 256, (spec['height'], spec['width'], 3), dtype=np.uint8)


def parsenet_page_net():
    """The ParseNet stand-in of PARSENET_PAGE_CASE: seeded, with the head scaled down and biased (see the case)."""
    import torch
    from pero_ocr_b200.synthetic import make_net
    spec = PARSENET_PAGE_CASE
    net = make_net('parsenet', seed=spec['seed'])
    with torch.no_grad():
        net.head.weight.mul_(spec['head_gain'])
        net.head.bias.copy_(torch.tensor(spec['head_bias']))
    return net


# autoregressive Transformer decoding: seeded encoder (pero_ocr_b200/synthetic.py) + seeded decoder
# (oracle/ar_oracle.py); classes = symbols + sentence boundary + ignore symbol (transformer_ocr_engine.py:16-19)
AR_CASE = dict(encoder_seed=3, decoder_seed=5, decoder_layers=2, classes=32, lines=3, width=1088,
               line_widths=[1088, 700, 420], input_seed=77, golden='ar_decoder.npz')
# a realistic alphabet (class count not a multiple of the kernels' 32-wide tiles) and a deeper decoder
AR_CASE_WIDE = dict(encoder_seed=4, decoder_seed=9, decoder_layers=3, classes=122, lines=2, width=1152,
                    line_widths=[1152, 777], input_seed=78, golden='ar_decoder_wide.npz')
AR_CASES = {'small': AR_CASE, 'wide': AR_CASE_WIDE}


def ar_inputs(spec=AR_CASE):
    """uint8 [N, 3, 40, W]: what TransformerEngineLineOCR.run_ocr hands to transcribe_batch (after its NHWC -> NCHW
    transpose and the centre padding to 1088 px, transformer_ocr_engine.py:33-40)."""
    rng = np.random.default_rng(spec['input_seed'])
    n, w = spec['lines'], spec['width']
    x = np.zeros((n, 3, 40, w), dtype=np.uint8)
    for i, wi in enumerate(spec['line_widths'][:n]):
        s = (w - wi) // 2
        x[i, :, :, s:s + wi] = rng.integers(0, 256, (1, 40, wi), dtype=np.uint8)
    return x

def ar_net_config(spec=AR_CASE):
    """transformer.build_net's config keys (transformer.py:12-20)."""
    return {'dim_model': 512, 'dim_ff': 2048, 'heads': 8, 'encoder_layers': 2,
            'decoder_layers': spec['decoder_layers'], 'conv_subsampling': [8, 4]}


AR_NET_CONFIG = ar_net_config()


def ar_state_dict(spec=AR_CASE):
    """The TransformerOCR checkpoint (reference key names) of an AR case: seeded encoder + seeded decoder."""
    from oracle.ar_oracle import ar_decoder_state
    from pero_ocr_b200.synthetic import make_net, transformer_ocr_state
    net = make_net('transformer', 120, seed=spec['encoder_seed'], layers=2)
    dec = ar_decoder_state(seed=spec['decoder_seed'], layers=spec['decoder_layers'], classes=spec['classes'])
    return net, dec, transformer_ocr_state(net, dec)


def write_ar_engine_json(tmpdir, max_line_width=None, spec=AR_CASE):
    import json
    import os
    path = os.path.join(str(tmpdir), 'ar_engine.json')
    cfg = {'line_px_height': 40, 'line_vertical_scale': 1.0, 'checkpoint': 'ar.pt',
           'characters': json_characters(spec['classes'] - 2), 'net_name': ar_net_config(spec)}
    if max_line_width is not None:
        cfg['max_line_width'] = max_line_width
    with open(path, 'w', encoding='utf8') as f:
        json.dump(cfg, f)
    return path


# host logic of process_lines for model_type "transformer" (line_ocr_engine.py:57-211): a deterministic stand-in for
# run_ocr whose transcription is a function of the pixel columns, so overlapping parts of a split line agree on
# their overlap the way a real recogniser's do.
AR_HOST_CASE = dict(max_line_width=256, batch_size=2, widths=[700, 256, 90, 257, 1000, 33], block=16, classes=12)


def ar_host_lines():
    rng = np.random.default_rng(91)
    out = []
    for w in AR_HOST_CASE['widths']:
        cols = rng.integers(1, 256, (1, w, 1), dtype=np.uint8)
        out.append(np.repeat(np.repeat(cols, 40, axis=0), 3, axis=2))
    return out


def ar_host_fake_run_ocr(characters):
    """-> run_ocr(batch_data) for the stand-in: one character per 16-px block that holds ink, chosen by the block's
    pixel sum; logits = seeded noise [len + 3, classes]."""
    import zlib
    block, classes = AR_HOST_CASE['block'], AR_HOST_CASE['classes']

    def run_ocr(batch_data, no_logits=False):
        texts, logits = [], []
        longest = 0
        for line in batch_data:
            row = line[0, :, 0].astype(np.int64)
            chars = []
            for b in range(0, row.shape[0] - block + 1, block):
                s = int(row[b:b + block].sum())
                if s:
                    chars.append(characters[s % (classes - 2)])
            texts.append(''.join(chars))
            longest = max(longest, len(chars))
        for t in texts:
            rng = np.random.default_rng(zlib.crc32(t.encode('utf8')))
            logits.append((rng.standard_normal((longest + 3, classes)) * 4).astype(np.float32))
        return texts, (None if no_logits else np.stack(logits))
    return run_ocr
