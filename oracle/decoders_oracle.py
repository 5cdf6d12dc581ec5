"""TEST INFRASTRUCTURE ONLY -- numpy restatement of pero_ocr.decoding (no language model).

Restates
  * GreedyDecoder.__call__                     pero_ocr/decoding/decoders.py:42-62
  * CTCPrefixLogRawNumpyDecoder.__call__       pero_ocr/decoding/decoders.py:220-299 (lm=None branch), with
    compute_Pnb :193-201, compute_Pb :207-208, get_reduced_Pc / get_reduced_last_chars :210-218,
    adjust_for_prefix_joining :138-155, find_new_prefixes :116-131, select_relevant_logits :166-167
  * multisort.top_k                            pero_ocr/decoding/multisort.py:4-15
  * BagOfHypotheses sort / best_hyp            pero_ocr/decoding/bag_of_hypotheses.py:19-20, 64-65
  * constructor validation                     decoders.py:12-33, 158-163

Parity is pinned by tests/golden/decoders.npz (outputs of the unmodified reference, oracle/make_golden.py)
and by the reference's own known-answer tests restated in tests/test_oracle_decoders.py
(test/test_decoding/test_decoders.py:23-166, 448-462).

Known freedom in the reference: multisort.top_k uses np.argpartition, whose choice among exactly tied
candidates at the beam boundary is unspecified.  This restatement keeps the candidates with the largest score
and breaks exact ties towards the lower flattened index; hypothesis *sets* can therefore differ from the
reference only when two candidates tie exactly at rank k.
"""
import numpy as np

BLANK_SYMBOL = '<BLANK>'
NEG_INF = -np.inf
RELEVANCE_THRESHOLD = -10.0            # decoders.py:166-167
MAX_UNNORMALIZATION = 1e-5             # decoders.py:49, 220


def validate_letters(letters):
    seen, dup = set(), []
    for x in letters:
        if x in seen:
            dup.append(x)
        seen.add(x)
    if dup:
        raise ValueError(f'Letters contain these duplicit elements: {dup}')
    where = letters.index(BLANK_SYMBOL)          # ValueError when missing, as in the reference
    if where != len(letters) - 1:
        raise ValueError(f'Expected {BLANK_SYMBOL} as the last of letters, it\'s instead at position {where}')


def validate_beam(k):
    if not isinstance(k, int):
        raise TypeError(f"Beam size 'k' has to be int, got {type(k)} instead (value: {k}).")
    if k < 1:
        raise ValueError(f"Beam size 'k' has to be positive, got {k} instead.")


def normalisation_error(logprobs):
    return float(np.max(np.abs(np.exp(logprobs).sum(axis=1) - 1)))


def _lse(v):
    m = np.max(v)
    if not np.isfinite(m):
        return m
    return m + np.log(np.sum(np.exp(v - m)))


def greedy(logprobs, letters, separator=''):
    """-> (transcript, score).  Score is logsumexp of the per-frame maxima (sic, decoders.py:60)."""
    if normalisation_error(logprobs) > MAX_UNNORMALIZATION:
        raise ValueError('Expected properly normalized logits')
    blank = len(letters) - 1
    best = logprobs.argmax(axis=1)
    maxes = logprobs.max(axis=1)
    keep = np.ones(len(best), dtype=bool)
    keep[1:] = best[1:] != best[:-1]
    ids = [int(b) for b in best[keep] if b != blank]
    return separator.join(letters[i] for i in ids), float(_lse(maxes.astype(np.float64))), ids


def prefix_beam(logprobs, k, blank=None):
    """CTC prefix beam search in the log domain.  -> list of (label id list, score) sorted by score desc.

    State per beam entry: prefix (list of ids), Pb (ends in blank), Pnb (ends in non-blank), last char.
    """
    T, C = logprobs.shape
    if blank is None:
        blank = C - 1
    if normalisation_error(logprobs) > MAX_UNNORMALIZATION:
        raise ValueError('Expected properly normalized logits')
    lp = np.asarray(logprobs, dtype=np.float64)
    prefixes = [[]]
    Pb = np.array([0.0])
    Pnb = np.array([NEG_INF])
    last = np.array([0], dtype=np.int64)                         # decoders.py:246 (0, not "none")
    for t in range(T):
        row = lp[t]
        p_blank = row[blank]
        sel = np.nonzero(row[:blank] > RELEVANCE_THRESHOLD)[0]
        if len(sel) == 0:                                        # :252-255
            Pb = np.logaddexp(Pb, Pnb) + p_blank
            Pnb = np.full_like(Pnb, NEG_INF)
            continue
        S = len(sel)
        pc = np.concatenate([row[sel], [NEG_INF]])               # slot S: "impossible" character
        pos = {int(c): i for i, c in enumerate(sel)}
        rlast = np.array([pos.get(int(c), S) for c in last])
        nb = len(prefixes)
        # candidate table [nb, S+2]: columns 0..S = extend with sel[j] (S = dummy), column S+1 = keep prefix
        from_blank = Pb[:, None] + pc[None, :]
        switch = Pnb[:, None] + pc[None, :]
        switch[np.arange(nb), rlast] = NEG_INF                   # a repeated char needs a blank in between
        table = np.concatenate([np.logaddexp(from_blank, switch), (Pnb + pc[rlast])[:, None]], axis=1)
        # prefix joining :138-155: mass of (parent + last char) moves to the child already in the beam
        for p, pref in enumerate(prefixes):
            if not pref:
                continue
            parents = [q for q, other in enumerate(prefixes) if other == pref[:-1]]
            if not parents:
                continue
            q = parents[0]
            table[p, -1] = np.logaddexp(table[p, -1], table[q, rlast[p]])
            table[q, rlast[p]] = NEG_INF
        new_Pb = np.logaddexp(Pb, Pnb) + p_blank
        score = table.copy()
        score[:, -1] = np.logaddexp(new_Pb, score[:, -1])
        flat = score.ravel()
        n_keep = min(k, int(np.isfinite(flat).sum()))
        if len(flat) <= n_keep:
            # multisort.py:7-8 returns arange(len(a)) here -- an index into rows only; reachable only when the
            # whole table is finite and not larger than k.  Keep every candidate (same hypothesis set).
            chosen = np.arange(len(flat))
        else:
            chosen = np.argsort(-flat, kind='stable')[:n_keep]
        rows, cols = np.unravel_index(chosen, score.shape)
        is_keep = cols == S + 1
        Pb = np.where(is_keep, new_Pb[rows], NEG_INF)
        Pnb = table[rows, cols]
        sel_ext = np.concatenate([sel, [-2, blank]])             # :267-268
        new_prefixes, new_last = [], np.empty(len(rows), dtype=np.int64)
        for i, (r, c) in enumerate(zip(rows, cols)):
            ch = int(sel_ext[c])
            if ch != blank:
                new_prefixes.append(prefixes[r] + [ch])
                new_last[i] = ch
            else:
                new_prefixes.append(prefixes[r])
                new_last[i] = last[r]
        prefixes, last = new_prefixes, new_last
    total = np.logaddexp(Pb, Pnb)
    order = sorted(range(len(prefixes)), key=lambda i: -total[i])    # BagOfHypotheses.sort (stable)
    return [(prefixes[i], float(total[i])) for i in order]


def best_hypothesis(hyps):
    """bag_of_hypotheses.py:64-65: first maximal vis_sc (lm_sc = 0 without a language model)."""
    best = None
    for ids, sc in hyps:
        if best is None or sc > best[1]:
            best = (ids, sc)
    return best


class GreedyDecoderOracle:
    def __init__(self, letters, symbol_separator=''):
        validate_letters(letters)
        self.letters = letters
        self.sep = symbol_separator

    def __call__(self, logprobs):
        text, score, _ = greedy(logprobs, self.letters, self.sep)
        return [(text, score)]


class PrefixBeamOracle:
    def __init__(self, letters, k, symbol_separator=''):
        validate_letters(letters)
        validate_beam(k)
        self.letters = letters
        self.k = k
        self.sep = symbol_separator

    def __call__(self, logprobs):
        hyps = prefix_beam(logprobs, self.k)
        return [(self.sep.join(self.letters[i] for i in ids), sc) for ids, sc in hyps]
