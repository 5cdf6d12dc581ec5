"""TEST INFRASTRUCTURE ONLY -- CPU restatement of pero-ocr's CTC forced alignment.

Follows pero_ocr/core/force_alignment.py: force_align (:13-35), hmm_trans_from_string (:38-60), complete_state_seq
(:63-75), initial_cost / final_cost (:78-101), compute_update (:118-130, the tie rule), viterbi_align (:133-149),
backtrack (:104-112), align_text (:152-165).  Pinned against the reference's own known-answer tests
(test/test_force_alignment.py:171-318, restated in tests/test_oracle_align.py) and against outputs of the unmodified
reference functions on seeded random inputs (tests/golden/align.npz, oracle/make_golden.py: golden_align).
"""
import numpy as np


def force_align(neg_logprobs, symbols, blank, return_seq_positions=False):
    """-> list of per-frame symbols (incl. blanks) of the cheapest path, or per-frame character indices (-1 = blank)."""
    symbols = [int(s) for s in symbols]
    if len(symbols) < 1:
        raise ValueError("Cannot construct a CTC 'HMM' from an empty string")
    if blank in symbols:
        raise ValueError(f'The blank symbol {blank} is present in the non blank seq {symbols}')
    neg = np.asarray(neg_logprobs)
    n_states = 2 * len(symbols) + 1
    state_symbol = [blank if s % 2 == 0 else symbols[s // 2] for s in range(n_states)]
    state_char = [-1 if s % 2 == 0 else s // 2 for s in range(n_states)]
    cost = np.full(n_states, np.inf)
    cost[0] = cost[1] = 0.0
    cost = cost + neg[0, state_symbol]                       # float64 + input dtype -> float64
    back = np.zeros((neg.shape[0], n_states), dtype=np.int64)
    for t in range(1, neg.shape[0]):
        frame = neg[t, state_symbol]
        new = np.full(n_states, np.inf)
        for s in range(n_states):
            sources = []
            if s % 2 == 1 and s >= 3 and symbols[s // 2] != symbols[s // 2 - 1]:
                sources.append(s - 2)                        # skip the blank between two different symbols
            if s >= 1:
                sources.append(s - 1)
            sources.append(s)
            for j in sources:                                # ascending source state, strict improvement only
                c = cost[j] + frame[s]
                if c < new[s]:
                    new[s] = c
                    back[t, s] = j
        cost = new
    final = np.full(n_states, np.inf)
    final[-1] = final[-2] = 0.0
    total = cost + final
    if np.amin(total) == np.inf:
        raise ValueError('It was not possible to align the states with the logits, best path has cost of np.inf')
    state = int(np.argmin(total))
    path = [state]
    for t in range(neg.shape[0] - 1, 0, -1):
        state = int(back[t, state])
        path.append(state)
    path.reverse()
    return [state_char[s] for s in path] if return_seq_positions else [state_symbol[s] for s in path]


def align_text(neg_logprobs, transcription, blank):
    """One frame per character: among the frames aligned to it, the one with the largest per-frame max probability."""
    neg = np.asarray(neg_logprobs)
    chars = np.asarray(force_align(neg, transcription, blank, return_seq_positions=True))
    frame_best = (-neg).max(axis=-1)
    out = np.zeros(len(transcription), dtype=np.int32)
    for i in range(len(transcription)):
        frames = np.nonzero(chars == i)[0]
        out[i] = frames[np.argmax(frame_best[frames])]
    return out


# ---- seeded cases shared by oracle/make_golden.py and the tests ---------------------------------------------------
def align_cases():
    """(name, neg_logprobs [T, C], labels, blank).  Peaky CTC-like outputs with the true text, a wrong text, repeated
    characters, a text as long as T allows, float32 and float64 inputs, exact ties."""
    from . import cases
    rng = np.random.default_rng(21)
    out = []
    lp = cases.peaky_logprobs(rng, 6, 64, 12, sharp=9.0)
    for i in range(4):
        best = lp[i].argmax(axis=1)
        text = [int(c) for k, c in enumerate(best) if c != 11 and (k == 0 or best[k - 1] != c)]
        if not text:
            text = [3]
        out.append((f'peaky_true_{i}', (-lp[i]).astype(np.float32 if i % 2 else np.float64), text, 11))
    out.append(('peaky_wrong_text', (-lp[4]).astype(np.float32), [1, 2, 3, 4, 5, 6, 7], 11))
    out.append(('repeats', (-lp[5]).astype(np.float32), [2, 2, 2, 5, 5, 1], 11))
    flat = -np.log(np.full((9, 4), 0.25))
    out.append(('all_ties', flat, [0, 1, 0], 3))                       # every path costs the same: pure tie rules
    out.append(('tight_fit', (rng.random((7, 5)) * 3).astype(np.float32), [0, 0, 1, 1], 4))   # T == minimal length
    out.append(('long_text', (rng.random((80, 30)) * 5).astype(np.float32), [int(v) for v in rng.integers(0, 29, 35)], 29))
    return out
