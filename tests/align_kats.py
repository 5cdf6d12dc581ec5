"""Known-answer tests of the reference's forced alignment (test/test_force_alignment.py:171-318), as data shared by
the oracle test (CPU) and the GPU parity test.  Viterbi-level cases given there on an explicit 3-state transition
matrix are expressed through force_align with the one-symbol transcription [1] (blank = 0) that builds that matrix."""
import numpy as np

INF = np.inf

# (name, neg_logprobs, transcription, blank, expected symbols or 'ValueError')
KATS = [
    ('trivial', [[0.0, 10.0], [10.0, 0.0]], [1], 0, [0, 1]),                                           # :282-288
    ('single_symbol_multi_blank', [[0.0, 10.0]] * 3 + [[10.0, 0.0]] + [[0.0, 10.0]] * 2, [1], 0,
     [0, 0, 0, 1, 0, 0]),                                                                              # :290-300
    ('multi_symbol_regression', [[0.0, 10.0, 10.0], [10.0, 10.0, 0.0], [5.0, 10.0, 5.0], [10.0, 10.0, 0.0]],
     [2, 2], 0, [0, 2, 0, 2]),                                                                         # :302-310
    ('skipping_first_regression', [[10.0, 10.0, 0.0], [0.0, 10.0, 10.0]], [1, 2], 0, [1, 2]),         # :312-318
    ('multi_frame_symbol', [[0.0, 10.0]] * 2 + [[10.0, 0.0]] * 3 + [[0.0, 10.0]], [1], 0,
     [0, 0, 1, 1, 1, 0]),                                                                              # :247-263
    ('respect_final_state', [[0.0, 10.0], [0.0, 8.0], [0.0, 10.0]], [1], 0, [0, 1, 0]),               # :265-278
    ('impossible_alignment', [[0.0, INF], [0.0, INF], [0.0, INF]], [1], 0, 'ValueError'),             # :214-227
    ('blank_in_transcription', [[0.0, 1.0, 2.0]] * 4, [1, 0, 2], 0, 'ValueError'),                    # :70-71
    ('empty_transcription', [[0.0, 1.0]] * 3, [], 0, 'ValueError'),                                   # :47-48
]
