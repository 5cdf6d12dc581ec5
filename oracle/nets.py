"""TEST INFRASTRUCTURE ONLY -- the synthetic net definitions live in pero_ocr_b200/synthetic.py (shared with
bench.py and smoke()); re-exported here for the oracle modules."""
from pero_ocr_b200.synthetic import *  # noqa: F401,F403
from pero_ocr_b200.synthetic import VGG_FRONTEND, make_net, seeded_state_dict  # noqa: F401
