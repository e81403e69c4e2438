"""Regenerates tests/golden/golden_rvq.npz: a small split residual vector quantiser, latents, and the codes / decoded latents
computed by the independent numpy restatement of the reference graphs in tests/test_mimi_rvq.py (NOT by the C oracle), so that the
oracle, the numpy restatement and the CUDA path are all held to one committed answer.
(The reference ships no golden vectors and cannot be built or imported here: SURVEY.md 4, 8c.)
    python tests/golden/make_golden_rvq.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import _pkgload  # noqa: E402
_pkgload.load()
from test_mimi_rvq import make_quantiser, np_split_decode, np_split_encode  # noqa: E402

rng = np.random.default_rng(2027)
qz = make_quantiser(rng, 1, 5, 96, 24, 40)
x = (rng.standard_normal((11, 40)) * 2.5).astype(np.float32)
codes = np_split_encode(qz, x, 6)
y = np_split_decode(qz, codes)
np.savez_compressed(os.path.join(HERE, "golden_rvq.npz"), cb_first=qz[0], cb_rest=qz[1], in_first=qz[2], in_rest=qz[3], out_first=qz[4],
                    out_rest=qz[5], x=x, codes=codes, y=y)
print("wrote golden_rvq.npz", codes.shape, y.shape)
