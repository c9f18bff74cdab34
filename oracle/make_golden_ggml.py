"""TEST INFRASTRUCTURE: golden vectors for the ggml block formats the reference re-permutes at load time
(Q4_0, TQ1_0, TQ2_0; 3rdparty/llama.cpp/ggml/src/ggml-tmac.cpp:98-236).  Generated in the build container by importing
the reference's own gguf-py (3rdparty/llama.cpp/gguf-py/gguf/quants.py, numpy only): quantised block bytes and their
dequantised fp32 values.  /root/reference does not exist on the GPU box, so the result is committed under
tests/golden/ggml_blocks.npz together with this script.

    python oracle/make_golden_ggml.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/3rdparty/llama.cpp/gguf-py"
sys.path.insert(0, REF)
from gguf import quants                                       # noqa: E402
from gguf.constants import GGMLQuantizationType as QT         # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ggml_blocks.npz")
rng = np.random.default_rng(20240923)
out = {}
for name, qt, rows, K in (("q4_0", QT.Q4_0, 64, 512), ("tq1_0", QT.TQ1_0, 128, 512), ("tq2_0", QT.TQ2_0, 128, 512)):
    if qt == QT.Q4_0:
        W = rng.standard_normal((rows, K)).astype(np.float32)
    else:                                                      # ternary weights with a per-row magnitude, like BitNet exports
        W = (rng.integers(-1, 2, size=(rows, K)) * np.abs(rng.standard_normal((rows, 1))) * 0.05).astype(np.float32)
    q = quants.quantize(W, qt)
    d = quants.dequantize(q, qt)
    out[name + "_type"] = np.int32(int(qt))
    out[name + "_bytes"] = q
    out[name + "_dequant"] = d.astype(np.float32)
np.savez_compressed(OUT, **out)
print("wrote", OUT, {k: getattr(v, "shape", v) for k, v in out.items()})
