"""TEST INFRASTRUCTURE: golden vectors for the GPTQ checkpoint unpack (python/t_mac/model_utils.py:95-129,
unpack_gptqv2), generated in the build container by importing the reference function itself.  model_utils imports
t_mac.weights, and t_mac/__init__ imports TVM (absent), so both modules are loaded by file path under a stub package.
Output: tests/golden/gptq_unpack.npz (committed; /root/reference does not exist on the GPU box).

    python oracle/make_golden_gptq.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/python/t_mac"
pkg = types.ModuleType("t_mac"); pkg.__path__ = [REF]; sys.modules["t_mac"] = pkg
for name in ("weights", "model_utils"):
    spec = importlib.util.spec_from_file_location("t_mac." + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec); sys.modules["t_mac." + name] = mod; spec.loader.exec_module(mod)
unpack_gptqv2 = sys.modules["t_mac.model_utils"].unpack_gptqv2

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gptq_unpack.npz")
rng = np.random.default_rng(77)
out = {}
for tag, bits, K, M, gs, v2 in (("w4_v2", 4, 512, 256, 128, True), ("w2_v1", 2, 512, 128, 64, False), ("w4_v1", 4, 256, 64, 32, False)):
    p = 32 // bits
    qweight = rng.integers(-2**31, 2**31, size=(K // p, M), dtype=np.int64).astype(np.int32)
    qzeros = rng.integers(-2**31, 2**31, size=(K // gs, M // p), dtype=np.int64).astype(np.int32)
    if not v2:   # AutoGPTQ stores zero - 1: keep every field <= 2^bits - 2 so that + 1 stays in range, as real checkpoints do
        fields = rng.integers(0, (1 << bits) - 1, size=(K // gs, M // p, p), dtype=np.int64)
        qzeros = (fields << (bits * np.arange(p))).sum(axis=-1).astype(np.uint32).view(np.int32)
    scales = (np.abs(rng.standard_normal((K // gs, M))) * 0.01 + 1e-3).astype(np.float16)
    w, s, z, b, g = unpack_gptqv2(qweight, scales, qzeros, v2)
    assert (b, g) == (bits, gs)
    out[tag + "_meta"] = np.array([bits, K, M, gs, int(v2)], np.int32)
    out[tag + "_qweight"] = qweight; out[tag + "_qzeros"] = qzeros; out[tag + "_scales"] = scales
    out[tag + "_w"] = np.ascontiguousarray(w); out[tag + "_s"] = np.ascontiguousarray(s).astype(np.float16); out[tag + "_z"] = np.ascontiguousarray(z).astype(np.float16)
np.savez_compressed(OUT, **out)
print("wrote", OUT, os.path.getsize(OUT))
