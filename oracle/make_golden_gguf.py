"""TEST INFRASTRUCTURE: a tiny .gguf written by the reference's own tooling -- GGUFWriter and quants of the vendored
gguf-py, preprocess_weights of python/t_mac/weights.py for the I2 blob (permuted weights || fp32 scales, exactly
preprocess_for_t_mac, model_utils.py:243-271) -- plus an .npz with what the reader must find.  Generated in the build
container; both files are committed (tests/golden/tiny_tmac.gguf, tiny_tmac_gguf.npz).

    python oracle/make_golden_gguf.py
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "3rdparty", "llama.cpp", "gguf-py"))
import gguf                                                    # noqa: E402
from gguf import quants                                        # noqa: E402
from gguf.constants import GGMLQuantizationType as QT          # noqa: E402

spec = importlib.util.spec_from_file_location("ref_weights", os.path.join(REF, "python", "t_mac", "weights.py"))
ref_weights = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_weights)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "tiny_tmac.gguf")
rng = np.random.default_rng(4242)
gold = {}

wr = gguf.GGUFWriter(OUT, "llama")
wr.add_name("tiny t-mac fixture")
wr.add_block_count(1)
wr.add_context_length(128)
wr.add_float32("tmac.test_float", 0.5)
wr.add_bool("tmac.test_bool", True)
wr.add_array("tmac.test_strings", ["alpha", "beta", "gamma"])
wr.add_array("tmac.test_ints", [3, 1, 4, 1, 5])

emb = rng.standard_normal((8, 16)).astype(np.float32)
wr.add_tensor("token_embd.weight", emb)
gold["token_embd"] = emb

# I2: W2 g128 zero-point 128 x 256, bm 256, kfactor 16 (the blob preprocess_for_t_mac returns)
M, K, bits, gs, bm, kf = 128, 256, 2, 128, 256, 16
w = rng.integers(0, 1 << bits, size=(M, K), dtype=np.uint8)
sc = (np.abs(rng.standard_normal((M, K // gs))) * 0.01 + 1e-3).astype(np.float16)
zr = (rng.standard_normal((M, K // gs)) * 0.01).astype(np.float16)
pw, ps = ref_weights.preprocess_weights(w, sc, zr, bits=bits, g=4, bm=bm, kfactor=kf, simd_n_in=16, simd_n_out=8)
blob = np.concatenate([pw.flatten(), ps.astype(np.float32).copy().view(np.uint8).flatten()])
wr.add_tensor("blk.0.attn_q.weight", blob, raw_dtype=QT.I2, raw_shape=gguf.quant_shape_to_byte_shape((M, K), QT.I2))
gold.update(i2_blob=blob, i2_w=w, i2_scales=sc.astype(np.float32), i2_zeros=zr.astype(np.float32))

for name, qt, rows, cols in (("blk.0.ffn_up.weight", QT.Q4_0, 64, 256), ("blk.0.ffn_down.weight", QT.TQ2_0, 128, 256), ("blk.0.attn_k.weight", QT.TQ1_0, 128, 256)):
    if qt == QT.Q4_0:
        W = rng.standard_normal((rows, cols)).astype(np.float32)
    else:
        W = (rng.integers(-1, 2, size=(rows, cols)) * np.abs(rng.standard_normal((rows, 1))) * 0.05).astype(np.float32)
    q = quants.quantize(W, qt)
    wr.add_tensor(name, q, raw_dtype=qt)
    key = qt.name.lower()
    gold[key + "_bytes"] = q
    gold[key + "_dequant"] = quants.dequantize(q, qt).astype(np.float32)

wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_tmac_gguf.npz"), **gold)
print("wrote", OUT, os.path.getsize(OUT))
