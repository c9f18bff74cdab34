"""TEST INFRASTRUCTURE: golden vectors for the converter-side quantisers of the reference
(3rdparty/llama.cpp/convert_hf_to_gguf.py: Model._t_mac_quantize_tensor_bitdistiller :409-452, BitnetModel.weight_quant
:1884-1893 followed by the T-MAC ternary rule :1909-1917), produced by calling the reference methods themselves (the script is
imported by file path; torch is present in the build container).  Output: tests/golden/quantizers.npz (committed).

    python oracle/make_golden_quantizers.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/3rdparty/llama.cpp/gguf-py")
spec = importlib.util.spec_from_file_location("ref_convert", "/root/reference/3rdparty/llama.cpp/convert_hf_to_gguf.py")
conv = importlib.util.module_from_spec(spec); spec.loader.exec_module(conv)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "quantizers.npz")
rng = np.random.default_rng(31337)
out = {}

# BitDistiller-style asymmetric group quantiser (the converter's path for fp16 BitDistiller checkpoints)
for tag, bits, gs, rows, cols in (("bd_w2_g128", 2, 128, 48, 512), ("bd_w4_g64", 4, 64, 32, 256), ("bd_w3_rowwise", 3, -1, 16, 384)):
    W = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
    W[3, :gs if gs > 0 else cols] = 0.125                                  # a constant group: (max - min) clamps to 1e-5
    w, s, z = conv.Model._t_mac_quantize_tensor_bitdistiller(None, torch.from_numpy(W.copy()), n_bit=bits, zero_point=True, q_group_size=gs)
    out[tag + "_meta"] = np.array([bits, gs, rows, cols], np.int32)
    out[tag + "_in"] = W; out[tag + "_w"] = w; out[tag + "_scales"] = s.astype(np.float32); out[tag + "_zeros"] = z.astype(np.float32)

# BitNet b1.58: absmean ternarisation, then codes = round(w / max|w| + 2), one scale
for tag, rows, cols in (("bitnet_a", 64, 320), ("bitnet_b", 33, 128)):
    W = (rng.standard_normal((rows, cols)) * 0.03).astype(np.float32)
    data = conv.BitnetModel.weight_quant(None, torch.from_numpy(W.copy())).numpy()
    scale = np.max(np.abs(data))
    codes = np.round(data / scale + 2).astype(np.uint8)
    out[tag + "_in"] = W; out[tag + "_codes"] = codes; out[tag + "_scale"] = np.float32(scale)

np.savez_compressed(OUT, **out)
print("wrote", OUT, os.path.getsize(OUT))
