"""GGUF reader (t-mac_b200/csrc/tmac_gguf.h behind tmac_b200_gguf_*), host-only: a file written by the reference's own
GGUFWriter / quants / preprocess_weights (oracle/make_golden_gguf.py) is parsed; metadata, directory and tensor bytes must be
what the writer put in; quantised tensors decode (block types) / re-encode (I2 blob) to the writer's inputs."""
import ctypes as C
import os

import numpy as np
import pytest

import tmac_b200 as tb
import tmac_oracle as T

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gg():
    lib = tb.load()
    h = lib.tmac_b200_gguf_open(os.path.join(GOLD, "tiny_tmac.gguf").encode())
    assert h > 0, lib.tmac_b200_last_error()
    yield lib, h
    assert lib.tmac_b200_gguf_close(h) == 0
    assert lib.tmac_b200_gguf_close(h) == -1


def info(lib, h, name):
    i = lib.tmac_b200_gguf_find_tensor(h, name.encode())
    assert i >= 0, name
    t = tb.GgufTensor()
    assert lib.tmac_b200_gguf_tensor_info(h, i, C.byref(t)) == 0
    return i, t


def bytes_of(t, n):
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(t.data)).copy()


def test_metadata_and_directory(gg):
    lib, h = gg
    assert lib.tmac_b200_gguf_tensor_count(h) == 5
    buf = C.create_string_buffer(64)
    assert lib.tmac_b200_gguf_meta_string(h, b"general.architecture", buf, 64) == 5 and buf.value == b"llama"
    assert lib.tmac_b200_gguf_meta_string(h, b"general.name", buf, 64) > 0 and buf.value == b"tiny t-mac fixture"
    v = C.c_double()
    assert lib.tmac_b200_gguf_meta_number(h, b"llama.block_count", C.byref(v)) == 0 and v.value == 1
    assert lib.tmac_b200_gguf_meta_number(h, b"llama.context_length", C.byref(v)) == 0 and v.value == 128
    assert lib.tmac_b200_gguf_meta_number(h, b"tmac.test_float", C.byref(v)) == 0 and v.value == 0.5
    assert lib.tmac_b200_gguf_meta_number(h, b"tmac.test_bool", C.byref(v)) == 0 and v.value == 1
    assert lib.tmac_b200_gguf_meta_number(h, b"tmac.test_strings", C.byref(v)) == -1      # arrays are skipped, not numbers
    assert lib.tmac_b200_gguf_meta_number(h, b"no.such.key", C.byref(v)) == -1
    assert lib.tmac_b200_gguf_find_tensor(h, b"no.such.tensor") == -1
    z = np.load(os.path.join(GOLD, "tiny_tmac_gguf.npz"))
    _, t = info(lib, h, "token_embd.weight")
    assert (t.ggml_type, t.n_dims, t.ne[0], t.ne[1]) == (0, 2, 16, 8)
    assert np.array_equal(bytes_of(t, 8 * 16 * 4).view(np.float32).reshape(8, 16), z["token_embd"])
    for name, typ, rows, cols, key in (("blk.0.ffn_up.weight", 2, 64, 256, "q4_0"), ("blk.0.ffn_down.weight", 35, 128, 256, "tq2_0"),
                                       ("blk.0.attn_k.weight", 34, 128, 256, "tq1_0")):
        _, t = info(lib, h, name)
        assert (t.ggml_type, t.n_dims, t.ne[0], t.ne[1]) == (typ, 2, cols, rows), name
        q = z[key + "_bytes"]
        assert t.nbytes >= q.size and t.offset % 32 == 0
        raw = bytes_of(t, q.size)
        assert np.array_equal(raw, q.reshape(-1))
        block = 32 if typ == 2 else 256
        bits = lib.ggml_tmac_get_type_bits(typ)
        w = np.zeros((rows, cols), np.uint8); sc = np.zeros((rows, cols // block), np.float32)
        assert lib.tmac_b200_debug_decode_ggml(typ, t.data, cols, rows, w.ctypes.data, sc.ctypes.data) == block
        real = (w.astype(np.float32) - float(1 << (bits - 1))) * np.repeat(sc, block, axis=1)
        assert np.array_equal(real, z[key + "_dequant"]), name


def test_i2_tensor_blob_is_the_reference_layout(gg):
    """The I2 tensor holds `permuted weights || fp32 scales` (model_utils.py:271): the mapped bytes equal the writer's blob,
    and pushing them through the reference-layout decoder gives the same stream bytes as encoding the plain weights."""
    lib, h = gg
    z = np.load(os.path.join(GOLD, "tiny_tmac_gguf.npz"))
    _, t = info(lib, h, "blk.0.attn_q.weight")
    M, K, bits = 128, 256, 2
    assert (t.ggml_type, t.ne[0], t.ne[1]) == (37, K, M)
    blob = z["i2_blob"]
    assert t.nbytes >= blob.size
    raw = bytes_of(t, blob.size)
    assert np.array_equal(raw, blob)
    cfg = T.Config(M, K, bits, bm=256, zero_point=True).resolved()
    A, S = T.pack_reference_layout(z["i2_w"], z["i2_scales"], z["i2_zeros"], cfg)
    assert np.array_equal(A.reshape(-1), blob[: M * K * bits // 8])
    assert np.array_equal(S.reshape(-1), blob[M * K * bits // 8:].view(np.float32))
    k = tb.make_kcfg(cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size, cfg.zero_point, cfg.one_scale)
    lay = (C.c_int * 12)()
    wptr = t.data
    sptr = t.data + M * K * bits // 8
    n = lib.tmac_b200_debug_encode(C.byref(k), wptr, sptr, None, 0, lay)
    assert n > 0
    a = np.zeros(n, np.uint8); b = np.zeros(n, np.uint8)
    assert lib.tmac_b200_debug_encode(C.byref(k), wptr, sptr, a.ctypes.data, n, lay) == n
    assert lib.tmac_b200_debug_encode(C.byref(k), A.ctypes.data, S.ctypes.data, b.ctypes.data, n, lay) == n
    assert np.array_equal(a, b)


def test_rejects_non_gguf_and_truncated_files(tmp_path):
    lib = tb.load()
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"NOPE" + bytes(64))
    assert lib.tmac_b200_gguf_open(str(bad).encode()) == -1 and b"magic" in lib.tmac_b200_last_error()
    src = open(os.path.join(GOLD, "tiny_tmac.gguf"), "rb").read()
    for cut in (10, 100, 400, len(src) - 5000):
        p = tmp_path / ("cut%d.gguf" % cut)
        p.write_bytes(src[:cut])
        assert lib.tmac_b200_gguf_open(str(p).encode()) == -1, cut
    assert lib.tmac_b200_gguf_open(b"/nonexistent/file.gguf") == -1
    # without a registered kcfg a quantised tensor cannot be uploaded -- and without a GPU nothing is ever uploaded
    h = lib.tmac_b200_gguf_open(os.path.join(GOLD, "tiny_tmac.gguf").encode())
    lib.tmac_b200_clear_kcfg()
    i = lib.tmac_b200_gguf_find_tensor(h, b"blk.0.attn_q.weight")
    assert lib.tmac_b200_gguf_load_tensor(h, i, None) == -1
    assert lib.tmac_b200_gguf_load_tensor(h, lib.tmac_b200_gguf_find_tensor(h, b"token_embd.weight"), None) == -1
    lib.tmac_b200_gguf_close(h)
