"""CPU suite, part 2: the host side of libtmac_b200.so without a GPU.
 * the shared library loads and exports every symbol include/tmac_b200.h declares;
 * kcfg registry / kcfg.ini parsing (host logic);
 * the reference-layout -> stream-layout transform (tmac_b200_debug_encode) decodes, through a
   numpy emulation of the kernel's PRMT/DP4A algebra (tests/stream_emul.py), to exactly the
   integer sums of the oracle, for symmetric and general LUTs;
 * no compute entry point works without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import stream_emul as E
import tmac_b200 as tb
import tmac_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = tb.load()
    hdr = open(os.path.join(ROOT, "include", "tmac_b200.h")).read()
    declared = set(re.findall(r"TMAC_B200_API\s+[\w\s\*]+?\b(\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "libtmac_b200.so does not export %s" % name
    assert declared == set(tb.EXPORTS), declared ^ set(tb.EXPORTS)
    assert lib.tmac_b200_version() >= 100


def test_kcfg_registry_and_ini(tmp_path):
    lib = tb.load()
    lib.tmac_b200_clear_kcfg()
    ini = tmp_path / "kcfg.ini"
    # the reference's own format (deploy/tuned/aarch64-hf-bitnet-3b/kcfg.ini, deploy/compile.py:156-165)
    ini.write_text("""[qgemm_lut_t1_int8_m6400_k8640_n1_b2]
bm = 128
simd_n_in = 16
simd_n_out = 8
kfactor = 16
group_size = 128
lut_scales_size = 135
scales_size = 1
n_tile_num = 50

[qgemm_lut_t1_int8_m8192_k4096_n1_b2]
bm = 128
simd_n_in = 16
simd_n_out = 8
kfactor = 16
group_size = 128
lut_scales_size = 64
scales_size = 262144
n_tile_num = 64
""")
    assert lib.tmac_b200_load_kcfg_file(str(ini).encode()) == 2
    c = tb.KCfg()
    assert lib.tmac_b200_find_kcfg(6400, 8640, 2, C.byref(c)) == 0
    assert (c.M, c.K, c.bits, c.bm, c.act_group_size, c.one_scale, c.zero_point) == (3200, 8640, 2, 128, 64, 1, 0)
    assert lib.tmac_b200_find_kcfg(8192, 4096, 2, C.byref(c)) == 0
    assert (c.M, c.group_size, c.act_group_size, c.one_scale, c.zero_point) == (4096, 128, 64, 0, 1)
    # a tile-sized m (bm) resolves to the same section, like the reference's dispatcher (kernels.h:21-37)
    assert lib.tmac_b200_find_kcfg(128, 4096, 2, C.byref(c)) == 0 and c.M == 4096
    assert lib.tmac_b200_find_kcfg(128, 4096, 4, C.byref(c)) == -1
    assert b"no kcfg" in lib.tmac_b200_last_error()
    # wsize / nbytes formulas of ggml-tmac.cpp:250-288
    assert lib.ggml_tmac_b200_mul_mat_get_wsize(4096, 4096, 1, 2) == ((4096 * 4 + 64 * 2 * 4 - 1) // 64 + 1) * 64
    assert lib.ggml_tmac_b200_get_nbytes(4096, 4096, 2) == 4096 * 4096 // 8 * 2 + 262144 * 4
    assert [lib.ggml_tmac_get_type_bits(t) for t in (36, 37, 38, 39, 2, 34, 35, 0)] == [1, 2, 3, 4, 4, 2, 2, 0]
    assert lib.ggml_tmac_b200_can_mul_mat(37, 1, 1, b"blk.0.attn_q.weight") == 1
    assert lib.ggml_tmac_b200_can_mul_mat(37, 1, 1, b"output.weight") == 0
    bad = tb.make_kcfg(100, 4096, 2, 128)
    assert lib.tmac_b200_register_kcfg(C.byref(bad)) == -1  # bm does not divide M*bits
    lib.tmac_b200_clear_kcfg()


CASES = [
    T.Config(128, 512, 2, zero_point=True), T.Config(64, 256, 4), T.Config(128, 256, 4, zero_point=True),
    T.Config(128, 256, 3), T.Config(256, 256, 1, zero_point=True), T.Config(160, 640, 2, one_scale=True),
    T.Config(64, 256, 4, kfactor=8, group_size=32, act_group_size=32),
    T.Config(192, 512, 2, bm=128, kfactor=8, group_size=64, act_group_size=32, zero_point=True),
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "w%d_%dx%d%s" % (c.bits, c.Mout, c.K, "_os" if c.one_scale else ""))
@pytest.mark.parametrize("sym", [True, False], ids=["symLUT", "generalLUT"])
def test_stream_layout_decodes_to_oracle_sums(oracle, cfg, sym):
    lib = tb.load()
    cfg = cfg.resolved()
    w, sc, z, x = T.make_problem(cfg, seed=21, N=1)
    A, S = T.pack_reference_layout(w, sc, z, cfg)
    k = tb.make_kcfg(cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size, cfg.zero_point, cfg.one_scale)
    lay = (C.c_int * 12)()
    n = lib.tmac_b200_debug_encode(C.byref(k), A.ctypes.data, S.ctypes.data, None, 0, lay)
    assert n > 0, tb.last_error()
    buf = np.zeros(n, np.uint8)
    assert lib.tmac_b200_debug_encode(C.byref(k), A.ctypes.data, S.ctypes.data, buf.ctypes.data, n, lay) == n
    if sym:
        qlut, _, _ = oracle.preprocessor(x, cfg.act_group_size)
    else:  # the reference's own verification feeds random, non-symmetric LUTs (python/t_mac/ops/qgemm.py:289)
        qlut = np.random.default_rng(5).integers(-127, 128, size=(1, cfg.K // 4, 16)).astype(np.int8)
    got = E.emulate_int_sums(buf, list(lay), qlut[0], cfg.Mout, cfg.bits, sym)
    cb = oracle.cbits(cfg, A, qlut)[0].astype(np.int64)          # [M*bits] reference plane layout
    rows = np.arange(cfg.Mout)
    want = np.zeros(cfg.Mout, np.int64)
    for b in range(cfg.bits):
        want += (1 << b) * cb[(rows // 8) * 8 * cfg.bits + b * 8 + rows % 8]
    assert np.array_equal(got, want)
    if not cfg.one_scale:
        s_dec, z_dec = E.decode_scales(buf, list(lay), cfg.Mout)
        per_chunk = np.repeat(sc, cfg.group_size // int(lay[4]), axis=1) if cfg.group_size > int(lay[4]) else sc
        assert np.array_equal(s_dec, per_chunk)
        if cfg.zero_point:
            assert np.array_equal(z_dec, z)


def test_scales_fall_back_to_fp32_when_not_fp16_exact(oracle):
    lib = tb.load()
    cfg = T.Config(64, 256, 4).resolved()
    w, sc, z, x = T.make_problem(cfg, seed=2)
    sc = (sc * np.float32(1.0001)).astype(np.float32)  # no longer fp16-representable
    A, S = T.pack_reference_layout(w, sc, z, cfg)
    k = tb.make_kcfg(cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size)
    lay = (C.c_int * 12)()
    n = lib.tmac_b200_debug_encode(C.byref(k), A.ctypes.data, S.ctypes.data, None, 0, lay)
    assert n > 0 and lay[7] == 4
    buf = np.zeros(n, np.uint8)
    lib.tmac_b200_debug_encode(C.byref(k), A.ctypes.data, S.ctypes.data, buf.ctypes.data, n, lay)
    s_dec, _ = E.decode_scales(buf, list(lay), cfg.Mout)
    assert np.array_equal(s_dec, sc)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    lib = tb.load()
    assert lib.tmac_b200_init(-1) == -1
    assert b"no CUDA device" in lib.tmac_b200_last_error() or b"cuda" in lib.tmac_b200_last_error().lower()
    x = np.zeros((1, 128), np.float32)
    ls = np.zeros(2, np.float32); lb = np.zeros(2, np.float32); q = np.zeros((32, 16), np.int8)
    assert lib.tmac_b200_preprocessor(128, 1, 64, 0, x.ctypes.data, ls.ctypes.data, lb.ctypes.data, q.ctypes.data) == -1


GGML_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ggml_blocks.npz")


@pytest.mark.parametrize("name,bits,block", [("q4_0", 4, 32), ("tq1_0", 2, 256), ("tq2_0", 2, 256)])
def test_ggml_block_decode_matches_reference_dequant(name, bits, block):
    """Q4_0 / TQ1_0 / TQ2_0 blocks (golden bytes + dequantised values produced by the reference's own gguf-py,
    oracle/make_golden_ggml.py) decode to codes w and scales d with (w - 2^(bits-1)) * d == dequant exactly: the
    semantics of the reference's accessors (ggml-tmac.cpp:98-236) + its dequant convention (tests/test_e2e.py:69-77)."""
    z = np.load(GGML_GOLDEN)
    q, deq, qt = z[name + "_bytes"], z[name + "_dequant"], int(z[name + "_type"])
    rows, K = deq.shape
    lib = tb.load()
    w = np.zeros((rows, K), np.uint8); sc = np.zeros((rows, K // block), np.float32)
    q = np.ascontiguousarray(q)
    assert lib.tmac_b200_debug_decode_ggml(qt, q.ctypes.data, K, rows, w.ctypes.data, sc.ctypes.data) == block
    assert lib.ggml_tmac_get_type_bits(qt) == bits
    assert w.max() < (1 << bits)
    if bits == 2:
        assert w.min() >= 1            # ternary: codes 1, 2, 3 <-> -1, 0, +1
    real = (w.astype(np.float32) - float(1 << (bits - 1))) * np.repeat(sc, block, axis=1)
    assert np.array_equal(real, deq)
    # K that is not a multiple of the block size is rejected
    assert lib.tmac_b200_debug_decode_ggml(qt, q.ctypes.data, K - 32 if block == 256 else K - 16, rows, w.ctypes.data, sc.ctypes.data) == -1


@pytest.mark.parametrize("tag", ["w4_v2", "w2_v1", "w4_v1"])
def test_gptq_unpack_matches_reference(tag):
    """GPTQ safetensors tensors -> (w, scales, biased zeros): bit-identical to the reference's unpack_gptqv2
    (python/t_mac/model_utils.py:95-129; goldens by oracle/make_golden_gptq.py), including the fp16 rounding of
    (z - 2^(bits-1)) * scale and AutoGPTQ's zero + 1."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gptq_unpack.npz"))
    bits, K, M, gs, v2 = [int(v) for v in z[tag + "_meta"]]
    qw, qz, sc = (np.ascontiguousarray(z[tag + k]) for k in ("_qweight", "_qzeros", "_scales"))
    w = np.zeros((M, K), np.uint8); s = np.zeros((M, K // gs), np.float32); zr = np.zeros_like(s)
    lib = tb.load()
    assert lib.tmac_b200_debug_unpack_gptq(qw.ctypes.data, sc.ctypes.data, qz.ctypes.data, K, M, bits, gs, v2, w.ctypes.data, s.ctypes.data, zr.ctypes.data) == 0
    assert np.array_equal(w, z[tag + "_w"])
    assert np.array_equal(s, z[tag + "_s"].astype(np.float32))
    assert np.array_equal(zr.view(np.uint32), z[tag + "_z"].astype(np.float32).view(np.uint32))
    assert lib.tmac_b200_debug_unpack_gptq(qw.ctypes.data, sc.ctypes.data, qz.ctypes.data, K, M, 3, gs, v2, w.ctypes.data, s.ctypes.data, zr.ctypes.data) == -1


def test_generated_kcfg_presets_load(tmp_path):
    """kcfg.ini files written by tools/make_kcfg.py (reference format) load through the reference-format reader and
    resolve every shape of the preset with the grouping the preset means."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_kcfg", os.path.join(ROOT, "tools", "make_kcfg.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    lib = tb.load()
    for preset, (bits, gs, ags, zp, one, shapes) in mk.PRESETS.items():
        lib.tmac_b200_clear_kcfg()
        path = mk.write(preset, str(tmp_path / preset / "kcfg.ini"))
        assert lib.tmac_b200_load_kcfg_file(path.encode()) == len(shapes), preset
        for mout, k in shapes:
            c = tb.KCfg()
            assert lib.tmac_b200_find_kcfg(mout * bits, k, bits, C.byref(c)) == 0, (preset, mout, k)
            assert (c.M, c.K, c.bits, c.group_size if not one else gs, c.one_scale, c.zero_point) == (mout, k, bits, gs, int(one), int(zp)), (preset, mout, k)
            assert c.act_group_size == (k if ags <= 0 else ags)
            assert (mout * bits) % c.bm == 0
    lib.tmac_b200_clear_kcfg()


def test_default_kcfg_follows_the_reference_rule():
    """tmac_b200_default_kcfg = first candidate of every knob of QGeMMLUTBitsCodegen._define_config (qgemm.py:98-115)."""
    lib = tb.load()
    c = tb.KCfg()
    cases = [  # (M, K, bits, gs, ags, zp, one) -> (bm, kfactor)
        ((4096, 4096, 2, 128, 64, 1, 0), (256, 16)), ((11008, 4096, 4, 128, 64, 1, 0), (256, 16)), ((4096, 11008, 2, 128, 64, 0, 0), (256, 16)),
        ((4096, 4096, 4, 32, 32, 0, 0), (256, 8)), ((3200, 8640, 2, 128, -1, 0, 1), (256, 8)), ((8640, 3200, 2, 128, -1, 0, 1), (128, 8)),
        ((384, 1024, 3, 128, 64, 1, 0), (192, 16)), ((160, 640, 2, 128, 64, 0, 0), (320, 16)),
    ]
    for args, (bm, kf) in cases:
        assert lib.tmac_b200_default_kcfg(*args, C.byref(c)) == 0, args
        assert (c.bm, c.kfactor, c.simd_n_in, c.simd_n_out) == (bm, kf, 16, 8), (args, c.bm, c.kfactor)
        assert c.act_group_size == (args[1] if args[4] <= 0 else args[4])
        assert lib.tmac_b200_register_kcfg(C.byref(c)) == 0                 # a default kcfg is always a valid one
        assert T.default_bm(args[0] * args[2], args[2]) == bm
    assert lib.tmac_b200_default_kcfg(100, 4096, 2, 128, 64, 0, 0, C.byref(c)) == -1   # no tile divides 200 plane rows
    lib.tmac_b200_clear_kcfg()


def test_converter_quantisers_match_reference():
    """BitDistiller group quantiser: bit-identical to the reference method; BitNet ternarisation: identical codes, scale within
    one ulp (goldens made by calling the reference's convert_hf_to_gguf.py methods, oracle/make_golden_quantizers.py)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "quantizers.npz"))
    lib = tb.load()
    for tag in ("bd_w2_g128", "bd_w4_g64", "bd_w3_rowwise"):
        bits, gs, rows, cols = [int(v) for v in z[tag + "_meta"]]
        W = np.ascontiguousarray(z[tag + "_in"])
        ng = cols // (gs if gs > 0 else cols)
        codes = np.zeros((rows, cols), np.uint8); sc = np.zeros((rows, ng), np.float32); zr = np.zeros_like(sc)
        assert lib.tmac_b200_quantize_bitdistiller(W.ctypes.data, rows, cols, bits, gs, codes.ctypes.data, sc.ctypes.data, zr.ctypes.data) == 0
        assert np.array_equal(codes, z[tag + "_w"]), tag
        assert np.array_equal(sc.view(np.uint32), z[tag + "_scales"].view(np.uint32)), tag
        assert np.array_equal(zr.view(np.uint32), z[tag + "_zeros"].view(np.uint32)), tag
        # dequantised with the T-MAC convention the weights come back within half a step (row 3 holds the degenerate constant
        # group of the fixture, which the reference algorithm itself cannot represent)
        real = (codes.astype(np.float32) - (1 << (bits - 1))) * np.repeat(sc, cols // ng, axis=1) - np.repeat(zr, cols // ng, axis=1)
        keep = np.arange(rows) != 3
        assert np.abs(real - W)[keep].max() <= 0.5001 * sc[keep].max() + 1e-7
    for tag in ("bitnet_a", "bitnet_b"):
        W = np.ascontiguousarray(z[tag + "_in"])
        codes = np.zeros(W.shape, np.uint8); s = np.zeros(1, np.float32)
        assert lib.tmac_b200_quantize_bitnet(W.ctypes.data, W.shape[0], W.shape[1], codes.ctypes.data, s.ctypes.data) == 0
        assert np.array_equal(codes, z[tag + "_codes"]), tag
        ref = np.float32(z[tag + "_scale"])
        assert abs(int(s.view(np.uint32)[0]) - int(ref.view(np.uint32))) <= 1, (s[0], ref)
        assert set(np.unique(codes)) <= {1, 2, 3}
    assert lib.tmac_b200_quantize_bitdistiller(W.ctypes.data, 4, 100, 2, 64, codes.ctypes.data, s.ctypes.data, s.ctypes.data) == -1
