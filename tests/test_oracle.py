"""CPU suite, part 1: pin the oracle (oracle/tmac_oracle.c + tmac_oracle.py) against
 (a) the known-answer vector of the reference's tests/test_lut_ctor.cc:12-24,
 (b) golden fixtures generated from the reference's own code (oracle/make_golden.py),
 (c) the reference build itself (oracle/_ref) when it is present."""
import glob
import os

import numpy as np
import pytest

import tmac_oracle as T


def _cfg_from(arr):
    Mout, K, bits, bm, kf, gs, ags, zp, os_ = [int(x) for x in arr]
    return T.Config(Mout, K, bits, bm, kf, gs, ags, bool(zp), bool(os_))


def test_kat_lut_ctor(oracle, golden_dir):
    # tests/test_lut_ctor.cc:12-24: b[i] = i, i < 32, lut_ctor(0, 4)
    b = np.arange(32, dtype=np.float32)
    ls = oracle.partial_max(b, 0.0)
    q, s, lb = oracle.lut_ctor(b, ls)
    assert np.float32(ls) == np.float32(118.0) / np.float32(127.0)
    assert lb == -496.0
    assert q[0].tolist() == [-6, -6, -4, -4, -2, -2, 0, 0, 0, 0, 2, 2, 4, 4, 6, 6]
    assert q[1].tolist() == [-24, -15, -13, -4, -11, -2, 0, 9, -9, 0, 2, 11, 4, 13, 15, 24]
    g = np.load(os.path.join(golden_dir, "kat_lut_ctor.npz"))
    assert np.array_equal(q, g["qlut"]) and np.float32(ls) == g["lut_scales"] and np.float32(lb) == g["lut_biases"]


GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "w*.npz")))


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_golden(oracle, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = _cfg_from(g["cfg"])
    # synthetic-input generator is deterministic
    w, sc, z, x = T.make_problem(cfg, seed=int({"w2_zp_g128": 1, "w4_sym_g128": 2, "w4_zp_g128": 3, "w3_sym_g128": 4,
                                                 "w1_zp_g128": 5, "w2_bitnet_int32": 6, "w4_q40_g32": 7,
                                                 "w2_zp_bm128_kf8": 8}[name]), N=g["x"].shape[0])
    assert np.array_equal(w, g["w"]) and np.array_equal(x, g["x"])
    # packer restatement == reference python/t_mac/weights.py output
    A, S = T.pack_reference_layout(w, sc, z, cfg)
    assert np.array_equal(A, g["A"]) and np.array_equal(S, g["S"])
    # preprocessor: bit exact
    q, ls, lb = oracle.preprocessor(x, cfg.act_group_size)
    assert np.array_equal(q, g["qlut"])
    assert np.array_equal(ls.view(np.uint32), g["lut_scales"].view(np.uint32))
    assert np.array_equal(lb.view(np.uint32), g["lut_biases"].view(np.uint32))
    # qgemm: bit exact on both paths (same op order, same FMA placement as the AVX2 kernels)
    Cout = oracle.qgemm(cfg, A, S, q, ls, lb)
    assert np.array_equal(Cout.view(np.uint32), g["C"].view(np.uint32))
    assert np.array_equal(oracle.cbits(cfg, A, q), g["cbits"])
    # reference's own sanity gate: NMSE <= 5e-4 vs dense dequant (python/t_mac/ops/qgemm.py:277-282)
    assert T.nmse(T.dense_reference(w, sc, z, x, cfg), Cout) < 5e-4


@pytest.mark.parametrize("cfg", [
    T.Config(256, 1024, 2, zero_point=True), T.Config(256, 1024, 4), T.Config(192, 512, 3, zero_point=True),
    T.Config(512, 256, 1), T.Config(320, 1280, 2, one_scale=True), T.Config(128, 512, 4, kfactor=8, group_size=32, act_group_size=32),
])
def test_oracle_vs_reference_build(oracle, cfg):
    ref = T.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    cfg = cfg.resolved()
    w, sc, z, x = T.make_problem(cfg, seed=11, N=3)
    A, S = T.pack_reference_layout(w, sc, z, cfg)
    qo, lso, lbo = oracle.preprocessor(x, cfg.act_group_size)
    qr, lsr, lbr = ref.preprocessor(x, cfg.act_group_size)
    assert np.array_equal(qo, qr) and np.array_equal(lso, lsr) and np.array_equal(lbo, lbr)
    assert np.array_equal(oracle.qgemm(cfg, A, S, qo, lso, lbo).view(np.uint32), ref.qgemm(cfg, A, S, qr, lsr, lbr).view(np.uint32))
    assert np.array_equal(oracle.cbits(cfg, A, qo), ref.cbits(cfg, A, qr))


def test_preprocessor_edge_cases(oracle):
    # all-zero activations: scale 0 -> t_scales 0 -> LUT all zero (lut_ctor.cc:124)
    q, ls, lb = oracle.preprocessor(np.zeros((1, 128), np.float32), 64)
    assert not q.any() and not ls.any() and not lb.any()
    # odd symmetry LUT[15-i] == -LUT[i] (lut_ctor.cc:153-155) and |q| <= 127
    x = np.random.default_rng(3).standard_normal((2, 256)).astype(np.float32) * 100
    q, ls, lb = oracle.preprocessor(x, 64)
    assert np.array_equal(q[..., ::-1].astype(np.int32), -q.astype(np.int32))
    assert np.abs(q.astype(np.int32)).max() == 127
    with pytest.raises(ValueError):
        oracle.preprocessor(np.zeros((1, 96), np.float32), 64)  # K % act_group_size != 0
