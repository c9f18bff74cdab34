"""make_golden.py -- regenerates tests/golden/*.npz (TEST INFRASTRUCTURE ONLY).

Runs IN THE BUILD CONTAINER only: imports the reference's own python/t_mac/weights.py by
file path (importing the t_mac package would pull in TVM) and calls the reference's own SIMD
kernels through oracle/_ref/libtmac_ref.so (built by oracle/Makefile from
/root/reference/python/t_mac/intrins).  The fixtures pin the oracle restatement and the CUDA
path on machines where /root/reference does not exist (the GPU box).

    python oracle/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tmac_oracle as T  # noqa: E402

REF_WEIGHTS = "/root/reference/python/t_mac/weights.py"
OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    # name: (Config, seed, N)
    "w2_zp_g128": (T.Config(128, 512, 2, zero_point=True), 1, 1),
    "w4_sym_g128": (T.Config(64, 512, 4), 2, 2),
    "w4_zp_g128": (T.Config(128, 256, 4, zero_point=True), 3, 1),
    "w3_sym_g128": (T.Config(128, 256, 3), 4, 1),
    "w1_zp_g128": (T.Config(256, 256, 1, zero_point=True), 5, 1),
    "w2_bitnet_int32": (T.Config(160, 640, 2, one_scale=True), 6, 2),
    "w4_q40_g32": (T.Config(64, 256, 4, kfactor=8, group_size=32, act_group_size=32), 7, 1),
    "w2_zp_bm128_kf8": (T.Config(128, 512, 2, bm=128, kfactor=8, group_size=64, act_group_size=32, zero_point=True), 8, 1),
}


def main():
    spec = importlib.util.spec_from_file_location("ref_weights", REF_WEIGHTS)
    refw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refw)
    ref = T.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    os.makedirs(OUT, exist_ok=True)
    for name, (cfg, seed, N) in CASES.items():
        cfg = cfg.resolved()
        w, sc, z, x = T.make_problem(cfg, seed, N)
        A, S = refw.preprocess_weights(w, sc, None if cfg.one_scale else z, bits=cfg.bits, bm=cfg.bm, kfactor=cfg.kfactor)
        A = np.ascontiguousarray(A, np.uint8)
        S = np.ascontiguousarray(S, np.float32).reshape(-1)
        qlut, ls, lb = ref.preprocessor(x, cfg.act_group_size)
        Cout = ref.qgemm(cfg, A, S, qlut, ls, lb)
        cbits = ref.cbits(cfg, A, qlut)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            cfg=np.array([cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size,
                          int(cfg.zero_point), int(cfg.one_scale)], np.int64),
            w=w, scales=sc, zeros=(z if z is not None else np.zeros(0, np.float32)), x=x,
            A=A, S=S, qlut=qlut, lut_scales=ls, lut_biases=lb, C=Cout, cbits=cbits)
        print(name, "C[0,:4] =", Cout[0, :4])
    # known-answer vector of tests/test_lut_ctor.cc:12-24 (b[i] = i, i < 32), produced by the reference build
    b = np.arange(32, dtype=np.float32)
    ls = ref.partial_max(b, 0.0)
    q, s, lbias = ref.lut_ctor(b, ls)
    np.savez_compressed(os.path.join(OUT, "kat_lut_ctor.npz"), b=b, lut_scales=np.float32(ls), lut_biases=np.float32(lbias), qlut=q)
    print("kat", ls, lbias)


if __name__ == "__main__":
    main()
