"""First-contact check of the DRAFT fp16-operand prefill tile (tmac_prefill16.cuh, opt-in knob "prefill16"): parity against
the oracle at north_star's tolerance (1e-3 of max|C|; the simulation predicts ~1.4e-4) on small shapes, then timing on the bench
shape next to the int8 tile.  Not part of tests/ until it has passed on hardware.

    python tools/pf16_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import tmac_b200 as tb            # noqa: E402
import tmac_oracle as T           # noqa: E402

lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
oracle = T.load_oracle()


def run(cfg, N, knob):
    tb.debug_set("prefill16", knob)
    w, sc, z, x = T.make_problem(cfg, seed=11, N=N)
    wt = tb.upload_plain(tb.make_kcfg(cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size, cfg.zero_point, cfg.one_scale), w, sc, z)
    dx = torch.from_numpy(x).cuda(); out = torch.zeros((N, cfg.Mout), device="cuda")
    tb.gemv(wt, N, dx, out); torch.cuda.synchronize()
    ll = tb.last_launch()
    A, S = T.pack_reference_layout(w, sc, z, cfg)
    qo, lso, lbo = oracle.preprocessor(x, cfg.act_group_size)
    Co = oracle.qgemm(cfg, A, S, qo, lso, lbo)
    err = np.abs(out.cpu().numpy() - Co).max() / np.abs(Co).max()
    wt.free()
    return err, ll


for cfg, N in ((T.Config(256, 1024, 2, zero_point=True), 64), (T.Config(384, 2048, 2), 130), (T.Config(512, 4096, 2, zero_point=True), 256),
               (T.Config(192, 512, 2, bm=128, zero_point=True), 300)):
    cfg = cfg.resolved()
    e16, ll = run(cfg, N, 1)
    e8, _ = run(cfg, N, 0)
    print("%dx%d N=%d: fp16 tile rel err %.2e (%s)   int8 tile %.2e   launch %s" % (cfg.Mout, cfg.K, N, e16, "OK" if e16 <= 1e-3 else "FAIL", e8, ll), flush=True)

import bench                      # noqa: E402
w, sc, z = bench.synth(9)
wt = tb.upload_plain(tb.make_kcfg(bench.MOUT, bench.K, bench.BITS, 128, 16, bench.GS, bench.AGS, bench.ZP, False), w, sc, z)
for NB in (128, 256, 512, 1024):
    xb = torch.randn((NB, bench.K), device="cuda"); ob = torch.zeros((NB, bench.MOUT), device="cuda")
    for knob in (0, 1):
        tb.debug_set("prefill16", knob)
        tb.gemv(wt, NB, xb, ob); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            tb.gemv(wt, NB, xb, ob)
        e1.record(st); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        print("N=%d %s tile: %.1f us  (%.0f dense-equivalent TFLOP/s)" % (NB, "fp16" if knob else "int8", us, 2.0 * NB * bench.MOUT * bench.K / us / 1e6), flush=True)
tb.debug_set("prefill16", 0)
