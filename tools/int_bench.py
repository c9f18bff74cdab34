"""Integer-path (BitNet grouping: one activation group = the row) GEMV launch chains: LUT built inside the GEMV (tmac_b200_gemv)
against the two reference calls (preprocessor + qgemm_lut), per BitNet-3B shape, 26 distinct tensors per chain in one CUDA graph.

    python tools/int_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import tmac_b200 as tb            # noqa: E402
import bench                      # noqa: E402

lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
L = 26
for (mout, k) in ((3200, 3200), (8640, 3200), (3200, 8640)):
    w, sc, z = bench.synth(7, mout, k, 2, 128, False, True)
    bm = 256 if (mout * 2) % 256 == 0 else (128 if (mout * 2) % 128 == 0 else 320)
    cfg = tb.make_kcfg(mout, k, 2, bm, 16, 128, k, False, True)
    base = tb.upload_plain(cfg, w, sc, z)
    hs = [base] + [tb.clone(base) for _ in range(L - 1)]
    nbytes = lib.tmac_b200_weights_nbytes(base.handle)
    with torch.cuda.stream(st):
        x = torch.randn((1, k), device="cuda"); o = torch.zeros((1, mout), device="cuda")
        q = torch.zeros((1, k // 4, 16), dtype=torch.int8, device="cuda"); l1 = torch.zeros((1, 1), device="cuda"); l2 = torch.zeros_like(l1)

    def fused():
        for h in hs:
            tb.gemv(h, 1, x, o)

    def two():
        for h in hs:
            tb.preprocessor(k, 1, k, x, l1, l2, q)
            tb.qgemm_lut(h, 1, q, l1, l2, o)

    def two_shared():
        tb.preprocessor(k, 1, k, x, l1, l2, q)
        for h in hs:
            tb.qgemm_lut(h, 1, q, l1, l2, o)

    res = {}
    for name, fn in (("fused", fused), ("preprocessor+qgemm", two), ("qgemm only (LUT given)", two_shared)):
        fn(); tb.check(lib.tmac_b200_sync(), "sync")
        tb.check(lib.tmac_b200_graph_begin(), "begin"); fn(); g = lib.tmac_b200_graph_end(); tb.check(g, "end")
        tb.check(lib.tmac_b200_graph_launch(g, 3), "warm"); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); tb.check(lib.tmac_b200_graph_launch(g, 20), "run"); e1.record(st); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 / L * 1e3
        res[name] = us
        lib.tmac_b200_graph_free(g)
    print("%5dx%-5d %.1f MB: " % (mout, k, nbytes / 1e6) + " | ".join("%s %.2f us (%.0f GB/s)" % (n, u, nbytes / u / 1e3) for n, u in res.items()), tb.last_launch(), flush=True)
    for h in hs:
        h.free()
