import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch, numpy as np
import tmac_b200 as tb, bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "ss")
w, sc, z = bench.synth(9)
cfg = tb.make_kcfg(bench.MOUT, bench.K, 2, 128, 16, 128, 64, True, False)
wt = tb.upload_plain(cfg, w, sc, z)
for N in (32, 64, 128, 256, 512):
    x = torch.randn((N, bench.K), device="cuda"); o = torch.zeros((N, bench.MOUT), device="cuda")
    q = torch.zeros((N, bench.K // 4, 16), dtype=torch.int8, device="cuda"); ls = torch.zeros((N, 64), device="cuda"); lb = torch.zeros_like(ls)
    tb.preprocessor(bench.K, N, 64, x, ls, lb, q)
    for name, fn in (("pre+tile", lambda: tb.gemv(wt, N, x, o)), ("tile only", lambda: tb.qgemm_lut(wt, N, q, ls, lb, o))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5): fn()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("N=%4d %-10s %8.1f us   dense-equivalent %7.1f TFLOP/s   int8 MMA rate %6.1f TOP/s  launch=%s" % (N, name, ms * 1e3, 2.0 * N * bench.MOUT * bench.K / ms / 1e9, 2.0 * ((N + 127) // 128 * 128) * 11008 * (2 * bench.K) / ms / 1e9, tb.last_launch()["batch"]))
