import os, sys
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch, tmac_b200 as tb, bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
w, sc, z = bench.synth(9)
wt = tb.upload_plain(tb.make_kcfg(bench.MOUT, bench.K, bench.BITS, 128, 16, bench.GS, bench.AGS, bench.ZP, False), w, sc, z)
for NB in (256, 512):
    xb = torch.randn((NB, bench.K), device="cuda"); ob = torch.zeros((NB, bench.MOUT), device="cuda")
    for name, p16, sk in (("int8", 0, 0), ("fp16 tile/CTA", 1, 0), ("fp16 stream-K", 1, 1)):
        tb.debug_set("prefill16", p16); tb.debug_set("pf_streamk", sk)
        tb.gemv(wt, NB, xb, ob); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            tb.gemv(wt, NB, xb, ob)
        e1.record(st); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print("N=%d %-14s %.1f us  %.0f dense-equivalent TFLOP/s  %s" % (NB, name, us, 2.0 * NB * bench.MOUT * bench.K / us / 1e6, tb.last_launch()), flush=True)
