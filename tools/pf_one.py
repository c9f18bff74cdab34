"""One prefill call (11008x4096 W2 g128 zp, N tokens) for ncu: python tools/pf_one.py [N] [prefill16] [pf_streamk]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch, tmac_b200 as tb, bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
tb.debug_set("prefill16", int(sys.argv[2]) if len(sys.argv) > 2 else 1)
tb.debug_set("pf_streamk", int(sys.argv[3]) if len(sys.argv) > 3 else 0)
w, sc, z = bench.synth(9)
wt = tb.upload_plain(tb.make_kcfg(bench.MOUT, bench.K, bench.BITS, 128, 16, bench.GS, bench.AGS, bench.ZP, False), w, sc, z)
xb = torch.randn((N, bench.K), device="cuda"); ob = torch.zeros((N, bench.MOUT), device="cuda")
for _ in range(3):
    tb.gemv(wt, N, xb, ob)
torch.cuda.synchronize()
print(tb.last_launch())
