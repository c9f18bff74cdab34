import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TMAC_B200_TRACE"] = "1"; os.environ["TMAC_B200_LIB"] = os.path.join(ROOT, "tools", "libtmac_trace.so")
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch, numpy as np
import tmac_b200 as tb, bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "ss")
w, sc, z = bench.synth(9)
cfg = tb.make_kcfg(bench.MOUT, bench.K, 2, 128, 16, 128, 64, True, False)
wt = tb.upload_plain(cfg, w, sc, z)
N = 128
x = torch.randn((N, bench.K), device="cuda"); o = torch.zeros((N, bench.MOUT), device="cuda")
q = torch.zeros((N, bench.K // 4, 16), dtype=torch.int8, device="cuda"); ls = torch.zeros((N, 64), device="cuda"); lb = torch.zeros_like(ls)
tb.preprocessor(bench.K, N, 64, x, ls, lb, q)
for _ in range(2): tb.qgemm_lut(wt, N, q, ls, lb, o)
torch.cuda.synchronize()
buf = np.zeros((8 * 12, 8), np.int64)
rc = lib.tmac_b200_debug_trace(buf.ctypes.data, 8 * 12); print('trace rc', rc, tb.last_error() if rc < 0 else '', 'lib', tb.LIB_PATH, 'launch', tb.last_launch(), 'nonzero', int((buf != 0).sum()))
t = buf.reshape(-1)[:3 * 32 * 4].reshape(3, 32, 4).astype(np.float64)
t0 = t[t > 0].min() if (t > 0).any() else 0
names = [("producer", ["loop top", "request done", "expand done", "data landed"]), ("mma", ["loop top", "full ok", "accempty ok", "committed"]), ("epilogue", ["loop top", "accfull ok", "fma done", "-"])]
for role in range(3):
    print(names[role][0], names[role][1])
    for step in list(range(0, 12)) + [20, 21, 30, 31]:
        print("  step %2d: " % step + "  ".join("%8.0f" % (v - t0) if v > 0 else "       -" for v in t[role, step]))
