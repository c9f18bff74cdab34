"""Debug: steady-state timeline (globaltimer ns) of consecutive integer-path (BitNet grouping) launches replayed in a CUDA graph.
Needs a library built with -DTMAC_ENABLE_TRACE (tools/libtmac_trace.so: TMAC_B200_OUT=tools/libtmac_trace.so ./build.sh -DTMAC_ENABLE_TRACE).
    python tools/trace_int.py MOUT K [fused]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TMAC_B200_TRACE"] = "1"
os.environ["TMAC_B200_LIB"] = os.path.join(ROOT, "tools", "libtmac_trace.so")
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch
import tmac_b200 as tb
import bench
mout, k = int(sys.argv[1]), int(sys.argv[2]); fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
_st = torch.cuda.Stream(); torch.cuda.set_stream(_st); tb.check(lib.tmac_b200_set_stream(_st.cuda_stream), "set_stream")
w, sc, z = bench.synth(7, mout, k, 2, 128, False, True)
bm = 256 if (mout * 2) % 256 == 0 else (128 if (mout * 2) % 128 == 0 else 320)
cfg = tb.make_kcfg(mout, k, 2, bm, 16, 128, k, False, True)
base = tb.upload_plain(cfg, w, sc, z)
layers = [base] + [tb.clone(base) for _ in range(15)]
x = torch.randn((1, k), device="cuda"); out = torch.zeros((1, mout), device="cuda")
q = torch.zeros((1, k // 4, 16), dtype=torch.int8, device="cuda"); ls = torch.zeros((1, 1), device="cuda"); lb = torch.zeros_like(ls)
tb.preprocessor(k, 1, k, x, ls, lb, q)
def step():
    for wt in layers:
        if fused:
            tb.gemv(wt, 1, x, out)
        else:
            tb.qgemm_lut(wt, 1, q, ls, lb, out)
step(); tb.check(lib.tmac_b200_sync(), "sync")
tb.check(lib.tmac_b200_graph_begin(), "gb"); step(); g = lib.tmac_b200_graph_end(); tb.check(g, "ge")
tb.check(lib.tmac_b200_graph_launch(g, 3), "run"); tb.check(lib.tmac_b200_sync(), "sync")
buf = np.zeros((8 * 4096, 8), np.int64)
nc = lib.tmac_b200_debug_trace(buf.ctypes.data, 8 * 4096)
t = buf[:8 * nc].reshape(8, nc, 8).astype(np.float64)
t0 = t[:, :, 0].min()
names = ["entry", "copies issued", "pdl wait done", "lut+data ready", "loop done", "cta reduced", "cluster synced", "stored(leader)"]
order = np.argsort(t[:, :, 0].min(axis=1))
print(mout, k, "fused" if fused else "qgemm only", tb.last_launch(), "ctas per launch", nc, "; ns relative to the first CTA entry of the oldest launch in the ring")
for li in order:
    row = []
    for s in range(8):
        v = t[li, :, s]; v = v[v > 0]
        row.append("%s %6.0f..%6.0f" % (names[s][:10], v.min() - t0, v.max() - t0) if len(v) else "%s   -" % names[s][:10])
    print(" | ".join(row))
