"""Debug: steady-state timeline (globaltimer ns) of 8 consecutive gemv3 launches replayed in a CUDA graph.
Needs a library built with -DTMAC_ENABLE_TRACE (tools/libtmac_trace.so) and TMAC_B200_TRACE=1."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TMAC_B200_TRACE"] = "1"
os.environ["TMAC_B200_LIB"] = os.path.join(ROOT, "tools", "libtmac_trace.so")
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch
import tmac_b200 as tb
import bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
_st = torch.cuda.Stream(); torch.cuda.set_stream(_st); tb.check(lib.tmac_b200_set_stream(_st.cuda_stream), "set_stream")
w, sc, z = bench.synth(1)
cfg = tb.make_kcfg(bench.MOUT, bench.K, 2, 128, 16, 128, 64, True, False)
base = tb.upload_plain(cfg, w, sc, z)
layers = [base] + [tb.clone(base) for _ in range(15)]
x = torch.randn((1, bench.K), device="cuda")
q = torch.zeros((1, bench.K // 4, 16), dtype=torch.int8, device="cuda")
ls = torch.zeros((1, 64), device="cuda"); lb = torch.zeros_like(ls); out = torch.zeros((16, bench.MOUT), device="cuda")
tb.preprocessor(bench.K, 1, 64, x, ls, lb, q)
def step():
    for i, wt in enumerate(layers):
        if os.environ.get("TRACE_FUSED", "0") == "1":
            tb.gemv(wt, 1, x, out[i])
        else:
            tb.qgemm_lut(wt, 1, q, ls, lb, out[i])
step(); tb.check(lib.tmac_b200_sync(), "sync")
tb.check(lib.tmac_b200_graph_begin(), "gb"); step(); g = lib.tmac_b200_graph_end(); tb.check(g, "ge")
tb.check(lib.tmac_b200_graph_launch(g, 3), "run"); tb.check(lib.tmac_b200_sync(), "sync")
buf = np.zeros((8 * 4096, 8), np.int64)
nc = lib.tmac_b200_debug_trace(buf.ctypes.data, 8 * 4096)
t = buf[:8 * nc].reshape(8, nc, 8).astype(np.float64)
t0 = t[:, :, 0].min()
names = ["entry", "copies issued", "pdl wait done", "lut+data ready", "loop done", "cta reduced", "cluster synced", "stored(leader)"]
FUSED = os.environ.get("TRACE_FUSED", "0") == "1"
order = np.argsort(t[:, :, 0].min(axis=1))
print("ctas per launch", nc, "; times in ns relative to the first CTA entry of the oldest launch in the ring")
for li in order:
    row = []
    for s in range(8):
        v = t[li, :, s]; v = v[v > 0]
        row.append("%s %6.0f..%6.0f" % (names[s][:10], v.min() - t0, v.max() - t0) if len(v) else "%s   -" % names[s][:10])
    print(" | ".join(row))
