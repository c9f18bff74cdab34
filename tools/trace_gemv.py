"""Debug helper: per-phase clock64 timeline of gemv2_kernel (TMAC_B200_TRACE=1)."""
import os, sys
import numpy as np
os.environ["TMAC_B200_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch
import tmac_b200 as tb
import bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
_st = torch.cuda.Stream(); torch.cuda.set_stream(_st); tb.check(lib.tmac_b200_set_stream(_st.cuda_stream), "set_stream")
w, sc, z = bench.synth(1)
cfg = tb.make_kcfg(bench.MOUT, bench.K, 2, 128, 16, 128, 64, True, False)
wt = tb.upload_plain(cfg, w, sc, z)
x = torch.randn((1, bench.K), device="cuda")
q = torch.zeros((1, bench.K // 4, 16), dtype=torch.int8, device="cuda")
ls = torch.zeros((1, 64), device="cuda"); lb = torch.zeros_like(ls); out = torch.zeros((1, bench.MOUT), device="cuda")
tb.preprocessor(bench.K, 1, 64, x, ls, lb, q)
for it in range(3):
    tb.qgemm_lut(wt, 1, q, ls, lb, out)
    tb.check(lib.tmac_b200_sync(), "sync")
buf = np.zeros((4096, 8), np.int64)
n = lib.tmac_b200_debug_trace(buf.ctypes.data, 4096)
t = buf[:n].astype(np.float64)
d = t - t[:, :1]
d[d < 0] = np.nan
import warnings; warnings.simplefilter("ignore")
names = ["entry", "loads issued", "pdl wait done", "lut staged", "loop done", "cta reduced", "cluster synced"]
g0 = buf[:n, 7].astype(np.float64); g0 -= g0.min()
print("CTA entry time spread (globaltimer ns): mean %.0f  p50 %.0f  p90 %.0f  max %.0f" % (g0.mean(), np.percentile(g0, 50), np.percentile(g0, 90), g0.max()))
print("ctas", n, "SM cycles since entry (mean / min / max):")
for i, nm in enumerate(names[:7]):
    print("  %-12s %8.0f %8.0f %8.0f" % (nm, np.nanmean(d[:, i]), np.nanmin(d[:, i]), np.nanmax(d[:, i])))
