"""A/B of the grouped launch (32 GEMVs per launch, bench workload) over stage buffers / warps per CTA / register variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, ROOT)
import torch
import tmac_b200 as tb
import bench
lib = tb.load(); tb.check(lib.tmac_b200_init(0), "init")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
w, sc, z = bench.synth(1)
cfg = tb.make_kcfg(bench.MOUT, bench.K, 2, 128, 16, 128, 64, True, False)
base = tb.upload_plain(cfg, w, sc, z)
L = 32
layers = [base] + [tb.clone(base) for _ in range(L - 1)]
x = torch.randn((L, bench.K), device="cuda")
q = torch.zeros((L, bench.K // 4, 16), dtype=torch.int8, device="cuda")
ls = torch.zeros((L, 64), device="cuda"); lb = torch.zeros_like(ls); out = torch.zeros((L, bench.MOUT), device="cuda")
for i in range(L):
    tb.preprocessor(bench.K, 1, 64, x[i], ls[i], lb[i], q[i])
def call():
    tb.qgemm_lut_grouped(layers, 1, [q[i] for i in range(L)], [ls[i] for i in range(L)], [lb[i] for i in range(L)], [out[i] for i in range(L)])
ref = None
for nbuf, wpc, minb in [(0, 0, 0), (1, 0, 0), (1, 0, 3), (2, 4, 0), (1, 4, 0), (1, 4, 3), (2, 0, 3)]:
    tb.debug_set("nbuf", nbuf); tb.debug_set("wpc", wpc); tb.debug_set("minb", minb)
    call(); tb.check(lib.tmac_b200_sync(), "sync")
    ll = tb.last_launch()
    o = out.clone()
    if ref is None: ref = o
    same = bool(torch.equal(o, ref))
    tb.check(lib.tmac_b200_graph_begin(), "gb"); call(); g = lib.tmac_b200_graph_end(); tb.check(g, "ge")
    tb.check(lib.tmac_b200_graph_launch(g, 5), "warm"); tb.check(lib.tmac_b200_sync(), "sync")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); tb.check(lib.tmac_b200_graph_launch(g, 30), "run"); e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 30 / L
    print("nbuf %d wpc %d minb %d -> %.3f us/GEMV  %.0f GB/s  same=%s  %s" % (nbuf, wpc, minb, us, 12741632 / us / 1e3, same, ll))
