"""A/B of the lone-launch chain (bench workload: 32 dependent-style launches per step, CUDA graph + PDL, next-tensor
L2 hints) over cluster size / warps per CTA / register variant / PDL trigger, fused (tmac_b200_gemv) and qgemm_lut only."""
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
HINT = os.environ.get("SWEEP_HINT", "1") == "1"
def fused():
    for i, wt in enumerate(layers):
        if HINT: lib.tmac_b200_hint_next_weights(layers[(i + 1) % L].handle)
        tb.gemv(wt, 1, x[i], out[i])
def plain():
    for i, wt in enumerate(layers):
        if HINT: lib.tmac_b200_hint_next_weights(layers[(i + 1) % L].handle)
        tb.qgemm_lut(wt, 1, q[i], ls[i], lb[i], out[i])
def time_chain(fn):
    fn(); tb.check(lib.tmac_b200_sync(), "sync")
    tb.check(lib.tmac_b200_graph_begin(), "gb"); fn(); g = lib.tmac_b200_graph_end(); tb.check(g, "ge")
    tb.check(lib.tmac_b200_graph_launch(g, 5), "warm"); tb.check(lib.tmac_b200_sync(), "sync")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); tb.check(lib.tmac_b200_graph_launch(g, 30), "run"); e1.record(st); torch.cuda.synchronize()
    lib.tmac_b200_graph_free(g)
    return e0.elapsed_time(e1) * 1e3 / 30 / L
cases = [(0, 0, 0, -1)]
for cs, wpc in [(4, 8), (8, 4), (2, 8), (8, 8), (4, 4)]:
    for minb in (3, 4):
        for late in (0, 1):
            cases.append((cs, wpc, minb, late))
for cs, wpc, minb, late in cases:
    tb.debug_set("cs", cs); tb.debug_set("wpc", wpc); tb.debug_set("minb", minb); tb.debug_set("pdl_late", late)
    try:
        tf = time_chain(fused); ll = tb.last_launch(); tp = time_chain(plain)
        print("cs %d wpc %d minb %d late %2d -> fused %.3f us  plain %.3f us   grid %d bpw %d" % (cs, wpc, minb, late, tf, tp, ll["grid_x"], ll["chunks_per_warp"]), flush=True)
    except Exception as e:
        print("cs %d wpc %d minb %d late %d -> %s" % (cs, wpc, minb, late, e), flush=True)
