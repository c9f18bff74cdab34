"""Profiling target: a few grouped launches (32 GEMVs per launch, bench workload).  Run under ncu -k regex:gemv3_kernel."""
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
for _ in range(4):
    tb.qgemm_lut_grouped(layers, 1, [q[i] for i in range(L)], [ls[i] for i in range(L)], [lb[i] for i in range(L)], [out[i] for i in range(L)])
tb.check(lib.tmac_b200_sync(), "sync")
print("ok", tb.last_launch())
