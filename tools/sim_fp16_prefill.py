"""CPU study for the next prefill tile: how far from the reference kernel is the formulation that
folds both scales into fp16 operands -- A = 0.5*scale[m][wg]*S[m][g][e], B = lut_scale[n][ag]*T8[n][g][e] -- and accumulates the
whole K in fp32, with the LUT-bias / zero-point terms kept exact?  Uses the oracle (TEST INFRASTRUCTURE) as the truth.

    python tools/sim_fp16_prefill.py         # W2 / W4, 512 x 4096, 8 tokens: max|dC|/max|C| 1.4e-4 / 2.4e-4 (bar: 1e-3)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmac_oracle as T   # noqa: E402

o = T.load_oracle()
for bits in (2, 4):
    cfg = T.Config(512, 4096, bits, zero_point=True).resolved()
    N = 8
    w, sc, z, x = T.make_problem(cfg, seed=3, N=N)
    A, S_ = T.pack_reference_layout(w, sc, z, cfg)
    q, ls, lb = o.preprocessor(x, cfg.act_group_size)
    Cref = o.qgemm(cfg, A, S_, q, ls, lb)
    M, K = cfg.Mout, cfg.K
    G = K // 4
    wb = np.stack([(w >> b) & 1 for b in range(bits)], axis=0).astype(np.int32)
    idx = (wb.reshape(bits, M, G, 4) * np.array([1, 2, 4, 8])).sum(-1)
    e, sg = np.where(idx < 8, idx, 15 - idx), np.where(idx < 8, 1, -1)
    S = np.zeros((M, G, 8), np.float32)                       # one-hot-signed expansion: sum_b 2*alpha_b * sign * [j == e]
    for b in range(bits):
        np.add.at(S, (np.arange(M)[:, None], np.arange(G)[None, :], e[b]), (sg[b] * (1 << b)).astype(np.float32))
    T8 = q[:, :, :8].astype(np.float32)
    gs, ags = cfg.group_size, cfg.act_group_size
    s_g, ls_g = np.repeat(sc, gs // 4, axis=1), np.repeat(ls, ags // 4, axis=1)
    Aop = (0.5 * s_g[:, :, None] * S).astype(np.float16).astype(np.float32).reshape(M, -1)
    Bop = (ls_g[:, :, None] * T8).astype(np.float16).astype(np.float32).reshape(N, -1)
    LB = lb.reshape(N, K // gs, gs // ags).sum(-1)
    rest = LB.astype(np.float64) @ (0.5 * sc + z).astype(np.float64).T
    C = (Bop @ Aop.T).astype(np.float64) + rest
    print("W%d fp16 operands, fp32 accumulate: max|dC|/max|C| = %.2e   nmse %.2e" % (bits, np.abs(C - Cref).max() / np.abs(Cref).max(), T.nmse(Cref, C.astype(np.float32))))
