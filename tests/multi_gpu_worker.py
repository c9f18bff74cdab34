"""Worker of tests/test_gpu_multi.py, launched with torch.distributed.run (one process per GPU, NCCL for the rendezvous).

Row-sharded GEMV with the all-gather fused into the kernel epilogue (tmac_b200_peer_outputs): every rank computes its
row shard (whole reference tiles, remainders spread) and stores the finished rows into every rank's output
vector over NVLink -- no collective launch.  Checks, on every rank: gathered vector == the un-sharded single-GPU result
BIT FOR BIT (same K decomposition pinned on both sides), and against the oracle (int32 path exact, fp path 2e-5)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "t-mac_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmac_b200 as tb            # noqa: E402
import tmac_oracle as T           # noqa: E402
from shard import row_partition   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = tb.load(); tb.check(lib.tmac_b200_init(local), "init")
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); tb.check(lib.tmac_b200_set_stream(st.cuda_stream), "set_stream")
    oracle = T.load_oracle()
    cases = [T.Config(11008, 4096, 2, zero_point=True), T.Config(4096, 4096, 4), T.Config(8640, 3200, 2, one_scale=True),
             T.Config(640, 1024, 2, bm=128, zero_point=True)]
    tb.debug_set("cs", 4); tb.debug_set("wpc", 8)        # one K decomposition on both sides: identical fp32 summation order per row
    for cfg in cases:
        cfg = cfg.resolved()
        w, sc, z, x = T.make_problem(cfg, seed=77)
        parts = row_partition(cfg.Mout, cfg.bm // cfg.bits, world)
        row0, rows = parts[rank]
        k_full = tb.make_kcfg(cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size, cfg.zero_point, cfg.one_scale)
        sv = tb.SharedVector(cfg.Mout, dist, rank, world)
        dx = torch.from_numpy(x).cuda()
        shard_wt = None
        if rows > 0:
            shard_wt = tb.upload_plain(k_full, w, sc, z, row0=row0, rows=rows)      # only this rank's rows become resident
        for rep in range(2):
            sv.local.zero_()
            sv.barrier()                                  # nobody stores into a vector that is still being cleared
            if rows > 0:
                tb.peer_outputs([sv.peer_ptr(q) + 4 * row0 for q in range(world) if q != rank])
                tb.gemv(shard_wt, 1, dx, sv.local[row0:row0 + rows])
            sv.barrier()                                  # stream-ordered: the producing launches of every rank are complete
            got = sv.local.clone()                        # (no host synchronisation between the GEMV and this read)
            torch.cuda.synchronize()
        full_wt = tb.upload_plain(k_full, w, sc, z)
        ref_gpu = torch.zeros((1, cfg.Mout), device="cuda")
        tb.gemv(full_wt, 1, dx, ref_gpu)
        torch.cuda.synchronize()
        assert torch.equal(got, ref_gpu[0]), "rank %d: gathered != un-sharded (%dx%d)" % (rank, cfg.Mout, cfg.K)
        A, S = T.pack_reference_layout(w, sc, z, cfg)
        q, ls, lb = oracle.preprocessor(x, cfg.act_group_size)
        Co = oracle.qgemm(cfg, A, S, q, ls, lb)[0]
        g = got.cpu().numpy()
        if cfg.one_scale:
            assert np.array_equal(g.view(np.uint32), Co.view(np.uint32)), "int32 path must be bit exact"
        else:
            assert np.abs(g - Co).max() <= 2e-5 * np.abs(Co).max()
        if rank == 0:
            print("OK %dx%d w%d%s world %d parts %s" % (cfg.Mout, cfg.K, cfg.bits, " bitnet" if cfg.one_scale else "", world, [r for _, r in parts]), flush=True)
        full_wt.free()
        if shard_wt is not None:
            shard_wt.free()
        sv.close(dist)
    tb.debug_set("cs", 0); tb.debug_set("wpc", 0)

    # ---- decode SEQUENCE with peer outputs (tmac_b200_seq_peer_outputs): every rank runs its own dependent chain (op 1 reads op 0's
    #      rows) in one persistent launch and stores the rows of both ops into every rank's [world][2][Mout] buffer from the kernel's
    #      epilogue; one flag exchange.  Gathered == an NCCL all-gather of the ranks' rows, == the peer-less sequence, and the oracle.
    cfgs = [T.Config(1024, 1024, 2, zero_point=True).resolved(), T.Config(1024, 512, 2, zero_point=True).resolved()]
    M = 1024
    probs = [T.make_problem(c, seed=200 + 10 * rank + i) for i, c in enumerate(cfgs)]      # different weights on every rank
    wts = [tb.upload_plain(tb.make_kcfg(c.Mout, c.K, c.bits, c.bm, c.kfactor, c.group_size, c.act_group_size, c.zero_point, c.one_scale), w, sc, z)
           for c, (w, sc, z, _) in zip(cfgs, probs)]
    sv = tb.SharedVector(world * 2 * M, dist, rank, world)
    gathered = sv.local.view(world, 2, M)
    mine = gathered[rank]
    plain = torch.zeros((2, M), device="cuda")
    dx = torch.from_numpy(probs[0][3][0]).cuda()
    seqs = []
    for dst, peers in ((mine, True), (plain, False)):
        sq = tb.Sequence()
        sq.add(wts[0], x=dx, out=dst[0])
        sq.add(wts[1], in_op=0, in_offset=0, out=dst[1])
        if peers:
            for i in range(2):
                sq.peer_outputs(i, [sv.peer_ptr(q) + 4 * (rank * 2 + i) * M for q in range(world) if q != rank])
        sq.build()
        assert sq.info()["ring_slots"] < 0, "the resident chain kernel must take this sequence"
        seqs.append(sq)
    for rep in range(2):
        sv.local.zero_()
        sv.barrier()
        seqs[0].launch()
        sv.barrier()
        got = gathered.clone()
        torch.cuda.synchronize()
    seqs[0].status()
    seqs[1].launch(); seqs[1].status()
    ref = torch.zeros((world, 2, M), device="cuda")
    dist.all_gather_into_tensor(ref.view(-1), plain.reshape(-1))
    torch.cuda.synchronize()
    assert torch.equal(got, ref), "rank %d: sequence gather != all-gather of the peer-less sequences" % rank
    o = got[rank].cpu().numpy()
    xin = probs[0][3][0]
    for i, c in enumerate(cfgs):
        w, sc, z, _ = probs[i]
        A, S = T.pack_reference_layout(w, sc, z, c)
        q, ls, lb = oracle.preprocessor(xin[None, :c.K], c.act_group_size)
        Co = oracle.qgemm(c, A, S, q, ls, lb)[0]
        assert np.abs(o[i] - Co).max() <= 2e-5 * np.abs(Co).max(), "rank %d op %d" % (rank, i)
        xin = o[i]
    if rank == 0:
        print("OK sequence with peer outputs, world %d" % world, flush=True)
    for sq in seqs:
        sq.free()
    for wt in wts:
        wt.free()
    sv.close(dist)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
