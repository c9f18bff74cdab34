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
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
