"""CPU suite, part 3: the N>1 path with world_size 2 over gloo.
Each rank owns a row shard (whole reference tiles); the per-rank compute is played by the oracle
(this is a test: the product path launches the CUDA kernels on its shard), the slices are
all-gathered and must equal the un-sharded result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import shard
import tmac_oracle as T


def test_row_partition_handles_remainders():
    parts = shard.row_partition(11008, 64, 8)       # 172 tiles over 8 ranks
    assert [r for _, r in parts] == [22 * 64] * 4 + [21 * 64] * 4
    assert parts[0][0] == 0 and sum(r for _, r in parts) == 11008
    assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(7))
    assert shard.row_partition(4096, 64, 8) == [(i * 512, 512) for i in range(8)]
    assert shard.row_partition(128, 64, 4) == [(0, 64), (64, 64), (128, 0), (128, 0)]
    with pytest.raises(ValueError):
        shard.row_partition(100, 64, 2)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = T.load_oracle()
        cfg = T.Config(384, 512, 2, bm=128, zero_point=True).resolved()     # 6 tiles of 64 rows -> 3 + 3
        w, sc, z, x = T.make_problem(cfg, seed=5, N=2)
        parts = shard.row_partition(cfg.Mout, cfg.bm // cfg.bits, world)
        row0, rows = parts[rank]
        sub = T.Config(rows, cfg.K, cfg.bits, bm=cfg.bm, zero_point=True).resolved()
        A, S = T.pack_reference_layout(w[row0:row0 + rows], sc[row0:row0 + rows], z[row0:row0 + rows], sub)
        q, ls, lb = oracle.preprocessor(x, cfg.act_group_size)             # every rank builds its own LUT
        local = torch.from_numpy(oracle.qgemm(sub, A, S, q, ls, lb))
        full = shard.gather_rows(local, cfg.Mout, parts)
        Af, Sf = T.pack_reference_layout(w, sc, z, cfg)
        want = oracle.qgemm(cfg, Af, Sf, q, ls, lb)
        ret[rank] = bool(np.array_equal(full.numpy().view(np.uint32), want.view(np.uint32)))
    finally:
        dist.destroy_process_group()


def test_sharded_gemv_allgather_matches_unsharded():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
