"""world_size-2 gloo test of the N>1 plumbing (barrier, max over ranks, batch sharding) on CPU."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from genre_shapehd_b200 import dist_util
    w, r, _ = dist_util.init(backend="gloo")
    assert (w, r) == (world, rank)
    lo, hi = dist_util.shard(65, w, r)
    dist_util.barrier()
    ms = dist_util.max_over_ranks(10.0 + 5.0 * rank)      # the slowest rank defines the step time
    total = dist_util.sum_over_ranks(hi - lo)
    out[rank] = (lo, hi, ms, total)
    dist_util.finalize()


def test_two_rank_gloo_plumbing():
    import sys
    from conftest import REPO
    sys.path.insert(0, REPO)
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][:2] == (0, 33) and res[1][:2] == (33, 65)       # contiguous, sizes differ by at most one
    assert res[0][2] == res[1][2] == 15.0
    assert res[0][3] == res[1][3] == 65.0


def test_shard_covers_everything():
    from genre_shapehd_b200 import dist_util
    for n in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [dist_util.shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
