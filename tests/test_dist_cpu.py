"""world_size-2 gloo tests (CPU): the host-side logic of the N>1 paths — the enumeration shard merge
(fplll_b200/dist.py) and bench.py's rank bookkeeping — without any GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fplll_b200.dist import merge_enum_results
    d = 6
    # rank 0 found a vector of length 10, rank 1 a shorter one (7) — and an equal-length competitor for the tie-break
    local = {0: dict(solutions=[(12.0, np.arange(d)), (10.0, np.ones(d))], nodes=np.arange(d, dtype=np.uint64),
                     stats=dict(leaves=3)),
             1: dict(solutions=[(7.0, np.array([0, 0, 1, -1, 0, 2.0]))], nodes=10 * np.ones(d, dtype=np.uint64),
                     stats=dict(leaves=4))}[rank]
    m = merge_enum_results(local, d)
    # second scenario: nobody found anything
    e = merge_enum_results(dict(solutions=[], nodes=np.zeros(d, np.uint64), stats={}), d)
    # third: a tie in length -> lexicographically smaller vector wins on every rank
    t = merge_enum_results(dict(solutions=[(5.0, np.array([1.0, 0, 0, 0, 0, rank]))], nodes=np.zeros(d, np.uint64),
                                stats={}), d)
    q.put((rank, m["solutions"][0][0], m["solutions"][0][1].tolist(), m["nodes"].tolist(), m["leaves"],
           len(e["solutions"]), t["solutions"][0][1].tolist()))
    dist.destroy_process_group()


def test_enum_shard_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dd, x, nodes, leaves, nempty, tie in outs:
        assert dd == 7.0 and x == [0, 0, 1, -1, 0, 2.0]          # the shorter vector, on both ranks
        assert nodes == [10 + k for k in range(6)] and leaves == 7  # sums
        assert nempty == 0
        assert tie == [1.0, 0, 0, 0, 0, 0]                          # deterministic tie-break
