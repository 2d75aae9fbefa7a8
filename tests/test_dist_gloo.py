"""world_size-2 gloo test of the N>1 host logic (sharding + counter reduction) -- CPU only."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from ldpc_3gpp_matlab_b200 import dist as D
    r, lr, w = D.init("gloo")
    lo, hi = D.shard_range(4097, r, w)
    c = D.sum_counters([hi - lo, r + 1, 10 * (r + 1), 0])
    m = D.max_over_ranks(1.5 + r)
    D.barrier()
    q.put((r, lo, hi, c.tolist(), m, D.rank_seed(0, r)))
    torch.distributed.destroy_process_group()


def test_shard_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, lo0, hi0, c0, m0, s0), (r1, lo1, hi1, c1, m1, s1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 2049, 2049, 4097)
    assert c0 == c1 == [4097, 3, 30, 0]
    assert m0 == m1 == 2.5
    assert s0 != s1


def test_shard_range_covers_everything():
    from ldpc_3gpp_matlab_b200.dist import shard_range
    for total in (0, 1, 7, 4096, 65536 + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
