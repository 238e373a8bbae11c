"""CPU: host logic of the NVLink-striped database (sift4g_b200/stripes.py) -- stripe boundaries, and the file-descriptor
exchange between the ranks of a box (world_size 2, gloo barriers; the descriptors here are plain pipes)."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from sift4g_b200 import stripes


def test_stripe_bounds_cover_and_align():
    G = 2 << 20
    for starts, total in (([0, 5 * G + 17, 9 * G - 3, 20 * G + 1], 31 * G + 5), ([0, 10, 20, 30, 40, 50, 60, 70], 80), ([0], 3 * G), ([0, G // 2], G)):
        P = stripes.stripe_bounds(starts, total, G)
        assert len(P) == len(starts) + 1 and P[0] == 0
        assert all(p % G == 0 for p in P) and all(b - a >= G for a, b in zip(P, P[1:]))
        assert P[-1] >= total + stripes.TAIL_PAD
        for s in range(1, len(starts)):            # the boundary sits within a granule of the shard start unless it had to move up
            assert P[s] >= starts[s] - G


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []
    for round_ in range(2):                        # two exchanges in a row (bench opens several views)
        r, w = os.pipe()
        os.write(w, b"stripe of rank %d round %d" % (rank, round_))
        os.close(w)
        fds = stripes.exchange_fds(r, rank, world, dist.barrier)
        assert fds[rank] == r and len(fds) == world
        for p in range(world):
            if p != rank:
                seen.append(os.read(fds[p], 100).decode())
                os.close(fds[p])
        os.close(r)
    q.put((rank, seen))
    dist.barrier()
    dist.destroy_process_group()


def test_fd_exchange_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == ["stripe of rank 1 round 0", "stripe of rank 1 round 1"]
    assert res[1] == ["stripe of rank 0 round 0", "stripe of rank 0 round 1"]
